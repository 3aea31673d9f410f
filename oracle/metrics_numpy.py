"""CPU oracle for the evaluation bookkeeping of evaluate-classification-metrics.py (numpy only).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as the other files in this directory).

``evaluate_model`` (code/disaster_detection/evaluate-classification-metrics.py:49-105) feeds ``output.argmax(dim=1)``
into torchmetrics ``Accuracy / F1Score / Precision / Recall (task="multiclass", num_classes=5)`` and
``ConfusionMatrix``.  torchmetrics is a third-party dependency absent from /root/reference and from this image
(requirements pin: torchmetrics in requirements-fyp.txt); its published semantics for these constructors are
``average="micro"``: with one label per sample, micro precision = micro recall = micro F1 = accuracy =
trace(cm) / sum(cm); ``ConfusionMatrix`` is cm[target, prediction].  ``compute_per_class_metrics``
(evaluate-classification-metrics.py:107-132) is restated line by line.
"""
from __future__ import annotations

import numpy as np

CLASSES = ['collapsed building', 'fire', 'flooded areas', 'normal', 'traffic incident']   # evaluate-classification-metrics.py:110


def confusion_matrix(pred, target, num_classes=5):
    cm = np.zeros((num_classes, num_classes), dtype=np.int64)
    for p, t in zip(np.asarray(pred).ravel(), np.asarray(target).ravel()):
        cm[int(t), int(p)] += 1
    return cm


def micro_metrics(cm):
    total = cm.sum()
    acc = float(np.trace(cm)) / float(total) if total else 0.0
    return {"accuracy": acc, "f1_score": acc, "precision": acc, "recall": acc}


def per_class_metrics(cm):
    out = {}
    for i, name in enumerate(CLASSES):
        tp = int(cm[i, i])
        fp = int(cm[:, i].sum()) - tp
        fn = int(cm[i, :].sum()) - tp
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0
        f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0
        out.update({f"{name}_precision": precision, f"{name}_recall": recall, f"{name}_f1": f1})
    return out
