"""Single-image and per-frame prediction (BASELINE.json configs[0] flavour): mirror of ``predict`` / ``main`` in
``code/disaster_detection/aider-predict.py:47-86,127-180`` and of ``run_inference`` in ``real-time-inference.py:80-108``
(``FrameClassifier``: pinned staging + CUDA-graph replay for a video stream).

``predict(model, image_path, transform, device)`` keeps the reference's signature and return value
``(class name, confidence in percent)``, including its quirk that the confidence is a *second* softmax over the model's
softmax output (aider-predict.py:80).  With ``transform=None`` (the fast path) the BGR frame that ``cv2.imread`` returns goes
straight to ``model.forward_frames(frame, bgr=True)``: channel swap, resize, centre crop, ToTensor and Normalize all run in the
ingest kernel on the GPU.  With a ``transform`` callable the reference's own sequence is followed (RGB -> PIL -> transform ->
reshape -> ``model(data)``).
"""
from __future__ import annotations

import argparse
import logging
import os
from typing import Tuple

import numpy as np
import torch

from .model import load_model

logger = logging.getLogger(__name__)
CLASSES = ['collapsed building', 'fire', 'flooded areas', 'normal', 'traffic incident']     # aider-predict.py:83


def predict(model, image_path: str, transform=None, device=None, use_trt: bool = False, quant: str = 'fp16') -> Tuple[str, float]:
    """Make prediction on a single image (aider-predict.py:47-86)."""
    import cv2
    image = cv2.imread(image_path)                                   # BGR, uint8 (aider-predict.py:57)
    if image is None:
        raise ValueError(f"Could not load image at {image_path}")
    device = torch.device(device) if device is not None else next(model.parameters()).device
    with torch.no_grad():
        if transform is None:
            frame = torch.from_numpy(np.ascontiguousarray(image)[None]).to(device)
            output = model.forward_frames(frame, bgr=True)
        else:
            from PIL import Image
            data = transform(Image.fromarray(cv2.cvtColor(image, cv2.COLOR_BGR2RGB)))
            hw = getattr(model, "IN_HW", 140)
            data = torch.reshape(data, (1, 3, hw, hw)).to(device)     # aider-predict.py:69-70
            if use_trt and quant == 'fp16':
                data = data.half()
            output = model(data)
        predicted_class = output.data.max(1, keepdim=True)[1]
        confidence = torch.nn.functional.softmax(output.float(), dim=1)[0][predicted_class].item() * 100
    return CLASSES[predicted_class.item()], confidence


class FrameClassifier:
    """Per-frame classification of a video stream (the loop body of real-time-inference.py:80-108) with everything
    preallocated: a pinned host staging frame, a device frame, the frames path captured into a CUDA graph
    (``model.graph_frames``) and a pinned result.  ``clf(frame)`` = one host->device copy of the frame, one graph launch,
    one 20-byte read-back.  ``frame``: uint8 (H,W,3) as OpenCV delivers it (BGR) unless ``bgr=False``."""

    def __init__(self, model, height, width, *, bgr=True, device=None):
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("FrameClassifier needs the model on a CUDA device (no CPU fallback)")
        self.shape = (int(height), int(width), 3)
        self._model, self._bgr = model, bgr              # strong reference: the captured graph points into the model's engine
        self._host = torch.empty((1,) + self.shape, dtype=torch.uint8).pin_memory()
        self._dev = torch.zeros((1,) + self.shape, dtype=torch.uint8, device=self.device)
        self._probs_host = torch.empty((1, 5), dtype=torch.float32).pin_memory()
        self._stream = torch.cuda.Stream(self.device)
        self._capture()

    def _token(self):
        """Identity of what the captured graph points at: the engine handle and the packed-weights fingerprint.
        ``_ensure_engine`` re-packs first when the weights, the precision or the device changed."""
        _lib_, h, _idx = self._model._ensure_engine()
        return (getattr(h, "value", h), getattr(self._model, "_fingerprint", None), getattr(self._model, "_loaded_version", None))

    def _capture(self):
        with torch.cuda.stream(self._stream):
            self._run = self._model.graph_frames(self._dev, bgr=self._bgr)
        self._stream.synchronize()
        self._tok = self._token()

    def probabilities(self, frame):
        """(5,) float32 numpy array of the model's softmax output for one frame."""
        arr = np.asarray(frame)
        if arr.dtype != np.uint8 or arr.shape != self.shape:
            raise ValueError(f"expected a uint8 frame of shape {self.shape}, got {arr.dtype} {arr.shape}")
        if self._token() != self._tok:                    # load_state_dict / .half() / set_precision since the capture:
            self._capture()                               # the old graph would replay against freed device memory
        self._host.numpy()[0] = arr
        with torch.cuda.stream(self._stream):
            self._dev.copy_(self._host, non_blocking=True)
            out = self._run()
            self._probs_host.copy_(out, non_blocking=True)
        self._stream.synchronize()
        return self._probs_host.numpy()[0].copy()

    def __call__(self, frame) -> Tuple[str, float]:
        p = self.probabilities(frame)
        k = int(p.argmax())                                           # output.data.max(1)[1], real-time-inference.py:97-98
        z = np.exp(p - p.max(), dtype=np.float32)
        confidence = float(z[k] / z.sum()) * 100                      # the second softmax, real-time-inference.py:101
        return CLASSES[k], confidence



def run_inference(model, frame, transform=None, input_shape=None, device=None, use_trt: bool = False, quant: str = 'fp16') -> Tuple[str, float]:
    """Run inference on a single BGR frame (real-time-inference.py:80-108; same arguments).  ``transform=None`` (the fast
    path) classifies through a cached ``FrameClassifier`` for this model and frame size."""
    if transform is not None:
        import cv2
        from PIL import Image
        hw = getattr(model, "IN_HW", 140)
        device = torch.device(device) if device is not None else next(model.parameters()).device
        tensor = transform(Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)))
        tensor = torch.reshape(tensor, input_shape or (1, 3, hw, hw)).to(device)        # real-time-inference.py:76-77
        with torch.no_grad():
            if use_trt and quant == 'fp16':
                tensor = tensor.half()
            output = model(tensor)
            predicted_class = output.data.max(1, keepdim=True)[1]
            confidence = torch.nn.functional.softmax(output.float(), dim=1)[0][predicted_class].item() * 100
        return CLASSES[predicted_class.item()], confidence
    # one runner per (model, frame size), stored ON the model: it lives exactly as long as the model does, and the runner
    # itself re-captures when the model's weights / precision change (FrameClassifier._token)
    cache = model.__dict__.setdefault("_frame_classifiers", {})
    key = tuple(np.asarray(frame).shape)
    clf = cache.get(key)
    if clf is None:
        if len(cache) > 8:
            cache.clear()
        clf = cache[key] = FrameClassifier(model, frame.shape[0], frame.shape[1], bgr=True, device=device)
    return clf(frame)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Predict disaster types from aerial images (B200 engine)')
    parser.add_argument('--model', type=str, default='ernet', choices=['ernet', 'squeeze-ernet', 'squeeze-redconv'])
    parser.add_argument('--image', type=str, required=True)
    parser.add_argument('--weights', type=str, default=None)
    parser.add_argument('--trt', action='store_true', help='also run the packed engine saved by build_engine (TRTModule)')
    parser.add_argument('--quant', type=str, default='fp16', choices=['fp16', 'bf16', 'fp32', 'int8'])
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(name)s - %(levelname)s - %(message)s')
    if not torch.cuda.is_available():
        raise RuntimeError('CUDA is not available: the B200 engine has no CPU path')
    device = torch.device('cuda')
    if not args.weights:                                              # aider-predict.py:139-146
        args.weights = {'ernet': 'weights/ernet-state_dict.pt', 'squeeze-ernet': 'weights/squeeze-ernet-state_dict.pt',
                        'squeeze-redconv': 'weights/squeeze-redconv-state_dict.pt'}[args.model]
    if not os.path.exists(args.weights):
        raise FileNotFoundError(f"No weights found at {args.weights}")
    model = load_model(args.model, args.weights, device)
    prediction, confidence = predict(model, args.image, None, device)
    logger.info(f"Prediction: {prediction} ({confidence:.1f}%)")
    if args.trt:                                                      # aider-predict.py:165-171
        from .build_engine import TRTModule
        trt_model = TRTModule()
        trt_model.load_state_dict(torch.load(f'tensorrt_state_dicts/{args.model}_{args.quant}_trt.pth', weights_only=True))
        trt_model = trt_model.to(device)
        trt_prediction, trt_confidence = predict(trt_model, args.image, None, device, True, args.quant)
        logger.info(f"Engine Prediction: {trt_prediction} ({trt_confidence:.1f}%)")
    return prediction, confidence


if __name__ == '__main__':
    main()
