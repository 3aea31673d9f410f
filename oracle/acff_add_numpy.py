"""CPU oracle for the ADD-fusion ACFF block of the detector half (numpy only).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may import it; the product path never routes through it.

Restates ``ACFF.forward`` of ``code/victim_localization/yolov3/models.py:293-315`` (modules built at
models.py:277-288): three depthwise 3x3 convolutions with dilation 1/2/3 and padding 0/1/2 (all produce H-2),
**added** (models.py:302 - the classifier's block concatenates instead, disaster_detection/model/acff.py:46), then
fused_conv 1x1 + bias (models.py:307), LeakyReLU(0.01) (models.py:308), BatchNorm2d in eval mode (models.py:309),
Dropout = identity in eval mode (models.py:310).

Parity pinning: ``tests/golden/make_golden_acff_add.py`` executes the reference's own ``class ACFF`` (its source is
read from /root/reference at generation time, never copied) in fp64 and fp32 on seeded inputs and commits the outputs
as ``tests/golden/acff_add_golden.npz``; ``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

import numpy as np

from .ernet_numpy import batch_norm_eval, conv2d_depthwise3x3, conv2d_pointwise, leaky_relu


def acff_add_depthwise(x, sd, prefix=""):
    """conv1(x) + conv2(x) + conv3(x), models.py:302 (left-to-right, like the Python expression)."""
    out = None
    for j, dil in ((1, 1), (2, 2), (3, 3)):
        y = conv2d_depthwise3x3(x, sd[f"{prefix}conv{j}.weight"], sd[f"{prefix}conv{j}.bias"], dil)
        out = y if out is None else out + y
    return out


def acff_add(x, sd, prefix=""):
    """Whole block, NCHW in -> NCHW out (models.py:293-315, eval mode)."""
    sd = {k: np.asarray(v).astype(x.dtype) if np.asarray(v).dtype.kind == "f" else np.asarray(v) for k, v in sd.items()}
    z = conv2d_pointwise(acff_add_depthwise(x, sd, prefix), sd[f"{prefix}fused_conv.weight"], sd[f"{prefix}fused_conv.bias"])
    z = leaky_relu(z)
    return batch_norm_eval(z, sd[f"{prefix}batch_norm.weight"], sd[f"{prefix}batch_norm.bias"],
                           sd[f"{prefix}batch_norm.running_mean"], sd[f"{prefix}batch_norm.running_var"])
