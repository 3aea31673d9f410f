"""CPU baseline arm: the reference's path restated with the same third-party library the reference
itself executes (torch.nn.functional on CPU -> ATen/oneDNN) plus the torchvision/Pillow eval transform.

THIS IS TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/ernet_numpy.py header).  It exists so
that `bench.py --impl reference` and the `cpu_baseline` leg time what the reference really runs on a
host CPU (stock nn.Conv2d / MaxPool2d / AvgPool2d / Linear / Softmax kernels, aider-predict.py:57-86)
rather than a slower numpy re-derivation.  The reference is Python and cannot travel to the GPU box, so
this restatement is pinned against the committed goldens (tests/test_oracle_golden.py).

  forward():  model/squeeze_ernet.py:24-46, model/squeeze_ernet_redconv.py:27-52, model/acff.py:37-59
  transform(): dataloaders/aider.py:421-426 (Resize(159) -> CenterCrop(140) -> ToTensor -> Normalize)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_torch_sd(sd, dtype=torch.float32):
    out = {}
    for k, v in sd.items():
        t = torch.as_tensor(np.asarray(v))
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out


def _acff(x, sd, p):
    c = x.shape[1]
    br = [F.conv2d(x, sd[f"{p}.conv{j}.weight"], sd[f"{p}.conv{j}.bias"], 1, d - 1, d, c)
          for j, d in ((1, 1), (2, 2), (3, 3))]
    out = torch.cat(br, 1)
    out = F.conv2d(out, sd[f"{p}.fused_conv.weight"], sd[f"{p}.fused_conv.bias"])
    out = F.leaky_relu(out, 0.01)
    return F.batch_norm(out, sd[f"{p}.batch_norm.running_mean"], sd[f"{p}.batch_norm.running_var"],
                        sd[f"{p}.batch_norm.weight"], sd[f"{p}.batch_norm.bias"], False, 0.1, 1e-5)


@torch.no_grad()
def forward(sd, x, arch):
    """sd: torch state_dict (see to_torch_sd); x: (B,3,140,140) tensor.  Returns (probs, logits)."""
    red = arch == "squeeze-redconv"
    out = F.conv2d(x, sd["conv1.weight"], None, 2)
    if red:
        out = F.conv2d(out, sd["conv_red1.weight"], sd["conv_red1.bias"])
    out = F.max_pool2d(_acff(out, sd, "acff1"), 2, 2)
    out = _acff(out, sd, "acff2")
    if red:
        out = F.conv2d(out, sd["conv_red2.weight"], sd["conv_red2.bias"])
    out = F.max_pool2d(out, 2, 2)
    out = F.max_pool2d(_acff(out, sd, "acff3"), 2, 2)
    if red:
        out = F.conv2d(out, sd["conv_red3.weight"], sd["conv_red3.bias"])
    out = _acff(out, sd, "acff4")
    out = F.conv2d(out, sd["conv2.weight"], None)
    out = F.avg_pool2d(out, 5, 1, 1)
    logits = F.linear(out.reshape(-1, 20), sd["fc.weight"], sd["fc.bias"])
    return F.softmax(logits, 1), logits


def make_transform():
    """The reference's eval transform built from torchvision (present in this image), else None."""
    try:
        from torchvision import transforms
    except Exception:
        return None
    return transforms.Compose([transforms.Resize(int(140 * 1.14)), transforms.CenterCrop(140), transforms.ToTensor(),
                               transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])


def transform_frames(frames):
    """(B,H,W,3) uint8 numpy -> (B,3,140,140) fp32 tensor, through PIL like aider-predict.py:57-66."""
    tf = make_transform()
    if tf is None:
        from oracle import ingest_numpy
        return torch.from_numpy(ingest_numpy.ingest(frames)), "numpy-oracle transform"
    from PIL import Image
    return torch.stack([tf(Image.fromarray(f)) for f in frames], 0), "torchvision/Pillow transform"
