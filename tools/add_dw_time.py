"""Add-fusion depthwise (detector ACFF, yolov3/models.py:302) against the HBM roofline at two batch sizes: algorithmic bytes
= C*H^2 in + C*(H-2)^2 out, fp32; L2 flushed before every launch.  (Round 2 also measured 12x12 tiles at 64 channels and
16x16 tiles at 32 channels against the shipped 8x8 x 64: 15-19 % slower on every shape - fewer resident CTAs cost more than
the smaller halo re-read saves.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from rtdm_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.7
flush = torch.empty(160 << 20, dtype=torch.float32, device=dev)
s = torch.cuda.current_stream()
for (Bn, C, H) in ((16, 128, 104), (16, 256, 52), (64, 128, 104), (64, 256, 52)):
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(Bn, H, H, C, generator=g).to(dev)
    w = (torch.randn(3, 9, C, generator=g) * 0.3).to(dev)
    b = (torch.randn(3, C, generator=g) * 0.1).to(dev)
    ref = None
    for ts in (8,):
        o = torch.empty(Bn, H - 2, H - 2, C, device=dev)
        tot = 0.0
        for it in range(12):
            flush.fill_(float(it))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            _lib.check(lib.ernet_acff_add_depthwise(x.data_ptr(), 0, Bn, H, H, C, H - 2, H - 2, w.data_ptr(), b.data_ptr(), o.data_ptr(), s.cuda_stream))
            e1.record(s)
            e1.synchronize()
            if it >= 2:
                tot += e0.elapsed_time(e1)
        ms = tot / 10
        nbytes = (x.numel() + o.numel()) * 4
        print(json.dumps({"row": "8f-4 add-fusion depthwise", "shape": [Bn, H, H, C], "tile": ts, "us": round(ms * 1e3, 2),
                          "gbs": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / peak, 3)}))
