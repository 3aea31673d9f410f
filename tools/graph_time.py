"""CUDA-graph replay of the frames path (torch.cuda.CUDAGraph around model.forward_frames): single-frame latency and batch-256
throughput, eager launches vs graph replay, results compared bitwise."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

dev = torch.device("cuda:0")
sd = fixtures.get_state_dict("squeeze-ernet", "shipped")


def timed(fn, iters=300, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


for prec in ("bf16", "fp32"):
    m = rtdm_b200.from_state_dict("squeeze-ernet", sd, dev, prec)
    for B in (1, 8, 256):
        frames = torch.randint(0, 256, (B, 240, 240, 3), dtype=torch.uint8, device=dev)
        eager = m.forward_frames(frames).clone()
        runner = m.graph_frames(frames)                      # captures forward_frames on the static buffer `frames`
        out = runner()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, eager))
        frames.copy_(torch.randint(0, 256, (B, 240, 240, 3), dtype=torch.uint8, device=dev))
        same2 = bool(torch.equal(runner(), m.forward_frames(frames)))
        ms_e = timed(lambda: m.forward_frames(frames))
        ms_g = timed(runner)
        print(json.dumps({"precision": prec, "batch": B, "eager_ms": round(ms_e, 4), "graph_ms": round(ms_g, 4),
                          "eager_img_s": round(B / ms_e * 1e3), "graph_img_s": round(B / ms_g * 1e3), "bit_identical": same and same2}))

# per-frame latency from HOST memory (real-time-inference.py's loop body): pinned staging copy + graph replay + 20-byte read-back
import time  # noqa: E402

import numpy as np  # noqa: E402
from rtdm_b200 import predict as P  # noqa: E402
m = rtdm_b200.from_state_dict("squeeze-ernet", sd, dev, "bf16")
clf = P.FrameClassifier(m, 240, 240)
fr = np.random.RandomState(0).randint(0, 256, (64, 240, 240, 3)).astype(np.uint8)
for i in range(20):
    clf(fr[i % 64])
t0 = time.perf_counter()
for i in range(500):
    clf(fr[i % 64])
dt = (time.perf_counter() - t0) / 500
print(json.dumps({"row": "single frame from host memory (FrameClassifier)", "precision": "bf16", "ms_per_frame": round(dt * 1e3, 4), "fps": round(1 / dt)}))
