"""Timings for the SURVEY 8f rows (run on the GPU box): ErNET frames path, add-fusion ACFF block at the detector's size,
evaluation loop from JPEG files (our loop vs a reference-style CPU loop on the oracle).  Prints one JSON object per line."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402
from rtdm_b200 import _lib, evaluate as EV  # noqa: E402

dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters=20, warm=3):
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


# ---- ErNET frames path
B = 256
sets = [torch.randint(0, 256, (B, 240, 240, 3), dtype=torch.uint8).to(dev) for _ in range(4)]
for prec in ("bf16", "fp16", "fp32"):
    m = rtdm_b200.from_state_dict("ernet", fixtures.get_state_dict("ernet", "shipped"), dev, prec)
    i = [0]

    def step():
        i[0] += 1
        return m.forward_frames(sets[i[0] % 4])
    ms = timed(step)
    x = m.ingest(sets[0], dtype={"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[prec])
    ms_model = timed(lambda: m(x))
    print(json.dumps({"row": "8f-1 ErNET frames path", "precision": prec, "batch": B, "frames_img_s": round(B / ms * 1e3),
                      "tensor_img_s": round(B / ms_model * 1e3), "ms_frames": round(ms, 4), "ms_tensor": round(ms_model, 4)}))
    del m

# ---- add-fusion ACFF block at the detector's size (yolov3-acffx.cfg:91, 416x416 input -> 104x104x128 map)
lib = _lib.load()
for (Bn, C, H, Cout) in ((16, 128, 104, 128), (16, 256, 52, 256)):
    blk = rtdm_b200.ACFF(C, Cout, 3).to(dev).eval()
    x = torch.randn(Bn, C, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    ms_blk = timed(lambda: blk(x))
    dw, db = blk._pack()[:2]
    xn = x.permute(0, 2, 3, 1).contiguous()
    s = torch.empty(Bn, H - 2, H - 2, C, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ms_dw = timed(lambda: _lib.check(lib.ernet_acff_add_depthwise(xn.data_ptr(), 0, Bn, H, H, C, H - 2, H - 2, dw.data_ptr(), db.data_ptr(), s.data_ptr(), st)))
    nbytes = (xn.numel() + s.numel()) * 4
    flops = 2.0 * Bn * (H - 2) ** 2 * C * Cout
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.nn.Sequential()
    import torch.nn as nn

    class RefACFF(nn.Module):                           # the same block expressed with stock torch modules on the GPU (cuDNN fp32)
        def __init__(s_):
            super().__init__()
            s_.c = nn.ModuleList([nn.Conv2d(C, C, 3, 1, d, d + 1, groups=C) for d in range(3)])
            s_.f, s_.a, s_.b = nn.Conv2d(C, Cout, 1), nn.LeakyReLU(0.01), nn.BatchNorm2d(Cout)

        def forward(s_, t):
            return s_.b(s_.a(s_.f(s_.c[0](t) + s_.c[1](t) + s_.c[2](t))))
    rm = RefACFF().to(dev).eval().to(memory_format=torch.channels_last)
    with torch.no_grad():
        ms_ref = timed(lambda: rm(x))
    print(json.dumps({"row": "8f-4 add-fusion ACFF", "shape": [Bn, C, H, H], "cout": Cout, "block_ms": round(ms_blk, 4),
                      "depthwise_ms": round(ms_dw, 4), "depthwise_gbs": round(nbytes / ms_dw / 1e6, 1), "depthwise_frac_hbm": round(nbytes / ms_dw / 1e6 / HBM, 3),
                      "pointwise_tflops_fp32": round(flops / max(ms_blk - ms_dw, 1e-6) / 1e9, 2), "torch_cudnn_fp32_ms": round(ms_ref, 4)}))

# ---- evaluation loop from JPEG files
from PIL import Image  # noqa: E402
N = 1024
tmp = tempfile.mkdtemp()
rs = np.random.RandomState(0)
fr = fixtures.smooth_frames(64, 240, 240, seed=77)
rows = []
for k in range(N):
    rel = f"img{k:05d}.jpg"
    Image.fromarray(np.roll(fr[k % 64], k, axis=1)).save(os.path.join(tmp, rel), quality=90)
    rows.append((os.path.join(tmp, rel), int(rs.randint(0, 5))))
sd = fixtures.get_state_dict("squeeze-ernet", "shipped")
workers = min(32, max(2, (os.cpu_count() or 8) - 2))
for prec in ("bf16", "fp32"):
    m = rtdm_b200.from_state_dict("squeeze-ernet", sd, dev, prec)
    EV.evaluate_model(m, EV.frame_batches(rows[:128], 64, workers), dev)
    t0 = time.perf_counter()
    met = EV.evaluate_model(m, EV.frame_batches(rows, 128, workers), dev)
    dt = time.perf_counter() - t0
    print(json.dumps({"row": "8f-2 evaluation loop", "precision": prec, "images": N, "decode_threads": workers, "wall_img_s": round(N / dt, 1),
                      "model_img_s_reference_definition": round(met["images_per_second"], 1), "accuracy": met["accuracy"]}))
# the same loop with the JPEGs decoded on the GPU (one batched nvJPEG call per chunk; host threads only read files)
N2 = 8192
rows2 = [rows[k % N] for k in range(N2)]
for prec in ("bf16",):
    m = rtdm_b200.from_state_dict("squeeze-ernet", sd, dev, prec)
    EV.evaluate_model(m, EV.frame_batches_device(rows2[:512], 256, dev, 4), dev)
    for bs, nw in ((256, 4), (256, workers), (128, workers), (64, workers)):
        t0 = time.perf_counter()
        met = EV.evaluate_model(m, EV.frame_batches_device(rows2, bs, dev, nw), dev)
        dt = time.perf_counter() - t0
        print(json.dumps({"row": "8f-2 evaluation loop, device JPEG decode", "precision": prec, "images": N2, "chunk": bs, "decode_threads": nw,
                          "wall_img_s": round(N2 / dt, 1), "model_img_s_reference_definition": round(met["images_per_second"], 1),
                          "accuracy": met["accuracy"]}))
    # agreement of the two decoders' predictions on the same files
    host = EV.evaluate_model(m, EV.frame_batches(rows, 128, workers), dev)["confusion_matrix"]
    devc = EV.evaluate_model(m, EV.frame_batches_device(rows, 256, dev, 4), dev)["confusion_matrix"]
    print(json.dumps({"row": "8f-2 host vs device decode", "confusion_matrix_abs_diff": int(np.abs(np.asarray(host) - np.asarray(devc)).sum()), "images": N}))
# decode alone (the floor of our loop) and the reference-style loop: PIL transform + CPU fp32 forward per batch
t0 = time.perf_counter()
n = sum(f.shape[0] for f, _ in EV.frame_batches(rows, 128, workers, pin_memory=False))
print(json.dumps({"row": "8f-2 decode only", "images": n, "threads": workers, "img_s": round(n / (time.perf_counter() - t0), 1)}))
from oracle import ernet_torch as T  # noqa: E402  (cpu_baseline leg: the reference's CPU path restated)
torch.set_num_threads(os.cpu_count() or 8)
sub = rows[:256]
t0 = time.perf_counter()
tf = T.make_transform()
tsd = T.to_torch_sd(sd)
for i0 in range(0, len(sub), 64):
    xb = torch.stack([tf(Image.open(p).convert("RGB")) for p, _ in sub[i0:i0 + 64]], 0)
    T.forward(tsd, xb, "squeeze-ernet")
print(json.dumps({"row": "8f-2 reference-style CPU loop", "images": len(sub), "threads": torch.get_num_threads(),
                  "img_s": round(len(sub) / (time.perf_counter() - t0), 1)}))
