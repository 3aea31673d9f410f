#!/usr/bin/env python
"""One small forward per engine, meant to run under compute-sanitizer (race / sync evidence for the
mbarrier, TMEM and cluster kernels):

    compute-sanitizer --tool racecheck  python tools/sanitize_check.py > profiles/r02_sanitizer_racecheck.log
    compute-sanitizer --tool synccheck  python tools/sanitize_check.py > profiles/r02_sanitizer_synccheck.log
    compute-sanitizer --tool memcheck   python tools/sanitize_check.py > profiles/r02_sanitizer_memcheck.log

B = 3 frames (odd: CTA pairs with a missing half) through the frames path (fused transform + conv1 + block 1) and the
tensor path, for bf16 / int8 / fp32 Squeeze_ErNET and fp16 Squeeze_RedConv; results are compared with the oracle so a
sanitizer-induced slowdown that trips a watchdog would be seen."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import fixtures
import rtdm_b200
from oracle import ernet_numpy as E, ingest_numpy as I
from rtdm_b200 import _lib

dev = torch.device("cuda:0")
frames = np.concatenate([fixtures.noise_frames(2, seed=5), fixtures.smooth_frames(1, seed=6)], 0)
x = I.ingest(frames)
only = sys.argv[1:] or None
for arch, prec, tol in (("squeeze-ernet", "bf16", 2e-2), ("squeeze-ernet", "int8", 0.15), ("squeeze-ernet", "fp32", 1e-4),
                        ("squeeze-redconv", "fp16", 2e-2)):
    if only and prec not in only:
        continue
    sd = fixtures.get_state_dict(arch, "shipped")
    ref = E.forward(sd, x, arch, dtype=np.float64)["logits"]
    m = rtdm_b200.from_state_dict(arch, sd, dev, prec)
    if prec == "int8":
        m.calibrate(np.concatenate([fixtures.noise_frames(8, seed=99), fixtures.smooth_frames(8, seed=98)], 0), batch=16)
    lf = m.forward_frames(torch.from_numpy(frames).to(dev), return_logits=True)[1]
    lt = m.logits(torch.from_numpy(x).to(dev))
    torch.cuda.synchronize()
    for name, lg in (("frames", lf), ("tensor", lt)):
        err = float(np.abs(lg.double().cpu().numpy() - ref).max() / np.abs(ref).max())
        print(f"{arch} {prec} {name}: rel logit err {err:.3e} (tolerance {tol})", flush=True)
        assert err <= tol
    assert _lib.load().ernet_check_watchdog() == 0
print("done")
