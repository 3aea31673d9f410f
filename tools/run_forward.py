"""Small driver for profiling: a few forward passes of BASELINE config 2 (Squeeze-ErNet bf16, B=256 frames)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "squeeze-ernet"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
m = rtdm_b200.from_state_dict(arch, fixtures.get_state_dict(arch, "shipped"), "cuda:0", prec)
g = torch.Generator().manual_seed(1234)
frames = torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8, generator=g).cuda()
if prec == "int8":
    m.calibrate()
for _ in range(iters):
    p = m.forward_frames(frames)
torch.cuda.synchronize()
# one more forward inside a profiler range: `ncu --profile-from-start off` captures exactly this one
torch.cuda.profiler.start()
p = m.forward_frames(frames)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(p.sum()))
