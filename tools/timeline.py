"""Study tool: per-CTA timeline of the persistent tensor-core block kernels (library built with -DERNET_TIMELINE).

    ERNET_NVCC_EXTRA=-DERNET_TIMELINE python -m real-time-disaster-management_b200.build   # then run this on a GPU
Stamps per unit k: 0 TMA issued, 1 input landed (MMA warp), 2 accumulator buffer free, 3 MMAs issued,
4 accumulators complete (epilogue warp 3), 5 epilogue done; unit slot 31: 6 = kernel body start, 7 = end."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402
from rtdm_b200 import _lib  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "shipped"), "cuda:0", "bf16")
frames = torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8).cuda()
for _ in range(5):
    m.forward_frames(frames)
torch.cuda.synchronize()
n = 3 * 148 * 32 * 8
buf = (C.c_ulonglong * n)()
_lib.check(_lib.load().ernet_debug_timeline(buf, n))
t = np.frombuffer(buf, dtype=np.uint64).astype(np.int64).reshape(3, 148, 32, 8)
for kern in range(3):
    for cta in (0, 1, 147):
        a = t[kern, cta]
        t0 = a[31, 6]
        print(f"--- block {kern + 1} cta {cta}: body {a[31, 7] - t0} cycles")
        for k in range(31):
            if a[k, 0] == 0 and a[k, 3] == 0:
                break
            r = [int(v - t0) if v else -1 for v in a[k, :6]]
            print(f"  unit {k:2d}: tma {r[0]:7d} landed {r[1]:7d} accfree {r[2]:7d} issued {r[3]:7d} accfull {r[4]:7d} epidone {r[5]:7d}")

print("--- acff4+head (cta 0, 1, 100): cycles from entry: pdl wait returned, input landed, depthwise done, accumulators complete, epilogue done")
for cta in (0, 1, 100):
    a = t[2, cta, 20]
    print(f"  cta {cta}: " + " ".join(str(int(v - a[0])) for v in a[1:6]))
