"""Study tool (-DERNET_TIMELINE build): global-timer view of one forward chain - when each kernel's first CTA entered,
when its PDL wait first returned, when its last CTA entered and left."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402
from rtdm_b200 import _lib  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "shipped"), "cuda:0", "bf16")
sets = [torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8).cuda() for _ in range(4)]
lib = _lib.load()
for i in range(8):
    m.forward_frames(sets[i % 4])
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
names = ["ingest+conv1", "block1", "block2", "block3", "acff4+head"]
for rep in range(3):
    _lib.check(lib.ernet_debug_chain(None, 1))
    for i in range(3):          # three back-to-back steps: stamps keep min entry of the first and max exit of the last per kernel
        m.forward_frames(sets[i % 4])
        if i == 0:
            torch.cuda.synchronize()
            _lib.check(lib.ernet_debug_chain(buf, 1))
            t = [int(v) for v in buf]
            t0 = t[0]
            print(f"--- single step {rep}")
            for k, n in enumerate(names):
                e, w, x, le = t[4 * k: 4 * k + 4]
                print(f"  {n:13s} first entry {(e - t0) / 1e3:8.1f} us  wait returned {(w - t0) / 1e3:8.1f}  last entry {(le - t0) / 1e3:8.1f}  last exit {(x - t0) / 1e3:8.1f}")
    torch.cuda.synchronize()
