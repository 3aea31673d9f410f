"""Study (-DERNET_TIMELINE build): two forward chains on two streams - do they overlap?"""
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

batch = 256
sd = fixtures.get_state_dict("squeeze-ernet", "shipped")
ms = [rtdm_b200.from_state_dict("squeeze-ernet", sd, "cuda:0", "bf16") for _ in range(2)]
st = [torch.cuda.Stream() for _ in range(2)]
sets = [torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8).cuda() for _ in range(4)]
lib = _lib.load()
for i in range(8):
    ms[i % 2].forward_frames(sets[i % 4], stream=st[i % 2])
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
names = ["ingest+conv1", "block1", "block2", "block3", "acff4+head"]
for rep in range(2):
    _lib.check(lib.ernet_debug_chain(None, 1))
    ms[0].forward_frames(sets[0], stream=st[0])
    ms[1].forward_frames(sets[1], stream=st[1])
    torch.cuda.synchronize()
    _lib.check(lib.ernet_debug_chain(buf, 1))
    t = [int(v) for v in buf]
    t0 = t[0]
    print(f"--- two chains on two streams, rep {rep} (min entry / max exit over both)")
    for k, n in enumerate(names):
        e, w, x, le = t[4 * k: 4 * k + 4]
        print(f"  {n:13s} first entry {(e - t0) / 1e3:8.1f} us  first wait-return {(w - t0) / 1e3:8.1f}  last entry {(le - t0) / 1e3:8.1f}  last exit {(x - t0) / 1e3:8.1f}")
