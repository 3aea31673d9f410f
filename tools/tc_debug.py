"""Debug driver for the tensor-core kernels: small forwards + the device watchdog record."""
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
from oracle import ernet_numpy as E  # noqa: E402
from rtdm_b200 import _lib  # noqa: E402


def status(reset=True):
    buf = (C.c_uint * 8)()
    _lib.check(_lib.load().ernet_debug_device_status(buf, 1 if reset else 0))
    return [hex(v) for v in buf]


arch = "squeeze-ernet"
sd = fixtures.get_state_dict(arch, "w3")
m = rtdm_b200.from_state_dict(arch, sd, "cuda:0", "bf16")
for B in [int(a) for a in sys.argv[1:]] or [1, 4]:
    x = fixtures.normal_tensors(B, seed=11)
    ref = E.forward(sd, x, arch, dtype=np.float64)["logits"]
    p, l = m.forward_with_logits(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    err = float(np.abs(l.double().cpu().numpy() - ref).max() / np.abs(ref).max())
    print(f"B={B} rel err {err:.3e} status {status()}", flush=True)

# schedules: 0 one image per CTA, 1 persistent single CTAs (bitwise equal to 0), 2 persistent + CTA pairs (other K order)
if hasattr(m, "set_persistent"):
    for B in (1, 2, 5, 300):
        x = torch.from_numpy(fixtures.normal_tensors(B, seed=12)).cuda()
        ref = E.forward(sd, x.cpu().numpy(), arch, dtype=np.float64)["logits"] if B <= 5 else None
        outs = {}
        for sched in (0, 1, 2):
            m.set_persistent(sched)
            _, l = m.forward_with_logits(x)
            torch.cuda.synchronize()
            outs[sched] = l.clone()
            st = status()
            e = float(np.abs(l.double().cpu().numpy() - ref).max() / np.abs(ref).max()) if ref is not None else float("nan")
            print(f"B={B} schedule {sched}: rel err vs oracle {e:.3e} status {st}", flush=True)
        print(f"   0==1 bitwise {bool((outs[0] == outs[1]).all())}; max|2-0| {float((outs[2] - outs[0]).abs().max()):.3e} of {float(outs[0].abs().max()):.3e}", flush=True)
    m.set_persistent(True)
