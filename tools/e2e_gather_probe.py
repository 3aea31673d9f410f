"""Host path: copy-engine transfer of the row range vs host_gather_kernel (footprint rows AND columns pulled over PCIe by a few
CTAs, csrc/host_gather.cuh) - end-to-end images/s of the streaming API for several CTA counts, and the raw rate of each."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "squeeze-ernet"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
g = torch.Generator().manual_seed(5)
sets = [torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(4)]
m = rtdm_b200.from_state_dict(arch, fixtures.get_state_dict(arch, "shipped"), "cuda:0", prec)
if prec == "int8":
    m.calibrate()


def run():
    for i in range(3):
        m.classify_host(sets[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = None
    for i in range(steps):
        nxt = m.classify_host_submit(sets[i % 4])
        if pending is not None:
            pending.result()
        pending = nxt
    p = pending.result()
    torch.cuda.synchronize()
    return batch * steps / (time.perf_counter() - t0), p


out = {"arch": arch, "precision": prec, "batch": batch, "steps": steps}
v, want = run()
out["copy_engine"] = {"img_s": round(v), "bytes_per_frame": m.host_copy_bytes_per_frame(240, 240)}
for ctas in (4, 8, 16, 32, 64, 148):
    m.set_host_gather(True, ctas=ctas)
    v, got = run()
    assert (got == want).all()
    out[f"gather_{ctas}"] = {"img_s": round(v), "bytes_per_frame": m.host_copy_bytes_per_frame(240, 240)}
m.set_host_gather(False)
v, got = run()
out["copy_engine_again"] = {"img_s": round(v)}
print(json.dumps(out))
