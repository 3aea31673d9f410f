"""Study: do two forward chains on two streams overlap (transform+conv1 of one under the tensor kernels of the other)?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = 400
sd = fixtures.get_state_dict("squeeze-ernet", "shipped")
sets = [torch.randint(0, 256, (B, 240, 240, 3), dtype=torch.uint8).cuda() for _ in range(8)]


def run(nstreams):
    models = [rtdm_b200.from_state_dict("squeeze-ernet", sd, "cuda:0", "bf16") for _ in range(nstreams)]
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for i in range(16):
        models[i % nstreams].forward_frames(sets[i % 8], stream=streams[i % nstreams])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        models[i % nstreams].forward_frames(sets[i % 8], stream=streams[i % nstreams])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{nstreams} stream(s): {steps * B / dt:,.0f} img/s  ({dt / steps * 1e6:.1f} us per step)", flush=True)


run(1)
run(2)
run(3)
B = 1
sets = [x[:1] for x in sets]
run(1)
