"""Single-frame latency (one 240x240 frame resident in HBM -> probabilities) through model.graph_frames, with the one-tile
pair-kernel units for small batches (CBlock2S / CBlock3S) on and off, plus batches 8 and 32 eager."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

g = torch.Generator().manual_seed(3)
frames = torch.randint(0, 256, (64, 240, 240, 3), dtype=torch.uint8, generator=g).cuda()
out = {}
ref = {}
for mode in ("0", "1"):
    os.environ["ERNET_SMALL_BATCH_UNITS"] = mode
    m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "shipped"), "cuda:0", "bf16")
    res = {}
    for B in (1, 8, 18, 32, 64):
        x = frames[:B]
        lg = m.forward_frames(x, return_logits=True)[1].clone()
        if mode == "0":
            ref[B] = lg
        else:
            assert torch.equal(lg, ref[B]), B
        for _ in range(10):
            m.forward_frames(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            m.forward_frames(x)
        e1.record()
        torch.cuda.synchronize()
        res[f"eager_b{B}_us"] = round(e0.elapsed_time(e1) / 200 * 1e3, 2)
    runner = m.graph_frames(frames[:1].clone())
    for _ in range(10):
        runner()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        runner()
    e1.record()
    torch.cuda.synchronize()
    res["graph_b1_us"] = round(e0.elapsed_time(e1) / 500 * 1e3, 2)
    out[mode] = res
print(json.dumps({"small_batch_units": out, "bit_identical": True}))
