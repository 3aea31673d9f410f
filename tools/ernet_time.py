"""Quick timing of the ErNET path (model(x) on device-resident (B,3,240,240) tensors)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = fixtures.get_state_dict("ernet", "shipped")
for prec in ("fp32", "bf16", "fp16"):
    m = rtdm_b200.from_state_dict("ernet", sd, "cuda:0", prec)
    xs = [torch.randn(B, 3, 240, 240, device="cuda", dtype={"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[prec]) for _ in range(3)]
    for i in range(3):
        m(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for i in range(n):
        m(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"ErNET {prec} B={B}: {ms:.3f} ms per batch, {B / ms * 1e3:,.0f} img/s (engine {m.engine})", flush=True)
