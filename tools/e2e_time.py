"""Host-buffer entry point only: images/s of classify_host on pinned 240x240 frames (bench.py's e2e leg, more steps)."""
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
m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "shipped"), "cuda:0", "bf16")
m.prepare_ingest(240, 240)
sets = [torch.randint(0, 256, (B, 240, 240, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
for i in range(5):
    m.classify_host(sets[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 100
for i in range(n):
    m.classify_host(sets[i % 4])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"classify_host B={B}: {n * B / dt:,.0f} img/s ({dt / n * 1e3:.3f} ms per call, {m.host_copy_bytes_per_frame(240, 240) * B / (dt / n) / 1e9:.1f} GB/s over PCIe)")
