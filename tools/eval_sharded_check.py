"""2-rank check of parallel.evaluate_sharded on GPUs (NCCL): every rank evaluates its shard of the same file list, the 5x5
confusion matrices are all-reduced, and both ranks must report the metrics a single process gets for the whole list.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/eval_sharded_check.py"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402
from rtdm_b200 import evaluate as EV, parallel  # noqa: E402
from PIL import Image  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tmp = tempfile.mkdtemp()
rs = np.random.RandomState(3)
rows = []
for k in range(101):                                     # every rank writes the same files into its own directory
    h, w = ((240, 240), (260, 300))[k % 2]
    path = os.path.join(tmp, f"img{k:03d}.png")
    Image.fromarray(fixtures.smooth_frames(1, h, w, seed=500 + k)[0]).save(path)
    rows.append((path, int(rs.randint(0, 5))))
model = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "w3"), dev, "bf16")
sharded = parallel.evaluate_sharded(model, rows, dev, batch_size=16, num_workers=2)
whole = EV.evaluate_model(model, EV.frame_batches(rows, 16, 2), dev)
ok = np.array_equal(sharded["confusion_matrix"], whole["confusion_matrix"]) and sharded["accuracy"] == whole["accuracy"] \
    and int(sharded["confusion_matrix"].sum()) == len(rows)
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "evaluate_sharded == evaluate_model on the whole list", "world": world, "images": len(rows),
                      "ok": bool(t.item() == 1.0), "accuracy": whole["accuracy"]}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
