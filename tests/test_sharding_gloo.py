"""Multi-rank host logic on CPU: world_size-2 gloo process group, shard bookkeeping and the optional
gather of per-rank probability blocks.  (The GPU path itself has no collective)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtdm_b200 import parallel


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 2, 7, 256, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(n, world)
    assert parallel.shard_bounds(8192, 8, 3) == (3072, 4096)        # BASELINE config 5: 1024 per GPU
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _fake_classify(frames):
    """Deterministic stand-in for model.forward_frames on CPU: a softmax of per-frame statistics."""
    f = frames.double().reshape(frames.shape[0], -1)
    z = torch.stack([f.mean(1), f.std(1), f[:, ::7].mean(1), f[:, 1::5].mean(1), f.max(1).values], 1) / 50.0
    return torch.softmax(z, 1).float()


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = torch.from_numpy(np.random.RandomState(5).randint(0, 256, (n, 12, 12, 3)).astype(np.uint8))
        full = parallel.classify_sharded(_fake_classify, frames, gather=True)
        local, (lo, hi) = parallel.classify_sharded(_fake_classify, frames, gather=False)
        ref = _fake_classify(frames)
        ok = torch.equal(full, ref) and torch.equal(local, ref[lo:hi])
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(float(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 2, 1])
def test_two_rank_sharded_classify_equals_single_process(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0


def _cm_worker(rank, world, port, n, out):
    """Sharded evaluation bookkeeping: each rank builds the confusion matrix of its shard, the matrices are all-reduced and
    every rank derives the metrics of the whole set."""
    from oracle import metrics_numpy as M
    from rtdm_b200 import evaluate as EV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(17)
        pred, targ = rs.randint(0, 5, n), rs.randint(0, 5, n)
        lo, hi = parallel.shard_bounds(n, world, rank)
        cm_local = torch.from_numpy(M.confusion_matrix(pred[lo:hi], targ[lo:hi]))
        got = EV.metrics_from_confusion(parallel.reduce_confusion(cm_local))
        cm_full = M.confusion_matrix(pred, targ)
        want = {**M.micro_metrics(cm_full), **M.per_class_metrics(cm_full)}
        ok = np.array_equal(got["confusion_matrix"], cm_full) and all(abs(got[k] - v) < 1e-12 for k, v in want.items())
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(float(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 1])
def test_two_rank_confusion_reduce_equals_single_process(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cm_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0
