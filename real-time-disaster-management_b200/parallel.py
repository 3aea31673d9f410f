"""Data-parallel sharding of the classification path: independent image slices, no collective on the
hot path.  One process per GPU; ``torch.distributed`` is plumbing only.

The reference has no multi-GPU classifier (`evaluate-classification-metrics.py:158` uses a single
`torch.device('cuda')`); this module is the new-functionality side of BASELINE config 5: a batch is cut
into contiguous slices, every rank classifies its slice with its own handle + weight copy, and the
(n_i, 5) probability blocks are optionally all-gathered (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of ``n`` items owned by ``rank``; sizes differ by at most one and
    concatenating the slices in rank order restores the original order."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} world={world} rank={rank}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def gather_rows(local, sizes, group=None):
    """All-gather row blocks of unequal height: ``local`` is (sizes[rank], k); returns (sum(sizes), k) on
    every rank, rows in rank order.  Blocks are padded to the tallest one for the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(sizes) != world or local.shape[0] != sizes[rank]:
        raise ValueError("sizes do not describe this process group")
    tallest = max(sizes)
    pad = local.new_zeros((tallest,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


def classify_sharded(classify, frames, gather=True, group=None):
    """Run ``classify`` (frames -> (n,5) tensor, e.g. ``model.forward_frames``) on this rank's slice of
    ``frames`` (same full batch visible on every rank, or any indexable).  Returns the full (N,5) result
    when ``gather`` is set, else this rank's block and its (lo, hi)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(frames)
    lo, hi = shard_bounds(n, world, rank)
    if hi > lo:
        local = classify(frames[lo:hi])
    else:                                       # more ranks than images: an empty block of the right width
        local = torch.zeros((0, 5), dtype=torch.float32,
                            device=frames.device if isinstance(frames, torch.Tensor) else "cpu")
    if not gather or world == 1:
        return local if gather else (local, (lo, hi))
    return gather_rows(local, shard_sizes(n, world), group)


def reduce_confusion(cm, group=None):
    """Sum the per-rank confusion matrices (int64 (nc,nc), on the rank's device) over the process group: the one collective
    of a sharded evaluation, 200 bytes (NCCL all-reduce on GPUs, gloo in the CPU tests).  No-op without a process group."""
    cm = torch.as_tensor(cm).to(torch.int64).clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group)
    return cm


def evaluate_sharded(model, samples, device, batch_size=64, num_workers=4, group=None):
    """Evaluation of ``samples`` ([(path, label)], the same list on every rank) split into contiguous shards: every rank
    decodes and classifies its shard with its own engine (evaluate.evaluate_model), the 5x5 confusion matrices are
    all-reduced, and every rank returns the metrics of the WHOLE set (evaluate-classification-metrics.py:89-103).  The
    timing entries describe this rank's shard."""
    from .evaluate import evaluate_model, frame_batches, metrics_from_confusion
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(samples), world, rank)
    local = evaluate_model(model, frame_batches(samples[lo:hi], batch_size, num_workers), device, allow_empty=True)
    cm = reduce_confusion(torch.as_tensor(local["confusion_matrix"]).to(device), group)
    out = metrics_from_confusion(cm.cpu())
    for k in ("avg_inference_time", "fps", "images_per_second"):
        out[k] = local[k]
    return out
