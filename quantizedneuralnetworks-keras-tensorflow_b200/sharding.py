"""Batch-sharded inference across the GPUs of one box.

Images are independent in every in-scope network type, so the path shards with NO data-path collective:
rank r of g runs the fused plan on its contiguous slice of the batch; the only exchange is the final
logit gather (fp32 [N/g, classes], <= 160 KB at N = 4096), a single all-gather over NCCL on GPUs (gloo in
the CPU tests).  One process per GPU, launched with torchrun; ``torch.distributed`` must be initialised.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int):
    """Contiguous, balanced slice [lo, hi) of n items for `rank`: the first n % world ranks get one extra."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank %d / world %d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gather the per-rank row blocks produced under :func:`shard_range` into the full [n_total, C] tensor
    (every rank receives it).  Ragged shards are padded to the largest shard for the collective and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    rows = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], hi - lo))
    buf = local
    if local.shape[0] != rows:
        buf = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous(), group=group)
    return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)], dim=0)


def predict_sharded(model, x, group=None, batch_size=None, forward=None):
    """Every rank passes the SAME global batch `x` (numpy / torch, host or device); each computes its shard with
    ``model.predict`` and all ranks return the full [N, classes] result.  ``forward`` overrides the local compute
    (used by the CPU tests, where no CUDA device exists)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(x.shape[0])
    lo, hi = shard_range(n, world, rank)
    xs = x[lo:hi]
    local = forward(xs) if forward is not None else model.predict(xs, batch_size=batch_size)
    if not isinstance(local, torch.Tensor):
        local = torch.as_tensor(local)
    if world == 1:
        return local
    backend = dist.get_backend(group)
    if backend == "nccl" and not local.is_cuda:
        local = local.cuda()
    return all_gather_rows(local, n, group)
