"""Batch-sharded inference across the GPUs of one box.

Images are independent in every in-scope network type, so the path shards with NO data-path collective:
rank r of g runs the fused plan on its contiguous slice of the batch; the only exchange is the final
logit gather (fp32 [N/g, classes], <= 160 KB at N = 4096), a single all-gather over NCCL on GPUs (gloo in
the CPU tests).  One process per GPU, launched with torchrun; ``torch.distributed`` must be initialised.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


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


class PeerGather:
    """The NVLink logit path: ``slots`` buffers of [world * rows, classes] fp32 that live on ``root``'s GPU and are
    mapped (CUDA IPC) into every other rank's process.  Rank r passes ``block(slot)`` -- rows [r * rows, (r+1) * rows)
    of slot ``slot`` -- as the output of its final dense kernel (``Plan.forward(x, out=...)``), so the shard's logits
    are written into the gathering rank's memory by the kernel's own stores over NVLink / NVSwitch and no collective
    runs per step.  ``gathered(slot)`` (root only) is the full tensor; it is complete once every rank has synchronised
    the stream it launched on and the ranks have met at a barrier (``fence()``).

    One process per GPU on one box; ``torch.distributed`` must be initialised (any backend: only the 64-byte handle
    travels through it)."""

    def __init__(self, rows, classes, slots=1, root=0, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.classes, self.slots, self.root = int(rows), int(classes), int(slots), int(root)
        self._slot_rows = self.world * self.rows
        nbytes = self.slots * self._slot_rows * self.classes * 4
        self._owned = self._mapped = None
        handle = (C.c_ubyte * L.PEER_HANDLE_BYTES)()
        err = None
        if self.rank == self.root:
            try:
                p = C.c_void_p()
                L.check(L.lib().qnnb_peer_alloc(nbytes, C.byref(p), handle))
                self._owned = p.value
            except Exception as exc:                    # the other ranks are waiting in the broadcast: tell them
                err = exc
        box = [bytes(handle) if (self.rank == self.root and err is None) else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, self.root) if group is not None else self.root, group=group)
        if box[0] is None:
            raise RuntimeError("PeerGather: the root rank could not export its buffer%s" % (": %s" % err if err else ""))
        if self.rank == self.root:
            base = self._owned
        else:
            raw = (C.c_ubyte * L.PEER_HANDLE_BYTES).from_buffer_copy(box[0])
            p = C.c_void_p()
            L.check(L.lib().qnnb_peer_open(raw, C.byref(p)))
            self._mapped = base = p.value
        self._buf = L.DeviceBuffer(base, (self.slots * self._slot_rows, self.classes), torch.float32)

    def block(self, slot=0):
        """This rank's row block of ``slot`` (a device address; on the root it is local memory)."""
        lo = (slot % self.slots) * self._slot_rows + self.rank * self.rows
        return self._buf.rows(lo, lo + self.rows)

    def gathered(self, slot=0):
        """Root only: the [world * rows, classes] tensor of ``slot`` (a torch view of the exported buffer)."""
        if self.rank != self.root:
            raise RuntimeError("gathered() is only available on the root rank")
        nbytes = self._slot_rows * self.classes * 4
        iface = {"shape": (self._slot_rows, self.classes), "typestr": "<f4", "version": 3,
                 "data": (self._owned + (slot % self.slots) * nbytes, False), "strides": None}
        holder = type("_PeerView", (), {"__cuda_array_interface__": iface, "_keep": self})()
        return torch.as_tensor(holder, device=torch.device("cuda", torch.cuda.current_device()))

    def fence(self):
        """Make every rank's stores visible to the root: drain this device, then meet."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def close(self):
        if self._mapped is not None:
            L.lib().qnnb_peer_close(C.c_void_p(self._mapped))
            self._mapped = None
        if dist.is_initialized():
            try:
                dist.barrier(group=self.group)          # nobody unmaps after the owner has freed
            except Exception:
                pass
        if self._owned is not None:
            L.lib().qnnb_peer_free(C.c_void_p(self._owned))
            self._owned = None
