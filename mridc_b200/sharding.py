"""Slice-sharded multi-GPU driver.

Every slice is an independent unit (no halo, no cross-slice state: the RIM hidden state is per sample and is
reset per cascade, cirim.py:148), so the N slices of a volume / batch are split into contiguous blocks of
ceil(N/G) slices per rank (keeps a volume's slices together, the output order the reference's
``test_epoch_end`` expects, reconstruction/models/base.py:576-581).  Weights are replicated; there is NO
collective on the data path; the only communication is one gather of the [n_local, h, w] complex64
reconstructions at the end -- to every rank (one all-gather straight into the output) or to one rank only
(point-to-point, what the reference's test_epoch_end needs), optionally asynchronous so that it overlaps with the
next batch (NCCL over NVLink on GPUs; gloo in the CPU unit tests).
One process per GPU (torchrun); nothing here is specific to a backend.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["partition", "shard_slices", "gather_reconstructions", "run_sharded", "PendingGather"]


def partition(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) block per rank; block size ceil(n/world); trailing ranks may be empty."""
    if n_items < 0 or world_size < 1:
        raise ValueError("partition: need n_items >= 0 and world_size >= 1")
    per = -(-n_items // world_size) if n_items else 0
    out = []
    for r in range(world_size):
        a = min(r * per, n_items)
        b = min(a + per, n_items)
        out.append((a, b))
    return out


def shard_slices(tensors: Sequence[Optional[torch.Tensor]], rank: int, world_size: int, n_items: Optional[int] = None):
    """Slice dim 0 of every tensor whose dim 0 equals n_items (broadcast tensors such as a [1,...] mask pass
    through untouched)."""
    if n_items is None:
        n_items = max(t.shape[0] for t in tensors if t is not None)
    a, b = partition(n_items, world_size)[rank]
    out = []
    for t in tensors:
        if t is None or t.shape[0] != n_items:
            out.append(t)
        else:
            out.append(t[a:b])
    return out, (a, b)


class PendingGather:
    """Handle of a gather issued with ``async_op=True``: the collective runs on the backend's own stream and overlaps
    with whatever the caller enqueues next; ``wait()`` orders the current stream after it and returns the result
    (``None`` on ranks that do not receive)."""

    def __init__(self, works, finish):
        self._works, self._finish = works, finish

    def wait(self):
        for w in self._works:
            w.wait()
        return self._finish()


def gather_reconstructions(local: torch.Tensor, n_items: int, group=None, dst: Optional[int] = None,
                           async_op: bool = False):
    """Gather per-rank [n_local, ...] results into [n_items, ...].

    ``dst=None``: every rank receives the result -- ONE ``all_gather_into_tensor`` written straight into the output
    when the blocks are even (the usual case), padded blocks otherwise.  ``dst=r``: only rank ``r`` receives (what
    the reference's ``test_epoch_end`` needs, models/base.py:576-587): point-to-point sends into slices of the
    output, no padding, no copy; the other ranks get ``None``.  ``async_op=True`` returns a ``PendingGather``."""
    if not dist.is_available() or not dist.is_initialized():
        return PendingGather([], lambda: local) if async_op else local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n_items, world)
    per = max(b - a for a, b in parts) if parts else 0
    if per == 0:
        return PendingGather([], lambda: local) if async_op else local
    is_c = local.is_complex()
    buf = (torch.view_as_real(local) if is_c else local).contiguous()
    tail = tuple(buf.shape[1:])

    def done(full):
        return torch.view_as_complex(full) if is_c else full

    if dst is not None:
        ops, out = [], None
        if rank == dst:
            out = torch.empty((n_items,) + tail, dtype=buf.dtype, device=buf.device)
            a, b = parts[rank]
            out[a:b].copy_(buf)
            for r, (a, b) in enumerate(parts):
                if r != rank and b > a:
                    ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
        elif buf.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, buf, dst, group))
        works = dist.batch_isend_irecv(ops) if ops else []
        pend = PendingGather(works, lambda: done(out) if out is not None else None)
        return pend if async_op else pend.wait()
    even = all(b - a == per for a, b in parts)
    if even:
        out = torch.empty((n_items,) + tail, dtype=buf.dtype, device=buf.device)
        w = dist.all_gather_into_tensor(out, buf, group=group, async_op=True)
        pend = PendingGather([w], lambda: done(out))
        return pend if async_op else pend.wait()
    pad = torch.zeros((per,) + tail, dtype=buf.dtype, device=buf.device)
    pad[: buf.shape[0]] = buf
    out = torch.empty((world * per,) + tail, dtype=buf.dtype, device=buf.device)
    w = dist.all_gather_into_tensor(out, pad, group=group, async_op=True)

    def finish():
        return done(torch.cat([out[r * per: r * per + (b - a)] for r, (a, b) in enumerate(parts)], 0).contiguous())

    pend = PendingGather([w], finish)
    return pend if async_op else pend.wait()


def run_sharded(fn: Callable[..., torch.Tensor], batch_tensors: Sequence[Optional[torch.Tensor]],
                n_items: Optional[int] = None, gather: bool = True, group=None, dst: Optional[int] = None,
                async_op: bool = False):
    """Run ``fn(*local_tensors) -> [n_local, ...]`` on this rank's block of slices and gather the result (to every
    rank, or to rank ``dst`` only; ``async_op`` as in ``gather_reconstructions``)."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if n_items is None:
        n_items = max(t.shape[0] for t in batch_tensors if t is not None)
    local, (a, b) = shard_slices(batch_tensors, rank, world, n_items)
    if b > a:
        res = fn(*local)
    else:  # this rank has no slice: contribute an empty block of the right trailing shape
        probe = fn(*[t[:1] if (t is not None and t.shape[0] == n_items) else t for t in batch_tensors])
        res = probe[:0]
    if not gather:
        return res
    return gather_reconstructions(res, n_items, group, dst=dst, async_op=async_op)
