"""Slice-sharded multi-GPU driver.

Every slice is an independent unit (no halo, no cross-slice state: the RIM hidden state is per sample and is
reset per cascade, cirim.py:148), so the N slices of a volume / batch are split into contiguous blocks of
ceil(N/G) slices per rank (keeps a volume's slices together, the output order the reference's
``test_epoch_end`` expects, reconstruction/models/base.py:576-581).  Weights are replicated; there is NO
collective on the data path; the only communication is one all-gather of the [n_local, h, w] complex64
reconstructions at the end (NCCL over NVLink on GPUs; gloo in the CPU unit tests).
One process per GPU (torchrun); nothing here is specific to a backend.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["partition", "shard_slices", "gather_reconstructions", "run_sharded"]


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


def gather_reconstructions(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-rank [n_local, ...] results into [n_items, ...] on every rank (one collective)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    parts = partition(n_items, world)
    per = max(b - a for a, b in parts) if parts else 0
    if per == 0:
        return local
    is_c = local.is_complex()
    buf = torch.view_as_real(local) if is_c else local
    pad = torch.zeros((per,) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
    pad[: buf.shape[0]] = buf
    out = torch.empty((world * per,) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    pieces = [out[r * per: r * per + (b - a)] for r, (a, b) in enumerate(parts)]
    full = torch.cat(pieces, 0)
    return torch.view_as_complex(full.contiguous()) if is_c else full


def run_sharded(fn: Callable[..., torch.Tensor], batch_tensors: Sequence[Optional[torch.Tensor]],
                n_items: Optional[int] = None, gather: bool = True, group=None) -> torch.Tensor:
    """Run ``fn(*local_tensors) -> [n_local, ...]`` on this rank's block of slices and gather the result."""
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
    return gather_reconstructions(res, n_items, group)
