"""Timeslice data parallelism: rank r of R owns t in [r*Lt/R, (r+1)*Lt/R); one gather at the
end, nothing per timeslice.  Works on NCCL (GPU tensors) and gloo (CPU tensors)."""
from __future__ import annotations

from typing import Optional, Tuple


def world(group=None) -> Tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def timeslice_range(Lt: int, rank: int, size: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first Lt % size ranks take one extra timeslice."""
    if not 0 <= rank < size:
        raise ValueError("rank out of range")
    base, extra = divmod(Lt, size)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def gather_timeslices(local, Lt: int, group=None, dst: Optional[int] = 0):
    """local: [t_local, ...] complex tensor of this rank's range -> [Lt, ...] on rank dst
    (on every rank if dst is None).  complex128 travels as float64 pairs (NCCL has no complex)."""
    import torch
    import torch.distributed as dist

    rank, size = world(group)
    if size == 1:
        return local
    tmax = -(-Lt // size)
    tail = tuple(local.shape[1:])
    real = torch.view_as_real(local.contiguous())
    padded = torch.zeros((tmax,) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
    padded[: real.shape[0]] = real
    everyone = dst is None
    if everyone:
        buf = torch.empty((size * tmax,) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
    else:
        # NCCL's gather needs the list only on dst; gloo's too
        parts = [torch.empty_like(padded) for _ in range(size)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
        buf = torch.cat(parts, 0)
    out = torch.empty((Lt,) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
    for r in range(size):
        t0, t1 = timeslice_range(Lt, r, size)
        out[t0:t1] = buf[r * tmax : r * tmax + (t1 - t0)]
    return torch.view_as_complex(out).reshape((Lt,) + tail)
