"""Timeslice data parallelism: rank r of R owns t in [r*Lt/R, (r+1)*Lt/R); the only exchange is the
gather of the results.  Works on NCCL (GPU tensors) and gloo (CPU tensors).

`TimesliceGatherer` streams every rank's finished timeslices, a few at a time, straight into ONE
preallocated [Lt, ...] buffer on the destination rank while the next ones are being computed: the
transfers run on the communicator's own stream, so only the last chunk is not overlapped, and the
destination holds the result once (its own timeslices are computed in place, no staging, no
concatenation)."""
from __future__ import annotations

from typing import List, Optional, Tuple


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


def _real(x):
    import torch

    return torch.view_as_real(x) if x.is_complex() else x  # NCCL has no complex type: complex128 travels as float64 pairs


class TimesliceGatherer:
    """Result buffer of a sharded run plus the sends / receives that fill it.

    Every rank calls `push(a, b)` when its local timeslices [a, b) (indices inside its own range) are
    final on the current stream, with the same chunk boundaries on every rank, and `finish()` at the
    end.  `local` is where a rank writes its timeslices: on the destination a view of the final buffer."""

    def __init__(self, Lt: int, tail_shape, dtype, device, group=None, dst: Optional[int] = 0, chunk: int = 4):
        import torch

        self.group, self.Lt, self.chunk = group, int(Lt), max(1, int(chunk))
        self.rank, self.size = world(group)
        self.everyone = dst is None
        self.dst = 0 if dst is None else int(dst)
        self.ranges: List[Tuple[int, int]] = [timeslice_range(self.Lt, r, self.size) for r in range(self.size)]
        self.t0, self.t1 = self.ranges[self.rank]
        self.n_local = self.t1 - self.t0
        tail = tuple(int(s) for s in tail_shape)
        self.is_dst = self.rank == self.dst
        if self.is_dst or self.everyone:
            self.out = torch.empty((self.Lt,) + tail, dtype=dtype, device=device)
            self.local = self.out[self.t0 : self.t1]
        else:
            self.out = None
            self.local = torch.empty((self.n_local,) + tail, dtype=dtype, device=device)
        self.works = []
        self._posted = [0] * self.size  # destination: receives posted for local indices [0, _posted[r]) of rank r

    def _global_rank(self, r: int) -> int:
        import torch.distributed as dist

        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def push(self, a: int, b: int):
        """Local timeslices [a, b) of this rank are final on the current stream and on their way to the destination.
        Every rank sends its range in chunks [k chunk, min((k+1) chunk, n_r)); the destination posts the matching
        receives - the senders' chunk boundaries, not its own - for every chunk that starts before b."""
        if self.size == 1 or b <= a:
            return
        import torch.distributed as dist

        ops = []
        if self.is_dst:
            for r, (r0, r1) in enumerate(self.ranges):
                if r == self.dst:
                    continue
                n_r = r1 - r0
                while self._posted[r] < min(b, n_r):
                    lo = self._posted[r]
                    hi = min((lo // self.chunk + 1) * self.chunk, n_r)
                    ops.append(dist.P2POp(dist.irecv, _real(self.out[r0 + lo : r0 + hi]), self._global_rank(r), self.group))
                    self._posted[r] = hi
        else:
            if a % self.chunk != 0 or (b % self.chunk != 0 and b != self.n_local):
                raise ValueError(f"push({a}, {b}): chunks must be [k*{self.chunk}, (k+1)*{self.chunk}) or end at {self.n_local}")
            lo = a
            while lo < min(b, self.n_local):
                hi = min(lo + self.chunk, self.n_local)
                ops.append(dist.P2POp(dist.isend, _real(self.local[lo:hi]), self._global_rank(self.dst), self.group))
                lo = hi
        if ops:
            self.works.extend(dist.batch_isend_irecv(ops))

    def finish(self):
        """Wait for the transfers; the [Lt, ...] result on the destination (every rank if dst was None), else None."""
        if self.size > 1:
            import torch.distributed as dist

            if self.is_dst:
                self.push(0, max(r1 - r0 for r0, r1 in self.ranges))  # whatever other ranks still owe
            for w in self.works:
                w.wait()
            self.works = []
            if self.everyone:
                dist.broadcast(_real(self.out), src=self._global_rank(self.dst), group=self.group)
        return self.out if (self.is_dst or self.everyone) else None


def gather_timeslices(local, Lt: int, group=None, dst: Optional[int] = 0):
    """local: [t_local, ...] tensor of this rank's range -> [Lt, ...] on rank dst (on every rank if
    dst is None).  One-shot form of `TimesliceGatherer` for results that already exist."""
    rank, size = world(group)
    if size == 1:
        return local
    g = TimesliceGatherer(Lt, local.shape[1:], local.dtype, local.device, group=group, dst=dst)
    if tuple(local.shape) != tuple(g.local.shape):
        raise ValueError(f"rank {rank} owns {g.n_local} timeslices, got {tuple(local.shape)}")
    g.local.copy_(local)
    g.push(0, g.n_local)
    return g.finish()
