"""Streamed input pipeline: while timeslice t is being contracted, the inputs of
t+1 are read ONCE from the handles into page-locked staging buffers and uploaded on a side
stream, and the result of t-1 drains to the host on a third stream.  The reference instead does
Ne separate open+mmap+copy calls per timeslice and a blocking upload (elemental.py:297-298,
filedata/ndarray.py:17-47)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import numpy as np

from . import _capi


_COPY_POOL = None
_COPY_THREADS = 8


def parallel_copy(dst: np.ndarray, src: np.ndarray, min_bytes: int = 8 << 20) -> None:
    """dst[...] = src with the leading axis split over a few threads (numpy releases the GIL while it copies).
    A result of 275 MB into freshly allocated pages is a page-fault-bound copy of ~0.2 s on one core, which is as
    long as a whole timeslice once the contraction is fast; four threads keep the host side out of the way."""
    global _COPY_POOL
    n = dst.shape[0] if dst.ndim else 0
    if dst.shape != src.shape:
        raise ValueError(f"parallel_copy: shapes differ, {dst.shape} vs {src.shape}")
    if dst.nbytes < min_bytes or n < 2:
        np.copyto(dst, src)
        return
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _COPY_POOL = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="edk-copy")
    parts = min(_COPY_THREADS, n)
    bounds = [n * k // parts for k in range(parts + 1)]
    futures = [_COPY_POOL.submit(np.copyto, dst[a:b], src[a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    for f in futures:
        f.result()  # re-raises a worker's exception


class TimeslicePipeline:
    """Double-buffered H2D / compute / D2H over a list of timeslices of one generator."""

    def __init__(self, gen):
        self.gen = gen
        eng = gen._engine
        torch = eng.torch
        self.torch = torch
        self.eng = eng
        Lx, Ly, Lz, Lt = (int(v) for v in gen.latt_size)
        V = Lx * Ly * Lz
        self.copy_stream = torch.cuda.Stream(device=eng.device)
        self.out_stream = torch.cuda.Stream(device=eng.device)
        self.U_pin = [torch.empty((V, 4, 3, 3), dtype=torch.complex128, pin_memory=True) for _ in range(2)]
        self.U_dev = [torch.empty((V, 4, 3, 3), dtype=torch.complex128, device=eng.device) for _ in range(2)]
        self.V_pin = [None, None]
        self.V_dev = [None, None]
        self.ev_h2d = [None, None]
        self.ev_consumed = [None, None]
        self.U_be = [False, False]  # slot holds a raw big-endian payload (swapped by the kernels that read it)
        self.V_be = [False, False]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        # In-memory sources are page-locked in place (cudaHostRegister) so their timeslices go to the
        # device by DMA straight from where they are; files / foreign handles go through the staging copy.
        self._registered = []
        self._U_direct = self._register(gen._U)
        ev = getattr(gen._eigenvector_data, "_a", None)
        self._V_direct = self._register(ev) if gen._eigvecs_view(0) is not None else False

    def _register(self, a) -> bool:
        if not isinstance(a, np.ndarray) or isinstance(a, np.memmap) or not a.flags.c_contiguous or a.nbytes == 0:
            return False
        if not a.flags.writeable or _capi.lib().edk_host_register(C.c_void_p(a.ctypes.data), a.nbytes) != _capi.EDK_OK:
            return False
        self._registered.append(a)
        return True

    def close(self):
        for a in self._registered:
            _capi.lib().edk_host_unregister(C.c_void_p(a.ctypes.data))
        self._registered = []

    def __del__(self):
        try:
            self.torch.cuda.synchronize(self.eng.device)
            self.close()
        except Exception:
            pass

    def _stage(self, b: int, t: int):
        """Host side of one timeslice: source handles -> pinned staging -> async upload."""
        torch = self.torch
        gen = self.gen
        if self.ev_h2d[b] is not None:
            self.ev_h2d[b].synchronize()  # the previous upload from this staging slot has left the host
        U_t = gen._U[t]
        V_t = gen._eigvecs_view(t)
        if V_t is None:
            V_t = gen._eigvecs_of(t)
        if isinstance(U_t, torch.Tensor) or isinstance(V_t, torch.Tensor):
            raise TypeError("the streamed pipeline is for host-resident inputs")
        U_t, self.U_be[b] = _capi.raw_view(U_t)
        V_t, self.V_be[b] = _capi.raw_view(V_t)
        tdt = torch.complex64 if V_t.dtype == np.complex64 else torch.complex128
        if self.V_dev[b] is None or self.V_dev[b].dtype != tdt:
            self.V_dev[b] = torch.empty(V_t.shape, dtype=tdt, device=self.eng.device)
        if self._U_direct:
            U_src = torch.from_numpy(U_t).reshape(self.U_dev[b].shape)
        else:
            np.copyto(self.U_pin[b].numpy().reshape(U_t.shape), U_t)
            U_src = self.U_pin[b]
        if self._V_direct:
            V_src = torch.from_numpy(V_t)
        else:
            if self.V_pin[b] is None or self.V_pin[b].dtype != tdt:
                self.V_pin[b] = torch.empty(V_t.shape, dtype=tdt, pin_memory=True)
            parallel_copy(self.V_pin[b].numpy(), V_t)  # file / memory-map reads of ~0.5 - 1 GB: a few threads, not one
            V_src = self.V_pin[b]
        with torch.cuda.stream(self.copy_stream):
            if self.ev_consumed[b] is not None:
                self.copy_stream.wait_event(self.ev_consumed[b])  # device slot free again
            self.U_dev[b].copy_(U_src, non_blocking=True)
            self.V_dev[b].copy_(V_src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.ev_h2d[b] = ev
        self.h2d_bytes += self.U_dev[b].numel() * 16 + self.V_dev[b].numel() * self.V_dev[b].element_size()

    def _compute(self, b: int, out):
        torch = self.torch
        cur = torch.cuda.current_stream(self.eng.device)
        cur.wait_event(self.ev_h2d[b])
        self.eng.set_links(self.U_dev[b], _capi.LINKS_FILE_T | (_capi.LINKS_BIG_ENDIAN if self.U_be[b] else 0))
        self.eng.set_eigvecs(self.V_dev[b], big_endian=self.V_be[b])
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ev_consumed[b] = ev
        return self.eng.calc(out)

    def run_device(self, timeslices: Iterable[int], out, on_done=None):
        """out[i] (device tensor [n, Nop, Nmom, Ne, Ne]) <- elementals of timeslices[i]; `on_done(i)` is called once the
        kernels of timeslice i are queued (the sharded run hands finished chunks to the gather from it)."""
        ts = list(timeslices)
        if not ts:
            return out
        self._stage(0, ts[0])
        for i, t in enumerate(ts):
            if i + 1 < len(ts):
                self._stage((i + 1) & 1, ts[i + 1])  # overlaps with the kernels of timeslice t-1 / t
            self._compute(i & 1, out[i])
            if on_done is not None:
                on_done(i)
        return out

    def run_host(self, timeslices: Iterable[int], out: np.ndarray):
        """out[i] (numpy [n, Nop, Nmom, Ne, Ne]) <- elementals of timeslices[i], D2H on a third stream."""
        torch = self.torch
        ts = list(timeslices)
        if not ts:
            return out
        dev = [torch.empty(self.eng.out_shape, dtype=torch.complex128, device=self.eng.device) for _ in range(2)]
        pin = [torch.empty(self.eng.out_shape, dtype=torch.complex128, pin_memory=True) for _ in range(2)]
        ev_out = [None, None]
        ev_drained = [None, None]
        cur = torch.cuda.current_stream(self.eng.device)

        def flush(j):  # host copy of result j once its D2H has completed
            ev_out[j & 1].synchronize()
            parallel_copy(out[j], pin[j & 1].numpy())

        self._stage(0, ts[0])
        for i, t in enumerate(ts):
            b = i & 1
            if i + 1 < len(ts):
                self._stage((i + 1) & 1, ts[i + 1])
            if ev_drained[b] is not None:
                cur.wait_event(ev_drained[b])  # result slot b has been copied out
            self._compute(b, dev[b])
            done = torch.cuda.Event()
            done.record(cur)
            if i >= 1:
                # Timeslice i is queued behind everything the device still has to do, so the host can now wait for
                # the download of i-1 (complete about when the kernels of i start) and copy it out while i computes;
                # this also frees the other pinned slot long before timeslice i+1 needs it.
                flush(i - 1)
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(done)
                pin[b].copy_(dev[b], non_blocking=True)
                e1 = torch.cuda.Event()
                e1.record(self.out_stream)
                ev_out[b] = e1
                ev_drained[b] = e1
            self.d2h_bytes += pin[b].numel() * 16
        flush(len(ts) - 1)
        return out
