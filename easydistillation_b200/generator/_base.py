"""Host logic shared by the two elemental generators: input handles, per-timeslice staging,
the generator-owned result buffer, batch/sharded iteration."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Tuple

import numpy as np

from .. import _capi
from ..constant import Nc, Nd
from ..engine import ElementalEngine


_RAW_C16 = (np.dtype("<c16"), np.dtype(">c16"))
_RAW_EIGVECS = (np.dtype("<c8"), np.dtype("<c16"), np.dtype(">c8"), np.dtype(">c16"))


class _TimesliceGenerator:
    _mode: int = None

    def _setup(self, latt_size, gauge_field, eigenvector, order: int, momentum_list, device):
        Lx, Ly, Lz, Lt = (int(v) for v in latt_size)
        self.latt_size = latt_size
        self.gauge_field = gauge_field
        self.eigenvector = eigenvector
        self.momentum_list = momentum_list
        self.num_momentum = len(momentum_list)
        self.Ne = eigenvector.Ne
        self._U = None
        self._gauge_field_data = None
        self._gauge_field_path = None
        self._eigenvector_data = None
        # device workspace + kernels; raises if the CUDA library or a GPU is missing
        self._engine = ElementalEngine((Lx, Ly, Lz), self.Ne, self._mode, order, list(momentum_list), device)
        # generator-owned, page-locked result buffer: calc() returns it and overwrites it next call
        # (reference: elemental.py:56,295,338), so `data[t] = gen.calc(t)` works unchanged
        self._VPV_pinned = _capi.PinnedBuffer(self._engine.out_shape, np.complex128)
        self._VPV = self._VPV_pinned.array
        self._V_pinned = None
        self._pipeline = None
        self._gauge_ops = []

    # ---- load(key): reference elemental.py:102-105 / displacement_elemental.py:73-76 -------------
    def load(self, key: str):
        if self._pipeline is not None:  # staging / page-locking is tied to the arrays of one configuration
            self._pipeline.close()
            self._pipeline = None
        if self._gauge_ops:  # a freshly loaded configuration is unsmeared, as in the reference
            self._gauge_ops = []
            self._engine.set_link_ops([])
        data = self.gauge_field.load(key)
        U = data[:]
        torch = self._engine.torch
        if isinstance(U, torch.Tensor) or getattr(U, "device_resident", False):
            self._U = U  # already in HBM: a tensor [Lt, ...] or per-timeslice tensors (preset.DeviceTimeslices)
        else:
            U = np.asarray(U)
            Lx, Ly, Lz, Lt = (int(v) for v in self.latt_size)
            if U.ndim == 5 and U.shape == (Lt, Lz * Ly * Lx, Nd, Nc, Nc):  # the flattened default shape of preset.py:142,152
                U = U.reshape(Lt, Lz, Ly, Lx, Nd, Nc, Nc)
            if U.dtype.kind != "c":
                raise ValueError(f"gauge field must be complex, got dtype {U.dtype} (GaugeFieldBinary defaults to '<f8' as in "
                                 "the reference, lattice/preset.py:162-170: pass dtype='<c16')")
            if U.ndim != 7 or U.shape[4:] != (Nd, Nc, Nc):
                raise ValueError(f"gauge field must be [Lt, Lz, Ly, Lx, {Nd}, {Nc}, {Nc}], got {U.shape}")
            if U.shape[:4] != (Lt, Lz, Ly, Lx):
                raise ValueError(f"gauge field shape {U.shape[:4]} does not match latt_size {self.latt_size}")
            # keep file order [Lt][Lz][Ly][Lx][4][3][3]: one timeslice is one contiguous block;
            # the time links are dropped on the device (the reference's [:Nd-1] view).  A big-endian
            # payload (ILDG) stays as read: the device swaps the bytes (filedata/ildg.py:70 does it here)
            self._U = np.ascontiguousarray(U) if U.dtype in _RAW_C16 else np.ascontiguousarray(U, dtype="<c16")
        self._gauge_field_path = getattr(data, "file", None)
        self._gauge_field_data = self._gauge_field_path
        self._eigenvector_data = self.eigenvector.load(key)

    # ---- one timeslice of inputs -----------------------------------------------------------------
    def _eigvecs_view(self, t: int):
        """Zero-copy view of timeslice t's eigenvectors when the handle wraps a host array (our
        ArrayData), else None: lets the streamed pipeline copy source -> pinned staging once."""
        a = getattr(self._eigenvector_data, "_a", None)
        if not isinstance(a, np.ndarray) or a.ndim < 3 or a.dtype not in _RAW_EIGVECS:
            return None
        Lx, Ly, Lz, Lt = (int(v) for v in self.latt_size)
        blk = a[t][: self.Ne]
        return blk.reshape((self.Ne, Lz, Ly, Lx, Nc)) if blk.flags.c_contiguous else None

    def _eigvecs_of(self, t: int):
        """[Ne, Lz, Ly, Lx, Nc] of timeslice t in ONE read when the handle allows it, else the
        reference's per-eigenvector loop (elemental.py:297-298)."""
        ev = self._eigenvector_data
        Lx, Ly, Lz, Lt = (int(v) for v in self.latt_size)
        shape = (self.Ne, Lz, Ly, Lx, Nc)
        torch = self._engine.torch
        if isinstance(ev, torch.Tensor):
            return ev[t, : self.Ne].reshape(shape).contiguous()
        if getattr(ev, "device_resident", False):
            return ev[t][: self.Ne].reshape(shape).contiguous()
        block = None
        try:
            block = np.asarray(ev[t])
        except Exception:
            block = None
        if block is None or block.ndim < 2 or block.shape[0] < self.Ne:
            block = np.stack([np.asarray(ev[t, e]) for e in range(self.Ne)])
        block = block[: self.Ne].reshape(shape)
        if block.dtype in _RAW_EIGVECS:  # big-endian records (QDP timeslice files) are swapped on the device
            return np.ascontiguousarray(block)
        return np.ascontiguousarray(block, dtype="<c16")

    def _check_loaded(self, t):
        if self._U is None or self._eigenvector_data is None:
            raise RuntimeError("call load(key) before calc(t)")
        Lt = int(self.latt_size[3])
        if not 0 <= t < Lt:
            raise IndexError(f"timeslice {t} out of range [0, {Lt})")

    # ---- calc(t): reference elemental.py:290-338 / displacement_elemental.py:78-96 ---------------
    def calc(self, t: int) -> np.ndarray:
        """Elementals of timeslice t, shape (Nop, Nmom, Ne, Ne) complex128, in the generator's own
        (page-locked) buffer: copy it before the next call, as with the reference."""
        self._check_loaded(t)
        eng = self._engine
        torch = eng.torch
        V_t = self._eigvecs_of(t)
        if not self._host_inputs() or isinstance(V_t, torch.Tensor):
            out = self.calc_device(t)
            self._VPV[...] = out.cpu().numpy()
            return self._VPV
        eng.calc_host(self._U[t], _capi.LINKS_FILE_T, V_t, self._VPV)
        return self._VPV

    def calc_device(self, t: int, out=None):
        """Same, result left on the GPU as a torch complex128 tensor (new tensor unless `out`)."""
        self._check_loaded(t)
        eng = self._engine
        torch = eng.torch
        U_t, u_be = self._U[t], False
        if not isinstance(U_t, torch.Tensor):
            U_t, u_be = _capi.raw_view(U_t)
            U_t = torch.from_numpy(U_t if U_t.flags.writeable else U_t.copy()).to(eng.device, non_blocking=False)
        V_t, v_be = self._eigvecs_of(t), False
        if not isinstance(V_t, torch.Tensor):
            V_t, v_be = _capi.raw_view(V_t)
            V_t = torch.from_numpy(V_t if V_t.flags.writeable else V_t.copy()).to(eng.device)
        eng.set_links(U_t.contiguous(), _capi.LINKS_FILE_T | (_capi.LINKS_BIG_ENDIAN if u_be else 0))
        eng.set_eigvecs(V_t.contiguous(), big_endian=v_be)
        return eng.calc(out)

    # ---- batch / sharded form (each rank owns a contiguous t-range) -------------------
    def _host_inputs(self) -> bool:
        torch = self._engine.torch

        def on_device(x):
            return isinstance(x, torch.Tensor) or getattr(x, "device_resident", False)

        return not on_device(self._U) and not on_device(self._eigenvector_data)

    def calc_range(self, t0: int, t1: int) -> np.ndarray:
        """(t1-t0, Nop, Nmom, Ne, Ne) for t in [t0, t1).  Host-resident inputs go through the
        streamed pipeline: upload of t+1 and download of t-1 overlap the kernels of t."""
        self._check_loaded(t0)
        if t1 > t0:
            self._check_loaded(t1 - 1)
        out = np.empty((max(t1 - t0, 0),) + self._engine.out_shape, np.complex128)
        if self._host_inputs():
            from ..pipeline import TimeslicePipeline

            if self._pipeline is None:
                self._pipeline = TimeslicePipeline(self)
            return self._pipeline.run_host(range(t0, t1), out)
        for i, t in enumerate(range(t0, t1)):
            out[i] = self.calc(t)
        return out

    def calc_to_file(self, elemental, key: str, t_range: Optional[Tuple[int, int]] = None, dtype: str = "<c16", group=None,
                     shard: bool = False):
        """Write the elemental file of configuration `key` in the reference's layout
        [Nop, Nmom, Lt, Ne, Ne] (tests/test_elemental.py:47, read back by ElementalNpy / lattice/data.py:26).

        `elemental` is an ElementalNpy-like handle; the file is pre-sized once and every batch of timeslices drops
        its slab in place on a writer thread while the next batch is being computed.  `dtype="<c8"` down-casts to the
        complex64 the reference declares for stored elementals (preset.py:176).

        Sharded runs: `shard=True` writes this rank's timeslice range of the torch.distributed `group`,
        `t_range=(t0, t1)` an explicit slab; ranks share the file.  Creation is separate from slab writing: under
        torch.distributed rank 0 makes sure the file exists (replacing one of another shape) and every rank waits at
        a barrier before it opens it; independent processes (no process group) create a missing file atomically and
        never truncate an existing one of the right shape, so slabs may be written in any order."""
        from ..sharding import timeslice_range, world

        Lt = int(self.latt_size[3])
        rank, size = world(group)
        if shard:
            if t_range is not None:
                raise ValueError("shard=True derives the timeslice range from the rank; do not pass t_range too")
            t_range = timeslice_range(Lt, rank, size)
        t0, t1 = (0, Lt) if t_range is None else t_range
        if not 0 <= t0 <= t1 <= Lt:
            raise IndexError(f"timeslice range {t_range} outside [0, {Lt}]")
        shape = (self._engine.out_shape[0], self._engine.out_shape[1], Lt, self.Ne, self.Ne)
        if not hasattr(elemental, "ensure"):  # a foreign handle with create / open_rw only: whole-file runs
            mm = elemental.create(key, shape, dtype) if t_range is None else elemental.open_rw(key, shape, dtype)
        elif size > 1:
            import torch.distributed as dist

            if rank == 0:
                elemental.ensure(key, shape, dtype, replace=True)
            dist.barrier(group)
            mm = elemental.open_rw(key, shape, dtype)
        else:
            elemental.ensure(key, shape, dtype, replace=t_range is None)
            mm = elemental.open_rw(key, shape, dtype)
        chunk = 4  # timeslices per streamed batch: bounds the host buffers (two batches alive at a time)

        def drop(a, b, block):
            # transposing (and down-casting) copy into the file mapping.  Fresh page-cache pages make this a page-fault-
            # bound copy (1.0 GB/s on one core, measured on the B200 box: slower than the GPU produces results); split over
            # the operator axis it reaches 2.4 GB/s, just ahead of config 5's 275 MB per 0.12 s.
            from ..pipeline import parallel_copy

            parallel_copy(mm[:, :, a:b], block.transpose(1, 2, 0, 3, 4))

        # one writer thread: batch k goes to the file while batch k+1 is being computed
        with ThreadPoolExecutor(max_workers=1) as writer:
            pending = None
            for a in range(t0, t1, chunk):
                b = min(a + chunk, t1)
                block = self.calc_range(a, b)  # [b-a, Nop, Nmom, Ne, Ne], a fresh array per batch
                if pending is not None:
                    pending.result()
                pending = writer.submit(drop, a, b, block)
            if pending is not None:
                pending.result()
        mm.flush()
        return mm

    def calc_all(self, group=None, dst: Optional[int] = 0, chunk: int = 4):
        """All Lt timeslices, sharded over the ranks of `group` (default: the world group if torch.distributed is
        initialised, else this process alone): rank r computes its contiguous range and its finished timeslices
        travel, `chunk` at a time, straight into the one [Lt, Nop, Nmom, Ne, Ne] buffer on rank `dst` while the next
        ones are computed (sharding.TimesliceGatherer) - the only exchange of the path.  Returns that device tensor on
        rank `dst` (on every rank if dst is None), None elsewhere."""
        from ..sharding import TimesliceGatherer

        torch = self._engine.torch
        Lt = int(self.latt_size[3])
        g = TimesliceGatherer(Lt, self._engine.out_shape, torch.complex128, self._engine.device, group=group, dst=dst, chunk=chunk)
        local, n = g.local, g.n_local
        if n:
            self._check_loaded(g.t0)
            self._check_loaded(g.t1 - 1)
        sent = 0

        def done(i):  # local timeslice i is queued on the current stream: hand over full chunks
            nonlocal sent
            if (i + 1) % g.chunk == 0 or i + 1 == n:
                g.push(sent, i + 1)
                sent = i + 1

        if self._host_inputs():
            from ..pipeline import TimeslicePipeline

            if self._pipeline is None:
                self._pipeline = TimeslicePipeline(self)
            self._pipeline.run_device(range(g.t0, g.t1), local, on_done=done)
        else:
            for i, t in enumerate(range(g.t0, g.t1)):
                self.calc_device(t, out=local[i])
                done(i)
        return g.finish()

    # ---- gauge preprocessing of the reference classes -----------------------------
    # The reference rewrites the whole loaded configuration at once (elemental.py:107-117,264-277).
    # Spatial smearing / projection never couple timeslices, so here the request is recorded and the
    # sm_100a kernels process each timeslice's links on the device right after they are uploaded.
    def stout_smear(self, nstep, rho):
        if self._U is None:
            raise RuntimeError("call load(key) before stout_smear")
        if nstep < 0:
            raise ValueError("nstep must be >= 0")
        self._gauge_ops.append(("stout", int(nstep), float(rho)))
        self._engine.set_link_ops(self._gauge_ops)

    def project_SU3(self):
        if self._U is None:
            raise RuntimeError("call load(key) before project_SU3")
        self._gauge_ops.append(("project",))
        self._engine.set_link_ops(self._gauge_ops)
