"""Thin object wrapper over one edk handle: one (GPU, lattice, Ne, operator set)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _capi


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class ElementalEngine:
    """Owns the device workspace of one generator and launches the sm_100a kernels.

    `mode`/`order`: (_capi.MODE_DERIVATIVE, num_nabla) or (_capi.MODE_DISPLACEMENT, distance).
    Tensors passed in must live on `device`; the current torch stream is used."""

    def __init__(self, latt3: Sequence[int], Ne: int, mode: int, order: int, momentum_list: List[Tuple[int]], device=None):
        torch = _capi.require_cuda()
        self.torch = torch
        self.lib = _capi.lib()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure torch's primary context exists first
        Lx, Ly, Lz = (int(v) for v in latt3)
        self.latt3 = (Lx, Ly, Lz)
        self.Ne = int(Ne)
        self.mode, self.order = int(mode), int(order)
        mom = np.ascontiguousarray(np.asarray(momentum_list, dtype=np.int32).reshape(-1, 3))
        self.nmom = mom.shape[0]
        self.momenta = [tuple(int(c) for c in m) for m in mom]
        h = C.c_void_p()
        rc = self.lib.edk_create(Lx, Ly, Lz, self.Ne, self.mode, self.order, self.nmom,
                                 mom.ctypes.data_as(C.POINTER(C.c_int)), self.device.index, C.byref(h))
        _capi.check(rc, "edk_create")
        self.h = h
        self.nop = self.lib.edk_num_operators(self.h)
        self.out_shape = (self.nop, self.nmom, self.Ne, self.Ne)
        self.field_shape = (self.Ne, Lz, Ly, Lx, 3)

    # -- lifetime -----------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.edk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.edk_workspace_bytes(self.h))

    @property
    def output_bytes(self) -> int:
        return int(self.lib.edk_output_bytes(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.edk_launch_count(self.h))

    # -- device-pointer path --------------------------------------------------------------
    def set_links(self, U, layout: int):
        torch = self.torch
        V = self.latt3[0] * self.latt3[1] * self.latt3[2]
        ndir = 4 if (layout & ~_capi.LINKS_BIG_ENDIAN) == _capi.LINKS_FILE_T else 3
        if U.dtype != torch.complex128 or U.numel() != ndir * V * 9 or not U.is_contiguous() or U.device != self.device:
            raise ValueError(f"links must be a contiguous complex128 tensor of {ndir}*V*9 elements on {self.device}")
        _capi.check(self.lib.edk_set_links(self.h, _ptr(U), layout, self._stream()), "edk_set_links")

    def set_eigvecs(self, Vt, big_endian: bool = False):
        """`big_endian`: the tensor holds the raw bytes of a big-endian file record."""
        torch = self.torch
        if Vt.dtype not in (torch.complex64, torch.complex128):
            raise ValueError("eigenvectors must be complex64 or complex128")
        if Vt.numel() != int(np.prod(self.field_shape)) or not Vt.is_contiguous() or Vt.device != self.device:
            raise ValueError(f"eigenvectors must be a contiguous tensor of shape {self.field_shape} on {self.device}")
        flags = (_capi.EIGVECS_C8 if Vt.dtype == torch.complex64 else 0) | (_capi.EIGVECS_BIG_ENDIAN if big_endian else 0)
        _capi.check(self.lib.edk_set_eigvecs(self.h, _ptr(Vt), flags, self._stream()), "edk_set_eigvecs")

    def set_link_ops(self, ops):
        """ops: list of ("stout", nstep, rho) / ("project",) applied to every timeslice's links."""
        n = len(ops)
        kinds = (C.c_int * max(n, 1))(*[1 if o[0] == "stout" else 2 for o in ops])
        nsteps = (C.c_int * max(n, 1))(*[int(o[1]) if o[0] == "stout" else 0 for o in ops])
        rhos = (C.c_double * max(n, 1))(*[float(o[2]) if o[0] == "stout" else 0.0 for o in ops])
        _capi.check(self.lib.edk_set_link_ops(self.h, n, kinds, nsteps, rhos), "edk_set_link_ops")

    def debug_links(self):
        Lx, Ly, Lz = self.latt3
        out = self.torch.empty((3, Lz, Ly, Lx, 3, 3), dtype=self.torch.complex128, device=self.device)
        _capi.check(self.lib.edk_debug_links(self.h, _ptr(out), self._stream()), "edk_debug_links")
        return out

    def set_blending(self, coeff):
        if coeff is None:
            _capi.check(self.lib.edk_set_blending(self.h, None, self._stream()))
            return
        torch = self.torch
        c = torch.as_tensor(np.ascontiguousarray(coeff, dtype=np.float64)).to(self.device)
        if tuple(c.shape) != (self.Ne, self.Ne):
            raise ValueError("blending matrix must be (Ne, Ne)")
        _capi.check(self.lib.edk_set_blending(self.h, _ptr(c), self._stream()), "edk_set_blending")
        torch.cuda.current_stream(self.device).synchronize()

    def calc(self, out=None):
        torch = self.torch
        if out is None:
            out = torch.empty(self.out_shape, dtype=torch.complex128, device=self.device)
        if out.dtype != torch.complex128 or tuple(out.shape) != self.out_shape or not out.is_contiguous():
            raise ValueError(f"out must be contiguous complex128 of shape {self.out_shape}")
        _capi.check(self.lib.edk_calc(self.h, _ptr(out), self._stream()), "edk_calc")
        return out

    def laplacian(self, F, out=None):
        """out[v] = (6 - hops) F[v] on the current links; F [nvec, Lz, Ly, Lx, 3] complex128 on the device."""
        torch = self.torch
        Lx, Ly, Lz = self.latt3
        if F.dtype != torch.complex128 or F.dim() != 5 or tuple(F.shape[1:]) != (Lz, Ly, Lx, 3) or not F.is_contiguous() \
                or F.device != self.device:
            raise ValueError(f"F must be a contiguous complex128 tensor [nvec, {Lz}, {Ly}, {Lx}, 3] on {self.device}")
        if out is None:
            out = torch.empty_like(F)
        _capi.check(self.lib.edk_laplacian(self.h, _ptr(F), _ptr(out), int(F.shape[0]), self._stream()), "edk_laplacian")
        return out

    # -- host-buffer path (what the generators' calc(t) and bench e2e use) ------------------
    def calc_host(self, U_host: np.ndarray, layout: int, V_host: np.ndarray, out_host: np.ndarray):
        """Big-endian numpy arrays (`>c16` links, `>c8`/`>c16` eigenvectors: raw file payloads) are
        uploaded as they are and byte-swapped by the kernels that read them."""
        U_host, u_be = _capi.raw_view(U_host)
        V_host, v_be = _capi.raw_view(V_host)
        layout = (layout & ~_capi.LINKS_BIG_ENDIAN) | (_capi.LINKS_BIG_ENDIAN if u_be or layout & _capi.LINKS_BIG_ENDIAN else 0)
        V = self.latt3[0] * self.latt3[1] * self.latt3[2]
        ndir = 4 if (layout & ~_capi.LINKS_BIG_ENDIAN) == _capi.LINKS_FILE_T else 3
        if U_host.dtype != np.complex128 or U_host.size != ndir * V * 9 or not U_host.flags.c_contiguous:
            raise ValueError(f"links must be C-contiguous complex128 with {ndir}*V*9 elements")
        if V_host.dtype not in (np.complex64, np.complex128) or V_host.size != int(np.prod(self.field_shape)) \
                or not V_host.flags.c_contiguous:
            raise ValueError(f"eigenvectors must be C-contiguous complex64/128 of shape {self.field_shape}")
        if out_host.dtype != np.complex128 or out_host.shape != self.out_shape or not out_host.flags.c_contiguous:
            raise ValueError(f"out must be C-contiguous complex128 of shape {self.out_shape}")
        vflags = (_capi.EIGVECS_C8 if V_host.dtype == np.complex64 else 0) | (_capi.EIGVECS_BIG_ENDIAN if v_be else 0)
        rc = self.lib.edk_calc_host(self.h, _np_ptr(U_host), layout, _np_ptr(V_host), vflags, _np_ptr(out_host),
                                    self._stream())
        _capi.check(rc, "edk_calc_host")
        return out_host

    # -- measurement / test hooks -------------------------------------------------------------
    def set_profiling(self, on: bool):
        _capi.check(self.lib.edk_set_profiling(self.h, int(on)))

    def get_profile(self):
        ms = (C.c_double * 4)()
        nl = (C.c_int * 4)()
        _capi.check(self.lib.edk_get_profile(self.h, ms, nl), "edk_get_profile")
        keys = ("prepare", "stencil", "contraction", "combine")
        return {k: {"ms": ms[i], "launches": nl[i]} for i, k in enumerate(keys)}

    def debug_field(self, idx: int):
        out = self.torch.empty(self.field_shape, dtype=self.torch.complex128, device=self.device)
        _capi.check(self.lib.edk_debug_field(self.h, idx, _ptr(out), self._stream()), "edk_debug_field")
        return out

    def debug_phase(self, ip: int):
        Lx, Ly, Lz = self.latt3
        out = self.torch.empty((Lz, Ly, Lx), dtype=self.torch.complex128, device=self.device)
        _capi.check(self.lib.edk_debug_phase(self.h, ip, _ptr(out), self._stream()), "edk_debug_phase")
        return out

    def debug_use_naive_gram(self, on: bool):
        _capi.check(self.lib.edk_debug_use_naive_gram(self.h, int(on)))

    def debug_gram_config(self, mfrag: int = 0, ksplit: int = 0):
        _capi.check(self.lib.edk_debug_gram_config(self.h, mfrag, ksplit), "edk_debug_gram_config")

    def debug_symmetry(self, mode: int):
        """-1 auto, 0 off, 1 force the Hermitian pairing of (left, right) field pairs."""
        _capi.check(self.lib.edk_debug_symmetry(self.h, int(mode)), "edk_debug_symmetry")

    def debug_loader(self, mode: int):
        """0 = TMA producer warp (default), 1 = cp.async loader inside the MMA warps."""
        _capi.check(self.lib.edk_debug_loader(self.h, int(mode)), "edk_debug_loader")

    def debug_algo(self, algo: int):
        """Ask for one contraction form instead of the planned one (tests / A-B measurements): 1 = GEMM form, 3M
        arithmetic, 0 = GEMM form, 4M, 2 = plane-wave factorised form, 3 = plane-wave form with centre-symmetric site
        pairs folded, 4 = separable form, -1 = back to the form the library plans (include/edk.h, edk_debug_algo)."""
        _capi.check(self.lib.edk_debug_algo(self.h, int(algo)), "edk_debug_algo")

    def query(self):
        q = lambda w: int(self.lib.edk_query(self.h, w))  # noqa: E731
        return {"hermitian_pairing": bool(q(0)), "internal_momenta": q(1), "pair_gemms_per_momentum": q(2),
                "ksplit": q(3), "mfrag": q(4), "jobs": q(5), "tma_stages": q(6),
                "real_mma_per_complex_block": q(7), "pair_momentum_gemms": q(8), "half_set_momenta": q(9),
                "contraction_form": q(10), "plane_wave_modes": q(11), "plane_wave_tile": q(12),
                "form_requested": q(13), "pairs_per_stage": q(14)}


def microbench_fp64(device: int = 0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) sustained on `device`."""
    _capi.require_cuda()
    a, b = C.c_double(), C.c_double()
    _capi.check(_capi.lib().edk_microbench_fp64(device, C.byref(a), C.byref(b)), "edk_microbench_fp64")
    return a.value, b.value
