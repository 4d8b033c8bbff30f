"""ctypes binding of include/edk.h (libedk_sm100a.so).  No CPU fallback: if the library
is missing, or there is no CUDA device when a kernel is asked for, this raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# EDK_LIBRARY points at another build of the same C ABI (A/B runs of an experimental kernel build)
LIB_PATH = os.environ.get("EDK_LIBRARY") or os.path.join(PKG, "libedk_sm100a.so")

EDK_OK = 0
EDK_ERR_ARG = -1
EDK_ERR_CUDA = -2
EDK_ERR_STATE = -3
EDK_ERR_NOMEM = -4
MODE_DERIVATIVE = 0
MODE_DISPLACEMENT = 1
LINKS_DIR_MAJOR = 0
LINKS_FILE_T = 1
LINKS_BIG_ENDIAN = 0x100  # OR-ed into the layout: raw big-endian file payload, swapped on the device
EIGVECS_C8 = 1
EIGVECS_BIG_ENDIAN = 2


def raw_view(a):
    """(little-endian-typed view of the same bytes, big_endian?) of a numpy complex array: big-endian
    file payloads go to the device as they are and are byte-swapped there."""
    if a.dtype.byteorder == ">":
        return a.view(a.dtype.newbyteorder("<")), True
    return a, False

# every symbol include/edk.h declares: (restype, argtypes)
_vp, _i, _sz, _dp = C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_double)
SIGNATURES = {
    "edk_version": (_i, []),
    "edk_last_error": (C.c_char_p, []),
    "edk_create": (_i, [_i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, C.POINTER(_vp)]),
    "edk_destroy": (_i, [_vp]),
    "edk_phase_table": (_i, [_i, _i, _i, _i, C.POINTER(_i), _vp, _i, _vp]),
    "edk_plan": (_i, [_i, _i, _i, C.POINTER(_i), _i, C.POINTER(_i)]),
    "edk_plan_modes": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "edk_plan_form": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "edk_plan_tiles": (_i, [_i, _i, C.POINTER(_i)]),
    "edk_num_operators": (_i, [_vp]),
    "edk_output_bytes": (_sz, [_vp]),
    "edk_workspace_bytes": (_sz, [_vp]),
    "edk_set_links": (_i, [_vp, _vp, _i, _vp]),
    "edk_set_eigvecs": (_i, [_vp, _vp, _i, _vp]),
    "edk_set_link_ops": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), _dp]),
    "edk_debug_links": (_i, [_vp, _vp, _vp]),
    "edk_set_blending": (_i, [_vp, _vp, _vp]),
    "edk_calc": (_i, [_vp, _vp, _vp]),
    "edk_calc_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "edk_laplacian": (_i, [_vp, _vp, _vp, _i, _vp]),
    "edk_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "edk_host_free": (_i, [_vp]),
    "edk_host_register": (_i, [_vp, _sz]),
    "edk_host_unregister": (_i, [_vp]),
    "edk_set_profiling": (_i, [_vp, _i]),
    "edk_get_profile": (_i, [_vp, _dp, C.POINTER(_i)]),
    "edk_launch_count": (C.c_longlong, [_vp]),
    "edk_debug_field": (_i, [_vp, _i, _vp, _vp]),
    "edk_debug_phase": (_i, [_vp, _i, _vp, _vp]),
    "edk_debug_use_naive_gram": (_i, [_vp, _i]),
    "edk_debug_gram_config": (_i, [_vp, _i, _i]),
    "edk_debug_symmetry": (_i, [_vp, _i]),
    "edk_debug_loader": (_i, [_vp, _i]),
    "edk_debug_algo": (_i, [_vp, _i]),
    "edk_query": (_i, [_vp, _i]),
    "edk_microbench_fp64": (_i, [_i, _dp, _dp]),
}

_lib = None


class EdkError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library once.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the sm_100a kernels first "
                "(python -m easydistillation_b200.build, or __graft_entry__.build()). "
                "easydistillation_b200 has no CPU or PyTorch fallback."
            )
        L = C.CDLL(LIB_PATH)
        if hasattr(L, "edk_host_emulator_build"):  # tests/emu: the kernels on host threads, for the CPU test suite only
            raise ImportError(f"{LIB_PATH} is the host-emulator test build of the C ABI; the package only runs on the "
                              "sm_100a library (no CPU path)")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "edk"):
    if rc == EDK_OK:
        return
    msg = lib().edk_last_error().decode(errors="replace")
    if rc == EDK_ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    if rc == EDK_ERR_NOMEM:
        raise MemoryError(f"{what}: {msg}")
    raise EdkError(f"{what} failed ({rc}): {msg}")


def plan(mode: int, order: int, momentum_list, sym_request: int = -1) -> dict:
    """Host-only contraction plan (edk_plan): pairing decision, momentum counts, pair-GEMM counts."""
    import numpy as np

    mom = np.ascontiguousarray(np.asarray(momentum_list, dtype=np.int32).reshape(-1, 3))
    out = (C.c_int * 8)()
    check(lib().edk_plan(mode, order, mom.shape[0], mom.ctypes.data_as(C.POINTER(C.c_int)), sym_request, out), "edk_plan")
    keys = ("hermitian_pairing", "internal_momenta", "half_set_momenta", "pairs_direct", "pairs_paired",
            "pair_momentum_gemms", "operators", "self_pairs")
    d = dict(zip(keys, list(out)))
    d["hermitian_pairing"] = bool(d["hermitian_pairing"])
    return d


def plan_form(latt3, mode: int, order: int, momentum_list) -> dict:
    """Host-only choice of the contraction form (edk_plan_form): what a handle of this shape will run."""
    import numpy as np

    mom = np.ascontiguousarray(np.asarray(momentum_list, dtype=np.int32).reshape(-1, 3))
    out = (C.c_int * 6)()
    check(lib().edk_plan_form(int(latt3[0]), int(latt3[1]), mode, order, mom.shape[0], mom.ctypes.data_as(C.POINTER(C.c_int)), out),
          "edk_plan_form")
    keys = ("form", "separable_available", "separable_qmax", "separable_r2", "pairs_per_stage", "separable_modes")
    d = dict(zip(keys, list(out)))
    d["separable_available"] = bool(d["separable_available"])
    return d


def plan_tiles(Ne: int):
    """CTA tiles of the separable contraction for Ne eigenvectors (edk_plan_tiles): int array [ntiles, 4] of
    (first row, first column, rows, columns)."""
    import numpy as np

    n = lib().edk_plan_tiles(int(Ne), 0, None)
    if n < 0:
        check(n, "edk_plan_tiles")
    out = np.zeros((n, 4), dtype=np.int32)
    lib().edk_plan_tiles(int(Ne), n, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def plan_modes(momentum_list):
    """Host-only mode plan of the plane-wave factorised contraction (edk_plan_modes):
    (modes [nmodes][3] = (qx, qy, kind), per-momentum [nmom][3] = (cos mode, sin mode or -1, sigma))."""
    import numpy as np

    mom = np.ascontiguousarray(np.asarray(momentum_list, dtype=np.int32).reshape(-1, 3))
    nmom = mom.shape[0]
    modes = np.zeros((2 * nmom, 3), dtype=np.int32)
    momode = np.zeros((nmom, 3), dtype=np.int32)
    n = C.c_int(0)
    ip = C.POINTER(C.c_int)
    check(lib().edk_plan_modes(nmom, mom.ctypes.data_as(ip), C.byref(n), modes.ctypes.data_as(ip), momode.ctypes.data_as(ip)),
          "edk_plan_modes")
    return modes[: n.value].copy(), momode


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise EdkError("no CUDA device: the elemental kernels are sm_100a-only and there is no CPU fallback")
    return torch


class PinnedBuffer:
    """Page-locked host memory from edk_host_alloc, exposed as a numpy array."""

    def __init__(self, shape, dtype):
        import numpy as np

        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().edk_host_alloc(C.byref(p), max(self.nbytes, 16)), "edk_host_alloc")
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().edk_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
