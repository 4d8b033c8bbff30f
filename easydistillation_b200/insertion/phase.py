"""Momentum phases on the device (reference: lattice/insertion/phase.py:6-46, `get` only).

`MomentumPhase(latt_size).get(p)` returns exp(+2 pi i (px x/Lx + py y/Ly + pz z/Lz)) as a
complex128 CUDA tensor of shape (Lz, Ly, Lx), cached per momentum like the reference.  The
table is produced by the `phase_table` kernel (p.x is reduced modulo L in integers, then
sincospi), not by numpy/cupy."""
import ctypes as C
from typing import List, Tuple

import numpy as np

from .. import _capi


class MomentumPhase:
    def __init__(self, latt_size: List[int], device=None) -> None:
        self.latt_size = latt_size
        self.device = device
        self.cache = {}

    def get(self, np_: Tuple[int]):
        key = tuple(int(v) for v in np_)
        if key not in self.cache:
            torch = _capi.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device() if self.device is None else self.device)
            Lx, Ly, Lz = (int(v) for v in self.latt_size[:3])
            out = torch.empty((Lz, Ly, Lx), dtype=torch.complex128, device=dev)
            mom = np.asarray(key, dtype=np.int32)
            rc = _capi.lib().edk_phase_table(Lx, Ly, Lz, 1, mom.ctypes.data_as(C.POINTER(C.c_int)),
                                             C.c_void_p(out.data_ptr()), dev.index,
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _capi.check(rc, "edk_phase_table")
            self.cache[key] = out
        return self.cache[key]
