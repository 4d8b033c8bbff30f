"""The eigensolver's gauge-covariant Laplacian on B200.

Reference: `_Laplacian` / `EigenvectorGenerator.load` of lattice/generator/eigenvector.py:11-26,49-53,
used there as the `matvec`/`matmat` of a scipy/cupyx `LinearOperator` (:234-257).  Only the operator is
provided here (the Lanczos driver is outside the elemental path); it runs on the same links, with the
same optional stout smearing / SU(3) projection, as the elemental generators."""
from __future__ import annotations

from typing import List

import numpy as np

from .. import _capi
from ..constant import Nc, Nd
from ..engine import ElementalEngine


class Laplacian:
    """`L = 6 - sum_d (U_d T_d + T_d^dagger U_d^dagger)` of one timeslice.

        lap = Laplacian(latt_size, gauge_field); lap.load(key); lap.set_timeslice(t)
        Y = lap.matmat(X)        # X [nvec, Lz, Ly, Lx, 3] complex128, torch CUDA tensor or numpy
    """

    def __init__(self, latt_size: List[int], gauge_field, *, device=None) -> None:
        Lx, Ly, Lz, Lt = (int(v) for v in latt_size)
        self.latt_size = latt_size
        self.gauge_field = gauge_field
        self._engine = ElementalEngine((Lx, Ly, Lz), 1, _capi.MODE_DERIVATIVE, 0, [(0, 0, 0)], device)
        self._U = None
        self._gauge_ops = []
        self.shape = (Lz * Ly * Lx * Nc, Lz * Ly * Lx * Nc)

    def load(self, key: str):
        U = np.asarray(self.gauge_field.load(key)[:])
        Lx, Ly, Lz, Lt = (int(v) for v in self.latt_size)
        if U.shape != (Lt, Lz, Ly, Lx, Nd, Nc, Nc):
            raise ValueError(f"gauge field must be [Lt, Lz, Ly, Lx, {Nd}, {Nc}, {Nc}], got {U.shape}")
        self._U = np.ascontiguousarray(U, dtype="<c16")
        self._gauge_ops = []
        self._engine.set_link_ops([])

    def stout_smear(self, nstep, rho):
        self._gauge_ops.append(("stout", int(nstep), float(rho)))
        self._engine.set_link_ops(self._gauge_ops)

    def project_SU3(self):
        self._gauge_ops.append(("project",))
        self._engine.set_link_ops(self._gauge_ops)

    def set_timeslice(self, t: int):
        if self._U is None:
            raise RuntimeError("call load(key) first")
        if not 0 <= t < int(self.latt_size[3]):
            raise IndexError(f"timeslice {t} out of range")
        torch = self._engine.torch
        self._engine.set_links(torch.from_numpy(self._U[t]).to(self._engine.device), _capi.LINKS_FILE_T)

    def matmat(self, X):
        torch = self._engine.torch
        if isinstance(X, torch.Tensor):
            return self._engine.laplacian(X.contiguous())
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype="<c16")).to(self._engine.device)
        return self._engine.laplacian(Xd).cpu().numpy()

    def matvec(self, x):
        return self.matmat(x[None])[0]
