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
        if U.ndim == 5 and U.shape == (Lt, Lz * Ly * Lx, Nd, Nc, Nc):  # the flattened default shape of the presets
            U = U.reshape(Lt, Lz, Ly, Lx, Nd, Nc, Nc)
        if U.shape != (Lt, Lz, Ly, Lx, Nd, Nc, Nc):
            raise ValueError(f"gauge field must be [Lt, Lz, Ly, Lx, {Nd}, {Nc}, {Nc}], got {U.shape}")
        # kept as loaded (a memory map stays a memory map): a timeslice is converted when it is selected
        self._U = U
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
        U_t, be = _capi.raw_view(np.ascontiguousarray(self._U[t]))
        if U_t.dtype != np.complex128:
            U_t, be = np.ascontiguousarray(U_t, dtype="<c16"), False
        self._engine.set_links(torch.from_numpy(U_t).to(self._engine.device),
                               _capi.LINKS_FILE_T | (_capi.LINKS_BIG_ENDIAN if be else 0))

    def _apply(self, X):
        """X [nvec, Lz, Ly, Lx, 3] (torch CUDA tensor or numpy) -> L X in the same container."""
        torch = self._engine.torch
        if isinstance(X, torch.Tensor):
            return self._engine.laplacian(X.to(self._engine.device, torch.complex128).contiguous())
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype="<c16")).to(self._engine.device)
        return self._engine.laplacian(Xd).cpu().numpy()

    def matmat(self, X):
        """Two conventions, told apart by the rank of X:
          * [nvec, Lz, Ly, Lx, 3] (vector index slowest, this package's field layout) -> the same shape;
          * (N, k) with N = Lz*Ly*Lx*3, vector index fastest - what scipy / cupyx `eigsh` hand to a LinearOperator and
            what the reference's `_Laplacian` reshapes with `F.reshape(Lz, Ly, Lx, Nc, -1)`
            (lattice/generator/eigenvector.py:11-26) -> (N, k)."""
        Lx, Ly, Lz, Lt = (int(v) for v in self.latt_size)
        if X.ndim == 2:
            if X.shape[0] != self.shape[0]:
                raise ValueError(f"flat vectors must have {self.shape[0]} rows, got {tuple(X.shape)}")
            k = X.shape[1]
            torch = self._engine.torch
            if isinstance(X, torch.Tensor):
                Y = self._apply(X.reshape(Lz, Ly, Lx, Nc, k).permute(4, 0, 1, 2, 3))
                return Y.permute(1, 2, 3, 4, 0).reshape(self.shape[0], k)
            Y = self._apply(np.moveaxis(np.asarray(X).reshape(Lz, Ly, Lx, Nc, k), -1, 0))
            return np.ascontiguousarray(np.moveaxis(Y, 0, -1)).reshape(self.shape[0], k)
        return self._apply(X)

    def matvec(self, x):
        """x (N,) flat, or [Lz, Ly, Lx, 3]."""
        if x.ndim == 1:
            return self.matmat(x.reshape(-1, 1)).reshape(-1)
        return self._apply(x[None])[0]

    def as_linear_operator(self):
        """scipy.sparse.linalg.LinearOperator over the selected timeslice (numpy in / out), as the reference builds for
        `eigsh` (lattice/generator/eigenvector.py:234-257)."""
        from scipy.sparse.linalg import LinearOperator

        return LinearOperator(self.shape, matvec=self.matvec, matmat=self.matmat, dtype=np.complex128)
