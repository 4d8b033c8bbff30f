"""Displacement elementals on B200 (drop-in for lattice/generator/displacement_elemental.py).

    E[k, p](t)[e, f] = sum_x V_e(x)^dagger exp(+i p.x) (D_k V_f)(x),   k = 0..distance

with D_k the average of the six straight Wilson lines of length k (displacement_elemental.py:
53-71).  Constructor, `load`, `calc` and attributes follow the reference (:12-51,73-96)."""
from __future__ import annotations

from typing import List, Tuple

from .. import _capi
from ._base import _TimesliceGenerator


class DisplacementElementalGenerator(_TimesliceGenerator):
    _mode = _capi.MODE_DISPLACEMENT

    def __init__(
        self,
        latt_size: List[int],
        gauge_field,
        eigenvector,
        distance: int = 0,
        momentum_list: List[Tuple[int]] = [(0, 0, 0)],
        *,
        device=None,
    ) -> None:
        if distance < 0:
            raise ValueError("distance must be >= 0")
        self.kernel = None
        self.distance = distance
        self._setup(latt_size, gauge_field, eigenvector, distance, momentum_list, device)
