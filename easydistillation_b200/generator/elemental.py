"""Derivative elementals on B200 (drop-in for lattice/generator/elemental.py of the reference).

    E[n, p](t)[e, f] = sum over left/right splits S of derivative(n):
        (-1)^|S|  sum_x  (nabla_{left} V_e)(x)^dagger  exp(+i p.x)  (nabla_{right} V_f)(x)

Same constructor, `load(key)`, `calc(t)` and public attributes as the reference class
(elemental.py:17-100,102-105,290-338); the arithmetic runs in hand-written sm_100a kernels
(stencil + DMMA contraction) through libedk_sm100a.so.  There is no numpy/cupy backend switch
and no CPU path: without the library and a GPU the constructor raises.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .. import _capi
from ..insertion.derivative import derivative, num_derivative
from ._base import _TimesliceGenerator


def blending_matrix(Ne: int, dilution: Tuple) -> np.ndarray:
    """Real (Ne, Ne) rescaling of the 'blending' stochastic option (elemental.py:61-95).

    dilution = (totNe_list, usedNe_list | int): block (i, j), i != j, carries
    tot_i tot_j / (used_i used_j); a diagonal block carries c1 = tot/used on its diagonal and
    c1 (tot-1)/(used-1) elsewhere."""
    tot = list(dilution[0])
    used = [dilution[1]] * len(tot) if isinstance(dilution[1], int) else list(dilution[1])
    assert len(used) == len(tot)
    assert all(u <= t for u, t in zip(used, tot))
    assert sum(used) == Ne
    edges = np.concatenate([[0], np.cumsum(used)])
    coeff = np.zeros((Ne, Ne))
    for i in range(len(tot)):
        for j in range(len(tot)):
            block = coeff[edges[i] : edges[i + 1], edges[j] : edges[j + 1]]
            if i == j:
                c1 = tot[i] / used[i]
                block[...] = c1 * (tot[i] - 1) / (used[i] - 1)
                np.fill_diagonal(block, c1)
            else:
                block[...] = tot[i] * tot[j] / used[i] / used[j]
    return coeff


class ElementalGenerator(_TimesliceGenerator):
    _mode = _capi.MODE_DERIVATIVE

    def __init__(
        self,
        latt_size: List[int],
        gauge_field,
        eigenvector,
        num_nabla: int = 0,
        momentum_list: List[Tuple[int]] = [(0, 0, 0)],
        dilution: Tuple = None,
        is_blending: bool = False,
        *,
        device=None,
    ) -> None:
        self.kernel = None  # the reference keeps its cupy stout kernel here; nothing to JIT in this build
        self.num_derivative = num_derivative(num_nabla)
        self.derivative_list = [derivative(n) for n in range(self.num_derivative)]
        if is_blending and dilution is None:
            raise ValueError("Dilution tuple is not defined.")
        self._setup(latt_size, gauge_field, eigenvector, num_nabla, momentum_list, device)
        if is_blending:
            self.stocastic_coeff = blending_matrix(self.Ne, dilution)
            self._engine.set_blending(self.stocastic_coeff)
        else:
            self.stocastic_coeff = None
