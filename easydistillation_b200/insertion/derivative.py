"""Derivative index map of the elemental path (reference: lattice/insertion/derivative.py:23-33).

`derivative(n)` turns the operator index stored along axis 0 of an elemental file into the
tuple of spatial directions (0=x, 1=y, 2=z), listed in the order the covariant differences are
applied: 0 -> (), 1..3 -> (0,),(1,),(2,), 4..12 -> (0,0),(0,1),(0,2),(1,0),...,(2,2), and so on
for higher orders (3^k operators of order k, numbered consecutively, most significant base-3
digit first)."""
from typing import Tuple


def num_derivative(num_nabla: int) -> int:
    """Operators of order <= num_nabla: (3^(n+1)-1)/2 (lattice/generator/elemental.py:48)."""
    if num_nabla < 0:
        raise ValueError("num_nabla must be >= 0")
    return (3 ** (num_nabla + 1) - 1) // 2


def derivative(n: int) -> Tuple[int, ...]:
    if not isinstance(n, int) or n < 0:
        raise ValueError("derivative index must be a non-negative int")
    order, block = 0, 1
    while n >= block:
        n -= block
        block *= 3
        order += 1
    seq = []
    for _ in range(order):
        block //= 3
        seq.append(n // block)
        n %= block
    return tuple(seq)
