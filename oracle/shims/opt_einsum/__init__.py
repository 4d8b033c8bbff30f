"""TEST INFRASTRUCTURE ONLY. Stand-in for the (absent) `opt_einsum` package so the
read-only reference at /root/reference can be imported in this container.

The reference (lattice/generator/elemental.py:4) only uses `contract`, which
chooses a pairwise order and dispatches to numpy.tensordot/einsum; numpy's own
`einsum(optimize=True)` does the same with the same BLAS underneath."""
import numpy


def contract(subscripts, *operands, **kwargs):
    return numpy.einsum(subscripts, *operands, optimize=True)
