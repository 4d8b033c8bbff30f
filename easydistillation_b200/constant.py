"""Colour, spin and dimension counts (reference: lattice/constant.py:1-3)."""
Nc = 3
Ns = 4
Nd = 4
