"""CPU oracle for the elemental-generation hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm and is imported
only by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs.
The product path (`easydistillation_b200`) never imports it and has no CPU
fallback.

Parity pinning: the reference's stored golden files (`tests/weak_field.*`) are
git-LFS pointers in the checkout available to this build, so the oracle is
pinned instead against *outputs of the reference itself run in the build
container* (`oracle/make_golden.py` imports /root/reference through
`oracle/shims` and writes `tests/golden/*.npz`; `tests/test_oracle.py` checks
every function below against those files).

Reference lines each function follows (relative to the reference root):
  derivative_tuple      lattice/insertion/derivative.py:23-33
  momentum_phase        lattice/insertion/phase.py:11-13,41-46
  covariant_hop/nabla   lattice/generator/elemental.py:279-288
  elemental_timeslice   lattice/generator/elemental.py:290-338
  blending_matrix       lattice/generator/elemental.py:61-100
  displacement_*        lattice/generator/displacement_elemental.py:53-96
  stout_smear_timeslice lattice/generator/elemental.py:175-241 (project_su3_timeslice :107-117)
  laplacian             lattice/generator/eigenvector.py:11-26

Array conventions (same as the reference): links `U[d, z, y, x, a, b]` for the
three spatial directions d = 0(x), 1(y), 2(z) of ONE timeslice, eigenvectors
`V[e, z, y, x, c]`, complex128 arithmetic, eigenvectors value-rounded through
complex64 before use (elemental.py:55,298).
"""
from __future__ import annotations

import numpy as np

Nc = 3


# --------------------------------------------------------------------------
# index maps and phases
# --------------------------------------------------------------------------
def num_derivative(num_nabla: int) -> int:
    """(3^(n+1) - 1) / 2 derivative operators up to order n (elemental.py:48)."""
    return (3 ** (num_nabla + 1) - 1) // 2


def derivative_tuple(n: int) -> tuple:
    """Index n -> tuple of directions, first element applied first.

    0 -> (), 1..3 -> (0,),(1,),(2,), 4..12 -> (0,0),(0,1),...,(2,2): orders are
    laid out consecutively and, inside one order, n is read as a base-3 number
    whose most significant digit comes first in the tuple."""
    order = 0
    while n >= 3**order:
        n -= 3**order
        order += 1
    digits = []
    for _ in range(order):
        digits.append(n % 3)
        n //= 3
    return tuple(reversed(digits))


def momentum_phase(latt_size, mom) -> np.ndarray:
    """exp(+2 pi i (px x/Lx + py y/Ly + pz z/Lz)) with shape (Lz, Ly, Lx)."""
    Lx, Ly, Lz = latt_size[:3]
    px, py, pz = mom
    gx = np.arange(Lx).reshape(1, 1, Lx) * 2j * np.pi / Lx
    gy = np.arange(Ly).reshape(1, Ly, 1) * 2j * np.pi / Ly
    gz = np.arange(Lz).reshape(Lz, 1, 1) * 2j * np.pi / Lz
    return np.exp(px * gx + py * gy + pz * gz + np.zeros((Lz, Ly, Lx)))


def round_through_c8(V: np.ndarray) -> np.ndarray:
    """The reference stages eigenvectors in a complex64 buffer."""
    return V.astype(np.complex64)


# --------------------------------------------------------------------------
# stencil
# --------------------------------------------------------------------------
def covariant_hop(V: np.ndarray, U: np.ndarray, d: int) -> np.ndarray:
    """(nabla_d V)(x) = U_d(x) V(x+d) - U_d(x-d)^dagger V(x-d); periodic; no 1/2.

    Written with the same four full-field temporaries as the reference (two
    rolls, two colour mat-vecs) so that its CPU cost is representative."""
    axis = 3 - d
    fwd = np.roll(V, -1, axis)
    u_fwd = np.einsum("zyxab,ezyxb->ezyxa", U[d], fwd, optimize=True)
    udag_v = np.einsum("zyxba,ezyxb->ezyxa", U[d].conj(), V, optimize=True)
    bwd = np.roll(udag_v, 1, axis)
    return u_fwd - bwd


def nabla_chain(V: np.ndarray, U: np.ndarray, directions) -> np.ndarray:
    """Apply covariant_hop for each direction in order (first element first)."""
    for d in directions:
        V = covariant_hop(V, U, d)
    return V


def gram(left: np.ndarray, right: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """G[e,f] = sum_{z,y,x,c} conj(left[e]) * phase * right[f]."""
    return np.einsum("zyx,ezyxc,fzyxc->ef", phase, left.conj(), right, optimize=True)


# --------------------------------------------------------------------------
# derivative elementals
# --------------------------------------------------------------------------
def blending_matrix(Ne: int, dilution) -> np.ndarray:
    """Real (Ne, Ne) rescaling used by the 'blending' stochastic option."""
    tot_list = list(dilution[0])
    used_list = [dilution[1]] * len(tot_list) if isinstance(dilution[1], int) else list(dilution[1])
    assert len(used_list) == len(tot_list)
    assert all(u <= t for u, t in zip(used_list, tot_list))
    assert sum(used_list) == Ne
    coeff = np.zeros((Ne, Ne))
    starts = np.concatenate([[0], np.cumsum(used_list)])
    for i, (ui, ti) in enumerate(zip(used_list, tot_list)):
        for j, (uj, tj) in enumerate(zip(used_list, tot_list)):
            blk = coeff[starts[i] : starts[i + 1], starts[j] : starts[j + 1]]
            if i != j:
                blk[...] = ti * tj / ui / uj
            else:
                c1 = ti / ui
                blk[...] = c1 * (ti - 1) / (ui - 1)
                np.fill_diagonal(blk, c1)
    return coeff


def elemental_timeslice(V_t, U_t, latt_size, num_nabla, momentum_list, stocastic_coeff=None):
    """All derivative elementals of one timeslice, the way the reference forms them.

    For every derivative tuple and every left/right split S of its positions:
    right = nabla over the picked directions in order, left = nabla over the
    remaining ones in reversed order, weight (-1)^|S| (elemental.py:309-329).
    Returns (num_derivative, Nmom, Ne, Ne) complex128.

    The rounded eigenvectors are held in complex128: the reference keeps them in its complex64 buffer and relies on
    the einsum path (phase x left first, then a complex128 tensordot) for FP64 accumulation, which is the path at
    every size its tests and the goldens use; for toy shapes (Ne = 2, V = 8) numpy's path optimiser may instead
    multiply the two complex64 operands first, which would make the zero-derivative block a float32 sum."""
    V = round_through_c8(np.asarray(V_t)).astype(np.complex128)
    Ne = V.shape[0]
    nder = num_derivative(num_nabla)
    out = np.zeros((nder, len(momentum_list), Ne, Ne), np.complex128)
    phases = [momentum_phase(latt_size, p) for p in momentum_list]
    for n in range(nder):
        dirs = derivative_tuple(n)
        for pick in range(2 ** len(dirs)):
            right_dirs = [d for i, d in enumerate(dirs) if (pick >> i) & 1]
            left_dirs = [d for i, d in enumerate(dirs) if not (pick >> i) & 1]
            sign = (-1) ** len(right_dirs)
            right = nabla_chain(V, U_t, right_dirs)
            left = nabla_chain(V, U_t, left_dirs[::-1])
            for ip, ph in enumerate(phases):
                out[n, ip] += gram(left, right, sign * ph)
    if stocastic_coeff is not None:
        out *= stocastic_coeff[None, None]
    return out


def elemental_timeslice_closed_form(V_t, U_t, latt_size, num_nabla, momentum_list, stocastic_coeff=None):
    """Same result through the 13 distinct fields / 34 distinct pairs.

    W1[a] = nabla_a W0, W2[a][b] = nabla_a nabla_b W0.  This is the factorisation
    the CUDA path uses; kept here so big cases can be checked without the 78-hop
    loop.  num_nabla <= 2."""
    assert num_nabla <= 2
    W0 = round_through_c8(np.asarray(V_t)).astype(np.complex128)
    Ne = W0.shape[0]
    nder = num_derivative(num_nabla)
    out = np.zeros((nder, len(momentum_list), Ne, Ne), np.complex128)
    W1 = [covariant_hop(W0, U_t, a) for a in range(3)] if num_nabla >= 1 else []
    W2 = [[covariant_hop(W1[b], U_t, a) for b in range(3)] for a in range(3)] if num_nabla >= 2 else []
    for ip, p in enumerate(momentum_list):
        ph = momentum_phase(latt_size, p)
        out[0, ip] = gram(W0, W0, ph)
        if num_nabla >= 1:
            for a in range(3):
                out[1 + a, ip] = gram(W1[a], W0, ph) - gram(W0, W1[a], ph)
        if num_nabla >= 2:
            for d1 in range(3):
                for d2 in range(3):
                    out[4 + 3 * d1 + d2, ip] = (
                        gram(W2[d1][d2], W0, ph)
                        - gram(W1[d2], W1[d1], ph)
                        - gram(W1[d1], W1[d2], ph)
                        + gram(W0, W2[d2][d1], ph)
                    )
    if stocastic_coeff is not None:
        out *= stocastic_coeff[None, None]
    return out


# --------------------------------------------------------------------------
# displacement elementals
# --------------------------------------------------------------------------
def displacement_fields(V_t, U_t, distance):
    """Yield D_k for k = 0..distance: mean of the six straight Wilson lines of
    length k ending at x (displacement_elemental.py:53-71).  Rounded eigenvectors are held in complex128, see
    elemental_timeslice."""
    W0 = round_through_c8(np.asarray(V_t)).astype(np.complex128)
    yield W0
    lines = None
    for k in range(1, distance + 1):
        new = np.zeros((6,) + W0.shape, np.complex128)
        for d in range(3):
            axis = 3 - d
            src_f = W0 if lines is None else lines[d]
            src_b = W0 if lines is None else lines[5 - d]
            new[d] = np.einsum("zyxab,ezyxb->ezyxa", U_t[d], np.roll(src_f, -1, axis), optimize=True)
            udag = np.einsum("zyxba,ezyxb->ezyxa", U_t[d].conj(), src_b, optimize=True)
            new[5 - d] = np.roll(udag, 1, axis)
        lines = new
        yield lines.mean(0)


def displacement_timeslice(V_t, U_t, latt_size, distance, momentum_list):
    """E[k, p] = G(W0, D_k, p), shape (distance+1, Nmom, Ne, Ne)."""
    W0 = round_through_c8(np.asarray(V_t)).astype(np.complex128)
    Ne = W0.shape[0]
    out = np.zeros((distance + 1, len(momentum_list), Ne, Ne), np.complex128)
    phases = [momentum_phase(latt_size, p) for p in momentum_list]
    for k, Dk in enumerate(displacement_fields(V_t, U_t, distance)):
        for ip, ph in enumerate(phases):
            out[k, ip] += gram(W0, Dk, ph)
    return out


# --------------------------------------------------------------------------
# gauge preprocessing of the generator classes
# --------------------------------------------------------------------------
def _shift(A, d, step):
    """A(x + step*d) for a one-timeslice array [Lz, Ly, Lx, ...]; d = 0(x), 1(y), 2(z)."""
    return np.roll(A, -step, 2 - d)


def _dag(A):
    return np.conj(np.swapaxes(A, -1, -2))


def exp_i_traceless_hermitian(Q):
    """exp(iQ) for traceless Hermitian 3x3 Q by Cayley-Hamilton (Morningstar-Peardon), with the
    same branch handling as lattice/generator/elemental.py:200-238 (c0 -> |c0| plus conjugation)."""
    Q2 = Q @ Q
    c0 = np.trace(Q @ Q2, axis1=-2, axis2=-1).real / 3
    c1 = np.trace(Q2, axis1=-2, axis2=-1).real / 2
    c0_max = 2 * (c1 / 3) ** 1.5
    neg = c0 < 0
    theta = np.arccos(np.abs(c0) / c0_max)
    u = np.sqrt(c1 / 3) * np.cos(theta / 3)
    w = np.sqrt(c1) * np.sin(theta / 3)
    u2, w2 = u * u, w * w
    xi0 = 1 - w2 / 6 * (1 - w2 / 20 * (1 - w2 / 42 * (1 - w2 / 72)))
    big = np.abs(w) > 0.05
    xi0[big] = np.sin(w[big]) / w[big]
    e2, em = np.exp(2j * u), np.exp(-1j * u)
    den = 1 / (9 * u2 - w2)
    f0 = ((u2 - w2) * e2 + em * (8 * u2 * np.cos(w) + 2j * u * (3 * u2 + w2) * xi0)) * den
    f1 = (2 * u * e2 - em * (2 * u * np.cos(w) - 1j * (3 * u2 - w2) * xi0)) * den
    f2 = (e2 - em * (np.cos(w) + 3j * u * xi0)) * den
    f0 = np.where(neg, np.conj(f0), f0)
    f1 = np.where(neg, -np.conj(f1), f1)
    f2 = np.where(neg, np.conj(f2), f2)
    return f0[..., None, None] * np.eye(3) + f1[..., None, None] * Q + f2[..., None, None] * Q2


def stout_smear_timeslice(U_t, nstep, rho):
    """Spatial stout smearing of one timeslice U_t[d, z, y, x, a, b] (elemental.py:175-241; the
    three spatial directions never couple different timeslices)."""
    U = np.array(U_t, dtype=np.complex128)
    for _ in range(nstep):
        new = np.empty_like(U)
        for mu in range(3):
            C = np.zeros_like(U[mu])
            for nu in range(3):
                if nu == mu:
                    continue
                C += U[nu] @ _shift(U[mu], nu, 1) @ _dag(_shift(U[nu], mu, 1))
                C += _dag(_shift(U[nu], nu, -1)) @ _shift(U[mu], nu, -1) @ _shift(_shift(U[nu], nu, -1), mu, 1)
            Om = rho * C @ _dag(U[mu])
            Q = 0.5j * (_dag(Om) - Om)
            Q = Q - np.trace(Q, axis1=-2, axis2=-1)[..., None, None] * np.eye(3) / 3
            new[mu] = exp_i_traceless_hermitian(Q) @ U[mu]
        U = new
    return U


def project_su3_timeslice(U_t, max_iter=100):
    """X <- (X + X^-dagger)/2 until X is unitary to 1e-15 (elemental.py:107-117)."""
    U = np.array(U_t, dtype=np.complex128)
    for _ in range(max_iter):
        Uinv = np.linalg.inv(U)
        if np.max(np.abs(U - _dag(Uinv))) <= 1e-15 and np.max(np.abs(U @ _dag(U) - np.eye(3))) <= 1e-15:
            break
        U = 0.5 * (U + _dag(Uinv))
    return U


def laplacian(V, U_t):
    """Gauge-covariant 3-d Laplacian of the eigensolver (lattice/generator/eigenvector.py:11-26):
    (L V)(x) = 6 V(x) - sum_d [ U_d(x) V(x+d) + U_d(x-d)^dagger V(x-d) ], V[e, z, y, x, c]."""
    out = 6.0 * V.astype(np.complex128)
    for d in range(3):
        axis = 3 - d
        out -= np.einsum("zyxab,ezyxb->ezyxa", U_t[d], np.roll(V, -1, axis), optimize=True)
        out -= np.roll(np.einsum("zyxba,ezyxb->ezyxa", U_t[d].conj(), V, optimize=True), 1, axis)
    return out


# --------------------------------------------------------------------------
# synthetic inputs: seeded, shared by tests and bench
# --------------------------------------------------------------------------
SEED0 = 20261017


def random_su3(rng, shape) -> np.ndarray:
    """Haar-ish random SU(3): QR of complex Gaussian, phase-fixed, det -> 1."""
    a = rng.standard_normal(shape + (3, 3)) + 1j * rng.standard_normal(shape + (3, 3))
    q, r = np.linalg.qr(a)
    dg = np.diagonal(r, axis1=-2, axis2=-1)
    q = q * (dg / np.abs(dg))[..., None, :]
    det = np.linalg.det(q)
    return q / (det ** (1.0 / 3.0))[..., None, None]


def weak_su3(rng, shape, eps=0.1) -> np.ndarray:
    """exp(i eps H), H random traceless Hermitian (weak-field variant)."""
    a = rng.standard_normal(shape + (3, 3)) + 1j * rng.standard_normal(shape + (3, 3))
    h = 0.5 * (a + np.conj(np.swapaxes(a, -1, -2)))
    h = h - np.trace(h, axis1=-2, axis2=-1)[..., None, None] * np.eye(3) / 3.0
    w, v = np.linalg.eigh(h)
    return np.einsum("...ab,...b,...cb->...ac", v, np.exp(1j * eps * w), v.conj())


def synthetic_links(latt_size, t, kind="random") -> np.ndarray:
    """Links of one timeslice in the reference's file order [Lz, Ly, Lx, Nd, Nc, Nc]."""
    Lx, Ly, Lz = latt_size[:3]
    rng = np.random.default_rng(SEED0 + t)
    gen = random_su3 if kind == "random" else weak_su3
    return np.ascontiguousarray(gen(rng, (Lz, Ly, Lx, 4)))


def synthetic_eigvecs(latt_size, Ne, t) -> np.ndarray:
    """[Ne, Lz, Ly, Lx, Nc] complex Gaussian, unit-normalised per vector."""
    Lx, Ly, Lz = latt_size[:3]
    rng = np.random.default_rng(SEED0 + 7919 * (t + 1))
    v = rng.standard_normal((Ne, Lz, Ly, Lx, Nc)) + 1j * rng.standard_normal((Ne, Lz, Ly, Lx, Nc))
    v /= np.sqrt((np.abs(v) ** 2).sum(axis=(1, 2, 3, 4), keepdims=True))
    return v


def momentum_set(count: int):
    """First `count` integer triples ordered by (|p|^2, p): 33 = all |p|^2 <= 4."""
    r = range(-3, 4)
    allp = sorted(((px * px + py * py + pz * pz, (px, py, pz)) for px in r for py in r for pz in r))
    return [p for _, p in allp[:count]]


def links_file_to_spatial(U_file_t: np.ndarray) -> np.ndarray:
    """[Lz,Ly,Lx,Nd,3,3] -> the reference's U[:, t] view [3, Lz, Ly, Lx, 3, 3]
    (elemental.py:103: transpose(4,0,1,2,3,5,6)[:Nd-1], then [:, t])."""
    return np.moveaxis(U_file_t, 3, 0)[:3]
