// Gauge preprocessing of the generator classes on the device, sm_100a.
//
//   stout_step  : one step of spatial stout smearing, U_mu <- exp(iQ_mu) U_mu with
//                 Q = traceless Hermitian part of rho * (staples of the two other spatial
//                 directions) * U_mu^dagger, exp(iQ) by Cayley-Hamilton.  Replaces the
//                 reference's only CUDA kernel, lattice/generator/stout_smear.cu:159-245
//                 (numpy form: lattice/generator/elemental.py:175-241).
//   project_su3 : X <- (X + X^-dagger)/2 until unitary to 1e-15, per link
//                 (lattice/generator/elemental.py:107-117).
//
// Both act on the handle's spatial links of ONE timeslice, [3][Lz][Ly][Lx][3][3] complex128:
// spatial smearing never couples timeslices, so it runs right after the timeslice's links are
// uploaded instead of on the whole configuration at load() time.  One thread per (site, mu);
// 13 link loads and ~13 3x3 products per thread, microseconds per step at every lattice size.
#include "edk_common.cuh"

namespace edk {

struct M3 {
    cplx m[9];
};

__device__ __forceinline__ cplx cmul(const cplx a, const cplx b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx cmul_conj_b(const cplx a, const cplx b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}
__device__ __forceinline__ cplx cadd(const cplx a, const cplx b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ M3 load_m3(const cplx* p) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = __ldg(p + i);
    return r;
}
__device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cplx s = cmul(a.m[3 * i], b.m[j]);
            s = cadd(s, cmul(a.m[3 * i + 1], b.m[3 + j]));
            s = cadd(s, cmul(a.m[3 * i + 2], b.m[6 + j]));
            r.m[3 * i + j] = s;
        }
    return r;
}
__device__ __forceinline__ M3 mul_bdag(const M3& a, const M3& b) {  // a * b^dagger
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cplx s = cmul_conj_b(a.m[3 * i], b.m[3 * j]);
            s = cadd(s, cmul_conj_b(a.m[3 * i + 1], b.m[3 * j + 1]));
            s = cadd(s, cmul_conj_b(a.m[3 * i + 2], b.m[3 * j + 2]));
            r.m[3 * i + j] = s;
        }
    return r;
}
__device__ __forceinline__ M3 adag_mul(const M3& a, const M3& b) {  // a^dagger * b
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cplx s = cmul_conj_b(b.m[j], a.m[i]);
            s = cadd(s, cmul_conj_b(b.m[3 + j], a.m[3 + i]));
            s = cadd(s, cmul_conj_b(b.m[6 + j], a.m[6 + i]));
            r.m[3 * i + j] = s;
        }
    return r;
}

__device__ __forceinline__ int site_of(int x, int y, int z, const Geom& g) { return (z * g.Ly + y) * g.Lx + x; }
__device__ __forceinline__ int shifted(int x, int y, int z, int d, int step, const Geom& g) {
    if (d == 0) x = (x + step + g.Lx) % g.Lx;
    if (d == 1) y = (y + step + g.Ly) % g.Ly;
    if (d == 2) z = (z + step + g.Lz) % g.Lz;
    return site_of(x, y, z, g);
}

__global__ void __launch_bounds__(128) stout_step_kernel(const cplx* __restrict__ Uin, cplx* __restrict__ Uout, double rho,
                                                         Geom g) {
    const int site = blockIdx.x * blockDim.x + threadIdx.x;
    const int mu = blockIdx.y;
    if (site >= g.V) return;
    const int x = site % g.Lx, y = (site / g.Lx) % g.Ly, z = site / (g.Lx * g.Ly);
    auto link = [&](int d, int s) { return load_m3(Uin + ((size_t)d * g.V + s) * 9); };

    const M3 Umu = link(mu, site);
    M3 C;
#pragma unroll
    for (int i = 0; i < 9; ++i) C.m[i] = make_double2(0.0, 0.0);
    const int x_mu = shifted(x, y, z, mu, +1, g);
    for (int nu = 0; nu < 3; ++nu) {
        if (nu == mu) continue;
        const int x_nu = shifted(x, y, z, nu, +1, g);
        const int x_mnu = shifted(x, y, z, nu, -1, g);
        int xm = x, ym = y, zm = z;  // x - nu + mu
        if (nu == 0) xm = (x - 1 + g.Lx) % g.Lx;
        if (nu == 1) ym = (y - 1 + g.Ly) % g.Ly;
        if (nu == 2) zm = (z - 1 + g.Lz) % g.Lz;
        const int x_mnu_mu = shifted(xm, ym, zm, mu, +1, g);
        // upper staple  U_nu(x) U_mu(x+nu) U_nu(x+mu)^dagger
        M3 t = mul(link(nu, site), link(mu, x_nu));
        t = mul_bdag(t, link(nu, x_mu));
#pragma unroll
        for (int i = 0; i < 9; ++i) C.m[i] = cadd(C.m[i], t.m[i]);
        // lower staple  U_nu(x-nu)^dagger U_mu(x-nu) U_nu(x-nu+mu)
        t = adag_mul(link(nu, x_mnu), link(mu, x_mnu));
        t = mul(t, link(nu, x_mnu_mu));
#pragma unroll
        for (int i = 0; i < 9; ++i) C.m[i] = cadd(C.m[i], t.m[i]);
    }
    // Omega = rho C U^dagger ; Q = (i/2)(Omega^dagger - Omega) - trace/3
    M3 Om = mul_bdag(C, Umu);
    M3 Q;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const cplx a = Om.m[3 * j + i];  // (Omega^dagger)_ij = conj(Omega_ji)
            const cplx b = Om.m[3 * i + j];
            const double dr = rho * (a.x - b.x), di = rho * (-a.y - b.y);  // (Omega^dagger - Omega)_ij
            Q.m[3 * i + j] = make_double2(-0.5 * di, 0.5 * dr);           // times i/2
        }
    const double tr_r = (Q.m[0].x + Q.m[4].x + Q.m[8].x) / 3.0, tr_i = (Q.m[0].y + Q.m[4].y + Q.m[8].y) / 3.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Q.m[4 * i].x -= tr_r;
        Q.m[4 * i].y -= tr_i;
    }
    const M3 Q2 = mul(Q, Q);
    // c0 = Re tr(Q^3)/3, c1 = Re tr(Q^2)/2
    double c0 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c0 += Q.m[3 * i + j].x * Q2.m[3 * j + i].x - Q.m[3 * i + j].y * Q2.m[3 * j + i].y;
    c0 /= 3.0;
    const double c1 = (Q2.m[0].x + Q2.m[4].x + Q2.m[8].x) / 2.0;
    cplx f0 = make_double2(1.0, 0.0), f1 = make_double2(0.0, 0.0), f2 = make_double2(0.0, 0.0);
    if (c1 > 0.0) {  // Q = 0 (unit plaquettes) leaves the link unchanged; the reference divides 0/0 there
        const double c0max = 2.0 * sqrt(c1 / 3.0) * (c1 / 3.0);
        const bool neg = c0 < 0.0;
        const double theta = acos(fmin(fabs(c0) / c0max, 1.0));
        double st, ct;
        sincos(theta / 3.0, &st, &ct);
        const double u = sqrt(c1 / 3.0) * ct, w = sqrt(c1) * st;
        const double u2 = u * u, w2 = w * w;
        double xi0;
        if (fabs(w) > 0.05)
            xi0 = sin(w) / w;
        else
            xi0 = 1.0 - w2 / 6.0 * (1.0 - w2 / 20.0 * (1.0 - w2 / 42.0 * (1.0 - w2 / 72.0)));
        double s1, c1u, s2, c2u;
        sincos(u, &s1, &c1u);
        sincos(2.0 * u, &s2, &c2u);
        const double cw = cos(w);
        const double den = 1.0 / (9.0 * u2 - w2);
        // h = a * e^{2iu} + e^{-iu} * (br + i bi)
        auto comb = [&](double a, double br, double bi) {
            return make_double2((a * c2u + c1u * br + s1 * bi) * den, (a * s2 - s1 * br + c1u * bi) * den);
        };
        f0 = comb(u2 - w2, 8.0 * u2 * cw, 2.0 * u * (3.0 * u2 + w2) * xi0);
        f1 = comb(2.0 * u, -2.0 * u * cw, (3.0 * u2 - w2) * xi0);
        f2 = comb(1.0, -cw, -3.0 * u * xi0);
        if (neg) {
            f0.y = -f0.y;
            f1.x = -f1.x;
            f2.y = -f2.y;
        }
    }
    M3 E;
#pragma unroll
    for (int i = 0; i < 9; ++i) E.m[i] = cadd(cmul(f1, Q.m[i]), cmul(f2, Q2.m[i]));
#pragma unroll
    for (int i = 0; i < 3; ++i) E.m[4 * i] = cadd(E.m[4 * i], f0);
    const M3 R = mul(E, Umu);
    cplx* po = Uout + ((size_t)mu * g.V + site) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) po[i] = R.m[i];
}

cudaError_t launch_stout_step(const cplx* Uin, cplx* Uout, double rho, Geom g, cudaStream_t s) {
    dim3 block(128), grid((g.V + 127) / 128, 3);
    EDK_LAUNCH(stout_step_kernel, grid, block, 0, s, Uin, Uout, rho, g);
    return cudaGetLastError();
}

__device__ __forceinline__ M3 inverse(const M3& a) {
    M3 c;  // cofactors, transposed (adjugate)
    auto minor2 = [&](int r0, int c0, int r1, int c1) {
        const cplx p = cmul(a.m[3 * r0 + c0], a.m[3 * r1 + c1]);
        const cplx q = cmul(a.m[3 * r0 + c1], a.m[3 * r1 + c0]);
        return make_double2(p.x - q.x, p.y - q.y);
    };
    c.m[0] = minor2(1, 1, 2, 2);
    c.m[1] = minor2(0, 2, 2, 1);
    c.m[2] = minor2(0, 1, 1, 2);
    c.m[3] = minor2(1, 2, 2, 0);
    c.m[4] = minor2(0, 0, 2, 2);
    c.m[5] = minor2(0, 2, 1, 0);
    c.m[6] = minor2(1, 0, 2, 1);
    c.m[7] = minor2(0, 1, 2, 0);
    c.m[8] = minor2(0, 0, 1, 1);
    cplx det = cmul(a.m[0], c.m[0]);
    det = cadd(det, cmul(a.m[1], c.m[3]));
    det = cadd(det, cmul(a.m[2], c.m[6]));
    const double n = det.x * det.x + det.y * det.y;
    const cplx inv = make_double2(det.x / n, -det.y / n);
#pragma unroll
    for (int i = 0; i < 9; ++i) c.m[i] = cmul(c.m[i], inv);
    return c;
}

__global__ void __launch_bounds__(128) project_su3_kernel(cplx* __restrict__ U, size_t nlinks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlinks) return;
    M3 X = load_m3(U + i * 9);
    for (int it = 0; it < 100; ++it) {
        const M3 Xi = inverse(X);
        double dev = 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const cplx d = make_double2(X.m[3 * r + c].x - Xi.m[3 * c + r].x, X.m[3 * r + c].y + Xi.m[3 * c + r].y);
                dev = fmax(dev, hypot(d.x, d.y));
            }
        const M3 XX = mul_bdag(X, X);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) dev = fmax(dev, hypot(XX.m[3 * r + c].x - (r == c ? 1.0 : 0.0), XX.m[3 * r + c].y));
        if (dev <= 1e-15) break;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                X.m[3 * r + c] = make_double2(0.5 * (X.m[3 * r + c].x + Xi.m[3 * c + r].x), 0.5 * (X.m[3 * r + c].y - Xi.m[3 * c + r].y));
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) U[i * 9 + k] = X.m[k];
}

cudaError_t launch_project_su3(cplx* U, Geom g, cudaStream_t s) {
    const size_t n = (size_t)3 * g.V;
    EDK_LAUNCH(project_su3_kernel, (unsigned)((n + 127) / 128), 128, 0, s, U, n);
    return cudaGetLastError();
}

}  // namespace edk
