// Prepare + stencil kernels of the elemental path (sm_100a).
//
//   round_eigvecs   : the complex64 staging of the reference (elemental.py:55,297-298)
//   reorder_links   : file-order timeslice -> direction-major spatial links (elemental.py:103)
//   phase_table     : exp(+2 pi i p.x/L)                        (insertion/phase.py:11-13,41-46)
//   nabla3          : all three covariant central differences of one field in one pass
//                     (nabla_d W)(x) = U_d(x) W(x+d) - U_d(x-d)^dagger W(x-d)   (elemental.py:279-288)
//   displace_step6  : extend the six straight Wilson lines by one link and average them
//                     (displacement_elemental.py:53-71)
//
// Data layout in HBM: fields [Ne][Lz][Ly][Lx][3] complex128 (48 B per site, x fastest after
// colour), links [3][Lz][Ly][Lx][3][3] complex128 (144 B per site and direction).
//
// Stencil work decomposition: one thread owns (site, direction) and keeps the two link
// matrices it needs, U_d(x) and U_d(x-d), in registers while it walks over a chunk of
// eigenvectors, so link traffic is amortised over the chunk and every field element is streamed.
#include "edk_common.cuh"

namespace edk {

__device__ __forceinline__ cplx ldg(const cplx* p) { return __ldg(p); }

// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, const cplx a, const cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc -= conj(a) * b
__device__ __forceinline__ void cfnma_conj(cplx& acc, const cplx a, const cplx b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfma_conj(cplx& acc, const cplx a, const cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}

__device__ __forceinline__ void site_coords(int site, const Geom& g, int& x, int& y, int& z) {
    x = site % g.Lx;
    int r = site / g.Lx;
    y = r % g.Ly;
    z = r / g.Ly;
}
// neighbour of (x,y,z) one step along direction d (0=x,1=y,2=z), sign = +1/-1, periodic
__device__ __forceinline__ int neighbour(int x, int y, int z, int d, int sign, const Geom& g) {
    if (d == 0) x = (x + sign + g.Lx) % g.Lx;
    if (d == 1) y = (y + sign + g.Ly) % g.Ly;
    if (d == 2) z = (z + sign + g.Lz) % g.Lz;
    return (z * g.Ly + y) * g.Lx + x;
}

// ---------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------
// `sum` receives Re + Im of every element: the A-side operand Lr + Li of the 3M contraction
// (rows of `sum` are padded to an even number of doubles so that a TMA tensor map can describe them)
// byte-order reversal of one IEEE word: big-endian file payloads (ILDG links, QDP eigenvector records:
// filedata/ildg.py:70, filedata/timeslice.py:96 convert them on the host) are swapped as they are read
__device__ __forceinline__ float bswap_f32(float v) { return __uint_as_float(__byte_perm(__float_as_uint(v), 0, 0x0123)); }
__device__ __forceinline__ double bswap_f64(double v) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    return __hiloint2double((int)__byte_perm(lo, 0, 0x0123), (int)__byte_perm(hi, 0, 0x0123));
}

template <bool C8, bool BE>
__global__ void round_eigvecs_kernel(const void* __restrict__ in, cplx* __restrict__ out, double* __restrict__ sum, size_t n,
                                     size_t row, size_t sum_row) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double re, im;
        if (C8) {
            float2 v = __ldg((const float2*)in + i);
            if (BE) {
                v.x = bswap_f32(v.x);
                v.y = bswap_f32(v.y);
            }
            re = (double)v.x;
            im = (double)v.y;
        } else {
            double2 v = __ldg((const double2*)in + i);
            if (BE) {
                v.x = bswap_f64(v.x);
                v.y = bswap_f64(v.y);
            }
            // round-to-nearest-even double -> float -> double: numpy's complex128 -> complex64 assignment
            re = (double)__double2float_rn(v.x);
            im = (double)__double2float_rn(v.y);
        }
        out[i] = make_double2(re, im);
        sum[(i / row) * sum_row + i % row] = re + im;
    }
}

cudaError_t launch_round_eigvecs(const void* V_in, int flags, cplx* W0, double* W0_sum, size_t n_cplx, size_t row,
                                 size_t sum_row, cudaStream_t s) {
    int block = 256;
    size_t want = (n_cplx + block - 1) / block;
    int grid = (int)(want < (size_t)148 * 16 ? (want ? want : 1) : (size_t)148 * 16);
    const bool c8 = flags & EDK_EIGVECS_C8, be = flags & EDK_EIGVECS_BIG_ENDIAN;
    auto kern = round_eigvecs_kernel<false, false>;
    if (c8 && be) kern = round_eigvecs_kernel<true, true>;
    else if (c8) kern = round_eigvecs_kernel<true, false>;
    else if (be) kern = round_eigvecs_kernel<false, true>;
    EDK_LAUNCH(kern, grid, block, 0, s, V_in, W0, W0_sum, n_cplx, row, sum_row);
    return cudaGetLastError();
}

__global__ void reorder_links_kernel(const cplx* __restrict__ in, int layout, int big_endian, cplx* __restrict__ out, int V) {
    // out[d][site][m], m = 3*a + b
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)3 * V * 9;
    if (i >= n) return;
    int m = (int)(i % 9);
    size_t r = i / 9;
    int site = (int)(r % V);
    int d = (int)(r / V);
    size_t src = (layout == EDK_LINKS_FILE_T) ? (((size_t)site * 4 + d) * 9 + m) : i;
    cplx v = __ldg(in + src);
    if (big_endian) {
        v.x = bswap_f64(v.x);
        v.y = bswap_f64(v.y);
    }
    out[i] = v;
}

cudaError_t launch_reorder_links(const cplx* U_in, int layout, int big_endian, cplx* U_out, Geom g, cudaStream_t s) {
    size_t n = (size_t)3 * g.V * 9;
    int block = 256;
    EDK_LAUNCH(reorder_links_kernel, (unsigned)((n + block - 1) / block), block, 0, s, U_in, layout, big_endian, U_out, g.V);
    return cudaGetLastError();
}

// `rot` (optional) receives -i * phase, the table the (f, im) columns of the contraction read
__global__ void phase_table_kernel(cplx* __restrict__ phase, cplx* __restrict__ rot, const int* __restrict__ mom3, int nmom,
                                   Geom g) {
    int site = blockIdx.x * blockDim.x + threadIdx.x;
    int ip = blockIdx.y;
    if (site >= g.Vpad) return;
    cplx v = make_double2(0.0, 0.0);  // padding sites contribute nothing
    if (site < g.V) {
        int x, y, z;
        site_coords(site, g, x, y, z);
        long long px = mom3[3 * ip + 0], py = mom3[3 * ip + 1], pz = mom3[3 * ip + 2];
        // reduce each p*x mod L in integers so the argument of sincospi stays in [0, 6)
        long long rx = ((px * x) % g.Lx + g.Lx) % g.Lx;
        long long ry = ((py * y) % g.Ly + g.Ly) % g.Ly;
        long long rz = ((pz * z) % g.Lz + g.Lz) % g.Lz;
        double turns = (double)rx / (double)g.Lx + (double)ry / (double)g.Ly + (double)rz / (double)g.Lz;
        double sn, cs;
        sincospi(2.0 * turns, &sn, &cs);
        v = make_double2(cs, sn);
    }
    phase[(size_t)ip * g.Vpad + site] = v;
    if (rot != nullptr) rot[(size_t)ip * g.Vpad + site] = make_double2(v.y, -v.x);
}

cudaError_t launch_phase_table(cplx* phase, cplx* rot, const int* mom3_dev, int nmom, Geom g, cudaStream_t s) {
    dim3 block(256), grid((g.Vpad + 255) / 256, nmom);
    EDK_LAUNCH(phase_table_kernel, grid, block, 0, s, phase, rot, mom3_dev, nmom, g);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// nabla3: out_d = nabla_d W  for d = x, y, z
//
// HBM-bound (192 B of field traffic per (eigenvector, site), 216 DFMA).  One thread owns
// (site, direction): U_d(x) and U_d(x-d) stay in registers while it walks over a chunk of ~50
// eigenvectors (the 288 B of links per thread cost 54 B per (e, site) at 16 eigenvectors per CTA
// and 17 B at 50: measured 3.94 -> 4.35 TB/s).  The neighbour colour vectors of the next two
// eigenvectors are prefetched straight into registers (two independent LDG batches in flight per
// thread); L1 serves the x/y neighbour reuse, L2 the z reuse.  A warp = 32 consecutive sites of
// one direction, so its results are 1536 contiguous bytes: they are staged through warp-private
// shared memory and every STG.128 writes 512 contiguous bytes (full sectors) instead of 16-byte
// pieces strided by 48 bytes.
// Measured alternatives (32^3 / 48^3, Ne=200): 4-deep cp.async ring in shared memory 3.94 TB/s at
// 16 eigenvectors per CTA (3 stages 3.96, 2 stages 3.26, 6 stages 2.99, 3 CTAs/SM 2.41: shared
// memory eats the L1 that serves the reuse); eigenvector chunk as the fast CTA index 3.69; CTAs
// grouped per wave 3.84; links pinned in L2 (access-policy window) no change.
// ---------------------------------------------------------------------------------------
constexpr int NABLA_SITES = 64;  // sites per CTA (threadIdx.x), threadIdx.y = direction
constexpr int NABLA_THREADS = NABLA_SITES * 3;

// PLANE: also write the real Re + Im plane of every output element (the third A operand of the 3M contraction);
// the plane-wave form of the contraction does not read it, and without it the pass moves 27 % fewer bytes.
template <bool PLANE>
__global__ void __launch_bounds__(NABLA_THREADS, 2)
nabla3_kernel(const cplx* __restrict__ W, cplx* __restrict__ o0, cplx* __restrict__ o1, cplx* __restrict__ o2,
              double* __restrict__ s0, double* __restrict__ s1, double* __restrict__ s2, size_t sum_row,
              const cplx* __restrict__ links, Geom g, int Ne, int chunk) {
    EDK_SHARED cplx stage[NABLA_THREADS / 32][96];
    const int site = blockIdx.x * NABLA_SITES + threadIdx.x;
    const int d = threadIdx.y;
    const int lane = threadIdx.x & 31;
    cplx* my_stage = stage[(threadIdx.y * NABLA_SITES + threadIdx.x) >> 5];
    const int warp_site0 = blockIdx.x * NABLA_SITES + (threadIdx.x & ~31);
    const int warp_sites = min(32, g.V - warp_site0);
    if (warp_sites <= 0) return;    // whole warp past the end of the volume (no block barrier below)
    const bool active = site < g.V;  // lanes past the end only help with the coalesced stores
    int x, y, z;
    site_coords(active ? site : 0, g, x, y, z);
    const int sf = neighbour(x, y, z, d, +1, g);
    const int sb = neighbour(x, y, z, d, -1, g);
    const int e0 = blockIdx.y * chunk;
    const int e1 = min(e0 + chunk, Ne);
    const size_t fs = (size_t)g.V * 3;
    const cplx* pf = W + (size_t)sf * 3;
    const cplx* pb = W + (size_t)sb * 3;

    cplx fa[3], ba[3], fb[3], bb[3];
    auto fetch = [&](int e, cplx (&f)[3], cplx (&b)[3]) {
        if (e < e1 && active) {
            const size_t off = (size_t)e * fs;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                f[c] = ldg(pf + off + c);
                b[c] = ldg(pb + off + c);
            }
        }
    };
    fetch(e0, fa, ba);
    fetch(e0 + 1, fb, bb);

    cplx U[9], Ub[9];
    const cplx* pu = links + ((size_t)d * g.V + (active ? site : 0)) * 9;
    const cplx* pl = links + ((size_t)d * g.V + sb) * 9;
#pragma unroll
    for (int m = 0; m < 9; ++m) {
        U[m] = ldg(pu + m);
        Ub[m] = ldg(pl + m);
    }
    cplx* out = (d == 0 ? o0 : (d == 1 ? o1 : o2)) + (size_t)warp_site0 * 3;  // the warp's block
    double* outs = PLANE ? (d == 0 ? s0 : (d == 1 ? s1 : s2)) + (size_t)warp_site0 * 3 : nullptr;  // its Re + Im plane
    const int warp_cplx = warp_sites * 3;
    auto apply = [&](int e, const cplx (&f)[3], const cplx (&b)[3]) {
        cplx r[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            r[a] = make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cfma(r[a], U[3 * a + c], f[c]);         // U_d(x) W(x+d)
                cfnma_conj(r[a], Ub[3 * c + a], b[c]);  // - U_d(x-d)^dagger W(x-d)
            }
        }
        __syncwarp();  // the previous round's reads of the staging block are done
#pragma unroll
        for (int a = 0; a < 3; ++a) my_stage[lane * 3 + a] = r[a];
        __syncwarp();
        cplx* po = out + (size_t)e * fs;
        double* ps = outs + (size_t)e * sum_row;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j * 32 + lane < warp_cplx) {
                const cplx v = my_stage[j * 32 + lane];
                po[j * 32 + lane] = v;
                if (PLANE) ps[j * 32 + lane] = v.x + v.y;
            }
    };
    for (int e = e0; e < e1; e += 2) {
        cplx f0[3], b0[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            f0[c] = fa[c];
            b0[c] = ba[c];
        }
        fetch(e + 2, fa, ba);
        apply(e, f0, b0);
        if (e + 1 < e1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                f0[c] = fb[c];
                b0[c] = bb[c];
            }
            fetch(e + 3, fb, bb);
            apply(e + 1, f0, b0);
        }
    }
}

// eigenvectors per CTA: ~56 at most, balanced over the chunks, but never so few CTAs that the
// 148 SMs (2 CTAs each) see less than ~3 waves
static int stencil_chunk(int V, int sites_per_cta, int Ne) {
    const int nblk = (V + sites_per_cta - 1) / sites_per_cta;
    int nchunk = (Ne + 55) / 56;
    const int want = (3 * 296 + nblk - 1) / nblk;
    if (nchunk < want) nchunk = want;
    if (nchunk > (Ne + 7) / 8) nchunk = (Ne + 7) / 8;
    if (nchunk < 1) nchunk = 1;
    return (Ne + nchunk - 1) / nchunk;
}

cudaError_t launch_nabla3(const cplx* W_in, cplx* out_x, cplx* out_y, cplx* out_z, double* sum_x, double* sum_y, double* sum_z,
                          size_t sum_row, const cplx* links, Geom g, int Ne, cudaStream_t s) {
    const int chunk = stencil_chunk(g.V, NABLA_SITES, Ne);
    dim3 block(NABLA_SITES, 3);
    dim3 grid((g.V + NABLA_SITES - 1) / NABLA_SITES, (Ne + chunk - 1) / chunk);
    if ((sum_x != nullptr) != (sum_y != nullptr) || (sum_x != nullptr) != (sum_z != nullptr)) return cudaErrorInvalidValue;
    auto kern = sum_x ? nabla3_kernel<true> : nabla3_kernel<false>;  // planes are written for all three outputs or none
    EDK_LAUNCH(kern, grid, block, 0, s, W_in, out_x, out_y, out_z, sum_x, sum_y, sum_z, sum_row, links, g, Ne, chunk);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// displacement step: six lines in, six lines out, plus their mean
//   line d   (d<3): F_d^k(x) = U_d(x)        F_d^{k-1}(x+d)
//   line 5-d      : B_d^k(x) = U_d(x-d)^dag  B_d^{k-1}(x-d)
// ---------------------------------------------------------------------------------------
constexpr int DISP_SITES = 64;

__global__ void __launch_bounds__(DISP_SITES * 6)
displace_step6_kernel(Ptr6 p, cplx* __restrict__ mean_out, const cplx* __restrict__ links, Geom g, int Ne, int chunk) {
    EDK_SHARED cplx red[6][DISP_SITES][3];
    const int site = blockIdx.x * DISP_SITES + threadIdx.x;
    const int line = threadIdx.y;
    const bool fwd = line < 3;
    const int d = fwd ? line : 5 - line;
    const bool active = site < g.V;
    int x = 0, y = 0, z = 0, sn = 0;
    cplx U[9];
    if (active) {
        site_coords(site, g, x, y, z);
        sn = neighbour(x, y, z, d, fwd ? +1 : -1, g);
        const cplx* pu = links + ((size_t)d * g.V + (fwd ? site : sn)) * 9;
#pragma unroll
        for (int m = 0; m < 9; ++m) U[m] = ldg(pu + m);
    }
    const cplx* src = p.src[line];
    cplx* dst = p.dst[line];
    const int e0 = blockIdx.y * chunk;
    const int e1 = min(e0 + chunk, Ne);
    const size_t fs = (size_t)g.V * 3;
    const int tid = threadIdx.y * DISP_SITES + threadIdx.x;
    for (int e = e0; e < e1; ++e) {
        cplx r[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) r[a] = make_double2(0.0, 0.0);
        if (active) {
            cplx w[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) w[c] = ldg(src + (size_t)e * fs + (size_t)sn * 3 + c);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    if (fwd)
                        cfma(r[a], U[3 * a + b], w[b]);
                    else
                        cfma_conj(r[a], U[3 * b + a], w[b]);
                }
            }
            cplx* po = dst + (size_t)e * fs + (size_t)site * 3;
#pragma unroll
            for (int a = 0; a < 3; ++a) po[a] = r[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) red[line][threadIdx.x][a] = r[a];
        __syncthreads();
        if (tid < DISP_SITES * 3) {
            const int sl = tid / 3, a = tid % 3;
            const int so = blockIdx.x * DISP_SITES + sl;
            if (so < g.V) {
                cplx acc = red[0][sl][a];
#pragma unroll
                for (int l = 1; l < 6; ++l) {
                    acc.x += red[l][sl][a].x;
                    acc.y += red[l][sl][a].y;
                }
                mean_out[(size_t)e * fs + (size_t)so * 3 + a] = make_double2(acc.x / 6.0, acc.y / 6.0);
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_displace_step6(Ptr6 p, cplx* mean_out, const cplx* links, Geom g, int Ne, cudaStream_t s) {
    const int chunk = stencil_chunk(g.V, DISP_SITES, Ne);
    dim3 block(DISP_SITES, 6);
    dim3 grid((g.V + DISP_SITES - 1) / DISP_SITES, (Ne + chunk - 1) / chunk);
    EDK_LAUNCH(displace_step6_kernel, grid, block, 0, s, p, mean_out, links, g, Ne, chunk);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Laplacian of the eigensolver (lattice/generator/eigenvector.py:11-26):
//   (L F)(x) = 6 F(x) - sum_d [ U_d(x) F(x+d) + U_d(x-d)^dagger F(x-d) ]
// Same decomposition as nabla3 (thread = (site, direction), links in registers across a block of
// vectors); the three hop sums are reduced through shared memory.  96 B of field traffic per
// (vector, site).
// ---------------------------------------------------------------------------------------
constexpr int LAP_SITES = 64;
constexpr int LAP_EB = 50;  // vectors per CTA at most: the 288 B of links a thread holds are amortised as in nabla3

// Thread (site, direction) keeps U_d(x) and U_d(x - d) in registers across the CTA's vectors.  The neighbour colour
// vectors and the centre value of vector e + 1 are fetched into registers before the partial sums of vector e are
// reduced over the three directions through shared memory (two buffers, so one barrier per vector).
__global__ void __launch_bounds__(LAP_SITES * 3)
laplacian_kernel(const cplx* __restrict__ F, cplx* __restrict__ out, const cplx* __restrict__ links, Geom g, int nvec, int eb) {
    EDK_SHARED cplx red[2][3][LAP_SITES][3];
    const int site = blockIdx.x * LAP_SITES + threadIdx.x;
    const int d = threadIdx.y;
    const bool active = site < g.V;
    int sf = 0, sb = 0;
    cplx U[9], Ub[9];
    if (active) {
        int x, y, z;
        site_coords(site, g, x, y, z);
        sf = neighbour(x, y, z, d, +1, g);
        sb = neighbour(x, y, z, d, -1, g);
        const cplx* pu = links + ((size_t)d * g.V + site) * 9;
        const cplx* pl = links + ((size_t)d * g.V + sb) * 9;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            U[m] = ldg(pu + m);
            Ub[m] = ldg(pl + m);
        }
    }
    const int e0 = blockIdx.y * eb;
    const int e1 = min(e0 + eb, nvec);
    const size_t fs = (size_t)g.V * 3;
    const int tid = threadIdx.y * LAP_SITES + threadIdx.x;
    // the element this thread finishes: component a of site sl of the CTA (consecutive threads, consecutive 16 bytes)
    const int sl = tid / 3, a_out = tid % 3;
    const int so = blockIdx.x * LAP_SITES + sl;
    const bool writer = so < g.V;
    const size_t out_off = (size_t)so * 3 + a_out;
    cplx wf[3], wb[3], fc = make_double2(0.0, 0.0);
    auto fetch = [&](int e, cplx(&f)[3], cplx(&b)[3], cplx& c) {
        const cplx* Fe = F + (size_t)e * fs;
        if (active) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                f[k] = ldg(Fe + (size_t)sf * 3 + k);
                b[k] = ldg(Fe + (size_t)sb * 3 + k);
            }
        }
        if (writer) c = ldg(Fe + out_off);
    };
    if (e0 < e1) fetch(e0, wf, wb, fc);
    int buf = 0;
    for (int e = e0; e < e1; ++e) {
        cplx nf[3], nb[3], nc = make_double2(0.0, 0.0);
        if (e + 1 < e1) fetch(e + 1, nf, nb, nc);
        cplx r[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) r[a] = make_double2(0.0, 0.0);
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    cfma(r[a], U[3 * a + b], wf[b]);
                    cfma_conj(r[a], Ub[3 * b + a], wb[b]);
                }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) red[buf][d][threadIdx.x][a] = r[a];
        __syncthreads();
        if (writer) {
            const double hx = red[buf][0][sl][a_out].x + red[buf][1][sl][a_out].x + red[buf][2][sl][a_out].x;
            const double hy = red[buf][0][sl][a_out].y + red[buf][1][sl][a_out].y + red[buf][2][sl][a_out].y;
            out[(size_t)e * fs + out_off] = make_double2(6.0 * fc.x - hx, 6.0 * fc.y - hy);
        }
        // no second barrier: buffer `buf` is written again two vectors on, after every thread has passed the next barrier
        buf ^= 1;
        if (e + 1 < e1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) wf[k] = nf[k], wb[k] = nb[k];
            fc = nc;
        }
    }
}

cudaError_t launch_laplacian(const cplx* F, cplx* out, const cplx* links, Geom g, int nvec, cudaStream_t s) {
    if (nvec < 1) return cudaSuccess;
    // balanced chunks of <= LAP_EB vectors, more of them on small lattices so that the 148 SMs (2 CTAs each) see ~3 waves
    const int nsb = (g.V + LAP_SITES - 1) / LAP_SITES;
    int nchunk = (nvec + LAP_EB - 1) / LAP_EB;
    const int want = (148 * 2 * 3 + nsb - 1) / nsb;
    if (nchunk < want) nchunk = want < nvec ? want : nvec;
    const int eb = (nvec + nchunk - 1) / nchunk;
    nchunk = (nvec + eb - 1) / eb;
    dim3 block(LAP_SITES, 3);
    dim3 grid(nsb, nchunk);
    EDK_LAUNCH(laplacian_kernel, grid, block, 0, s, F, out, links, g, nvec, eb);
    return cudaGetLastError();
}

}  // namespace edk
