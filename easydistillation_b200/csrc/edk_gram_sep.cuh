// Separable contraction ("algo 4"): the kernels, as templates.  Included by edk_gram_sep.cu (host logic, z fold,
// launch dispatch) and by edk_gram_sep_s11.cu / _s12.cu / _s24.cu, which instantiate the kernels of one mode structure each
// so that the three compile side by side.
#pragma once
#include <type_traits>
#include <utility>

#include "edk_common.cuh"
#include "edk_pipe.cuh"

namespace edk {

// ---- mode structure -------------------------------------------------------------------------------------
// x modes: 0 = constant, 2q-1 = c_q, 2q = s_q (q = 1 .. QMAX); y weights are numbered the same way.
// xy-modes in the order of the loops below: (qx, x kind), then (qy, y kind) with qx^2 + qy^2 <= R2.
template <int QMAX, int R2>
struct SepModes {
    static constexpr int NX = 2 * QMAX + 1;
    static __host__ __device__ constexpr int find(int want_k, int wqx, int wxk, int wqy, int wyk, int what) {
        // what = 0: number of modes, 1: x mode of mode want_k, 2: y weight of mode want_k, 3: index of (wqx, wxk, wqy, wyk)
        int n = 0;
        for (int qx = 0; qx <= QMAX; ++qx)
            for (int xk = 0; xk < (qx ? 2 : 1); ++xk)
                for (int qy = 0; qy <= QMAX; ++qy) {
                    if (qx * qx + qy * qy > R2) continue;
                    for (int yk = 0; yk < (qy ? 2 : 1); ++yk) {
                        if (what == 1 && n == want_k) return qx ? 2 * qx - 1 + xk : 0;
                        if (what == 2 && n == want_k) return qy ? 2 * qy - 1 + yk : 0;
                        if (what == 3 && qx == wqx && xk == wxk && qy == wqy && yk == wyk) return n;
                        ++n;
                    }
                }
        return what == 0 ? n : -1;
    }
    static constexpr int N = find(0, 0, 0, 0, 0, 0);
    static __host__ __device__ constexpr int xm(int k) { return find(k, 0, 0, 0, 0, 1); }
    static __host__ __device__ constexpr int yw(int k) { return find(k, 0, 0, 0, 0, 2); }
    static __host__ __device__ constexpr int index(int qx, int xk, int qy, int yk) { return find(0, qx, xk, qy, yk, 3); }
};

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{})
template <class F, int... Is>
__host__ __device__ __forceinline__ void sep_static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__host__ __device__ __forceinline__ void sep_static_for(F&& f) {
    sep_static_for_impl(f, std::make_integer_sequence<int, N>{});
}

constexpr int SEP_THREADS = (SEP_WARPS + 4) * 32;  // 8 compute warps + one producer warpgroup (one warp of it works)
constexpr int SEP_REGS_CONSUMER = 232;
constexpr int SEP_REGS_PRODUCER = 40;
constexpr int SEP_MAX_STAGES = 8;
constexpr int SEP_CHUNKS = 3;                              // 128-byte chunks of a row of 8 sites (24 complex)
constexpr int SEP_TILE_L = SEP_CHUNKS * SEP_TE * 128;      // [chunk][row][128 bytes, swizzled]
constexpr int SEP_TILE_R = SEP_CHUNKS * SEP_TF * 128;
constexpr int SEP_HALF = SEP_TILE_L + SEP_TILE_R;          // front set; the back set follows
constexpr int SEP_STAGE = 2 * SEP_HALF;
constexpr int SEP_WX_BYTES = SEP_MAX_PR * 4 * 8;
constexpr int SEP_TAIL = SEP_WX_BYTES + 16 * SEP_MAX_STAGES + (int)sizeof(GramJob) + 64;
static_assert(SEP_TILE_L % 1024 == 0 && SEP_TILE_R % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");
static_assert(SEP_WARPS == 8 && SEP_TE == 16 && SEP_TF == 32, "warp grid 4 x 2, warp tile 4 x 16");

template <int QMAX, int R2, int PAIRS>
__global__ void __launch_bounds__(SEP_THREADS, 1) gram_sep_kernel(const SepParams P, const __grid_constant__ SepTma Tm) {
    using M = SepModes<QMAX, R2>;
    constexpr int NX = M::NX, NM = M::N;
    static_assert(PAIRS >= 1 && PAIRS <= 8, "a stage holds 8 front and 8 back sites");
    extern __shared__ __align__(1024) unsigned char smem[];
    // the swizzle pattern is a function of the shared-memory address: tiles must sit on 1024-byte boundaries
    unsigned char* sm = smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem) & 1023u)) & 1023u);
    const int nst = Tm.nstages;
    unsigned char* tail = sm + (size_t)nst * SEP_STAGE;
    double* swx = reinterpret_cast<double*>(tail);  // [Lx/2][4]
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail + SEP_WX_BYTES);
    const uint32_t bar_empty = bar_full + 8 * SEP_MAX_STAGES;
    GramJob* sjob = reinterpret_cast<GramJob*>(tail + SEP_WX_BYTES + 16 * SEP_MAX_STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // work item = (job, z-plane, e-tile, f-tile), f-tile fastest: CTAs that run together read the same plane
    int item = blockIdx.x;
    const int ft = item % P.n_ft;
    item /= P.n_ft;
    const int et = item % P.n_et;
    item /= P.n_et;
    const int z = item % P.Lz;
    const int job_id = item / P.Lz;
    const int e0 = et * SEP_TE, f0 = ft * SEP_TF;
    const int Ne = P.Ne;

    // padded work of edge tiles is skipped where it is warp-uniform: warp (we, wf) has rows 4 we .. 4 we + 3 of L
    // and two blocks of 8 rows of R from 16 wf; a warp without a valid row or block leaves right after setmaxnreg
    auto blocks_of = [&](int w) {
        const int rows_ok = e0 + (w & 3) * 4 < Ne;
        const int jn = (Ne - (f0 + (w >> 2) * 16) + 7) / 8;
        return rows_ok ? (jn < 0 ? 0 : (jn > 2 ? 2 : jn)) : 0;
    };
    int n_active = 0;
#pragma unroll
    for (int w = 0; w < SEP_WARPS; ++w) n_active += blocks_of(w) > 0;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, n_active > 0 ? n_active : 1);
        }
        mbar_init_fence();
    }
    const int PR = P.Lx >> 1;
    for (int i = tid; i < PR * 4; i += SEP_THREADS) swx[i] = P.wx[i];
    __syncthreads();
    // self pair (L == R): the site product is Hermitian in (e, f) and the weights are real, so
    // Y[m][e][f] = conj(Y[m][f][e]); tiles entirely below the diagonal are left to the fold kernel's mirror read.
    // (No early return here: an exit ahead of setmaxnreg makes ptxas spill the accumulators in the stage loop.)
    const bool skip_tile = sjob->nseg == 1 && sjob->Lf[0] == sjob->Rf[0] && e0 > f0 + SEP_TF - 1;
    const int SR = P.SR;
    const int T = skip_tile ? 0 : sjob->nseg * P.Ly * SR;  // stages: every segment walks the plane once
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);

    if (warp >= SEP_WARPS) {
        // ================================ producer warpgroup ================================
        warpgroup_reg_dealloc<SEP_REGS_PRODUCER>();
        if (warp != SEP_WARPS) return;
        const int plane_site0 = z * P.Lx * P.Ly;
        // lanes 0-11: tile = lane / 3 (front L, front R, back L, back R), 128-byte chunk = lane % 3
        const int tile = lane / 3, chunk = lane - 3 * tile;
        const int back = tile >> 1, is_r = tile & 1;
        const uint32_t dst_off = (uint32_t)(back * SEP_HALF + (is_r ? SEP_TILE_L + chunk * (SEP_TF * 128) : chunk * (SEP_TE * 128)));
        int seg = 0, y = 0, k = 0, s = 0;
        uint32_t par = 1;  // the first pass over the ring finds every slot free
        for (int it = 0; it < T; ++it) {
            mbar_wait(bar_empty + 8 * s, par);
            const uint32_t full = bar_full + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)SEP_STAGE);
            __syncwarp();
            if (lane < 12) {
                // 8 sites from the front of the row's untouched part, 8 sites up to its back; sites a box reads past the
                // row (or zero-filled past the end of the field) belong to no pair of this stage
                const int site0 = plane_site0 + y * P.Lx + (back ? P.Lx - 8 - PAIRS * k : PAIRS * k);
                tma_load_3d(sm_base + (uint32_t)(s * SEP_STAGE) + dst_off, is_r ? (const void*)Tm.mapR : (const void*)Tm.mapL, full,
                            site0 * 6 + 16 * chunk, is_r ? f0 : e0, is_r ? sjob->Rf[seg] : sjob->Lf[seg]);
            }
            if (++k == SR) {
                k = 0;
                if (++y == P.Ly) {
                    y = 0;
                    ++seg;
                }
            }
            if (++s == nst) {
                s = 0;
                par ^= 1;
            }
        }
        return;
    }

    // ================================== compute warps ==================================
    warpgroup_reg_alloc<SEP_REGS_CONSUMER>();
    const int jn = blocks_of(warp);
    if (jn == 0) return;  // not counted in the empty barriers
    const int we = warp & 3, wf = warp >> 2;
    // lane -> (row of L, row of R): every aligned group of 4 lanes touches only 2 distinct L rows and 2 distinct R rows.
    // The shared-memory pipe retires a 128-bit warp load in 2 cycles if no such group asks for more than two 16-byte
    // addresses, in 4 otherwise (tools/microbench/lds_patterns.cu: lane & 7 costs 4, this mapping 2).
    const int le = (lane >> 1) & 3, lf = ((lane >> 3) << 1) | (lane & 1);
    const int row_l = we * 4 + le;
    const uint32_t off_l = (uint32_t)(row_l * 128), xor_l = (uint32_t)((row_l & 7) << 4);
    const uint32_t off_r = (uint32_t)(SEP_TILE_L + (wf * 16 + lf) * 128), xor_r = (uint32_t)(lf << 4);

    double xr[2][NX], xi[2][NX], yr[2][NM], yi[2][NM];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int m = 0; m < NX; ++m) xr[j][m] = xi[j][m] = 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) yr[j][m] = yi[j][m] = 0.0;
    }

    int cur_sign = 1;
    int seg = 0, y = 0, k = 0, s = 0;
    uint32_t par = 0;
    for (int it = 0; it < T; ++it) {
        if (y == 0 && k == 0) {  // a new segment: fold its sign by flipping the running sums (X is zero at a row start)
            const int sgn = sjob->sign[seg];
            if (sgn != cur_sign) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        yr[j][m] = flip_sign(yr[j][m]);
                        yi[j][m] = flip_sign(yi[j][m]);
                    }
                cur_sign = sgn;
            }
        }
        mbar_wait(bar_full + 8 * s, par);
        const unsigned char* stage = sm + (size_t)s * SEP_STAGE;
        const double* wrow = swx + (size_t)(k * PAIRS) * 4;
        auto stage_body = [&](auto JN) {
            constexpr int NJ = decltype(JN)::value;
#pragma unroll
            for (int t = 0; t < PAIRS; ++t) {
                // weights of pair t of the stage (the same for every lane: broadcast reads)
                double wc[QMAX + 1], ws[QMAX + 1];
#pragma unroll
                for (int q = 1; q <= QMAX; ++q) {
                    const double2 w = *reinterpret_cast<const double2*>(wrow + 4 * t + 2 * (q - 1));
                    wc[q] = w.x;
                    ws[q] = w.y;
                }
                // front site t of the stage and its partner, local site 7 - t of the back set
                cplx lfr[3], lbk[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int kf = 3 * t + c, kb = 3 * (7 - t) + c;
                    lfr[c] = *reinterpret_cast<const cplx*>(stage + (kf >> 3) * (SEP_TE * 128) + off_l + ((uint32_t)((kf & 7) << 4) ^ xor_l));
                    lbk[c] = *reinterpret_cast<const cplx*>(stage + SEP_HALF + (kb >> 3) * (SEP_TE * 128) + off_l +
                                                            ((uint32_t)((kb & 7) << 4) ^ xor_l));
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    cplx rfr[3], rbk[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int kf = 3 * t + c, kb = 3 * (7 - t) + c;
                        rfr[c] = *reinterpret_cast<const cplx*>(stage + (kf >> 3) * (SEP_TF * 128) + off_r + j * 1024 +
                                                                ((uint32_t)((kf & 7) << 4) ^ xor_r));
                        rbk[c] = *reinterpret_cast<const cplx*>(stage + SEP_HALF + (kb >> 3) * (SEP_TF * 128) + off_r + j * 1024 +
                                                                ((uint32_t)((kb & 7) << 4) ^ xor_r));
                    }
                    // conj(L) . R over the three colours, front site and back site
                    double fr = lfr[0].x * rfr[0].x, fi = lfr[0].x * rfr[0].y;
                    double br = lbk[0].x * rbk[0].x, bi = lbk[0].x * rbk[0].y;
                    fr = fma(lfr[0].y, rfr[0].y, fr);
                    fi = fma(-lfr[0].y, rfr[0].x, fi);
                    br = fma(lbk[0].y, rbk[0].y, br);
                    bi = fma(-lbk[0].y, rbk[0].x, bi);
#pragma unroll
                    for (int c = 1; c < 3; ++c) {
                        fr = fma(lfr[c].x, rfr[c].x, fr);
                        fi = fma(lfr[c].x, rfr[c].y, fi);
                        br = fma(lbk[c].x, rbk[c].x, br);
                        bi = fma(lbk[c].x, rbk[c].y, bi);
                        fr = fma(lfr[c].y, rfr[c].y, fr);
                        fi = fma(-lfr[c].y, rfr[c].x, fi);
                        br = fma(lbk[c].y, rbk[c].y, br);
                        bi = fma(-lbk[c].y, rbk[c].x, bi);
                    }
                    const double sr = fr + br, si = fi + bi, dr = fr - br, di = fi - bi;
                    xr[j][0] += sr;
                    xi[j][0] += si;
#pragma unroll
                    for (int q = 1; q <= QMAX; ++q) {
                        xr[j][2 * q - 1] = fma(wc[q], sr, xr[j][2 * q - 1]);
                        xi[j][2 * q - 1] = fma(wc[q], si, xi[j][2 * q - 1]);
                        xr[j][2 * q] = fma(ws[q], dr, xr[j][2 * q]);
                        xi[j][2 * q] = fma(ws[q], di, xi[j][2 * q]);
                    }
                }
            }
        };
        if (jn == 2)
            stage_body(std::integral_constant<int, 2>{});
        else
            stage_body(std::integral_constant<int, 1>{});
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == nst) {
            s = 0;
            par ^= 1;
        }
        if (++k == SR) {
            // ---- end of row y: y stage, then X = 0 ----
            k = 0;
            double wyv[NX];
            wyv[0] = 1.0;
#pragma unroll
            for (int q = 1; q <= QMAX; ++q) {
                const double2 w = *reinterpret_cast<const double2*>(P.wy + 4 * y + 2 * (q - 1));
                wyv[2 * q - 1] = w.x;
                wyv[2 * q] = w.y;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < jn) {
                    sep_static_for<NM>([&](auto K) {
                        constexpr int m = decltype(K)::value;
                        constexpr int mx = M::xm(m), my = M::yw(m);
                        if constexpr (my == 0) {
                            yr[j][m] += xr[j][mx];
                            yi[j][m] += xi[j][mx];
                        } else {
                            yr[j][m] = fma(wyv[my], xr[j][mx], yr[j][m]);
                            yi[j][m] = fma(wyv[my], xi[j][mx], yi[j][m]);
                        }
                    });
                }
#pragma unroll
                for (int m = 0; m < NX; ++m) xr[j][m] = xi[j][m] = 0.0;
            }
            if (++y == P.Ly) {
                y = 0;
                ++seg;
            }
        }
    }

    // ---- epilogue: lane holds (e, f) = (e0 + row_l, f0 + 16 wf + lf + 8 j), every separable mode ----
    if (skip_tile) return;
    const double fs = (double)cur_sign;
    const size_t mat = (size_t)Ne * Ne;
    const int e = e0 + row_l;
    if (e >= Ne) return;
    cplx* Yp = P.Y + ((size_t)job_id * P.Lz + z) * (size_t)NM * mat + (size_t)e * Ne;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int f = f0 + wf * 16 + lf + 8 * j;
        if (j < jn && f < Ne) {
#pragma unroll
            for (int m = 0; m < NM; ++m) Yp[(size_t)m * mat + f] = make_double2(fs * yr[j][m], fs * yi[j][m]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gram_sepx_kernel: lane tile 2 x 2, y-stage accumulators in TENSOR MEMORY, CTA tiles from a host-built table.
//
// What bounds gram_sep_kernel is not the FP64 pipe but the shared-memory data pipe and its own registers:
//  * a 128-bit warp load costs two cycles of the SM's one data pipe at best (it writes 512 bytes of registers), four if
//    an aligned group of 4 lanes asks for more than two 16-byte addresses (tools/microbench/lds_patterns.cu).  A lane of
//    the 1 x 2 tile needs 20 loads per 76 FP64 operations and four sub-partitions share the pipe: more than one pipe
//    cycle per FP64-pipe cycle.  With 2 x 2 elements per lane it is 24 loads per 152 operations, and the lane -> row
//    map (le, lf) below keeps every load at two cycles; the x weights come through the constant cache (kernel
//    parameter) instead of shared memory.
//  * Y (13 complex per element, touched once per row) would be 208 of a lane's registers.  Tensor memory is 256 KB per
//    SM that this path (FP64: no tcgen05.mma kind) leaves unused; with the 32x32b shape of tcgen05.ld / tcgen05.st every
//    thread owns a row of columns, so Y costs no register between rows and no shared-memory bandwidth: per row and
//    element a lane reads 4 x 16 words, applies its multiply-adds and writes them back, the load of the next transfer in
//    flight meanwhile.  Layout: lane quadrant 32 (warp % 4), columns 256 (warp / 4) + 128 i + 64 j + 4 m + {0, 1: Re,
//    2, 3: Im} for mode m of element (i, j).  The producer warp allocates the 512 columns and frees them after the
//    last compute warp has arrived on bar_done.
//  * tiles: 8 compute warps arranged 4 x 2 (32 x 32 elements), 1 x 8 (8 x 128) or 8 x 1 (64 x 16) over the host's tile
//    table (SepTile, sep_build_tiles), one launch per shape or one in all (see the kernels below).  One fixed 32 x 32
//    tile leaves, at Ne = 200, a ragged row and column of tiles whose CTAs run two to four of their eight warps; the
//    flat shapes cover those edges.
// Lane (le, lf) of warp (a, b) owns rows e0 + 8 a + le + {0, 4} of L and rows f0 + 16 b + lf + {0, 8} of R.
// Measured at config 5 (48^3, Ne = 200, 33 momenta): 116 ms per timeslice against 172 ms of gram_sep_kernel and 179 ms
// of gram_pwf_kernel; FP64 pipe 71 % busy over the launch of 32 x 32 tiles, about 84 % inside the stage loop.
// ---------------------------------------------------------------------------------------------------------
struct SepxGeom {
    static constexpr int WARPS = 8, THREADS = (WARPS + 4) * 32;
    static constexpr int REGS_CONSUMER = 232, REGS_PRODUCER = 40;
    static constexpr int TAIL = 16 * SEP_MAX_STAGES + 16 + (int)sizeof(GramJob) + 64;
};

template <int QMAX, int R2, int PAIRS, int WE>
__device__ __forceinline__ void sepx_cta(const SepParams& P, const SepTmaX& Tm, const SepWeights& Wx) {
    using M = SepModes<QMAX, R2>;
    using G = SepxGeom;
    constexpr int NX = M::NX, NM = M::N;
    constexpr int NH = (NM + 7) / 8;
    static_assert(PAIRS >= 1 && PAIRS <= 8 && NH <= 2, "a stage holds 8 front and 8 back sites; 64 TMEM columns per element");
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (1024u - (smem_s & 1023u)) & 1023u;
    const uint32_t sm_base = smem_s + pad;  // swizzled tiles sit on 1024-byte boundaries
    unsigned char* tail = smem + pad + (227 * 1024 - 1024 - G::TAIL);
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail);
    const uint32_t bar_empty = bar_full + 8 * SEP_MAX_STAGES;
    const uint32_t bar_done = bar_empty + 8 * SEP_MAX_STAGES;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tail + 16 * SEP_MAX_STAGES + 8);
    GramJob* sjob = reinterpret_cast<GramJob*>(tail + 16 * SEP_MAX_STAGES + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // work item = (job, z-plane, tile), tile fastest: CTAs that run together read the same plane
    int item = blockIdx.x;
    const int tile_id = item % P.ntiles;
    item /= P.ntiles;
    const int z = item % P.Lz;
    const int job_id = item / P.Lz;
    const SepTile tile = P.tiles[tile_id];
    const int e0 = tile.e0, f0 = tile.f0;
    constexpr int shape = WE == 4 ? 0 : (WE == 1 ? 1 : 2);  // WE warps along e, 8 / WE along f: one launch per shape
    static_assert(WE == 4 || WE == 1 || WE == 8, "tile shapes 32 x 32, 8 x 128, 64 x 16");
    const int Ne = P.Ne;
    constexpr int cl = 8 * WE * 128, cr = (128 / WE) * 128;  // bytes of one 128-byte chunk of the L / R tile
    constexpr int half = SEP_CHUNKS * (cl + cr), stage_bytes = 2 * half;
    const int nst = Tm.nstages[shape];

    // warp w = (a = w % WE, b = w / WE): two blocks of 4 rows of L from e0 + 8 a, two blocks of 8 rows of R from f0 + 16 b
    const int e1 = tile.e1, f1 = tile.f1;  // the tile's own limits: a strip tile must not reach into the next region
    auto blocks_i = [&](int w) {
        const int n = (e1 - (e0 + (w % WE) * 8) + 3) / 4;
        return n < 0 ? 0 : (n > 2 ? 2 : n);
    };
    auto blocks_j = [&](int w) {
        const int n = (f1 - (f0 + (w / WE) * 16) + 7) / 8;
        return n < 0 ? 0 : (n > 2 ? 2 : n);
    };
    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    __syncthreads();
    // self pair (L == R): the site product is Hermitian in (e, f) and the weights are real, so Y[m][e][f] =
    // conj(Y[m][f][e]).  Blocks of 8 x 8 elements strictly below the diagonal are left to the fold kernel's mirror read:
    // a whole tile below it does nothing, inside a tile the warps whose 8 rows lie below all of their 16 columns.
    const bool self_pair = sjob->nseg == 1 && sjob->Lf[0] == sjob->Rf[0];
    auto warp_needed = [&](int w) {
        if (blocks_i(w) == 0 || blocks_j(w) == 0) return false;
        return !(self_pair && e0 + (w % WE) * 8 > f0 + (w / WE) * 16 + 15);
    };
    int n_active = 0;
#pragma unroll
    for (int w = 0; w < G::WARPS; ++w) n_active += warp_needed(w);
    const int SR = P.SR;
    const int T = n_active ? sjob->nseg * P.Ly * SR : 0;  // stages: every segment walks the plane once

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, n_active > 0 ? n_active : 1);
        }
        mbar_init(bar_done, n_active > 0 ? n_active : 1);
        mbar_init_fence();
    }
    if (warp == G::WARPS) tmem_alloc((uint32_t)__cvta_generic_to_shared(tmem_base_s), 512);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp >= G::WARPS) {
        // ================================ producer warpgroup ================================
        warpgroup_reg_dealloc<G::REGS_PRODUCER>();
        if (warp != G::WARPS) return;
        const int plane_site0 = z * P.Lx * P.Ly;
        // lanes 0-11: tile part = lane / 3 (front L, front R, back L, back R), 128-byte chunk = lane % 3
        const int part = lane / 3, chunk = lane - 3 * part;
        const int back = part >> 1, is_r = part & 1;
        const uint32_t dst_off = (uint32_t)(back * half + (is_r ? SEP_CHUNKS * cl + chunk * cr : chunk * cl));
        const void* map = is_r ? (const void*)Tm.mapR[shape] : (const void*)Tm.mapL[shape];
        int seg = 0, y = 0, k = 0, s = 0;
        uint32_t par = 1;  // the first pass over the ring finds every slot free
        for (int it = 0; it < T; ++it) {
            mbar_wait(bar_empty + 8 * s, par);
            const uint32_t full = bar_full + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)stage_bytes);
            __syncwarp();
            if (lane < 12) {
                const int site0 = plane_site0 + y * P.Lx + (back ? P.Lx - 8 - PAIRS * k : PAIRS * k);
                tma_load_3d(sm_base + (uint32_t)(s * stage_bytes) + dst_off, map, full, site0 * 6 + 16 * chunk, is_r ? f0 : e0,
                            is_r ? sjob->Rf[seg] : sjob->Lf[seg]);
            }
            if (++k == SR) {
                k = 0;
                if (++y == P.Ly) {
                    y = 0;
                    ++seg;
                }
            }
            if (++s == nst) {
                s = 0;
                par ^= 1;
            }
        }
        // tensor memory goes back once every compute warp is through with it
        if (n_active > 0) mbar_wait(bar_done, 0);
        tmem_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
        return;
    }

    // ================================== compute warps ==================================
    warpgroup_reg_alloc<G::REGS_CONSUMER>();
    if (!warp_needed(warp)) return;  // not counted in the barriers
    const int ni = blocks_i(warp), nj = blocks_j(warp);
    const int wa = warp % WE, wb = warp / WE;
    // every aligned group of 4 lanes touches 2 distinct L rows and 2 distinct R rows: 2-cycle 128-bit loads
    const int le = (lane >> 1) & 3, lf = ((lane >> 3) << 1) | (lane & 1);
    const int row_l = wa * 8 + le;  // + 4 i; the swizzle term (row & 7) is le, then le + 4
    const int row_r = wb * 16 + lf;  // + 8 j; (row & 7) is lf
    const uint32_t off_l = (uint32_t)(row_l * 128), xor_l = (uint32_t)(le << 4);
    const uint32_t off_r = (uint32_t)(SEP_CHUNKS * cl + row_r * 128), xor_r = (uint32_t)(lf << 4);
    const uint32_t ty = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));  // + 128 i + 64 j + 32 half

    double xr[2][2][NX], xi[2][2][NX];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int m = 0; m < NX; ++m) xr[i][j][m] = xi[i][j][m] = 0.0;
    {
        uint32_t zero[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) zero[q] = 0u;
#pragma unroll
        for (int q = 0; q < 4 * NH; ++q) tmem_st32(ty + 64 * (q / NH) + 32 * (q % NH), zero);
        tmem_wait_st();
    }

    int cur_sign = 1;
    int seg = 0, y = 0, k = 0, s = 0;
    uint32_t par = 0;
    for (int it = 0; it < T; ++it) {
        if (y == 0 && k == 0) {  // a new segment: fold its sign by flipping the running sums (X is zero at a row start)
            const int sgn = sjob->sign[seg];
            if (sgn != cur_sign) {
                tmem_wait_st();
                uint32_t yb[32];
#pragma unroll
                for (int q = 0; q < 4 * NH; ++q) {
                    tmem_ld32(ty + 64 * (q / NH) + 32 * (q % NH), yb);
                    tmem_wait_ld();
#pragma unroll
                    for (int w = 1; w < 32; w += 2) yb[w] ^= 0x80000000u;
                    tmem_st32(ty + 64 * (q / NH) + 32 * (q % NH), yb);
                }
                cur_sign = sgn;
            }
        }
        mbar_wait(bar_full + 8 * s, par);
        const unsigned char* st_l = smem + pad + (size_t)s * stage_bytes + off_l;  // this lane's row of the front L tile, chunk 0
        const unsigned char* st_r = smem + pad + (size_t)s * stage_bytes + off_r;
        const int pair0 = k * PAIRS;
        auto stage_body = [&](auto NI_, auto NJ_) {
            constexpr int NI = decltype(NI_)::value, NJ = decltype(NJ_)::value;
            sep_static_for<PAIRS>([&](auto TT) {
                constexpr int t = decltype(TT)::value;
                // x weights of this pair through the constant cache
                double wc[QMAX + 1], ws[QMAX + 1];
#pragma unroll
                for (int q = 1; q <= QMAX; ++q) {
                    wc[q] = Wx.w[(pair0 + t) * 4 + 2 * (q - 1)];
                    ws[q] = Wx.w[(pair0 + t) * 4 + 2 * (q - 1) + 1];
                }
                double cr_[2][NI][NJ], ci_[2][NI][NJ];  // site products [front / back]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // front site t of the stage, then its partner, local site 7 - t of the back set
                    cplx lv[NI][3], rv[NJ][3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int kc = h ? 3 * (7 - t) + c : 3 * t + c;
                        const uint32_t sl = (uint32_t)((kc & 7) << 4);
                        const uint32_t ol = (uint32_t)(h * half + (kc >> 3) * cl), orr = (uint32_t)(h * half + (kc >> 3) * cr);
#pragma unroll
                        for (int i = 0; i < NI; ++i)  // row + 4 i: 512 bytes on, slot XOR 4
                            lv[i][c] = *reinterpret_cast<const cplx*>(st_l + ol + (uint32_t)(i * 512) + ((sl ^ xor_l) ^ (uint32_t)(i << 6)));
#pragma unroll
                        for (int j = 0; j < NJ; ++j) rv[j][c] = *reinterpret_cast<const cplx*>(st_r + orr + (uint32_t)(j * 1024) + (sl ^ xor_r));
                    }
#pragma unroll
                    for (int i = 0; i < NI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            double ar = lv[i][0].x * rv[j][0].x, ai = lv[i][0].x * rv[j][0].y;
                            ar = fma(lv[i][0].y, rv[j][0].y, ar);
                            ai = fma(-lv[i][0].y, rv[j][0].x, ai);
#pragma unroll
                            for (int c = 1; c < 3; ++c) {
                                ar = fma(lv[i][c].x, rv[j][c].x, ar);
                                ai = fma(lv[i][c].x, rv[j][c].y, ai);
                                ar = fma(lv[i][c].y, rv[j][c].y, ar);
                                ai = fma(-lv[i][c].y, rv[j][c].x, ai);
                            }
                            cr_[h][i][j] = ar;
                            ci_[h][i][j] = ai;
                        }
                }
#pragma unroll
                for (int i = 0; i < NI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const double sr = cr_[0][i][j] + cr_[1][i][j], si = ci_[0][i][j] + ci_[1][i][j];
                        const double dr = cr_[0][i][j] - cr_[1][i][j], di = ci_[0][i][j] - ci_[1][i][j];
                        xr[i][j][0] += sr;
                        xi[i][j][0] += si;
#pragma unroll
                        for (int q = 1; q <= QMAX; ++q) {
                            xr[i][j][2 * q - 1] = fma(wc[q], sr, xr[i][j][2 * q - 1]);
                            xi[i][j][2 * q - 1] = fma(wc[q], si, xi[i][j][2 * q - 1]);
                            xr[i][j][2 * q] = fma(ws[q], dr, xr[i][j][2 * q]);
                            xi[i][j][2 * q] = fma(ws[q], di, xi[i][j][2 * q]);
                        }
                    }
            });
        };
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        if (ni == 2 && nj == 2)
            stage_body(I2{}, I2{});
        else if (ni == 2)
            stage_body(I2{}, I1{});
        else if (nj == 2)
            stage_body(I1{}, I2{});
        else
            stage_body(I1{}, I1{});
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == nst) {
            s = 0;
            par ^= 1;
        }
        if (++k == SR) {
            // ---- end of row y: y stage on the accumulators in tensor memory, then X = 0.  The transfers of 16 words (4
            // modes of one element) are walked in one sequence with two buffers: the load of transfer n + 1 is in flight
            // while the multiply-adds of transfer n run (tcgen05.wait::ld waits for every outstanding load, so it comes
            // first).  32-word transfers would cost 64 registers here, which ptxas takes from the stage loop (spills). ----
            k = 0;
            double wyv[NX];
            wyv[0] = 1.0;
#pragma unroll
            for (int q = 1; q <= QMAX; ++q) {
                const double2 w = *reinterpret_cast<const double2*>(P.wy + 4 * y + 2 * (q - 1));
                wyv[2 * q - 1] = w.x;
                wyv[2 * q] = w.y;
            }
            tmem_wait_st();  // the stores of the previous row
            constexpr int NQ = (NM + 3) / 4;  // transfers of 16 words (4 complex modes) per element
            constexpr int NTR = 4 * NQ;       // transfer n = (element n / NQ = 2 i + j, modes 4 (n % NQ) ..)
            auto wanted = [&](int n) { return (n / NQ) / 2 < ni && (n / NQ) % 2 < nj; };
            auto taddr = [&](int n) { return ty + (uint32_t)(64 * (n / NQ) + 16 * (n % NQ)); };
            // multiply-adds of transfer n on the words in `in`, results to tensor memory from a buffer of their own (in
            // place, ptxas copies all 16 registers into the store's operand tuple)
            auto update = [&](auto N_, const uint32_t(&in)[16]) {
                constexpr int n = decltype(N_)::value;
                constexpr int i = (n / NQ) / 2, j = (n / NQ) % 2, m0 = 4 * (n % NQ), m1 = (NM < m0 + 4 ? NM : m0 + 4);
                uint32_t ys[16];
                sep_static_for<(m1 - m0)>([&](auto K) {
                    constexpr int m = m0 + decltype(K)::value;
                    constexpr int mx = M::xm(m), my = M::yw(m);
                    double re = tmem_get_f64(in, 2 * (m - m0)), im = tmem_get_f64(in, 2 * (m - m0) + 1);
                    if constexpr (my == 0) {
                        re += xr[i][j][mx];
                        im += xi[i][j][mx];
                    } else {
                        re = fma(wyv[my], xr[i][j][mx], re);
                        im = fma(wyv[my], xi[i][j][mx], im);
                    }
                    tmem_put_f64(ys, 2 * (m - m0), re);
                    tmem_put_f64(ys, 2 * (m - m0) + 1, im);
                });
                if constexpr (m1 - m0 < 4) {
#pragma unroll
                    for (int w = 4 * (m1 - m0); w < 16; ++w) ys[w] = 0u;  // columns of no mode
                }
                tmem_st16(taddr(n), ys);
            };
            uint32_t yb[2][16];
            tmem_ld16(taddr(0), yb[0]);
            sep_static_for<NTR>([&](auto N_) {
                constexpr int n = decltype(N_)::value;
                tmem_wait_ld();
                if constexpr (n + 1 < NTR) {
                    if (wanted(n + 1)) tmem_ld16(taddr(n + 1), yb[(n + 1) & 1]);
                }
                if (wanted(n)) update(N_, yb[n & 1]);
            });
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int m = 0; m < NX; ++m) xr[i][j][m] = xi[i][j][m] = 0.0;
            if (++y == P.Ly) {
                y = 0;
                ++seg;
            }
        }
    }

    // ---- epilogue: lane holds (e, f) = (e0 + 8 a + le + 4 i, f0 + 16 b + lf + 8 j), every separable mode, in tensor memory ----
    {
        tmem_wait_st();
        const double fs = (double)cur_sign;
        const size_t mat = (size_t)Ne * Ne;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (i < ni && j < nj) {  // warp-uniform: the TMEM loads are .sync.aligned
                    const int e = e0 + row_l + 4 * i, f = f0 + row_r + 8 * j;
                    cplx* Yp = P.Y + ((size_t)job_id * P.Lz + z) * (size_t)NM * mat + (size_t)(e < e1 ? e : 0) * Ne;
#pragma unroll
                    for (int hh = 0; hh < NH; ++hh) {
                        uint32_t yb[32];
                        tmem_ld32(ty + 128 * i + 64 * j + 32 * hh, yb);
                        tmem_wait_ld();
                        if (e < e1 && f < f1) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (8 * hh + q < NM)
                                    Yp[(size_t)(8 * hh + q) * mat + f] = make_double2(fs * tmem_get_f64(yb, 2 * q), fs * tmem_get_f64(yb, 2 * q + 1));
                        }
                    }
                }
            }
    }
    tmem_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_done);
}

// One launch over the whole tile table: the CTA takes the warp grid of its tile (a block-uniform branch into one of three
// bodies whose strides are compile-time constants).  Used for small problems, where two launches and two tails less
// count (config 2: 0.36 -> 0.33 ms); at config 4 / 5 the strips running among the 32 x 32 tiles cost more than their
// L2 hits bring (117.7 against 116.0 ms) and the host launches one shape at a time (build_sep).
template <int QMAX, int R2, int PAIRS>
__global__ void __launch_bounds__(SepxGeom::THREADS, 1) gram_sepx_kernel(const SepParams P, const __grid_constant__ SepTmaX Tm,
                                                                         const __grid_constant__ SepWeights Wx) {
    const int shape = P.tiles[blockIdx.x % P.ntiles].shape;
    if (shape == 0)
        sepx_cta<QMAX, R2, PAIRS, 4>(P, Tm, Wx);
    else if (shape == 1)
        sepx_cta<QMAX, R2, PAIRS, 1>(P, Tm, Wx);
    else
        sepx_cta<QMAX, R2, PAIRS, 8>(P, Tm, Wx);
}
// One launch per shape: P.tiles holds tiles of one shape only.
template <int QMAX, int R2, int PAIRS, int WE>
__global__ void __launch_bounds__(SepxGeom::THREADS, 1) gram_sepx1_kernel(const SepParams P, const __grid_constant__ SepTmaX Tm,
                                                                          const __grid_constant__ SepWeights Wx) {
    sepx_cta<QMAX, R2, PAIRS, WE>(P, Tm, Wx);
}

#ifndef EDK_EMU_NO_LAUNCHERS
// ---- launch helpers of one mode structure (instantiated in edk_gram_sep_s*.cu) --------------------------------
template <int QMAX, int R2, int PAIRS>
cudaError_t launch_gram_sep_p(const SepParams& P, const SepTma& T, int bytes, unsigned items, cudaStream_t s) {
    auto kern = gram_sep_kernel<QMAX, R2, PAIRS>;
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    EDK_LAUNCH(kern, items, SEP_THREADS, bytes, s, P, T);
    return cudaGetLastError();
}
template <int QMAX, int R2>
cudaError_t launch_gram_sep_q(const SepParams& P, const SepTma& T, int pairs, int bytes, unsigned items, cudaStream_t s) {
    if (P.nmodes != SepModes<QMAX, R2>::N) return cudaErrorInvalidValue;
    if (pairs == 8) return launch_gram_sep_p<QMAX, R2, 8>(P, T, bytes, items, s);
    if (pairs == 6) return launch_gram_sep_p<QMAX, R2, 6>(P, T, bytes, items, s);
    return launch_gram_sep_p<QMAX, R2, 4>(P, T, bytes, items, s);
}
template <int QMAX, int R2, int PAIRS>
cudaError_t launch_gram_sepx_p(const SepParams& P, const SepTmaX& T, const SepWeights& W, int shape, int bytes, unsigned items,
                               cudaStream_t s) {
    auto kern = shape == 0 ? gram_sepx1_kernel<QMAX, R2, PAIRS, 4>
                           : (shape == 1 ? gram_sepx1_kernel<QMAX, R2, PAIRS, 1>
                                         : (shape == 2 ? gram_sepx1_kernel<QMAX, R2, PAIRS, 8> : gram_sepx_kernel<QMAX, R2, PAIRS>));
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    EDK_LAUNCH(kern, items, SepxGeom::THREADS, bytes, s, P, T, W);
    return cudaGetLastError();
}
template <int QMAX, int R2>
cudaError_t launch_gram_sepx_q(const SepParams& P, const SepTmaX& T, const SepWeights& W, int pairs, int shape, int bytes, unsigned items,
                               cudaStream_t s) {
    if (P.nmodes != SepModes<QMAX, R2>::N) return cudaErrorInvalidValue;
    if (pairs == 8) return launch_gram_sepx_p<QMAX, R2, 8>(P, T, W, shape, bytes, items, s);
    if (pairs == 6) return launch_gram_sepx_p<QMAX, R2, 6>(P, T, W, shape, bytes, items, s);
    return launch_gram_sepx_p<QMAX, R2, 4>(P, T, W, shape, bytes, items, s);
}
#endif  // EDK_EMU_NO_LAUNCHERS
// the three instantiated mode structures: (max |px|, |py|; max px^2 + py^2) = (1, 1), (1, 2), (2, 4)
cudaError_t launch_gram_sep_s11(const SepParams& P, const SepTma& T, int pairs, int bytes, unsigned items, cudaStream_t s);
cudaError_t launch_gram_sep_s12(const SepParams& P, const SepTma& T, int pairs, int bytes, unsigned items, cudaStream_t s);
cudaError_t launch_gram_sep_s24(const SepParams& P, const SepTma& T, int pairs, int bytes, unsigned items, cudaStream_t s);
cudaError_t launch_gram_sepx_s11(const SepParams& P, const SepTmaX& T, const SepWeights& W, int pairs, int shape, int bytes, unsigned items, cudaStream_t s);
cudaError_t launch_gram_sepx_s12(const SepParams& P, const SepTmaX& T, const SepWeights& W, int pairs, int shape, int bytes, unsigned items, cudaStream_t s);
cudaError_t launch_gram_sepx_s24(const SepParams& P, const SepTmaX& T, const SepWeights& W, int pairs, int shape, int bytes, unsigned items, cudaStream_t s);

}  // namespace edk
