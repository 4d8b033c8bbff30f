// Plane-wave factorised contraction ("algo 2"), sm_100a.
//
//   G[p][e][f] = sum_seg sign_seg sum_{x,c} conj(L_seg[e][x][c]) phase_p(x) R_seg[f][x][c]
//   (the einsum "zyx,ezyxc,fzyxc->ef" of lattice/generator/elemental.py:322-329 and
//    lattice/generator/displacement_elemental.py:94-95 for every momentum of the list)
//
// The GEMM form (edk_gram.cu) spends 8 Ne^2 3V flops per (pair, momentum) although the momentum only
// enters through a phase that factorises over the lattice axes:
//   phase_p(x,y,z) = [cos(theta_q(x,y)) + i sigma_p sin(theta_q(x,y))] * exp(2 pi i pz z / Lz),
//   theta_q = 2 pi (qx x/Lx + qy y/Ly),  (px,py) = sigma_p (qx,qy).
// Here the colour-summed site product
//   C(x)[e][f] = sum_c conj(L[e][x][c]) R[f][x][c]
// is formed once per site, whatever the number of momenta, and transformed over one xy-plane against
// the REAL mode functions w_m in {cos(theta_q), sin(theta_q)} (13 of them for the 33 momenta with
// |p|^2 <= 4) on the FP64 tensor pipe:
//   gram_pw_kernel<MB> : Y[job][z][m][e][f] = sum_{(x,y)} w_m(x,y) C(x,y,z)[e][f]
//                        one DMMA.8x8x4 = (8 modes) x (4 sites) x (8 f): A operand = weights, B operand =
//                        the C values the lanes have just formed with 12 DFMA each, accumulator = Y.
//                        Lane (site s = lane%4, column n = lane/4) owns EL e-rows {EL warp + i} and FL f-columns
//                        {n + 8 j} of an (8 EL) x (8 FL) tile (16 x 32 or 16 x 40); TMA producer warp + mbarrier
//                        ring as in gram_tma_kernel, same [k-group][row][4 complex] shared-memory tiles.
//   pw_zfold_kernel    : G[job][p] = sum_z exp(2 pi i pz z/Lz) (Y[z][mc(p)] + i sigma_p Y[z][ms(p)])
//                        written as split 0 of the partial-sum buffer, so combine_kernel is unchanged.
// FP64-pipe issue slots per (pair, e, f, site): 12/32 (DFMA) + 2 MB 8/32 (DMMA) = 1.4 at MB = 2, against
// 8.3 for the 3M GEMM form with the Hermitian pairing and the half set (563/19 momenta x 9/32).
// More than 16 modes run as several passes over the same fields (mb0 = first m-block of the pass).
// Self pairs (L == R) have a Hermitian site product: their tiles below the diagonal are skipped and the fold
// kernel reads the mirror element conjugated.  tests/test_pw_model.py is a lane-level numpy transcription of
// the index arithmetic below (tile layout, fragment ownership, plane-boundary weights, mirror reads).
#include <type_traits>

#include "edk_common.cuh"
#include "edk_pipe.cuh"

namespace edk {

constexpr int PW_THREADS = (PW_WARPS + 4) * 32;  // 8 MMA warps + one producer warpgroup (one warp of it works)
constexpr int PW_REGS_CONSUMER = 232;
constexpr int PW_REGS_PRODUCER = 40;
constexpr int PW_KG = 6;                                // k-groups of 4 complex per stage (= 8 sites x 3 colours)
constexpr int PW_W_BYTES = 2 * PW_MAX_MB * 256;         // [group of 4 sites][m-block][32 lanes] doubles
constexpr int PW_MAX_STAGES = 8;
constexpr int PW_TAIL_BYTES = 16 * PW_MAX_STAGES + (int)sizeof(GramJob) + 64;
__host__ __device__ constexpr int pw_stage_bytes(int el, int fl) { return PW_KG * (PW_WARPS * el + 8 * fl) * 64 + PW_W_BYTES; }

int pw_plan_smem(int el, int fl, int* nstages, int* smem_bytes) {
    const int stage = pw_stage_bytes(el, fl);
    int nst = (227 * 1024 - PW_TAIL_BYTES) / stage;
    if (nst > PW_MAX_STAGES) nst = PW_MAX_STAGES;
    if (nst < 2) return -1;
    *nstages = nst;
    *smem_bytes = nst * stage + PW_TAIL_BYTES;
    return 0;
}

// Tile shape with the least FP64-pipe time for this Ne (ties: the first = widest tile, fewest CTAs, least operand
// traffic).  f-blocks past Ne are skipped by the kernels, so only the e side pads: an e-tile costs two warp slots per
// SM sub-partition (8 MMA warps on 4 sub-partitions), the last one a single slot if at most 4 of its warps own real
// rows.  (1, 7) = 8 x 56 halves the row granularity; it reads 1.6x the operand bytes per site product, so it has to
// win by more than 5 %.
static const int kPwTiles[][3] = {{2, 5, 100}, {2, 4, 100}, {1, 7, 105}};
void pw_pick_tile(int Ne, bool folded, int* el, int* fl) {
    long long best = -1;
    for (const auto& t : kPwTiles) {
        const int rl = PW_WARPS * t[0];
        const int n_et = (Ne + rl - 1) / rl;
        const int active_last = (Ne - (n_et - 1) * rl + t[0] - 1) / t[0];  // warps of the last e-tile with real rows
        const long long cost = (long long)((n_et - 1) * 2 + (active_last > 4 ? 2 : 1)) * t[0] * t[2];
        // the folded 16 x 40 instance pays for its register budget with program-ordered operand loads: at equal cost
        // (16 x 32 and 16 x 40 always tie now that f-blocks past Ne are skipped) form 3 takes 16 x 32
        const bool take = best < 0 || cost < best || (folded && cost == best && *el == 2 && *fl == 5 && t[0] == 2 && t[1] == 4);
        if (take) {
            best = cost;
            *el = t[0];
            *fl = t[1];
        }
    }
}

bool pw_tile_available(int el, int fl) {
    for (const auto& t : kPwTiles)
        if (t[0] == el && t[1] == fl) return true;
    return false;
}

template <int MB, int PW_EL, int PW_FL>
__global__ void __launch_bounds__(PW_THREADS, 1) gram_pw_kernel(const PwParams P, const __grid_constant__ PwTma Tm) {
    constexpr int PW_ROWS_L = PW_WARPS * PW_EL, PW_ROWS_R = 8 * PW_FL;
    constexpr int PW_L_BYTES = PW_KG * PW_ROWS_L * 64;  // [kg][row][4 complex]
    constexpr int PW_R_BYTES = PW_KG * PW_ROWS_R * 64;
    constexpr int PW_STAGE_BYTES = pw_stage_bytes(PW_EL, PW_FL);
    static_assert(PW_STAGE_BYTES % 128 == 0 && PW_L_BYTES % 128 == 0 && (PW_ROWS_L * 64) % 128 == 0 && (PW_ROWS_R * 64) % 128 == 0,
                  "TMA destinations are 128-byte aligned");
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nst = Tm.nstages;
    unsigned char* tail = smem + (size_t)nst * PW_STAGE_BYTES;
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail);
    const uint32_t bar_empty = bar_full + 8 * PW_MAX_STAGES;
    GramJob* sjob = reinterpret_cast<GramJob*>(tail + 16 * PW_MAX_STAGES);

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
    const int e0 = et * PW_ROWS_L, f0 = ft * PW_ROWS_R;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, PW_WARPS);
        }
        mbar_init_fence();
    }
    __syncthreads();
    // self pair (L == R): the site product is Hermitian in (e, f) and the weights are real, so
    // Y[m][e][f] = conj(Y[m][f][e]); tiles entirely below the diagonal are left to the fold kernel's mirror read.
    // (No early return here: an exit ahead of setmaxnreg makes ptxas spill the accumulators in the stage loop.)
    const bool skip_tile = sjob->nseg == 1 && sjob->Lf[0] == sjob->Rf[0] && e0 > f0 + PW_ROWS_R - 1;
    const int T = skip_tile ? 0 : sjob->nseg * P.kplane;  // stages of 8 sites: every segment walks the plane once
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    if (warp >= PW_WARPS) {
        // ================================ producer warpgroup ================================
        warpgroup_reg_dealloc<PW_REGS_PRODUCER>();
        if (warp != PW_WARPS) return;
        const uint32_t tx_bytes = (uint32_t)(PW_L_BYTES + PW_R_BYTES + 2 * MB * 256);
        const int plane_site0 = z * P.A;
        int seg = 0, kstep = 0, s = 0;
        uint32_t par = 1;  // the first pass over the ring finds every slot free
        for (int it = 0; it < T; ++it) {
            mbar_wait(bar_empty + 8 * s, par);
            const uint32_t full = bar_full + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(full, tx_bytes);
            __syncwarp();
            const uint32_t st = smem_base + (uint32_t)(s * PW_STAGE_BYTES);
            // first double of the stage's 8 sites; sites past the end of the plane belong to the next
            // plane (or are zero-filled past the end of the row) and meet zero weights
            const int kd = (plane_site0 + 8 * kstep) * 6;
            if (lane < PW_KG) {
                tma_load_3d(st + lane * (PW_ROWS_L * 64), Tm.mapL, full, kd + 8 * lane, e0, sjob->Lf[seg]);
            } else if (lane >= 8 && lane < 8 + PW_KG) {
                tma_load_3d(st + PW_L_BYTES + (lane - 8) * (PW_ROWS_R * 64), Tm.mapR, full, kd + 8 * (lane - 8), f0,
                            sjob->Rf[seg]);
            } else if (lane == 16 || lane == 17) {
                const int grp = lane - 16;
                const double* src = P.wtiles + (((size_t)kstep * 2 + grp) * P.mbtot + P.mb0) * 32;
                bulk_load(st + PW_L_BYTES + PW_R_BYTES + grp * (MB * 256), src, MB * 256, full);
            }
            if (++kstep == P.kplane) {
                kstep = 0;
                ++seg;
            }
            if (++s == nst) {
                s = 0;
                par ^= 1;
            }
        }
        return;
    }

    // ================================== consumer warps ==================================
    warpgroup_reg_alloc<PW_REGS_CONSUMER>();
    const int sidx = lane & 3, n = lane >> 2;  // site inside a group of 4 (MMA k), column (MMA n)
    // byte offsets of this lane's three colours of site 4 grp + sidx inside the L and R tiles
    uint32_t offL[2][3], offR[2][3];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int kc = (4 * grp + sidx) * 3 + c;  // complex k inside the stage
            const int kg = kc >> 2, kin = kc & 3;
            offL[grp][c] = (uint32_t)(kg * (PW_ROWS_L * 64) + warp * PW_EL * 64 + kin * 16);
            offR[grp][c] = (uint32_t)(PW_L_BYTES + kg * (PW_ROWS_R * 64) + n * 64 + kin * 16);
        }
    const uint32_t offW = (uint32_t)(PW_L_BYTES + PW_R_BYTES + lane * 8);
    // padded work of edge tiles is skipped where whole warps / whole f-blocks are padding: a warp whose rows are all
    // past Ne only keeps the ring moving, and f-blocks past Ne are not formed (both conditions are warp-uniform)
    const bool warp_active = e0 + warp * PW_EL < P.Ne;
    const int jmax = min(PW_FL, (P.Ne - f0 + 7) / 8);

    double yre[PW_EL][PW_FL][MB][2], yim[PW_EL][PW_FL][MB][2];
#pragma unroll
    for (int i = 0; i < PW_EL; ++i)
#pragma unroll
        for (int j = 0; j < PW_FL; ++j)
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) yre[i][j][mb][0] = yre[i][j][mb][1] = yim[i][j][mb][0] = yim[i][j][mb][1] = 0.0;

    int cur_sign = 1;
    int cs_seg = 0, cs_kstep = 0;
    int s = 0;
    uint32_t par = 0;
    for (int it = 0; it < T; ++it) {
        const int sgn = sjob->sign[cs_seg];
        if (++cs_kstep == P.kplane) {
            cs_kstep = 0;
            ++cs_seg;
        }
        if (sgn != cur_sign) {  // uniform: fold the segment sign by flipping the running sums
#pragma unroll
            for (int i = 0; i < PW_EL; ++i)
#pragma unroll
                for (int j = 0; j < PW_FL; ++j)
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            yre[i][j][mb][q] = flip_sign(yre[i][j][mb][q]);
                            yim[i][j][mb][q] = flip_sign(yim[i][j][mb][q]);
                        }
            cur_sign = sgn;
        }
        mbar_wait(bar_full + 8 * s, par);
        const unsigned char* stage = smem + (size_t)s * PW_STAGE_BYTES;
        // full tiles run the unguarded body (one basic block per stage: DFMAs, DMMAs and loads of different f-blocks
        // interleave freely); only edge tiles pay for the per-f-block guard
        auto stage_body = [&](auto guarded) {
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
            double w[MB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) w[mb] = *reinterpret_cast<const double*>(stage + offW + (grp * MB + mb) * 256);
            cplx lv[PW_EL][3];
#pragma unroll
            for (int i = 0; i < PW_EL; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) lv[i][c] = *reinterpret_cast<const cplx*>(stage + offL[grp][c] + i * 64);
#pragma unroll
            for (int j = 0; j < PW_FL; ++j) {
                if (decltype(guarded)::value && j >= jmax) break;
                cplx rv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) rv[c] = *reinterpret_cast<const cplx*>(stage + offR[grp][c] + j * 512);
#pragma unroll
                for (int i = 0; i < PW_EL; ++i) {
                    // conj(L) . R over the three colours of this lane's site
                    double cr = lv[i][0].x * rv[0].x;
                    double ci = lv[i][0].x * rv[0].y;
                    cr = fma(lv[i][0].y, rv[0].y, cr);
                    ci = fma(-lv[i][0].y, rv[0].x, ci);
#pragma unroll
                    for (int c = 1; c < 3; ++c) {
                        cr = fma(lv[i][c].x, rv[c].x, cr);
                        ci = fma(lv[i][c].x, rv[c].y, ci);
                        cr = fma(lv[i][c].y, rv[c].y, cr);
                        ci = fma(-lv[i][c].y, rv[c].x, ci);
                    }
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
                        dmma884(yre[i][j][mb][0], yre[i][j][mb][1], w[mb], cr);
                        dmma884(yim[i][j][mb][0], yim[i][j][mb][1], w[mb], ci);
                    }
                }
            }
        }
        };
        if (warp_active) {
            if (jmax == PW_FL)
                stage_body(std::false_type{});
            else
                stage_body(std::true_type{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == nst) {
            s = 0;
            par ^= 1;
        }
    }

    // ---- epilogue: lane holds modes (mb0 + mb) 8 + n of columns f = f0 + 8 j + 2 sidx + {0, 1} -------------
    if (skip_tile) return;
    const double fs = (double)cur_sign;
    const int Ne = P.Ne;
    const size_t mat = (size_t)Ne * Ne;
    cplx* Yp = P.Y + ((size_t)job_id * P.Lz + z) * (size_t)P.nmodes * mat;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int mode = (P.mb0 + mb) * 8 + n;
        if (mode >= P.nmodes) continue;
#pragma unroll
        for (int i = 0; i < PW_EL; ++i) {
            const int e = e0 + warp * PW_EL + i;
            if (e >= Ne) continue;
            cplx* row = Yp + (size_t)mode * mat + (size_t)e * Ne;
#pragma unroll
            for (int j = 0; j < PW_FL; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int f = f0 + 8 * j + 2 * sidx + q;
                    if (f < Ne) row[f] = make_double2(fs * yre[i][j][mb][q], fs * yim[i][j][mb][q]);
                }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// Folded variant ("algo 3"): centre-symmetric site pairs.
// With the modes taken about the centre of the plane, w^c_q(s) = cos(theta_q(s) - delta_q) is even and
// w^s_q(s) = sin(theta_q(s) - delta_q) is odd under s -> A-1-s (the site (Lx-1-x, Ly-1-y)); the constant phase
// exp(i sigma delta_q) of a momentum goes into the z-phase table.  So for the pair (s, sbar = A-1-s)
//   sum over both sites of w^c C = w^c(s) (C(s) + C(sbar)),   of w^s C = w^s(s) (C(s) - C(sbar)):
// a lane forms the site products of BOTH sites of its pair (24 DFMA), their sum and difference (4 DADD) and feeds the
// sum to one block of <= 8 cos-modes and the difference to one block of <= 8 sin-modes - half the DMMAs per site.
// A stage = 8 pairs: the 8 front sites 8k..8k+7 and the 8 back sites A-8-8k..A-1-8k of the plane (two sets of TMA
// boxes; sites of the last stage that belong to no pair meet zero weights).  The host only selects this kernel for
// planes of at least 8 sites, so a back run never starts before its plane.  For an odd
// plane the middle site is its own partner and carries half the cos weight.  More than 8 {+q,-q} couples run as
// passes.  The operands of two sites are live at once, which fits the register budget with 16 x 32 tiles only.
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pwf_stage_bytes(int el, int fl) { return 2 * PW_KG * (PW_WARPS * el + 8 * fl) * 64 + 2 * 2 * 256; }

int pwf_plan_smem(int el, int fl, int* nstages, int* smem_bytes) {
    const int stage = pwf_stage_bytes(el, fl);
    int nst = (227 * 1024 - PW_TAIL_BYTES) / stage;
    if (nst > PW_MAX_STAGES) nst = PW_MAX_STAGES;
    if (nst < 2) return -1;
    *nstages = nst;
    *smem_bytes = nst * stage + PW_TAIL_BYTES;
    return 0;
}

template <int PW_EL, int PW_FL>
__global__ void __launch_bounds__(PW_THREADS, 1) gram_pwf_kernel(const PwParams P, const __grid_constant__ PwTma Tm) {
    constexpr int PW_ROWS_L = PW_WARPS * PW_EL, PW_ROWS_R = 8 * PW_FL;
    constexpr int PW_L_BYTES = PW_KG * PW_ROWS_L * 64;  // [kg][row][4 complex], one set of 8 sites
    constexpr int PW_R_BYTES = PW_KG * PW_ROWS_R * 64;
    constexpr int PW_HALF_BYTES = PW_L_BYTES + PW_R_BYTES;  // front set, then back set, then the weights
    constexpr int PW_STAGE_BYTES = pwf_stage_bytes(PW_EL, PW_FL);
    static_assert(PW_STAGE_BYTES % 128 == 0 && PW_HALF_BYTES % 128 == 0 && PW_L_BYTES % 128 == 0 && (PW_ROWS_L * 64) % 128 == 0 &&
                      (PW_ROWS_R * 64) % 128 == 0,
                  "TMA destinations are 128-byte aligned");
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nst = Tm.nstages;
    unsigned char* tail = smem + (size_t)nst * PW_STAGE_BYTES;
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail);
    const uint32_t bar_empty = bar_full + 8 * PW_MAX_STAGES;
    GramJob* sjob = reinterpret_cast<GramJob*>(tail + 16 * PW_MAX_STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    int item = blockIdx.x;
    const int ft = item % P.n_ft;
    item /= P.n_ft;
    const int et = item % P.n_et;
    item /= P.n_et;
    const int z = item % P.Lz;
    const int job_id = item / P.Lz;
    const int e0 = et * PW_ROWS_L, f0 = ft * PW_ROWS_R;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, PW_WARPS);
        }
        mbar_init_fence();
    }
    __syncthreads();
    const bool skip_tile = sjob->nseg == 1 && sjob->Lf[0] == sjob->Rf[0] && e0 > f0 + PW_ROWS_R - 1;
    const int T = skip_tile ? 0 : sjob->nseg * P.kplane;  // stages of 8 pairs
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    if (warp >= PW_WARPS) {
        // ================================ producer warpgroup ================================
        warpgroup_reg_dealloc<PW_REGS_PRODUCER>();
        if (warp != PW_WARPS) return;
        const uint32_t tx_bytes = (uint32_t)(2 * PW_HALF_BYTES + 2 * 2 * 256);
        const int plane_site0 = z * P.A;
        int seg = 0, kstep = 0, s = 0;
        uint32_t par = 1;
        for (int it = 0; it < T; ++it) {
            mbar_wait(bar_empty + 8 * s, par);
            const uint32_t full = bar_full + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(full, tx_bytes);
            __syncwarp();
            const uint32_t st = smem_base + (uint32_t)(s * PW_STAGE_BYTES);
            // lanes 0-5 / 8-13: L and R boxes of the front sites, lanes 16-21 / 24-29: of the back sites
            const int half = lane >> 4, sub = lane & 15;
            const int site0 = half ? plane_site0 + P.A - 8 - 8 * kstep : plane_site0 + 8 * kstep;
            const int kd = site0 * 6;
            const uint32_t hb = st + (uint32_t)(half * PW_HALF_BYTES);
            if (sub < PW_KG) {
                tma_load_3d(hb + sub * (PW_ROWS_L * 64), Tm.mapL, full, kd + 8 * sub, e0, sjob->Lf[seg]);
            } else if (sub >= 8 && sub < 8 + PW_KG) {
                tma_load_3d(hb + PW_L_BYTES + (sub - 8) * (PW_ROWS_R * 64), Tm.mapR, full, kd + 8 * (sub - 8), f0, sjob->Rf[seg]);
            } else if (sub == 6) {  // lanes 6 and 22: the weights of one group of 4 pairs (cos block, sin block)
                const double* src = P.wtiles + (((size_t)kstep * 2 + half) * P.mbtot + P.mb0) * 32;
                bulk_load(st + 2 * PW_HALF_BYTES + half * 512, src, 512, full);
            }
            if (++kstep == P.kplane) {
                kstep = 0;
                ++seg;
            }
            if (++s == nst) {
                s = 0;
                par ^= 1;
            }
        }
        return;
    }

    // ================================== consumer warps ==================================
    warpgroup_reg_alloc<PW_REGS_CONSUMER>();
    const int sidx = lane & 3, n = lane >> 2;  // pair inside a group of 4 (MMA k), column (MMA n)
    const uint32_t offW = (uint32_t)(2 * PW_HALF_BYTES + lane * 8);
    const bool warp_active = e0 + warp * PW_EL < P.Ne;  // see gram_pw_kernel: padding-only warps and f-blocks are skipped
    const int jmax = min(PW_FL, (P.Ne - f0 + 7) / 8);

    double yc_re[PW_EL][PW_FL][2], yc_im[PW_EL][PW_FL][2], ys_re[PW_EL][PW_FL][2], ys_im[PW_EL][PW_FL][2];
#pragma unroll
    for (int i = 0; i < PW_EL; ++i)
#pragma unroll
        for (int j = 0; j < PW_FL; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) yc_re[i][j][q] = yc_im[i][j][q] = ys_re[i][j][q] = ys_im[i][j][q] = 0.0;

    int cur_sign = 1;
    int cs_seg = 0, cs_kstep = 0;
    int s = 0;
    uint32_t par = 0;
    for (int it = 0; it < T; ++it) {
        const int sgn = sjob->sign[cs_seg];
        if (++cs_kstep == P.kplane) {
            cs_kstep = 0;
            ++cs_seg;
        }
        if (sgn != cur_sign) {
#pragma unroll
            for (int i = 0; i < PW_EL; ++i)
#pragma unroll
                for (int j = 0; j < PW_FL; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        yc_re[i][j][q] = flip_sign(yc_re[i][j][q]);
                        yc_im[i][j][q] = flip_sign(yc_im[i][j][q]);
                        ys_re[i][j][q] = flip_sign(ys_re[i][j][q]);
                        ys_im[i][j][q] = flip_sign(ys_im[i][j][q]);
                    }
            cur_sign = sgn;
        }
        mbar_wait(bar_full + 8 * s, par);
        const unsigned char* stage = smem + (size_t)s * PW_STAGE_BYTES;
        // full tiles run the unguarded body (one basic block per stage: DFMAs, DMMAs and loads of different f-blocks
        // interleave freely); only edge tiles pay for the per-f-block guard
        auto stage_body = [&](auto guarded) {
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
            const double wc = *reinterpret_cast<const double*>(stage + offW + grp * 512);
            const double ws = *reinterpret_cast<const double*>(stage + offW + grp * 512 + 256);
            // local site of this lane's pair in the front set, and of its partner in the back set (stored ascending)
            const int tf = 4 * grp + sidx, tb = 7 - tf;
            uint32_t oLf[3], oLb[3], oRf[3], oRb[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int kf = tf * 3 + c, kb = tb * 3 + c;
                oLf[c] = (uint32_t)((kf >> 2) * (PW_ROWS_L * 64) + warp * PW_EL * 64 + (kf & 3) * 16);
                oLb[c] = (uint32_t)(PW_HALF_BYTES + (kb >> 2) * (PW_ROWS_L * 64) + warp * PW_EL * 64 + (kb & 3) * 16);
                oRf[c] = (uint32_t)(PW_L_BYTES + (kf >> 2) * (PW_ROWS_R * 64) + n * 64 + (kf & 3) * 16);
                oRb[c] = (uint32_t)(PW_HALF_BYTES + PW_L_BYTES + (kb >> 2) * (PW_ROWS_R * 64) + n * 64 + (kb & 3) * 16);
            }
#pragma unroll
            for (int j = 0; j < PW_FL; ++j) {
                if (decltype(guarded)::value && j >= jmax) break;
                // front site first, then its partner: only one site's R fragment is live at a time.
                // 16 x 40 tiles: the L fragments are re-read for every f-block (broadcast loads, one wavefront each)
                // instead of being kept across the j loop, which is what keeps that instance inside 232 registers.
                auto site_products = [&](const uint32_t (&oL)[3], const uint32_t (&oR)[3], double (&pr)[PW_EL], double (&pi)[PW_EL]) {
                    cplx r[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        r[c] = (PW_EL * PW_FL > 8) ? lds128_again(stage + oR[c] + j * 512)
                                                   : *reinterpret_cast<const cplx*>(stage + oR[c] + j * 512);
#pragma unroll
                    for (int i = 0; i < PW_EL; ++i) {
                        cplx l[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            l[c] = (PW_EL * PW_FL > 8) ? lds128_again(stage + oL[c] + i * 64)
                                                       : *reinterpret_cast<const cplx*>(stage + oL[c] + i * 64);
                        // conj(L) . R over the three colours of the site
                        double cr = l[0].x * r[0].x, ci = l[0].x * r[0].y;
                        cr = fma(l[0].y, r[0].y, cr);
                        ci = fma(-l[0].y, r[0].x, ci);
#pragma unroll
                        for (int c = 1; c < 3; ++c) {
                            cr = fma(l[c].x, r[c].x, cr);
                            ci = fma(l[c].x, r[c].y, ci);
                            cr = fma(l[c].y, r[c].y, cr);
                            ci = fma(-l[c].y, r[c].x, ci);
                        }
                        pr[i] = cr;
                        pi[i] = ci;
                    }
                };
                double fr[PW_EL], fi[PW_EL], br[PW_EL], bi[PW_EL];
                site_products(oLf, oRf, fr, fi);
                site_products(oLb, oRb, br, bi);
#pragma unroll
                for (int i = 0; i < PW_EL; ++i) {
                    dmma884(yc_re[i][j][0], yc_re[i][j][1], wc, fr[i] + br[i]);
                    dmma884(yc_im[i][j][0], yc_im[i][j][1], wc, fi[i] + bi[i]);
                    dmma884(ys_re[i][j][0], ys_re[i][j][1], ws, fr[i] - br[i]);
                    dmma884(ys_im[i][j][0], ys_im[i][j][1], ws, fi[i] - bi[i]);
                }
            }
        }
        };
        if (warp_active) {
            if (jmax == PW_FL)
                stage_body(std::false_type{});
            else
                stage_body(std::true_type{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == nst) {
            s = 0;
            par ^= 1;
        }
    }

    // ---- epilogue: lane holds rows n of the cos block and of the sin block, columns f0 + 8 j + 2 sidx + {0, 1} ----
    if (skip_tile) return;
    const double fs = (double)cur_sign;
    const int Ne = P.Ne;
    const size_t mat = (size_t)Ne * Ne;
    cplx* Yp = P.Y + ((size_t)job_id * P.Lz + z) * (size_t)P.nmodes * mat;
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
        const int mode = P.slotmode[(P.mb0 + blk) * 8 + n];  // compact mode index of this row, -1: unused
        if (mode < 0) continue;
#pragma unroll
        for (int i = 0; i < PW_EL; ++i) {
            const int e = e0 + warp * PW_EL + i;
            if (e >= Ne) continue;
            cplx* row = Yp + (size_t)mode * mat + (size_t)e * Ne;
#pragma unroll
            for (int j = 0; j < PW_FL; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int f = f0 + 8 * j + 2 * sidx + q;
                    if (f < Ne)
                        row[f] = blk ? make_double2(fs * ys_re[i][j][q], fs * ys_im[i][j][q])
                                     : make_double2(fs * yc_re[i][j][q], fs * yc_im[i][j][q]);
                }
        }
    }
}

#ifndef EDK_EMU_NO_LAUNCHERS
template <int MB, int EL, int FL>
static cudaError_t launch_gram_pw_t(const PwParams& P, const PwTma& T, int bytes, unsigned items, cudaStream_t s) {
    auto kern = gram_pw_kernel<MB, EL, FL>;
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    EDK_LAUNCH(kern, items, PW_THREADS, bytes, s, P, T);
    return cudaGetLastError();
}

cudaError_t launch_gram_pwf(const PwParams& P, const PwTma& T, int el, int fl, cudaStream_t s) {
    int nst, bytes;
    if (pwf_plan_smem(el, fl, &nst, &bytes) != 0 || T.nstages < 2 || T.nstages > nst || !P.slotmode || !pw_tile_available(el, fl)) return cudaErrorInvalidValue;
    const long long items = (long long)P.njobs * P.Lz * P.n_et * P.n_ft;
    if (items < 1 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
    auto kern = el == 1 ? gram_pwf_kernel<1, 7> : (fl == 4 ? gram_pwf_kernel<2, 4> : gram_pwf_kernel<2, 5>);
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    EDK_LAUNCH(kern, (unsigned)items, PW_THREADS, bytes, s, P, T);
    return cudaGetLastError();
}

cudaError_t launch_gram_pw(const PwParams& P, const PwTma& T, int MB, int el, int fl, cudaStream_t s) {
    int nst, bytes;
    if (pw_plan_smem(el, fl, &nst, &bytes) != 0 || T.nstages < 2 || T.nstages > nst || MB < 1 || MB > PW_MAX_MB) return cudaErrorInvalidValue;
    const long long items = (long long)P.njobs * P.Lz * P.n_et * P.n_ft;
    if (items < 1 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned n = (unsigned)items;
    if (el == 2 && fl == 4) return MB == 1 ? launch_gram_pw_t<1, 2, 4>(P, T, bytes, n, s) : launch_gram_pw_t<2, 2, 4>(P, T, bytes, n, s);
    if (el == 2 && fl == 5) return MB == 1 ? launch_gram_pw_t<1, 2, 5>(P, T, bytes, n, s) : launch_gram_pw_t<2, 2, 5>(P, T, bytes, n, s);
    if (el == 1 && fl == 7) return MB == 1 ? launch_gram_pw_t<1, 1, 7>(P, T, bytes, n, s) : launch_gram_pw_t<2, 1, 7>(P, T, bytes, n, s);
    return cudaErrorInvalidValue;
}
#endif  // EDK_EMU_NO_LAUNCHERS

// weights of the real xy-modes, laid out as the A fragments of the plane transform:
//   wtiles[kstep][grp][m-block][lane] = w_{8 mblock + lane/4}(site 8 kstep + 4 grp + lane%4 of the plane)
// modes3[m] = (qx, qy, kind): kind 0 -> cos(theta_q), 1 -> sin(theta_q); zero outside the plane / mode list
__global__ void pw_weights_kernel(double* __restrict__ wtiles, const int* __restrict__ modes3, int nmodes, int mbtot, int kplane,
                                  Geom g) {
    const size_t total = (size_t)kplane * 2 * mbtot * 32;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int lane = (int)(idx & 31);
    size_t r = idx >> 5;
    const int mbt = (int)(r % mbtot);
    r /= mbtot;
    const int grp = (int)(r & 1);
    const int kstep = (int)(r >> 1);
    const int mode = mbt * 8 + (lane >> 2);
    const int sxy = 8 * kstep + 4 * grp + (lane & 3);
    double v = 0.0;
    if (mode < nmodes && sxy < g.Lx * g.Ly) {
        const int x = sxy % g.Lx, y = sxy / g.Lx;
        const long long qx = modes3[3 * mode + 0], qy = modes3[3 * mode + 1];
        // q.x reduced mod L in integers, as in phase_table_kernel
        const long long rx = ((qx * x) % g.Lx + g.Lx) % g.Lx;
        const long long ry = ((qy * y) % g.Ly + g.Ly) % g.Ly;
        const double turns = (double)rx / (double)g.Lx + (double)ry / (double)g.Ly;
        double sn, cs;
        sincospi(2.0 * turns, &sn, &cs);
        v = modes3[3 * mode + 2] ? sn : cs;
    }
    wtiles[idx] = v;
}

// weights of the folded variant, about the centre of the plane:
//   wtiles[kstep][grp][block][lane] = w_{slotmode[block][lane/4]}(front site of pair 8 kstep + 4 grp + lane%4)
// with w = cos(theta_q - delta_q) or sin(theta_q - delta_q), theta_q - delta_q = pi (qx (2x - Lx + 1)/Lx + qy (2y - Ly + 1)/Ly);
// zero for unused rows and past the last pair; the middle site of an odd plane is its own partner: half the cos weight.
__global__ void pwf_weights_kernel(double* __restrict__ wtiles, const int* __restrict__ modes3, const int* __restrict__ slotmode,
                                   int mbtot, int kplane, Geom g) {
    const size_t total = (size_t)kplane * 2 * mbtot * 32;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int lane = (int)(idx & 31);
    size_t r = idx >> 5;
    const int mbt = (int)(r % mbtot);
    r /= mbtot;
    const int grp = (int)(r & 1);
    const int kstep = (int)(r >> 1);
    const int mode = slotmode[mbt * 8 + (lane >> 2)];
    const int A = g.Lx * g.Ly;
    const int pair = 8 * kstep + 4 * grp + (lane & 3);
    double v = 0.0;
    if (mode >= 0 && 2 * pair < A) {  // pairs 0 .. ceil(A/2) - 1
        const int x = pair % g.Lx, y = pair / g.Lx;
        const long long qx = modes3[3 * mode + 0], qy = modes3[3 * mode + 1];
        const long long mx = 2LL * g.Lx, my = 2LL * g.Ly;
        const long long rx = ((qx * (2 * x - g.Lx + 1)) % mx + mx) % mx;  // in units of pi / Lx
        const long long ry = ((qy * (2 * y - g.Ly + 1)) % my + my) % my;
        double sn, cs;
        sincospi((double)rx / (double)g.Lx + (double)ry / (double)g.Ly, &sn, &cs);
        v = modes3[3 * mode + 2] ? sn : cs;
        if (2 * pair == A - 1) v *= 0.5;
    }
    wtiles[idx] = v;
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_pwf_weights(double* wtiles, const int* modes3_dev, const int* slotmode_dev, int mbtot, int kplane, Geom g,
                               cudaStream_t s) {
    const size_t total = (size_t)kplane * 2 * mbtot * 32;
    EDK_LAUNCH(pwf_weights_kernel, (unsigned)((total + 255) / 256), 256, 0, s, wtiles, modes3_dev, slotmode_dev, mbtot, kplane, g);
    return cudaGetLastError();
}

cudaError_t launch_pw_weights(double* wtiles, const int* modes3_dev, int nmodes, int mbtot, int kplane, Geom g, cudaStream_t s) {
    const size_t total = (size_t)kplane * 2 * mbtot * 32;
    EDK_LAUNCH(pw_weights_kernel, (unsigned)((total + 255) / 256), 256, 0, s, wtiles, modes3_dev, nmodes, mbtot, kplane, g);
    return cudaGetLastError();
}

#endif  // EDK_EMU_NO_LAUNCHERS

// G[job][p][e][f] = sum_z zphase[p][z] (Y[job][z][mc][e][f] + i sigma_p Y[job][z][ms][e][f]).
// One block folds one {+q, -q} couple of one (job, 256 matrix elements): the couple's two planes of Y are read once
// per z and feed all its momenta (up to 8 accumulators at a time; 5-6 momenta per couple for the |p|^2 <= 4 set), so Y
// is read once instead of once per momentum.  Blocks are indexed by mode; those of a sin mode have nothing to do.
constexpr int PW_FOLD_THREADS = 256;
constexpr int PW_FOLD_MAXM = 8;
__global__ void __launch_bounds__(PW_FOLD_THREADS) pw_zfold_kernel(const PwFold F) {
    const size_t mat = (size_t)F.Ne * F.Ne;
    const int nblk = (int)((mat + PW_FOLD_THREADS - 1) / PW_FOLD_THREADS);
    int b = blockIdx.x;
    const int mc = b % F.nmodes;
    b /= F.nmodes;
    const int blk = b % nblk;
    const int job = b / nblk;
    const GramJob& J = F.jobs[job];
    const size_t ef = (size_t)blk * PW_FOLD_THREADS + threadIdx.x;
    if (ef >= mat) return;
    // tiles of a self pair below the diagonal were not computed: read the mirror element, conjugated
    const int e = (int)(ef / F.Ne), f = (int)(ef - (size_t)e * F.Ne);
    const bool mirror = J.nseg == 1 && J.Lf[0] == J.Rf[0] && (e / F.rows_l) * F.rows_l > (f / F.rows_r) * F.rows_r + F.rows_r - 1;
    const double cj = mirror ? -1.0 : 1.0;
    const cplx* Yj = F.Y + (size_t)job * F.Lz * F.nmodes * mat + (mirror ? (size_t)f * F.Ne + e : ef);
    // momenta of this couple among the first J.nmom (self pairs contract the half set only), 8 at a time
    for (int base = 0; base < J.nmom;) {
        int pl[PW_FOLD_MAXM];
        double sg[PW_FOLD_MAXM];
        int cnt = 0, ms = -1, p = base;
        for (; p < J.nmom && cnt < PW_FOLD_MAXM; ++p) {
            if (F.momode[3 * p] != mc) continue;
            ms = F.momode[3 * p + 1];
#pragma unroll
            for (int k = 0; k < PW_FOLD_MAXM; ++k)
                if (k == cnt) {
                    pl[k] = p;
                    sg[k] = (double)F.momode[3 * p + 2];
                }
            ++cnt;
        }
        base = p;
        if (cnt == 0) continue;
        double ar[PW_FOLD_MAXM], ai[PW_FOLD_MAXM];
#pragma unroll
        for (int k = 0; k < PW_FOLD_MAXM; ++k) ar[k] = ai[k] = 0.0;
        for (int z = 0; z < F.Lz; ++z) {
            const cplx yc = Yj[((size_t)z * F.nmodes + mc) * mat];
            const double cr = yc.x, ci = cj * yc.y;
            double sr = 0.0, si = 0.0;  // i * (sin plane), up to the sign sigma_p
            if (ms >= 0) {
                const cplx ys = Yj[((size_t)z * F.nmodes + ms) * mat];
                sr = -cj * ys.y;
                si = ys.x;
            }
#pragma unroll
            for (int k = 0; k < PW_FOLD_MAXM; ++k) {
                if (k < cnt) {
                    const double ur = fma(sg[k], sr, cr), ui = fma(sg[k], si, ci);
                    const cplx ph = F.zphase[(size_t)pl[k] * F.Lz + z];
                    ar[k] = fma(ph.x, ur, ar[k]);
                    ar[k] = fma(-ph.y, ui, ar[k]);
                    ai[k] = fma(ph.x, ui, ai[k]);
                    ai[k] = fma(ph.y, ur, ai[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PW_FOLD_MAXM; ++k)
            if (k < cnt) F.partial[((size_t)job * F.nmom_int + pl[k]) * mat + ef] = make_double2(ar[k], ai[k]);
    }
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_pw_zfold(const PwFold& F, cudaStream_t s) {
    const size_t mat = (size_t)F.Ne * F.Ne;
    const long long nblk = (long long)((mat + PW_FOLD_THREADS - 1) / PW_FOLD_THREADS);
    const long long blocks = nblk * F.njobs * F.nmodes;
    if (blocks < 1 || blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    EDK_LAUNCH(pw_zfold_kernel, (unsigned)blocks, PW_FOLD_THREADS, 0, s, F);
    return cudaGetLastError();
}

#endif  // EDK_EMU_NO_LAUNCHERS

}  // namespace edk
