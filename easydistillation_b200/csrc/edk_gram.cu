// Momentum-phased Ne x Ne contraction on the FP64 tensor pipe (DMMA.8x8x4), sm_100a.
//
//   G[p][e][f] = sum_seg sign_seg * sum_{x,c} conj(L_seg[e][x][c]) * phase_p(x) * R_seg[f][x][c]
//   (reference: the einsum "zyx,ezyxc,fzyxc->ef" of lattice/generator/elemental.py:322-329 and
//    lattice/generator/displacement_elemental.py:94-95, with the left/right split sum of
//    elemental.py:309-321 folded into job lists, see edk_api.cu)
//
// tcgen05 has no f64 kind, so the FP64 tensor path is mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4.
// Kernels in this file:
//   gram_tma_kernel<MF, ALGO>  product path: TMA producer warp + mbarrier full/empty ring, 8 MMA
//                              warps; ALGO 1 = 3M arithmetic (3 real MMAs per complex block),
//                              ALGO 0 = 4M (cross-check)
//   gram_dmma_kernel<MF>       earlier variant: cp.async ring issued by the MMA warps themselves, 4M
//                              (A/B cross-check, selected by edk_debug_loader)
//   gram_naive_kernel          one thread per output element (tests only)
//   combine_kernel             split-K sum, +-weights, Hermitian / half-set reads, blending matrix
//
// 4M mapping onto real m8n8k4 MMAs:
//   A operand  = rows e of L, k-slots hold Re L (first MMA) or Im L (second MMA) of 4 complex k,
//   B operand  = columns (f, re|im), k-slots hold the matching part of P = phase * R:
//                  column (f,re): (Pr | Pi)      column (f,im): (Pi | -Pr)
//   so that  C[e][(f,re)] += Lr*Pr + Li*Pi = Re(conj(L) P),  C[e][(f,im)] += Lr*Pi - Li*Pr = Im(conj(L) P)
//   and the accumulator fragment (2 doubles per lane) is one complex128 in natural (re, im) order.
//   The (f,im) columns read a pre-rotated table -i*phase, so there is no select or negation.
// 3M mapping: B columns are 8 complex f; T1 += Lr*Pr, T2 += Li*Pi, T3 += (Lr+Li)*(Pi-Pr);
//   Re = T1 + T2, Im = T3 + T1 - T2 in the epilogue.
// In both, the phase multiply is fused into the B-fragment build (a few FP64 ops per 2-3*MF MMAs).
//
// Shared memory per stage (8 sites = 24 complex k): A [6 k-groups][8*MF rows][4 complex] so a
// warp's fragment load is 512 contiguous bytes (bank-conflict free LDS.128), B likewise, plus the
// phase tile [n-fragment][8 sites].  All warps of a CTA share the 8*MF rows of L, so no B value is
// built twice; n-fragments (4 or 8 consecutive f at one momentum) are flattened f-fragment-major.
#include "edk_common.cuh"
#include "edk_pipe.cuh"

namespace edk {

constexpr int GRAM_NWARP = 8;
constexpr int GRAM_NTHREADS = GRAM_NWARP * 32;
constexpr int GRAM_NF = 2;                      // n-fragments per warp
constexpr int GRAM_NT = GRAM_NWARP * GRAM_NF;   // n-fragments per CTA
constexpr int GRAM_BROWS = 4 * GRAM_NT;         // max rows of R a CTA can need
constexpr int GRAM_KG = 6;                      // k-groups of 4 complex per stage (= 8 sites)
constexpr int GRAM_STAGES = 3;
constexpr int GRAM_MAX_MF = 13;

template <int MF>
struct GramSmem {
    static constexpr int ROWS_A = 8 * MF;
    static constexpr int A_BYTES = GRAM_KG * ROWS_A * 64;
    static constexpr int B_BYTES = GRAM_KG * GRAM_BROWS * 64;
    static constexpr int PH_BYTES = 2 * GRAM_NT * 8 * 16;  // phase and -i*phase
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + PH_BYTES;
    static constexpr int JOB_OFF = GRAM_STAGES * STAGE_BYTES;
    static constexpr int TOTAL = JOB_OFF + (int)sizeof(GramJob);
};

// One pipeline stage (8 sites = 6 k-groups of 4 complex k) of one warp: 2*MF*NF DMMAs per k-group.
//   a_s : this lane's slot in A[kg][row][4]  (+ (kg*ROWS_A + 8 i)*64 selects fragment i)
//   b_s : B[kg][row][4], b_kg_stride bytes per k-group; my_boff = this lane's (row, k) offset
//   p_s : this lane's phase tile [n-fragment][8 sites] (already offset to phase or -i*phase)
// MFL <= MF is the number of live m-fragments of this tile (row tiles are balanced, so the last
// tiles of a column may own one fragment less; its MMAs are simply not issued).
template <int MF, int MFL>
__device__ __forceinline__ void gram_compute_stage(double (&acc)[MF][GRAM_NF][2], const unsigned char* a_s,
                                                   const unsigned char* b_s, const int b_kg_stride,
                                                   const unsigned char* p_s, const uint32_t (&my_boff)[GRAM_NF],
                                                   const int kk) {
    constexpr int ROWS_A = 8 * MF;
    {
        // B fragments of k-group kg: P' = phase' * R with phase' = phase (re columns) or -i*phase
        // (im columns); b1 = Re P' pairs with Re L, b2 = Im P' pairs with Im L.
        cplx rr[GRAM_NF], pp[GRAM_NF];
        double b1[GRAM_NF], b2[GRAM_NF];
#pragma unroll
        for (int n = 0; n < GRAM_NF; ++n) {
            rr[n] = *reinterpret_cast<const cplx*>(b_s + my_boff[n]);
            pp[n] = *reinterpret_cast<const cplx*>(p_s + (n * 8 + kk / 3) * 16);
        }
#pragma unroll
        for (int n = 0; n < GRAM_NF; ++n) {
            b1[n] = fma(pp[n].x, rr[n].x, -(pp[n].y * rr[n].y));
            b2[n] = fma(pp[n].x, rr[n].y, pp[n].y * rr[n].x);
        }
#pragma unroll
        for (int kg = 0; kg < GRAM_KG; ++kg) {
            // raw operands of the next k-group are fetched now and multiplied in the shadow of the MMAs
            if (kg + 1 < GRAM_KG) {
#pragma unroll
                for (int n = 0; n < GRAM_NF; ++n) {
                    rr[n] = *reinterpret_cast<const cplx*>(b_s + (kg + 1) * (b_kg_stride) + my_boff[n]);
                    pp[n] = *reinterpret_cast<const cplx*>(p_s + (n * 8 + (4 * (kg + 1) + kk) / 3) * 16);
                }
            }
            double nb1[GRAM_NF], nb2[GRAM_NF];
            // m-fragments in groups of <= 5: both MMAs of an accumulator are >= 2*group MMAs apart
            constexpr int GRP = 5;
#pragma unroll
            for (int i0 = 0; i0 < MFL; i0 += GRP) {
                cplx a[GRP];
#pragma unroll
                for (int ii = 0; ii < GRP; ++ii)
                    if (i0 + ii < MFL) a[ii] = *reinterpret_cast<const cplx*>(a_s + (kg * ROWS_A + 8 * (i0 + ii)) * 64);
#pragma unroll
                for (int ii = 0; ii < GRP; ++ii)
                    if (i0 + ii < MFL) {
#pragma unroll
                        for (int n = 0; n < GRAM_NF; ++n) dmma884(acc[i0 + ii][n][0], acc[i0 + ii][n][1], a[ii].x, b1[n]);
                    }
                if (i0 == 0 && kg + 1 < GRAM_KG) {
#pragma unroll
                    for (int n = 0; n < GRAM_NF; ++n) {
                        nb1[n] = fma(pp[n].x, rr[n].x, -(pp[n].y * rr[n].y));
                        nb2[n] = fma(pp[n].x, rr[n].y, pp[n].y * rr[n].x);
                    }
                }
#pragma unroll
                for (int ii = 0; ii < GRP; ++ii)
                    if (i0 + ii < MFL) {
#pragma unroll
                        for (int n = 0; n < GRAM_NF; ++n) dmma884(acc[i0 + ii][n][0], acc[i0 + ii][n][1], a[ii].y, b2[n]);
                    }
            }
            if (kg + 1 < GRAM_KG) {
#pragma unroll
                for (int n = 0; n < GRAM_NF; ++n) {
                    b1[n] = nb1[n];
                    b2[n] = nb2[n];
                }
            }
        }
        }
}

// Balanced row tiling: ceil(Ne/8) m-fragments over n_mt tiles, the first (frags % n_mt) tiles own
// one fragment more.  Returns the first row of tile mt and its live fragment count.
__device__ __forceinline__ int gram_row_tile(int Ne, int n_mt, int mt, int& mf_live) {
    const int frags = (Ne + 7) >> 3;
    const int q = frags / n_mt, r = frags - q * n_mt;
    mf_live = q + (mt < r ? 1 : 0);
    return 8 * (mt * q + min(mt, r));
}

template <int MF>
__global__ void __launch_bounds__(GRAM_NTHREADS, 1) gram_dmma_kernel(const GramParams P) {
    using S = GramSmem<MF>;
    constexpr int ROWS_A = S::ROWS_A;
    constexpr int AJ = (ROWS_A + 31) / 32;  // loader passes over the rows of A
    extern __shared__ __align__(128) unsigned char smem[];
    GramJob* sjob = reinterpret_cast<GramJob*>(smem + S::JOB_OFF);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, kk = lane & 3;

    const int2 cta = P.cta_map[blockIdx.x];
    const int job_id = cta.x;
    const int split = blockIdx.y;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    __syncthreads();
    const int nseg = sjob->nseg;

    // ---- column bookkeeping -------------------------------------------------------------
    const int Ne = P.Ne, nmom = sjob->nmom;  // this job's momenta (a prefix of the internal list)
    const int nfrag_f = (Ne + 3) >> 2;
    const int N_flat = nfrag_f * nmom;
    const int n_nt = (N_flat + GRAM_NT - 1) / GRAM_NT;
    const int mt = cta.y / n_nt, nt = cta.y - mt * n_nt;
    const int nflat0 = nt * GRAM_NT;
    const int nflat_last = min(nflat0 + GRAM_NT, N_flat) - 1;
    const int ff0 = nflat0 / nmom;
    const int nrows_b = 4 * (nflat_last / nmom - ff0 + 1);
    int my_ffrag[GRAM_NF], my_p[GRAM_NF];
    bool my_valid[GRAM_NF];
    uint32_t my_boff[GRAM_NF];  // byte offset of this lane's R element inside one k-group of B
#pragma unroll
    for (int n = 0; n < GRAM_NF; ++n) {
        const int nf_raw = nflat0 + warp * GRAM_NF + n;
        my_valid[n] = nf_raw < N_flat;
        const int nf = min(nf_raw, N_flat - 1);
        my_ffrag[n] = nf / nmom;
        my_p[n] = nf - my_ffrag[n] * nmom;
        my_boff[n] = (uint32_t)(((4 * (my_ffrag[n] - ff0) + (g >> 1)) * 4 + kk) * 16);
    }
    // columns (f, im) of the B operand are the components of (-i * phase) * R: those lanes read
    // the second (pre-rotated) phase tile, so the fragment build has no select and no negation
    const uint32_t my_phoff = (uint32_t)(((g & 1) * GRAM_NT + warp * GRAM_NF) * 8 * 16);

    // ---- k range of this split --------------------------------------------------------------
    const int T_all = nseg * P.ksteps;
    const int T0 = (int)(((long long)T_all * split) / P.ksplit);
    const int T1 = (int)(((long long)T_all * (split + 1)) / P.ksplit);
    const int T = T1 - T0;

    // ---- loader state: everything that does not depend on the stage --------------------------
    // thread -> (row lrow + 32 j, 16-byte chunk lkq + 8 m) of a 24-chunk (8-site) row segment:
    // 8 consecutive threads fetch one full 128-byte line of a row.
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    int mf_live;
    const int row0 = gram_row_tile(Ne, P.n_mt, mt, mf_live);
    const int lrow = tid >> 3, lkq = tid & 7;
    const uint32_t a_dst0 = (uint32_t)(((lkq >> 2) * ROWS_A + lrow) * 64 + (lkq & 3) * 16);
    const uint32_t b_dst0 = (uint32_t)(S::A_BYTES + ((lkq >> 2) * GRAM_BROWS + lrow) * 64 + (lkq & 3) * 16);
    const size_t a_goff = (size_t)(row0 + lrow) * P.Kc + lkq;
    const size_t b_goff = (size_t)(4 * ff0 + lrow) * P.Kc + lkq;
    const size_t rstride = (size_t)32 * P.Kc;
    unsigned a_issue = 0, a_ok = 0, b_issue = 0, b_ok = 0;  // bit j: row pass j exists / is inside the matrix
#pragma unroll
    for (int j = 0; j < AJ; ++j) {
        if (lrow + 32 * j < ROWS_A) a_issue |= 1u << j;
        if (row0 + lrow + 32 * j < Ne) a_ok |= 1u << j;
    }
#pragma unroll
    for (int j = 0; j < GRAM_BROWS / 32; ++j) {
        if (lrow + 32 * j < nrows_b) b_issue |= 1u << j;
        if (4 * ff0 + lrow + 32 * j < Ne) b_ok |= 1u << j;
    }
    // phase tiles: threads 0..127 fetch phase, 128..255 fetch -i*phase, 16 n-fragments x 8 sites each
    const cplx* ph_src;
    {
        const int t2 = tid & (GRAM_NT * 8 - 1);
        const int nf = min(nflat0 + (t2 >> 3), N_flat - 1);
        ph_src = P.phase + ((size_t)(tid >> 7) * P.nmom + (nf % nmom)) * P.Vpad + (t2 & 7);
    }
    const uint32_t ph_dst0 = (uint32_t)(S::A_BYTES + S::B_BYTES + tid * 16);

    int ld_seg = T0 / P.ksteps;             // segment / k-step of the NEXT stage to be issued
    int ld_kstep = T0 - ld_seg * P.ksteps;
    auto issue_stage = [&](int buf) {
        const int kbase = ld_kstep * 24;
        const cplx* Lp = sjob->L[ld_seg];
        const cplx* Rp = sjob->R[ld_seg];
        const uint32_t st = smem_base + buf * S::STAGE_BYTES;
        bool kok[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) kok[m] = kbase + lkq + 8 * m < P.Kc;
        const cplx* ap = Lp + a_goff + kbase;
#pragma unroll
        for (int j = 0; j < AJ; ++j) {
            if (a_issue >> j & 1) {
                const bool rok = a_ok >> j & 1;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const bool ok = rok && kok[m];
                    cp_async16(st + a_dst0 + (uint32_t)(m * 2 * ROWS_A * 64 + j * 32 * 64), ok ? ap + j * rstride + 8 * m : Lp, ok);
                }
            }
        }
        const cplx* bp = Rp + b_goff + kbase;
#pragma unroll
        for (int j = 0; j < GRAM_BROWS / 32; ++j) {
            if (b_issue >> j & 1) {
                const bool rok = b_ok >> j & 1;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const bool ok = rok && kok[m];
                    cp_async16(st + b_dst0 + (uint32_t)(m * 2 * GRAM_BROWS * 64 + j * 32 * 64), ok ? bp + j * rstride + 8 * m : Rp, ok);
                }
            }
        }
        cp_async16(st + ph_dst0, ph_src + ld_kstep * 8, true);
        if (++ld_kstep == P.ksteps) {
            ld_kstep = 0;
            ++ld_seg;
        }
    };

    double acc[MF][GRAM_NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int n = 0; n < GRAM_NF; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;
    int cur_sign = 1;

#pragma unroll
    for (int s = 0; s < GRAM_STAGES - 1; ++s) {
        if (s < T) issue_stage(s);
        cp_async_commit();
    }

    int cs_seg = T0 / P.ksteps;             // segment / k-step of the stage being computed
    int cs_kstep = T0 - cs_seg * P.ksteps;
    int buf = 0;
    for (int it = 0; it < T; ++it) {
        cp_async_wait<GRAM_STAGES - 2>();
        __syncthreads();
        {
            int nb = buf + GRAM_STAGES - 1;
            if (nb >= GRAM_STAGES) nb -= GRAM_STAGES;
            if (it + GRAM_STAGES - 1 < T) issue_stage(nb);
            cp_async_commit();
        }
        const int sgn = sjob->sign[cs_seg];
        if (++cs_kstep == P.ksteps) {
            cs_kstep = 0;
            ++cs_seg;
        }
        if (sgn != cur_sign) {  // uniform: fold the segment sign by flipping the running sum
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int n = 0; n < GRAM_NF; ++n) {
                    acc[i][n][0] = flip_sign(acc[i][n][0]);
                    acc[i][n][1] = flip_sign(acc[i][n][1]);
                }
            cur_sign = sgn;
        }

        const unsigned char* stage = smem + buf * S::STAGE_BYTES;
        const unsigned char* a_s = stage + lane * 16;
        const unsigned char* b_s = stage + S::A_BYTES;
        const unsigned char* p_s = b_s + S::B_BYTES + my_phoff;
        if (++buf == GRAM_STAGES) buf = 0;

        gram_compute_stage<MF, MF>(acc, a_s, b_s, GRAM_BROWS * 64, p_s, my_boff, kk);
    }
    cp_async_wait<0>();

    // ---- epilogue: lane holds C[e = row0+8i+g][f = 4*ffrag+kk] = (re, im) ------------------------
    const double fs = (double)cur_sign;
    cplx* outj = P.partial + ((size_t)split * P.njobs + job_id) * (size_t)P.nmom * Ne * Ne;
#pragma unroll
    for (int n = 0; n < GRAM_NF; ++n) {
        const int f = 4 * my_ffrag[n] + kk;
        if (!my_valid[n] || f >= Ne) continue;
        cplx* outp = outj + (size_t)my_p[n] * Ne * Ne;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
            const int e = row0 + 8 * i + g;
            if (i < mf_live && e < Ne) outp[(size_t)e * Ne + f] = make_double2(fs * acc[i][n][0], fs * acc[i][n][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// TMA-fed, warp-specialised variant: warps 0..7 only wait / MMA / release, warp 8 is the producer.
//   producer : per stage 6 A boxes + 6*nbb B boxes (cp.async.bulk.tensor.3d over the field array
//              viewed as [field][row][2K doubles], box = 8 doubles x rows) and the phase tile
//              (cp.async.bulk runs of consecutive momenta), all completing on full[stage]
//   consumers: mbarrier.try_wait(full) -> gram_compute_stage -> arrive(empty)
// No CTA-wide barrier and no loader code in the MMA warps, so they drift apart instead of hitting
// the same bubbles in lock-step.  Out-of-range rows / k are zero-filled by the TMA unit.  The
// producer warpgroup gives its registers away (setmaxnreg), the MMA warps run with 232.
// ---------------------------------------------------------------------------------------
constexpr int GT_CONSUMERS = GRAM_NWARP;             // two consumer warpgroups
constexpr int GT_THREADS = (GT_CONSUMERS + 4) * 32;  // + one producer warpgroup (one warp of it works)
constexpr int GT_REGS_CONSUMER = 232;                // setmaxnreg: 256 x 232 + 128 x 40 <= 64 K registers
constexpr int GT_REGS_PRODUCER = 40;

// Column organisation of the two arithmetic variants.
//   ALGO 0 ("4M"): an n-fragment is 4 complex f, its 8 MMA columns are (f, re|im); 2 n-fragments per
//                  warp, 16 per CTA; phase tile holds phase and -i*phase.  4 real MMAs per complex block.
//   ALGO 1 ("3M"): an n-fragment is 8 complex f = the 8 MMA columns; 1 per warp, 8 per CTA; three
//                  accumulators T1 = Lr.Pr, T2 = Li.Pi, T3 = (Lr+Li).(Pi-Pr), combined in the epilogue as
//                  Re = T1 + T2, Im = T3 + T1 - T2 (Gauss / 3M).  3 real MMAs per complex block.
template <int ALGO>
struct GtCols {
    static constexpr int FW = ALGO ? 8 : 4;                 // complex f per n-fragment
    static constexpr int NF = ALGO ? 1 : 2;                 // n-fragments per warp
    static constexpr int NT = GT_CONSUMERS * NF;            // n-fragments per CTA
    static constexpr int NTAB = ALGO ? 1 : 2;               // phase tables staged per stage
    static constexpr int PH_BYTES = NTAB * NT * 8 * 16;
    static constexpr int NACC = ALGO ? 3 : 2;               // accumulator fragments per (m-fragment, warp)
    static constexpr int LS_PER_ROW = ALGO ? 32 : 0;        // bytes per (row, k-group) of the Re+Im plane
};

// 3M stage: same operand tiles as gram_compute_stage plus the plane Ls = Re L + Im L that the
// field-producing kernels wrote and the producer fetched as a third box (ls_s = this lane's slot
// in Ls[kg][row][4]); three MMAs per (m-fragment, k-group), no FP64 add in the MMA warps.
template <int MF, int MFL>
__device__ __forceinline__ void gram_compute_stage_3m(double (&acc)[MF][3][2], const unsigned char* a_s,
                                                      const unsigned char* ls_s, const unsigned char* b_s,
                                                      const int b_kg_stride, const unsigned char* p_s,
                                                      const uint32_t my_boff, const int kk) {
    constexpr int ROWS_A = 8 * MF;
    cplx rr = *reinterpret_cast<const cplx*>(b_s + my_boff);
    cplx pp = *reinterpret_cast<const cplx*>(p_s + (kk / 3) * 16);
    double pr = fma(pp.x, rr.x, -(pp.y * rr.y));
    double pi = fma(pp.x, rr.y, pp.y * rr.x);
    double pd = pi - pr;
#pragma unroll
    for (int kg = 0; kg < GRAM_KG; ++kg) {
        if (kg + 1 < GRAM_KG) {
            rr = *reinterpret_cast<const cplx*>(b_s + (kg + 1) * b_kg_stride + my_boff);
            pp = *reinterpret_cast<const cplx*>(p_s + ((4 * (kg + 1) + kk) / 3) * 16);
        }
        double npr = 0.0, npi = 0.0, npd = 0.0;
        constexpr int GRP = 4;
#pragma unroll
        for (int i0 = 0; i0 < MFL; i0 += GRP) {
            cplx a[GRP];
            double as[GRP];
#pragma unroll
            for (int ii = 0; ii < GRP; ++ii)
                if (i0 + ii < MFL) {
                    a[ii] = *reinterpret_cast<const cplx*>(a_s + (kg * ROWS_A + 8 * (i0 + ii)) * 64);
                    as[ii] = *reinterpret_cast<const double*>(ls_s + (kg * ROWS_A + 8 * (i0 + ii)) * 32);
                }
#pragma unroll
            for (int ii = 0; ii < GRP; ++ii)
                if (i0 + ii < MFL) {
                    dmma884(acc[i0 + ii][0][0], acc[i0 + ii][0][1], a[ii].x, pr);
                    dmma884(acc[i0 + ii][1][0], acc[i0 + ii][1][1], a[ii].y, pi);
                    dmma884(acc[i0 + ii][2][0], acc[i0 + ii][2][1], as[ii], pd);
                }
            if (i0 == 0 && kg + 1 < GRAM_KG) {
                npr = fma(pp.x, rr.x, -(pp.y * rr.y));
                npi = fma(pp.x, rr.y, pp.y * rr.x);
                npd = npi - npr;
            }
        }
        if (kg + 1 < GRAM_KG) {
            pr = npr;
            pi = npi;
            pd = npd;
        }
    }
}

template <int MF, int ALGO>
__global__ void __launch_bounds__(GT_THREADS, 1) gram_tma_kernel(const GramParams P, const __grid_constant__ GramTma Tm) {
    using C = GtCols<ALGO>;
    constexpr int ROWS_A = 8 * MF;
    constexpr int A_BYTES = GRAM_KG * ROWS_A * 64;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nst = Tm.nstages;
    const int b_kg_stride = Tm.brows_alloc * 64;
    constexpr int LS_BYTES = GRAM_KG * ROWS_A * C::LS_PER_ROW;
    const int ls_off = A_BYTES + GRAM_KG * b_kg_stride + C::PH_BYTES;  // stage = A | B | phase tile(s) | Ls plane (3M)
    const int stage_bytes = ls_off + LS_BYTES;
    unsigned char* tail = smem + (size_t)nst * stage_bytes;
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail);  // nst x 8 bytes
    const uint32_t bar_empty = bar_full + 8 * nst;
    GramJob* sjob = reinterpret_cast<GramJob*>(tail + 16 * 8);  // room for up to 8 stages of barriers

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, kk = lane & 3;

    const int2 cta = P.cta_map[blockIdx.x];
    const int job_id = cta.x;
    const int split = blockIdx.y;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, GT_CONSUMERS);
        }
        mbar_init_fence();
    }
    __syncthreads();
    const int nseg = sjob->nseg;

    const int Ne = P.Ne, nmom = sjob->nmom;  // this job's momenta (a prefix of the internal list)
    const int nfrag_f = (Ne + C::FW - 1) / C::FW;
    const int N_flat = nfrag_f * nmom;
    const int n_nt = (N_flat + C::NT - 1) / C::NT;
    const int mt = cta.y / n_nt, nt = cta.y - mt * n_nt;
    const int nflat0 = nt * C::NT;
    const int nflat_last = min(nflat0 + C::NT, N_flat) - 1;
    const int ff0 = nflat0 / nmom;
    const int nrows_b = C::FW * (nflat_last / nmom - ff0 + 1);
    int mf_live;
    const int row0 = gram_row_tile(Ne, P.n_mt, mt, mf_live);

    const int T_all = nseg * P.ksteps;
    const int T0 = (int)(((long long)T_all * split) / P.ksplit);
    const int T1 = (int)(((long long)T_all * (split + 1)) / P.ksplit);
    const int T = T1 - T0;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    if (warp >= GT_CONSUMERS) {
        // ================================ producer warpgroup ================================
        // hands its registers to the MMA warpgroups; only its first warp issues copies
        warpgroup_reg_dealloc<GT_REGS_PRODUCER>();
        if (warp != GT_CONSUMERS) return;
        const int nvalid = nflat_last - nflat0 + 1;  // n-fragments of this tile that exist
        const int p0 = nflat0 - ff0 * nmom;
        // phase tile = runs of consecutive momenta: slot s holds momentum (p0 + s) mod nmom
        int nruns = 0;
        for (int slot = 0, p = p0; slot < nvalid;) {
            const int len = min(nvalid - slot, nmom - p);
            ++nruns;
            slot += len;
            p = 0;
        }
        int run_slot = 0, run_p = 0, run_len = 0, run_tab = 0;
        if (lane < C::NTAB * nruns) {
            run_tab = lane / nruns;
            const int r = lane - run_tab * nruns;
            int i = 0;
            for (int slot = 0, p = p0; slot < nvalid; ++i) {
                const int len = min(nvalid - slot, nmom - p);
                if (i == r) {
                    run_slot = slot;
                    run_p = p;
                    run_len = len;
                }
                slot += len;
                p = 0;
            }
        }
        const int nbb = (nrows_b + 7) >> 3;
        const uint32_t tx_bytes = (uint32_t)(A_BYTES + LS_BYTES + GRAM_KG * nbb * 512 + C::NTAB * nvalid * 128);
        int seg = T0 / P.ksteps;
        int kstep = T0 - seg * P.ksteps;
        int s = 0;
        uint32_t par = 1;  // the first pass over the ring finds every slot free
        for (int it = 0; it < T; ++it) {
            mbar_wait(bar_empty + 8 * s, par);
            const uint32_t full = bar_full + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(full, tx_bytes);
            __syncwarp();
            const uint32_t st = smem_base + (uint32_t)(s * stage_bytes);
            const int kd = kstep * 48;  // first double of this stage's 24 complex k
            if (lane < GRAM_KG) tma_load_3d(st + lane * (ROWS_A * 64), Tm.mapA, full, kd + 8 * lane, row0, sjob->Lf[seg]);
            if (ALGO == 1 && lane >= 8 && lane < 8 + GRAM_KG)
                tma_load_3d(st + ls_off + (lane - 8) * (ROWS_A * 32), Tm.mapS, full, kstep * 24 + 4 * (lane - 8), row0,
                            sjob->Lf[seg]);
            for (int i = lane; i < GRAM_KG * nbb; i += 32) {
                const int kg = i / nbb, j = i - kg * nbb;
                tma_load_3d(st + A_BYTES + kg * b_kg_stride + j * 512, Tm.mapB, full, kd + 8 * kg, C::FW * ff0 + 8 * j,
                            sjob->Rf[seg]);
            }
            if (lane < C::NTAB * nruns) {
                // table 0 = phase, table 1 = -i*phase (global tile layout [kstep][2][nmom][8])
                const cplx* src = Tm.phase_tiles + (((size_t)kstep * 2 + run_tab) * P.nmom + run_p) * 8;
                bulk_load(st + A_BYTES + GRAM_KG * b_kg_stride + (run_tab * C::NT + run_slot) * 128, src, run_len * 128, full);
            }
            if (++kstep == P.ksteps) {
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
    warpgroup_reg_alloc<GT_REGS_CONSUMER>();
    int my_ffrag[C::NF], my_p[C::NF];
    bool my_valid[C::NF];
    uint32_t my_boff[C::NF];
#pragma unroll
    for (int n = 0; n < C::NF; ++n) {
        const int nf_raw = nflat0 + warp * C::NF + n;
        my_valid[n] = nf_raw < N_flat;
        const int nf = min(nf_raw, N_flat - 1);
        my_ffrag[n] = nf / nmom;
        my_p[n] = nf - my_ffrag[n] * nmom;
        const int brow = ALGO ? 8 * (my_ffrag[n] - ff0) + g : 4 * (my_ffrag[n] - ff0) + (g >> 1);
        my_boff[n] = (uint32_t)((brow * 4 + kk) * 16);
    }
    // slot of n-fragment nl in the phase tile is nl itself; n-fragments past the end of the tail
    // tile read whatever the slot holds, their columns are never stored
    const uint32_t my_phoff = (uint32_t)(((ALGO ? 0 : (g & 1)) * C::NT + warp * C::NF) * 8 * 16);

    double acc[MF][C::NACC][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int n = 0; n < C::NACC; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;
    int cur_sign = 1;
    int cs_seg = T0 / P.ksteps;
    int cs_kstep = T0 - cs_seg * P.ksteps;
    int s = 0;
    uint32_t par = 0;
    for (int it = 0; it < T; ++it) {
        const int sgn = sjob->sign[cs_seg];
        if (++cs_kstep == P.ksteps) {
            cs_kstep = 0;
            ++cs_seg;
        }
        if (sgn != cur_sign) {
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int n = 0; n < C::NACC; ++n) {
                    acc[i][n][0] = flip_sign(acc[i][n][0]);
                    acc[i][n][1] = flip_sign(acc[i][n][1]);
                }
            cur_sign = sgn;
        }
        mbar_wait(bar_full + 8 * s, par);
        const unsigned char* stage = smem + (size_t)s * stage_bytes;
        const unsigned char* b_s = stage + A_BYTES;
        const unsigned char* p_s = b_s + GRAM_KG * b_kg_stride + my_phoff;
        constexpr int MFM = MF > 1 ? MF - 1 : 1;
        if constexpr (ALGO == 0) {
            if (MF > 1 && mf_live == MF - 1)  // uniform per CTA
                gram_compute_stage<MF, MFM>(acc, stage + lane * 16, b_s, b_kg_stride, p_s, my_boff, kk);
            else
                gram_compute_stage<MF, MF>(acc, stage + lane * 16, b_s, b_kg_stride, p_s, my_boff, kk);
        } else {
            if (MF > 1 && mf_live == MF - 1)
                gram_compute_stage_3m<MF, MFM>(acc, stage + lane * 16, stage + ls_off + lane * 8, b_s, b_kg_stride, p_s,
                                               my_boff[0], kk);
            else
                gram_compute_stage_3m<MF, MF>(acc, stage + lane * 16, stage + ls_off + lane * 8, b_s, b_kg_stride, p_s,
                                              my_boff[0], kk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++s == nst) {
            s = 0;
            par ^= 1;
        }
    }

    const double fs = (double)cur_sign;
    cplx* outj = P.partial + ((size_t)split * P.njobs + job_id) * (size_t)P.nmom * Ne * Ne;
    if constexpr (ALGO == 0) {
        // lane holds C[e = row0+8i+g][f = 4*ffrag+kk] = (re, im)
#pragma unroll
        for (int n = 0; n < C::NF; ++n) {
            const int f = 4 * my_ffrag[n] + kk;
            if (!my_valid[n] || f >= Ne) continue;
            cplx* outp = outj + (size_t)my_p[n] * Ne * Ne;
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const int e = row0 + 8 * i + g;
                if (i < mf_live && e < Ne) outp[(size_t)e * Ne + f] = make_double2(fs * acc[i][n][0], fs * acc[i][n][1]);
            }
        }
    } else {
        // lane holds T1,T2,T3 of C[e = row0+8i+g][f = 8*ffrag + 2kk + {0,1}]
        const int f = 8 * my_ffrag[0] + 2 * kk;
        if (my_valid[0]) {
            cplx* outp = outj + (size_t)my_p[0] * Ne * Ne;
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const int e = row0 + 8 * i + g;
                if (i < mf_live && e < Ne) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (f + c < Ne) {
                            const double t1 = acc[i][0][c], t2 = acc[i][1][c], t3 = acc[i][2][c];
                            outp[(size_t)e * Ne + f + c] = make_double2(fs * (t1 + t2), fs * (t3 + (t1 - t2)));
                        }
                    }
                }
            }
        }
    }
}

// shared-memory plan of the TMA variant: rows of R per k-group, ring depth, total bytes
int gram_tma_plan(int algo, int mfrag, int nmom, int Ne, int* brows_alloc, int* nstages, int* smem_bytes) {
    const int FW = algo ? 8 : 4, NT = algo ? 8 : 16;
    const int nfrag_f = (Ne + FW - 1) / FW;
    int maxff = (NT - 1) / nmom + 2;  // distinct f-fragments NT consecutive n-fragments can touch
    if (maxff > NT) maxff = NT;
    if (maxff > nfrag_f) maxff = nfrag_f;
    int rows = ((FW * maxff + 7) / 8) * 8;
    if (rows > GRAM_BROWS) rows = GRAM_BROWS;
    const int ph = (algo ? 1 : 2) * NT * 128;
    const int stage = GRAM_KG * 8 * mfrag * 64 + GRAM_KG * rows * 64 + ph + (algo ? GRAM_KG * 8 * mfrag * 32 : 0);
    const int tail = 16 * 8 + (int)sizeof(GramJob) + 64;
    int nst = (227 * 1024 - tail) / stage;
    if (nst > 8) nst = 8;
    if (nst < 2) return -1;
    *brows_alloc = rows;
    *nstages = nst;
    *smem_bytes = nst * stage + tail;
    return 0;
}

int gram_nfrag_per_tile(int algo) { return algo ? 8 : 16; }
int gram_fwidth(int algo) { return algo ? 8 : 4; }

#ifndef EDK_EMU_NO_LAUNCHERS
template <int MF, int ALGO>
static cudaError_t launch_gram_tma_mf(const GramParams& P, const GramTma& T, cudaStream_t s) {
    int rows, nst, bytes;
    if (gram_tma_plan(ALGO, MF, P.nmom, P.Ne, &rows, &nst, &bytes) != 0 || rows != T.brows_alloc || nst != T.nstages)
        return cudaErrorInvalidValue;
    auto kern = gram_tma_kernel<MF, ALGO>;
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)P.ncta, (unsigned)P.ksplit);
    EDK_LAUNCH(kern, grid, GT_THREADS, bytes, s, P, T);
    return cudaGetLastError();
}

cudaError_t launch_gram_tma(const GramParams& P, const GramTma& T, int mfrag, int algo, cudaStream_t s) {
#define EDK_TMA_CASE(M) \
    case M: return algo ? launch_gram_tma_mf<M, 1>(P, T, s) : launch_gram_tma_mf<M, 0>(P, T, s);
    switch (mfrag) {
        EDK_TMA_CASE(2) EDK_TMA_CASE(3) EDK_TMA_CASE(4) EDK_TMA_CASE(5) EDK_TMA_CASE(6) EDK_TMA_CASE(7)
        EDK_TMA_CASE(8) EDK_TMA_CASE(9) EDK_TMA_CASE(10) EDK_TMA_CASE(11) EDK_TMA_CASE(12) EDK_TMA_CASE(13)
        default: return cudaErrorInvalidValue;
    }
#undef EDK_TMA_CASE
}
#endif  // EDK_EMU_NO_LAUNCHERS

// phase[2][nmom][Vpad] -> tiles[kstep][2][nmom][8]
__global__ void phase_tiles_kernel(const cplx* __restrict__ phase2, cplx* __restrict__ tiles, int nmom, int Vpad) {
    const size_t n = (size_t)2 * nmom * Vpad;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int site = (int)(i % Vpad);
    const size_t r = i / Vpad;  // tab * nmom + p
    const int p = (int)(r % nmom), tab = (int)(r / nmom);
    tiles[((((size_t)(site >> 3)) * 2 + tab) * nmom + p) * 8 + (site & 7)] = phase2[i];
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_phase_tiles(const cplx* phase2, cplx* tiles, int nmom, int Vpad, cudaStream_t s) {
    const size_t n = (size_t)2 * nmom * Vpad;
    EDK_LAUNCH(phase_tiles_kernel, (unsigned)((n + 255) / 256), 256, 0, s, phase2, tiles, nmom, Vpad);
    return cudaGetLastError();
}
#endif  // EDK_EMU_NO_LAUNCHERS

// available tile heights (m-fragments of 8 rows per CTA)
static const int kMfragAvail[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13};

int gram_pick_mfrag(int Ne) {
    const int frags = (Ne + 7) / 8;
    const int n_mt = (frags + GRAM_MAX_MF - 1) / GRAM_MAX_MF;
    const int need = (frags + n_mt - 1) / n_mt;
    for (int v : kMfragAvail)
        if (v >= need) return v;
    return GRAM_MAX_MF;
}
int gram_rows_per_tile(int mfrag) { return 8 * mfrag; }

#ifndef EDK_EMU_NO_LAUNCHERS
template <int MF>
static cudaError_t launch_gram_mf(const GramParams& P, cudaStream_t s) {
    using S = GramSmem<MF>;
    auto kern = gram_dmma_kernel<MF>;
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)P.ncta, (unsigned)P.ksplit);
    EDK_LAUNCH(kern, grid, GRAM_NTHREADS, S::TOTAL, s, P);
    return cudaGetLastError();
}

cudaError_t launch_gram_dmma(const GramParams& P, int mfrag, cudaStream_t s) {
#define EDK_DMMA_CASE(M) \
    case M: return launch_gram_mf<M>(P, s);
    switch (mfrag) {
        EDK_DMMA_CASE(2) EDK_DMMA_CASE(3) EDK_DMMA_CASE(4) EDK_DMMA_CASE(5) EDK_DMMA_CASE(6) EDK_DMMA_CASE(7)
        EDK_DMMA_CASE(8) EDK_DMMA_CASE(9) EDK_DMMA_CASE(10) EDK_DMMA_CASE(11) EDK_DMMA_CASE(12) EDK_DMMA_CASE(13)
        default: return cudaErrorInvalidValue;
    }
#undef EDK_DMMA_CASE
}
#endif  // EDK_EMU_NO_LAUNCHERS

bool gram_mfrag_available(int mfrag) {
    for (int v : kMfragAvail)
        if (v == mfrag) return true;
    return false;
}

// ---------------------------------------------------------------------------------------
// scalar cross-check (tests only): one thread per output element, split 0 only
// ---------------------------------------------------------------------------------------
__global__ void gram_naive_kernel(const GramParams P) {
    const int Ne = P.Ne;
    const size_t per_job = (size_t)P.nmom * Ne * Ne;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_job * P.njobs) return;
    const int job_id = (int)(idx / per_job);
    size_t r = idx - (size_t)job_id * per_job;
    const int p = (int)(r / ((size_t)Ne * Ne));
    r -= (size_t)p * Ne * Ne;
    const int e = (int)(r / Ne), f = (int)(r % Ne);
    const GramJob& job = P.jobs[job_id];
    double sr = 0.0, si = 0.0;
    const int V = P.Kc / 3;
    for (int s = 0; s < (p < job.nmom ? job.nseg : 0); ++s) {
        const cplx* L = job.L[s] + (size_t)e * P.Kc;
        const cplx* R = job.R[s] + (size_t)f * P.Kc;
        double tr = 0.0, ti = 0.0;
        for (int x = 0; x < V; ++x) {
            const cplx ph = P.phase[(size_t)p * P.Vpad + x];
            double ur = 0.0, ui = 0.0;
            for (int c = 0; c < 3; ++c) {
                const cplx l = L[3 * x + c], rr = R[3 * x + c];
                ur += l.x * rr.x + l.y * rr.y;
                ui += l.x * rr.y - l.y * rr.x;
            }
            tr += ph.x * ur - ph.y * ui;
            ti += ph.x * ui + ph.y * ur;
        }
        sr += job.sign[s] * tr;
        si += job.sign[s] * ti;
    }
    P.partial[idx] = make_double2(sr, si);
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_gram_naive(const GramParams& P, cudaStream_t s) {
    const size_t n = (size_t)P.njobs * P.nmom * P.Ne * P.Ne;
    EDK_LAUNCH(gram_naive_kernel, (unsigned)((n + 127) / 128), 128, 0, s, P);
    return cudaGetLastError();
}
#endif  // EDK_EMU_NO_LAUNCHERS

// ---------------------------------------------------------------------------------------
// combine: out[op][p] = coeff * sum_terms w * sum_split (partial[job][p]  or  partial[job][-p]^dagger)
// (the caller's momenta are the first nmom_out entries of the internal list of nmom_int)
// ---------------------------------------------------------------------------------------
__global__ void combine_kernel(const CombineOp* __restrict__ ops, int nop, const cplx* __restrict__ partial, int njobs,
                               int ksplit, int nmom_int, int nmom_out, const int* __restrict__ pmap,
                               const int* __restrict__ negidx, int n_half, int Ne, const double* __restrict__ coeff,
                               cplx* __restrict__ out) {
    const size_t mat = (size_t)Ne * Ne;
    const size_t per_op = (size_t)nmom_out * mat;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_op * nop) return;
    const int op = (int)(idx / per_op);
    size_t r = idx - (size_t)op * per_op;
    const int p_out = (int)(r / mat);
    r -= (size_t)p_out * mat;
    const int e = (int)(r / Ne), f = (int)(r - (size_t)e * Ne);
    const int p = pmap[p_out];  // internal index of the caller's momentum
    const CombineOp& o = ops[op];
    const size_t per_job = (size_t)nmom_int * mat;
    double sr = 0.0, si = 0.0;
    for (int t = 0; t < o.nterm; ++t) {
        bool herm = o.herm[t] != 0;
        int q = herm ? negidx[p] : p;
        if (o.half[t] && q >= n_half) {  // self pair kept for one momentum of each +-p couple only
            q = negidx[q];
            herm = !herm;
        }
        const size_t off = (size_t)q * mat + (herm ? (size_t)f * Ne + e : r);
        double tr = 0.0, ti = 0.0;
        for (int s = 0; s < ksplit; ++s) {
            const cplx v = partial[((size_t)s * njobs + o.job[t]) * per_job + off];
            tr += v.x;
            ti += v.y;
        }
        sr = fma(o.weight[t], tr, sr);
        si = fma(herm ? -o.weight[t] : o.weight[t], ti, si);
    }
    if (coeff != nullptr) {
        const double c = coeff[r];
        sr *= c;
        si *= c;
    }
    out[idx] = make_double2(sr, si);
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_combine(const CombineOp* ops_dev, int nop, const cplx* partial, int njobs, int ksplit, int nmom_int,
                           int nmom_out, const int* pmap, const int* negidx, int n_half, int Ne, const double* coeff, cplx* out,
                           cudaStream_t s) {
    const size_t n = (size_t)nop * nmom_out * Ne * Ne;
    EDK_LAUNCH(combine_kernel, (unsigned)((n + 255) / 256), 256, 0, s, ops_dev, nop, partial, njobs, ksplit, nmom_int, nmom_out, pmap,
               negidx, n_half, Ne, coeff, out);
    return cudaGetLastError();
}
#endif  // EDK_EMU_NO_LAUNCHERS

// ---------------------------------------------------------------------------------------
// FP64 pipe micro-benchmarks (roofline denominators measured on the box)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* sink, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

#ifndef EDK_HOST_EMU
cudaError_t microbench_fp64(double* dmma_tflops, double* dfma_tflops) {
    double* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 2, threads = 256;
    float ms = 0.f;
    // DMMA: per warp-iteration 16 MMAs x 512 flop
    const int it_mma = 20000;
    dmma_peak_kernel<<<blocks, threads>>>(sink, 2000);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<blocks, threads>>>(sink, it_mma);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = (double)blocks * (threads / 32) * (double)it_mma * 16.0 * 512.0;
        best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    *dmma_tflops = best;
    const int it_fma = 40000;
    dfma_peak_kernel<<<blocks, threads>>>(sink, 2000);
    best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(sink, it_fma);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = (double)blocks * threads * (double)it_fma * 16.0 * 2.0;
        best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    *dfma_tflops = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e = cudaDeviceSynchronize();
    cudaFree(sink);
    return e;
}
#elif !defined(EDK_EMU_NO_LAUNCHERS)
cudaError_t microbench_fp64(double*, double*) { return cudaErrorNotSupported; }  // timing loops make no sense on the emulator
#endif  // EDK_HOST_EMU

}  // namespace edk
