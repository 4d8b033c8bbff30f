// Momentum-phased Ne x Ne contraction on the FP64 tensor pipe (DMMA.8x8x4), sm_100a.
//
//   G[p][e][f] = sum_seg sign_seg * sum_{x,c} conj(L_seg[e][x][c]) * phase_p(x) * R_seg[f][x][c]
//   (reference: the einsum "zyx,ezyxc,fzyxc->ef" of lattice/generator/elemental.py:322-329 and
//    lattice/generator/displacement_elemental.py:94-95, with the left/right split sum of
//    elemental.py:309-321 folded into the K loop as signed segments)
//
// Mapping onto real m8n8k4 MMAs (tcgen05 has no f64 kind; mma.sync -> SASS DMMA.8x8x4):
//   A operand  = rows e of L, k-slots hold Re L (first MMA) or Im L (second MMA) of 4 complex k,
//   B operand  = columns (f, re|im), k-slots hold the matching part of P = phase * R:
//                  column (f,re): (Pr | Pi)      column (f,im): (Pi | -Pr)
//   so that  C[e][(f,re)] += Lr*Pr + Li*Pi = Re(conj(L) P),  C[e][(f,im)] += Lr*Pi - Li*Pr = Im(conj(L) P)
//   and the accumulator fragment (2 doubles per lane) is one complex128 in natural (re, im) order.
//   The phase multiply is fused into the B-fragment build: 4 FP64 ops per 2*MF MMAs; the
//   (f,im) columns read a pre-rotated table -i*phase, so there is no select or negation.
//
// CTA tile: 8*MF rows (all warps share them) x 16 "n-fragments"; an n-fragment is 4 consecutive
// f at one momentum, n-fragments are flattened f-fragment-major so one CTA needs few rows of R.
// Shared memory per stage (8 sites = 24 complex k): A [6 kgroups][8*MF rows][4 complex] so a
// warp's fragment load is 512 contiguous bytes (bank-conflict free LDS.128), B likewise, plus the
// 16 x 8 phase entries.  3-stage cp.async pipeline, one __syncthreads per stage.
#include "edk_common.cuh"

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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> 16 zero bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// sign flip on the integer pipe: the FP64 pipe is the one the DMMAs need
__device__ __forceinline__ double flip_sign(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
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

    int bid = blockIdx.x;
    const int nt = bid % P.n_nt;
    bid /= P.n_nt;
    const int mt = bid % P.n_mt;
    const int job_id = bid / P.n_mt;
    const int split = blockIdx.y;

    if (tid < (int)(sizeof(GramJob) / sizeof(int))) {
        reinterpret_cast<int*>(sjob)[tid] = reinterpret_cast<const int*>(P.jobs + job_id)[tid];
    }
    __syncthreads();
    const int nseg = sjob->nseg;

    // ---- column bookkeeping -------------------------------------------------------------
    const int Ne = P.Ne, nmom = P.nmom;
    const int nfrag_f = (Ne + 3) >> 2;
    const int N_flat = nfrag_f * nmom;
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
    const int row0 = mt * ROWS_A;
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
        ph_src = P.phase + ((size_t)(tid >> 7) * nmom + (nf % nmom)) * P.Vpad + (t2 & 7);
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
                    rr[n] = *reinterpret_cast<const cplx*>(b_s + (kg + 1) * (GRAM_BROWS * 64) + my_boff[n]);
                    pp[n] = *reinterpret_cast<const cplx*>(p_s + (n * 8 + (4 * (kg + 1) + kk) / 3) * 16);
                }
            }
            double nb1[GRAM_NF], nb2[GRAM_NF];
            // m-fragments in groups of <= 5: both MMAs of an accumulator are >= 2*group MMAs apart
            constexpr int GRP = 5;
#pragma unroll
            for (int i0 = 0; i0 < MF; i0 += GRP) {
                cplx a[GRP];
#pragma unroll
                for (int ii = 0; ii < GRP; ++ii)
                    if (i0 + ii < MF) a[ii] = *reinterpret_cast<const cplx*>(a_s + (kg * ROWS_A + 8 * (i0 + ii)) * 64);
#pragma unroll
                for (int ii = 0; ii < GRP; ++ii)
                    if (i0 + ii < MF) {
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
                    if (i0 + ii < MF) {
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
    cp_async_wait<0>();

    // ---- epilogue: lane holds C[e = row0+8i+g][f = 4*ffrag+kk] = (re, im) ------------------------
    const double fs = (double)cur_sign;
    cplx* outj = P.partial + ((size_t)split * P.njobs + job_id) * (size_t)nmom * Ne * Ne;
#pragma unroll
    for (int n = 0; n < GRAM_NF; ++n) {
        const int f = 4 * my_ffrag[n] + kk;
        if (!my_valid[n] || f >= Ne) continue;
        cplx* outp = outj + (size_t)my_p[n] * Ne * Ne;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
            const int e = row0 + 8 * i + g;
            if (e < Ne) outp[(size_t)e * Ne + f] = make_double2(fs * acc[i][n][0], fs * acc[i][n][1]);
        }
    }
}

// available tile heights (m-fragments of 8 rows per CTA)
static const int kMfragAvail[] = {2, 4, 5, 7, 9, 10, 11, 13};

int gram_pick_mfrag(int Ne) {
    const int frags = (Ne + 7) / 8;
    const int n_mt = (frags + GRAM_MAX_MF - 1) / GRAM_MAX_MF;
    const int need = (frags + n_mt - 1) / n_mt;
    for (int v : kMfragAvail)
        if (v >= need) return v;
    return GRAM_MAX_MF;
}
int gram_rows_per_tile(int mfrag) { return 8 * mfrag; }
int gram_nfrag_per_tile() { return GRAM_NT; }

template <int MF>
static cudaError_t launch_gram_mf(const GramParams& P, cudaStream_t s) {
    using S = GramSmem<MF>;
    static bool configured = false;  // per template instance; attribute is per device function
    cudaError_t e = cudaFuncSetAttribute(gram_dmma_kernel<MF>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    configured = true;
    (void)configured;
    dim3 grid((unsigned)(P.njobs * P.n_mt * P.n_nt), (unsigned)P.ksplit);
    gram_dmma_kernel<MF><<<grid, GRAM_NTHREADS, S::TOTAL, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_gram_dmma(const GramParams& P, int mfrag, cudaStream_t s) {
    switch (mfrag) {
        case 2: return launch_gram_mf<2>(P, s);
        case 4: return launch_gram_mf<4>(P, s);
        case 5: return launch_gram_mf<5>(P, s);
        case 7: return launch_gram_mf<7>(P, s);
        case 9: return launch_gram_mf<9>(P, s);
        case 10: return launch_gram_mf<10>(P, s);
        case 11: return launch_gram_mf<11>(P, s);
        case 13: return launch_gram_mf<13>(P, s);
        default: return cudaErrorInvalidValue;
    }
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
    for (int s = 0; s < job.nseg; ++s) {
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

cudaError_t launch_gram_naive(const GramParams& P, cudaStream_t s) {
    const size_t n = (size_t)P.njobs * P.nmom * P.Ne * P.Ne;
    gram_naive_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// combine: out[op][p] = coeff * sum_terms w * sum_split (partial[job][p]  or  partial[job][-p]^dagger)
// (the caller's momenta are the first nmom_out entries of the internal list of nmom_int)
// ---------------------------------------------------------------------------------------
__global__ void combine_kernel(const CombineOp* __restrict__ ops, int nop, const cplx* __restrict__ partial, int njobs,
                               int ksplit, int nmom_int, int nmom_out, const int* __restrict__ negidx, int Ne,
                               const double* __restrict__ coeff, cplx* __restrict__ out) {
    const size_t mat = (size_t)Ne * Ne;
    const size_t per_op = (size_t)nmom_out * mat;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_op * nop) return;
    const int op = (int)(idx / per_op);
    size_t r = idx - (size_t)op * per_op;
    const int p = (int)(r / mat);
    r -= (size_t)p * mat;
    const int e = (int)(r / Ne), f = (int)(r - (size_t)e * Ne);
    const CombineOp& o = ops[op];
    const size_t per_job = (size_t)nmom_int * mat;
    double sr = 0.0, si = 0.0;
    for (int t = 0; t < o.nterm; ++t) {
        const bool herm = o.herm[t] != 0;
        const size_t off = herm ? (size_t)negidx[p] * mat + (size_t)f * Ne + e : (size_t)p * mat + r;
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

cudaError_t launch_combine(const CombineOp* ops_dev, int nop, const cplx* partial, int njobs, int ksplit, int nmom_int,
                           int nmom_out, const int* negidx, int Ne, const double* coeff, cplx* out, cudaStream_t s) {
    const size_t n = (size_t)nop * nmom_out * Ne * Ne;
    combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ops_dev, nop, partial, njobs, ksplit, nmom_int, nmom_out, negidx,
                                                               Ne, coeff, out);
    return cudaGetLastError();
}

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

cudaError_t microbench_fp64(double* dmma_tflops, double* dfma_tflops) {
    double* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 256;
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

}  // namespace edk
