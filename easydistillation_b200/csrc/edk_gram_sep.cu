// Separable contraction ("algo 4"), sm_100a.
//
//   G[p][e][f] = sum_seg sign_seg sum_{x,c} conj(L_seg[e][x][c]) phase_p(x) R_seg[f][x][c]
//   (the einsum "zyx,ezyxc,fzyxc->ef" of lattice/generator/elemental.py:322-329 and
//    lattice/generator/displacement_elemental.py:94-95 for every momentum of the list)
//
// The plane-wave forms (edk_gram_pw.cu) form the colour-summed site product C(x)[e][f] once per site and transform
// an xy-plane against the 13 real xy-modes on the DMMA: 2 x 16 mode rows per site, 30 FP64-pipe operations per
// (pair, e, f, site) in the folded form.  DMMA.8x8x4 and DFMA run on ONE FP64 pipe at the same rate on B200
// (tools/microbench/fp64_pipes.cu), so the pipe time is the operation count - and the phase factorises further,
// between x and y.  About the centre of the lattice,
//   exp(2 pi i (px x/Lx + py y/Ly)) = k_p (c_|px|(x) + i sgn(px) s_|px|(x)) (c_|py|(y) + i sgn(py) s_|py|(y)),
//   c_q(x) = cos(pi q (2x - Lx + 1)/Lx), s_q(x) = sin(pi q (2x - Lx + 1)/Lx): c_q even, s_q odd under x -> Lx-1-x.
// gram_sep_kernel transforms a plane row by row, in registers, every lane owning its (e, f) elements privately:
//   per pair of sites (x, xbar = Lx-1-x) of a row:   C(x), C(xbar)                       24 DFMA
//                                                   S = C(x) + C(xbar), D = C(x) - C(xbar) 4 DADD
//       x stage   X[1] += S,  X[c_q] += c_q(x) S,  X[s_q] += s_q(x) D                   2 DADD + 4 QMAX DFMA
//   once per row: y stage   Y[(m, c_q')] += c_q'(y) X[m],  Y[(m, s_q')] += s_q'(y) X[m],  X = 0
// = 19 operations per site for the 33 momenta with |p|^2 <= 4 (5 x-modes, 13 separable xy-modes), 12 of them the
// site product itself.  sep_zfold_kernel folds z and recombines the separable modes into the momenta:
//   G[job][p] = sum_z zphase_p(z) (Y[cc] + i sy Y[cs] + i sx Y[sc] - sx sy Y[ss]),  sx = sgn(px), sy = sgn(py).
// All lanes of a warp walk the same sites, so the operands are broadcast reads: lane (le = lane / 8, lf = lane % 8) of
// warp (we, wf) owns row 4 we + le of the 16-row L tile and rows 16 wf + lf + {0, 8} of the 32-row R tile.  The tiles
// are TMA boxes of 16 doubles (128 bytes = 8 complex) x rows with the 128-byte swizzle: 16-byte slot s of row r sits
// at slot s ^ (r % 8), so the 8 rows a load touches fall into 8 different bank groups (a row of 8 sites = 24 complex =
// 3 chunks of 128 bytes).  Stage = PAIRS site pairs of one row: the 8 front sites from x = PAIRS k and the 8 back
// sites up to x = Lx - 1 - PAIRS k, as two sets of boxes; TMA producer warp + mbarrier ring as in gram_pwf_kernel.
// Self pairs (L == R) skip their tiles below the diagonal (Hermitian site product); the fold reads the mirror.
#include "edk_gram_sep.cuh"

namespace edk {

int sep_num_modes(int qmax, int r2) {
    if (qmax == 1 && r2 == 1) return SepModes<1, 1>::N;
    if (qmax == 1 && r2 == 2) return SepModes<1, 2>::N;
    if (qmax == 2 && r2 == 4) return SepModes<2, 4>::N;
    return 0;
}
int sep_mode_index(int qmax, int r2, int qx, int xk, int qy, int yk) {
    if (qx < 0 || qy < 0 || qx > qmax || qy > qmax || (xk && !qx) || (yk && !qy)) return -1;
    if (qmax == 1 && r2 == 1) return SepModes<1, 1>::index(qx, xk, qy, yk);
    if (qmax == 1 && r2 == 2) return SepModes<1, 2>::index(qx, xk, qy, yk);
    if (qmax == 2 && r2 == 4) return SepModes<2, 4>::index(qx, xk, qy, yk);
    return -1;
}

int sep_plan_smem(int* nstages, int* smem_bytes) {
    int nst = (227 * 1024 - 1024 - SEP_TAIL) / SEP_STAGE;  // 1024: the kernel aligns its tiles itself
    if (nst > SEP_MAX_STAGES) nst = SEP_MAX_STAGES;
    if (nst < 2) return -1;
    *nstages = nst;
    *smem_bytes = nst * SEP_STAGE + SEP_TAIL + 1024;
    return 0;
}

int sepx_shape(int shape, int* we, int* wf) {
    static const int kShapes[SEP_NSHAPES][2] = {{4, 2}, {1, 8}, {8, 1}};
    if (shape < 0 || shape >= SEP_NSHAPES) return -1;
    *we = kShapes[shape][0];
    *wf = kShapes[shape][1];
    return 0;
}

int sepx_plan_smem(int shape, int* nstages, int* smem_bytes) {
    int we, wf;
    if (sepx_shape(shape, &we, &wf) != 0) return -1;
    const int stage = 2 * SEP_CHUNKS * (8 * we + 16 * wf) * 128;
    int nst = (227 * 1024 - 1024 - SepxGeom::TAIL) / stage;
    if (nst > SEP_MAX_STAGES) nst = SEP_MAX_STAGES;
    if (nst < 2) return -1;
    *nstages = nst;
    *smem_bytes = 227 * 1024;  // one allocation for every shape: the ring depth follows the shape of the CTA's tile
    return 0;
}

#ifndef EDK_EMU_NO_LAUNCHERS
int sep_variant_rows(int variant) { return variant == 6 ? 8 : 16; }

int sep_variant_plan(int variant, int* nstages, int* smem_bytes) {
    if (variant == 0) return sep_plan_smem(nstages, smem_bytes);
    return -1;
}

// shape 0 .. 2: P.tiles / P.ntiles are the tiles of that shape only (sep_build_tiles keeps them grouped; A/B reference);
// shape 3 = SEP_NSHAPES: the whole tile table in one launch (the product).
cudaError_t launch_gram_sepx(const SepParams& P, const SepTmaX& T, const SepWeights& W, int qmax, int r2, int pairs, int shape, cudaStream_t s) {
    int nst, bytes = 0;
    if (shape < 0 || shape > SEP_NSHAPES) return cudaErrorInvalidValue;
    for (int sh = 0; sh < SEP_NSHAPES; ++sh) {
        if (shape != SEP_NSHAPES && shape != sh) continue;
        if (sepx_plan_smem(sh, &nst, &bytes) != 0 || T.nstages[sh] < 2 || T.nstages[sh] > nst) return cudaErrorInvalidValue;
    }
    if ((pairs != 8 && pairs != 6 && pairs != 4) || (P.Lx & 1) || P.Lx < 8 || (P.Lx / 2) % pairs != 0 || P.Lx / 2 > SEP_MAX_PR ||
        P.SR != P.Lx / 2 / pairs || !P.tiles || P.ntiles < 1)
        return cudaErrorInvalidValue;
    const long long items = (long long)P.njobs * P.Lz * P.ntiles;
    if (items < 1 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (qmax == 1 && r2 == 1) return launch_gram_sepx_s11(P, T, W, pairs, shape, bytes, (unsigned)items, s);
    if (qmax == 1 && r2 == 2) return launch_gram_sepx_s12(P, T, W, pairs, shape, bytes, (unsigned)items, s);
    if (qmax == 2 && r2 == 4) return launch_gram_sepx_s24(P, T, W, pairs, shape, bytes, (unsigned)items, s);
    return cudaErrorInvalidValue;
}

// variant 0 = gram_sep_kernel (1 x 2 elements per lane, Y in registers: the first version, kept as an A/B reference);
// variant 6 = gram_sepx_kernel (the product) is launched by launch_gram_sepx
cudaError_t launch_gram_sep(const SepParams& P, const SepTma& T, const SepWeights&, int qmax, int r2, int pairs, int variant, cudaStream_t s) {
    int nst, bytes;
    if (variant != 0 || sep_variant_plan(variant, &nst, &bytes) != 0 || T.nstages < 2 || T.nstages > nst) return cudaErrorInvalidValue;
    if ((pairs != 8 && pairs != 6 && pairs != 4) || (P.Lx & 1) || P.Lx < 8 || (P.Lx / 2) % pairs != 0 || P.Lx / 2 > SEP_MAX_PR ||
        P.SR != P.Lx / 2 / pairs)
        return cudaErrorInvalidValue;
    if (P.n_et != (P.Ne + SEP_TE - 1) / SEP_TE || P.n_ft != (P.Ne + SEP_TF - 1) / SEP_TF) return cudaErrorInvalidValue;
    const long long items = (long long)P.njobs * P.Lz * P.n_et * P.n_ft;
    if (items < 1 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (qmax == 1 && r2 == 1) return launch_gram_sep_s11(P, T, pairs, bytes, (unsigned)items, s);
    if (qmax == 1 && r2 == 2) return launch_gram_sep_s12(P, T, pairs, bytes, (unsigned)items, s);
    if (qmax == 2 && r2 == 4) return launch_gram_sep_s24(P, T, pairs, bytes, (unsigned)items, s);
    return cudaErrorInvalidValue;
}
#endif  // EDK_EMU_NO_LAUNCHERS

// G[job][p][e][f] = sum_z zphase[p][z] (Y[cc] + i sy Y[cs] + i sx Y[sc] - sx sy Y[ss])[job][z][e][f].
// One block folds one class of momenta - those with the same (|px|, |py|), which read the same <= 4 separable modes -
// of one (job, 128 matrix elements), up to SEP_FOLD_MAXM momenta at a time (12 = the largest class of the |p|^2 <= 4
// set), so every plane of Y is read once: the kernel is a stream over Y (7.6 GB at config 5).  Per plane the four sign
// combinations are formed once,
//   P = cc - ss, Q = cc + ss, R = cs + sc, S = sc - cs:   u = P + i sigma R (sx sy >= 0, sigma = sy or sx),  u = Q + i sx S (sx sy < 0),
// so a momentum costs two multiply-adds for u and four for zphase * u; the z phases of the block's momenta sit in
// shared memory.  CNT (momenta of this pass) is a template parameter: no predicated-off work, and the accumulators of
// a small class do not cost the registers of a large one.
constexpr int SEP_FOLD_THREADS = 128;
constexpr int SEP_FOLD_MAXM = 12;

template <int CNT>
__device__ __forceinline__ void sep_fold_pass(const SepFold& F, const cplx* Yj, double cj, const SepClass& K, const int* m3,
                                              const cplx* sph, size_t mat, size_t out_base) {
    // m3: (internal momentum, sgn px, sgn py) of the CNT momenta of this pass; sph[k][z] their z phases
    double ar[CNT], ai[CNT], sig[CNT];
    bool useq[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        ar[k] = ai[k] = 0.0;
        const int sx = m3[3 * k + 1], sy = m3[3 * k + 2];
        useq[k] = sx * sy < 0;
        sig[k] = useq[k] ? (double)sx : (double)(sy != 0 ? sy : sx);
    }
    const size_t zstride = (size_t)F.nmodes * mat;
    const bool has1 = K.mode[1] >= 0, has2 = K.mode[2] >= 0, has3 = K.mode[3] >= 0;
    const cplx* p0 = Yj + (size_t)K.mode[0] * mat;
    const cplx* p1 = Yj + (size_t)(has1 ? K.mode[1] : K.mode[0]) * mat;
    const cplx* p2 = Yj + (size_t)(has2 ? K.mode[2] : K.mode[0]) * mat;
    const cplx* p3 = Yj + (size_t)(has3 ? K.mode[3] : K.mode[0]) * mat;
    const cplx zero = make_double2(0.0, 0.0);
#pragma unroll(CNT > 6 ? 2 : 4)
    for (int z = 0; z < F.Lz; ++z) {
        const cplx cc = p0[(size_t)z * zstride];
        const cplx cs = has1 ? p1[(size_t)z * zstride] : zero;
        const cplx sc = has2 ? p2[(size_t)z * zstride] : zero;
        const cplx ss = has3 ? p3[(size_t)z * zstride] : zero;
        // (the conjugation of a mirror read, cj = -1, acts on the imaginary parts)
        const double Pr = cc.x - ss.x, Pi = cj * (cc.y - ss.y), Qr = cc.x + ss.x, Qi = cj * (cc.y + ss.y);
        const double Rr = cs.x + sc.x, Ri = cj * (cs.y + sc.y), Sr = sc.x - cs.x, Si = cj * (sc.y - cs.y);
#pragma unroll
        for (int k = 0; k < CNT; ++k) {
            const double br = useq[k] ? Qr : Pr, bi = useq[k] ? Qi : Pi, rr = useq[k] ? Sr : Rr, ri = useq[k] ? Si : Ri;
            const double ur = fma(-sig[k], ri, br), ui = fma(sig[k], rr, bi);  // base + i sigma rot
            const cplx ph = sph[k * F.Lz + z];
            ar[k] = fma(ph.x, ur, ar[k]);
            ar[k] = fma(-ph.y, ui, ar[k]);
            ai[k] = fma(ph.x, ui, ai[k]);
            ai[k] = fma(ph.y, ur, ai[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < CNT; ++k) F.partial[out_base + (size_t)m3[3 * k] * mat] = make_double2(ar[k], ai[k]);
}

__global__ void __launch_bounds__(SEP_FOLD_THREADS, 2) sep_zfold_kernel(const SepFold F) {
    extern __shared__ __align__(1024) unsigned char smem[];
    cplx* sph = reinterpret_cast<cplx*>(smem);  // [SEP_FOLD_MAXM][Lz]
    const size_t mat = (size_t)F.Ne * F.Ne;
    const int nblk = (int)((mat + SEP_FOLD_THREADS - 1) / SEP_FOLD_THREADS);
    int b = blockIdx.x;
    const int cls = b % F.nclass;
    b /= F.nclass;
    const int blk = b % nblk;
    const int job = b / nblk;
    const GramJob& J = F.jobs[job];
    const SepClass K = F.classes[cls];
    // the job's momenta are the first J.nmom of the internal list (self pairs contract the half set only); the class
    // lists its momenta in ascending order, so the ones this job needs are a prefix
    int nvalid = 0;
    for (int i = 0; i < K.count; ++i) nvalid += F.mom[3 * (K.first + i)] < J.nmom;
    const size_t ef_raw = (size_t)blk * SEP_FOLD_THREADS + threadIdx.x;
    const bool live = ef_raw < mat;
    const size_t ef = live ? ef_raw : 0;
    // blocks of a self pair below the diagonal were not computed: read the mirror element, conjugated
    const int e = (int)(ef / F.Ne), f = (int)(ef - (size_t)e * F.Ne);
    const bool mirror = J.nseg == 1 && J.Lf[0] == J.Rf[0] && (e / F.rows_l) * F.rows_l > (f / F.rows_r) * F.rows_r + F.rows_r - 1;
    const double cj = mirror ? -1.0 : 1.0;
    const cplx* Yj = F.Y + (size_t)job * F.Lz * F.nmodes * mat + (mirror ? (size_t)f * F.Ne + e : ef);
    const size_t out_base = (size_t)job * F.nmom_int * mat + ef;
    for (int c0 = 0; c0 < nvalid; c0 += SEP_FOLD_MAXM) {
        const int cnt = min(SEP_FOLD_MAXM, nvalid - c0);
        const int* m3 = F.mom + 3 * (K.first + c0);
        __syncthreads();  // the previous pass is through with the table
        for (int i = threadIdx.x; i < cnt * F.Lz; i += SEP_FOLD_THREADS) sph[i] = F.zphase[(size_t)m3[3 * (i / F.Lz)] * F.Lz + i % F.Lz];
        __syncthreads();
        if (live) {
            switch (cnt) {
#define EDK_FOLD_CASE(N) \
    case N: sep_fold_pass<N>(F, Yj, cj, K, m3, sph, mat, out_base); break;
                EDK_FOLD_CASE(1) EDK_FOLD_CASE(2) EDK_FOLD_CASE(3) EDK_FOLD_CASE(4) EDK_FOLD_CASE(5) EDK_FOLD_CASE(6)
                EDK_FOLD_CASE(7) EDK_FOLD_CASE(8) EDK_FOLD_CASE(9) EDK_FOLD_CASE(10) EDK_FOLD_CASE(11) EDK_FOLD_CASE(12)
#undef EDK_FOLD_CASE
                default: break;
            }
        }
    }
}

#ifndef EDK_EMU_NO_LAUNCHERS
cudaError_t launch_sep_zfold(const SepFold& F, cudaStream_t s) {
    const size_t mat = (size_t)F.Ne * F.Ne;
    const long long nblk = (long long)((mat + SEP_FOLD_THREADS - 1) / SEP_FOLD_THREADS);
    const long long blocks = nblk * F.njobs * F.nclass;
    const size_t smem = (size_t)SEP_FOLD_MAXM * F.Lz * sizeof(cplx);
    if (blocks < 1 || blocks > 0x7fffffffLL || smem > 160 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)sep_zfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    EDK_LAUNCH(sep_zfold_kernel, (unsigned)blocks, SEP_FOLD_THREADS, smem, s, F);
    return cudaGetLastError();
}
#endif  // EDK_EMU_NO_LAUNCHERS

}  // namespace edk
