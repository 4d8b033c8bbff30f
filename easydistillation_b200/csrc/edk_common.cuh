// Shared declarations of the elemental-distillation kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "edk.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "edk kernels are written for sm_100a (B200) only"
#endif

// Kernel launches go through one macro so that tests/emu can run the whole library (host glue, launchers and
// kernel sources) on a machine without a GPU: under EDK_HOST_EMU (test builds by g++ only, never the shipped
// library) a launch runs the kernel body on host threads, one per CUDA thread.
#ifdef EDK_HOST_EMU
#define EDK_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) \
    ::edk::emu_launch(dim3(grid), dim3(block), (size_t)(smem_bytes), [&] { kernel(__VA_ARGS__); })
#define EDK_SHARED static
#else
#define EDK_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) kernel<<<(grid), (block), (smem_bytes), (stream)>>>(__VA_ARGS__)
#define EDK_SHARED __shared__
#endif

namespace edk {

typedef double2 cplx;  // (re, im)

// ---- error plumbing (thread-local message, C-ABI status codes) ----------------------
void set_error(const char* fmt, ...);
#define EDK_CUDA_TRY(expr)                                                                       \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            edk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return EDK_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

// ---- lattice geometry of one timeslice ------------------------------------------------
struct Geom {
    int Lx, Ly, Lz;
    int V;     // Lx*Ly*Lz
    int Vpad;  // V rounded up to a multiple of 8 sites (k-stage granularity of the contraction)
};

// ---- contraction job description --------------------------------------------------------
// One job = sum over segments s of  sign_s * L_s^dagger diag(phase_p) R_s  for every momentum p,
// written to partial[split][job][p][e][f].
#define EDK_MAX_SEG 8
struct GramJob {
    int nseg;
    int nmom;                    // momenta contracted by this job: the first nmom of the internal list
    int sign[EDK_MAX_SEG];
    int Lf[EDK_MAX_SEG];         // field indices (third TMA coordinate) of L and R
    int Rf[EDK_MAX_SEG];
    const cplx* L[EDK_MAX_SEG];  // [Ne][3V]
    const cplx* R[EDK_MAX_SEG];  // [Ne][3V]
};

struct GramParams {
    const GramJob* jobs;
    int njobs;
    int Ne;
    int nmom;
    int Kc;       // 3*V complex entries per row
    int ksteps;   // Vpad/8 : stages of 8 sites (24 complex) per segment
    int Vpad;     // row stride of the phase table
    int ksplit;   // split-K factor
    int n_mt;     // tiles along e (rows); tiles along the flattened (f-fragment, momentum) axis depend on the job
    int ncta;     // CTAs per split = entries of cta_map
    const int2* cta_map;  // per CTA: (job, tile index inside the job = mt * n_nt(job) + nt)
    const cplx* phase;  // [2][nmom][Vpad]: phase, then -i*phase
    cplx* partial;      // [ksplit][njobs][nmom][Ne][Ne]
};

// Extra launch state of the TMA-fed contraction: tensor maps over the field array viewed as
// [nfield][Ne][2*Kc doubles] (box = 8 doubles x rows x 1 field) and the phase table re-tiled as
// [kstep][2][nmom][8 sites] so that consecutive momenta of one 8-site stage are contiguous.
struct GramTma {
    alignas(64) unsigned char mapA[128];  // CUtensorMap, box rows = 8*mfrag
    alignas(64) unsigned char mapB[128];  // CUtensorMap, box rows = 8
    alignas(64) unsigned char mapS[128];  // CUtensorMap over the Re+Im planes [nfield][Ne][Kc doubles], box 4 x 8*mfrag x 1
    const cplx* phase_tiles;
    int brows_alloc;  // rows of R kept per k-group in shared memory (multiple of 8, <= 64)
    int nstages;      // depth of the full/empty ring
};

// Output operator n at momentum p = coeff * sum_terms weight * sum_split X, with
//   X = partial[split][job][p]                        (herm = 0)
//   X = partial[split][job][index of -p]^dagger        (herm = 1: G(L,R,p) = G(R,L,-p)^dagger)
#define EDK_MAX_TERMS 8
//   half = 1: the job is a self pair (L == R) contracted only for the first n_half momenta of the
//   internal list (one of every +-p couple); the other momenta are read as G(L,L,-p)^dagger.
struct CombineOp {
    int nterm;
    int job[EDK_MAX_TERMS];
    int herm[EDK_MAX_TERMS];
    int half[EDK_MAX_TERMS];
    double weight[EDK_MAX_TERMS];
};

// ---- plane-wave factorised contraction (algo 2, edk_gram_pw.cu) -----------------------------
// The phase of momentum p factorises over the lattice axes,
//   phase_p(x,y,z) = [cos(theta_q(x,y)) + i sigma_p sin(theta_q(x,y))] * exp(2 pi i pz z / Lz),
// with q the {+(px,py), -(px,py)} couple of p.  So the colour-summed site product
//   C(x)[e][f] = sum_seg sign_seg sum_c conj(L_seg[e][x][c]) R_seg[f][x][c]
// is formed ONCE per site (12 DFMA) and folded over an xy-plane against the few REAL mode
// functions cos(theta_q), sin(theta_q) by DMMA (A = modes x 4 sites, B = 4 sites x 8 f):
//   Y[job][z][m][e][f] = sum_{x,y} w_m(x,y) C(x,y,z)[e][f]
// (13 modes for the 33 momenta |p|^2 <= 4), and a small second kernel folds z:
//   G[job][p] = sum_z exp(2 pi i pz z/Lz) (Y[z][mc(p)] + i sigma_p Y[z][ms(p)]).
// FP64-pipe issue slots per (pair, e, f, site): 12/32 + 2*8*MB/32 = 1.4 against 8.3 of the 3M GEMM form.
// A lane owns EL e-rows x FL f-columns of its warp's (EL) x (8 FL) tile; the 8 MMA warps are stacked along e, so a
// CTA tile is 8 EL rows of L by 8 FL rows of R.  Instantiated shapes: (2,4) = 16 x 32, (2,5) = 16 x 40, (1,7) = 8 x 56.
constexpr int PW_WARPS = 8;             // MMA warps, stacked along e
constexpr int PW_MAX_MB = 2;            // m-blocks (8 modes) per pass

struct PwParams {
    const GramJob* jobs;
    int njobs;
    int Ne;
    int Lz;
    int A;        // sites of one xy-plane (Lx*Ly)
    int kplane;   // stages of 8 sites per plane = ceil(A/8)
    int n_et, n_ft;  // tiles of (8 EL) x (8 FL) rows
    int nmodes;   // real xy-modes kept in Y
    int mb0;      // first m-block of this pass
    int mbtot;    // m-blocks of the weight tiles = ceil(nmodes/8)
    const double* wtiles;  // [kplane][2 groups of 4 sites][mbtot][32 lanes]: lane = (mode%8)*4 + site%4
    cplx* Y;      // [njobs][Lz][nmodes][Ne][Ne]
    const int* slotmode;  // folded variant only: [mbtot][8] compact mode index held by a weight-tile row, -1 = unused
};
struct PwTma {
    alignas(64) unsigned char mapL[128];  // CUtensorMap over [nfield][Ne][2*Kc doubles], box 8 x (8 EL) x 1
    alignas(64) unsigned char mapR[128];  // box 8 x (8 FL) x 1
    int nstages;
};
struct PwFold {
    const GramJob* jobs;
    int njobs, Ne, Lz, nmodes, nmom_int;
    int rows_l, rows_r;   // tile shape of the plane kernel (self pairs: tiles below the diagonal are mirror reads)
    const cplx* Y;
    const cplx* zphase;   // [nmom_int][Lz]
    const int* momode;    // [nmom_int][3]: cos mode, sin mode (-1: none), sigma
    cplx* partial;        // [njobs][nmom_int][Ne][Ne] (split 0 of the partial-sum buffer)
};

// ---- separable contraction (algo 4, edk_gram_sep.cu) ------------------------------------------
// The phase also factorises between x and y.  With the modes taken about the centre of the lattice,
//   exp(2 pi i (px x/Lx + py y/Ly)) = k_p (c_|px|(x) + i sgn(px) s_|px|(x)) (c_|py|(y) + i sgn(py) s_|py|(y)),
//   c_q(x) = cos(pi q (2x - Lx + 1)/Lx),  s_q(x) = sin(pi q (2x - Lx + 1)/Lx),  k_p a constant of modulus 1,
// the transform of the site product C over a plane is done row by row in registers, with plain DFMAs:
//   x stage, per pair of sites (x, Lx-1-x) of a row:  S = C(x) + C(xbar), D = C(x) - C(xbar),
//       X[1] += S,  X[c_q] += c_q(x) S,  X[s_q] += s_q(x) D                      (2 QMAX + 1 complex accumulators)
//   y stage, once per row:  Y[(m, c_q')] += c_q'(y) X[m],  Y[(m, s_q')] += s_q'(y) X[m],  X = 0
// i.e. 12 (site product) + 7 FP64 operations per (pair, e, f, site) for the 33 momenta with |p|^2 <= 4 against
// 12 + 2 + 16 of the folded plane-wave form (DMMA and DFMA share one FP64 pipe at the same rate on B200, so the
// operation count is the cost).  Every lane owns its (e, f) elements privately - no MMA fragments - and all lanes of
// a warp walk the same sites, so operands are broadcast reads of 128-byte-swizzled TMA tiles.
// Y[job][z][m][e][f] holds the NM separable modes (m = (x mode, y mode), SepModes below); sep_zfold_kernel folds z
// and recombines them into the momenta.
constexpr int SEP_WARPS = 8;              // compute warps: 4 along e x 2 along f, warp tile 4 x 16
constexpr int SEP_TE = 16, SEP_TF = 32;   // CTA tile: rows of L x rows of R
constexpr int SEP_MAX_PR = 64;            // site pairs per row (Lx/2) the shared-memory weight table holds
constexpr int SEP_DEFAULT_VARIANT = 6;    // kernel variant of a new handle (launch_gram_sep); EDK_SEP_VARIANT overrides (A/B hook)

// CTA tiles of gram_sepx_kernel: 8 compute warps arranged we x wf over (8 we) rows of L and (16 wf) rows of R.  The host
// covers the Ne x Ne elements with 32 x 32 tiles where they fit and with 8 x 128 / 64 x 16 tiles along the ragged edges,
// so that a CTA with idle warps is the exception at any Ne (sep_build_tiles).
struct SepTile {
    int e0, f0;
    int shape;   // 0: 4 x 2 warps = 32 x 32, 1: 1 x 8 = 8 x 128, 2: 8 x 1 = 64 x 16
    int e1, f1;  // the tile's elements are e0 <= e < e1, f0 <= f < f1 (clipped to its region and to Ne)
    int pad;
};
struct SepParams {
    const GramJob* jobs;
    int njobs;
    int Ne;
    int Lx, Ly, Lz;
    int SR;          // stages per row = (Lx/2) / pairs per stage
    int n_et, n_ft;  // tiles of SEP_TE x SEP_TF
    int nmodes;      // separable xy-modes kept in Y (5 / 9 / 13)
    const double* wx;  // [Lx/2][4]: c_1, s_1, c_2, s_2 at the front site x of pair (x, Lx-1-x)
    const double* wy;  // [Ly][4]: c_1, s_1, c_2, s_2 at y
    cplx* Y;           // [njobs][Lz][nmodes][Ne][Ne]
    const SepTile* tiles;  // gram_sepx_kernel: the CTA tiles (of the launch's shape) of one (job, plane); grid = njobs * Lz * ntiles
    int ntiles;
};
constexpr int SEP_NSHAPES = 3;
struct SepTmaX {
    alignas(64) unsigned char mapL[SEP_NSHAPES][128];  // boxes 16 doubles x {32, 8, 64} rows, 128-byte swizzle
    alignas(64) unsigned char mapR[SEP_NSHAPES][128];  // boxes 16 doubles x {32, 128, 16} rows
    int nstages[SEP_NSHAPES];
};
struct SepWeights {  // the x weights again, as a kernel parameter: read through the constant cache
    double w[SEP_MAX_PR * 4];
};
struct SepTma {
    alignas(64) unsigned char mapL[128];  // CUtensorMap over [nfield][Ne][6V doubles], box 16 x SEP_TE x 1, 128-byte swizzle
    alignas(64) unsigned char mapR[128];  // box 16 x SEP_TF x 1
    int nstages;
};
// z fold: one block folds one class of momenta (same |px|, |py|: they read the same <= 4 separable modes)
struct SepClass {
    int mode[4];  // cc, cs, sc, ss mode index in Y, -1 = absent
    int first, count;  // its momenta: entries [first, first + count) of `mom`
};
struct SepFold {
    const GramJob* jobs;
    int njobs, Ne, Lz, nmodes, nmom_int;
    int rows_l, rows_r;      // tile shape of the plane kernel (self pairs: tiles below the diagonal are mirror reads)
    const cplx* Y;
    const cplx* zphase;      // [nmom_int][Lz]: exp(2 pi i pz z/Lz) k_p
    int nclass;
    const SepClass* classes;
    const int* mom;          // per entry: internal momentum index, sgn(px), sgn(py); ascending momentum index inside a class
    cplx* partial;           // [njobs][nmom_int][Ne][Ne]
};

// ---- launchers (defined in the .cu files) -------------------------------------------------
// prepare
cudaError_t launch_round_eigvecs(const void* V_in, int flags, cplx* W0, double* W0_sum, size_t n_cplx, size_t row,
                                 size_t sum_row, cudaStream_t s);
cudaError_t launch_reorder_links(const cplx* U_in, int layout, int big_endian, cplx* U_out, Geom g, cudaStream_t s);
cudaError_t launch_phase_table(cplx* phase, cplx* rot, const int* mom3_dev, int nmom, Geom g, cudaStream_t s);
// stencil
cudaError_t launch_nabla3(const cplx* W_in, cplx* out_x, cplx* out_y, cplx* out_z, double* sum_x, double* sum_y, double* sum_z,
                          size_t sum_row, const cplx* links, Geom g, int Ne, cudaStream_t s);
struct Ptr6 {
    const cplx* src[6];
    cplx* dst[6];
};
cudaError_t launch_displace_step6(Ptr6 p, cplx* mean_out, const cplx* links, Geom g, int Ne, cudaStream_t s);
cudaError_t launch_laplacian(const cplx* F, cplx* out, const cplx* links, Geom g, int nvec, cudaStream_t s);
// gauge preprocessing
cudaError_t launch_stout_step(const cplx* Uin, cplx* Uout, double rho, Geom g, cudaStream_t s);
cudaError_t launch_project_su3(cplx* U, Geom g, cudaStream_t s);
// contraction
cudaError_t launch_gram_dmma(const GramParams& P, int mfrag, cudaStream_t s);
cudaError_t launch_gram_naive(const GramParams& P, cudaStream_t s);
cudaError_t launch_gram_tma(const GramParams& P, const GramTma& T, int mfrag, int algo, cudaStream_t s);
cudaError_t launch_phase_tiles(const cplx* phase2, cplx* tiles, int nmom, int Vpad, cudaStream_t s);
int gram_tma_plan(int algo, int mfrag, int nmom, int Ne, int* brows_alloc, int* nstages, int* smem_bytes);
int gram_fwidth(int algo);
int gram_pick_mfrag(int Ne);
int gram_rows_per_tile(int mfrag);
bool gram_mfrag_available(int mfrag);
int gram_nfrag_per_tile(int algo);
cudaError_t launch_combine(const CombineOp* ops_dev, int nop, const cplx* partial, int njobs, int ksplit, int nmom_int,
                           int nmom_out, const int* pmap, const int* negidx, int n_half, int Ne, const double* coeff, cplx* out,
                           cudaStream_t s);
// plane-wave factorised contraction
int pw_plan_smem(int el, int fl, int* nstages, int* smem_bytes);
void pw_pick_tile(int Ne, bool folded, int* el, int* fl);
bool pw_tile_available(int el, int fl);
cudaError_t launch_pw_weights(double* wtiles, const int* modes3_dev, int nmodes, int mbtot, int kplane, Geom g, cudaStream_t s);
cudaError_t launch_gram_pw(const PwParams& P, const PwTma& T, int MB, int el, int fl, cudaStream_t s);
// folded variant (centre-symmetric site pairs): m-blocks come in (cos, sin) pairs, one pair per pass
int pwf_plan_smem(int el, int fl, int* nstages, int* smem_bytes);
cudaError_t launch_gram_pwf(const PwParams& P, const PwTma& T, int el, int fl, cudaStream_t s);
cudaError_t launch_pwf_weights(double* wtiles, const int* modes3_dev, const int* slotmode_dev, int mbtot, int kplane, Geom g,
                               cudaStream_t s);
cudaError_t launch_pw_zfold(const PwFold& F, cudaStream_t s);
// separable contraction: qmax / r2 select the mode structure (max |px|,|py| and max px^2 + py^2 covered),
// pairs = site pairs per stage (8, 6 or 4, a divisor of Lx/2)
int sep_num_modes(int qmax, int r2);                    // 0 = structure not instantiated
int sep_mode_index(int qmax, int r2, int qx, int xk, int qy, int yk);  // xk, yk: 0 cos, 1 sin; -1 = not in the structure
int sep_plan_smem(int* nstages, int* smem_bytes);
// variant 0 = accumulators in registers, 1 x 2 elements per lane (gram_sep_kernel); variant 6 = y-stage accumulators in tensor
// memory, 2 x 2 elements per lane, tile table (gram_sepx_kernel, launch_gram_sepx)
int sep_variant_rows(int variant);  // rows of L per CTA tile of variant 0 (16); variant 6: 8, the unit of its tile table
int sepx_shape(int shape, int* we, int* wf);                       // warp arrangement of a tile shape; -1 if unknown
int sepx_plan_smem(int shape, int* nstages, int* smem_bytes);      // ring depth of a shape, dynamic shared memory of the kernel
cudaError_t launch_gram_sepx(const SepParams& P, const SepTmaX& T, const SepWeights& W, int qmax, int r2, int pairs, int shape, cudaStream_t s);
int sep_variant_plan(int variant, int* nstages, int* smem_bytes);
cudaError_t launch_gram_sep(const SepParams& P, const SepTma& T, const SepWeights& W, int qmax, int r2, int pairs, int variant, cudaStream_t s);
cudaError_t launch_sep_zfold(const SepFold& F, cudaStream_t s);
// microbench
cudaError_t microbench_fp64(double* dmma_tflops, double* dfma_tflops);

}  // namespace edk

#ifdef EDK_HOST_EMU
#include "edk_emu.h"  // tests/emu: host implementations of threadIdx, barriers, launches (test builds only)
#endif
