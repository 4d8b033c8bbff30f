// Shared declarations of the elemental-distillation kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "edk.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "edk kernels are written for sm_100a (B200) only"
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
// microbench
cudaError_t microbench_fp64(double* dmma_tflops, double* dfma_tflops);

}  // namespace edk
