/*
 * edk.h - C ABI of the B200 elemental-distillation kernels (libedk_sm100a.so).
 *
 * This is the drop-in boundary for ONE path of IHEP-LQCD/EasyDistillation:
 * per-timeslice elemental generation.  The reference has no FFI for it (the
 * path is pure numpy/cupy behind two Python classes), so every entry point
 * cites the reference interface it stands in for; `INTEGRATION.md` shows the
 * ctypes stub a reference maintainer would add.  Paths are relative to the
 * reference root.
 *
 * Conventions
 *   - complex128 = two doubles (re, im); complex64 = two floats.
 *   - all arrays are C-contiguous, slowest index first.
 *   - "device" pointers are CUDA device pointers on the handle's device,
 *     "host" pointers are ordinary (ideally page-locked) host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns EDK_OK (0) or a negative error; edk_last_error()
 *     gives the message of the calling thread's last failure.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef EDK_H
#define EDK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDK_OK 0
#define EDK_ERR_ARG -1    /* bad argument (-> ValueError on the Python side) */
#define EDK_ERR_CUDA -2   /* CUDA runtime / launch failure                  */
#define EDK_ERR_STATE -3  /* call order violated (inputs not set)           */
#define EDK_ERR_NOMEM -4  /* workspace does not fit                         */

#define EDK_MODE_DERIVATIVE 0   /* ElementalGenerator              */
#define EDK_MODE_DISPLACEMENT 1 /* DisplacementElementalGenerator  */

/* link layouts accepted by edk_set_links / edk_calc_host */
#define EDK_LINKS_DIR_MAJOR 0 /* [3][Lz][Ly][Lx][3][3]   = the reference's U[:, t] made contiguous            */
#define EDK_LINKS_FILE_T 1    /* [Lz][Ly][Lx][4][3][3]   = one timeslice of the file order, time links skipped */
/* OR-ed into `layout`: the doubles are big-endian, as in the payload of an ILDG file; they are
 * byte-swapped on the device instead of the host-side `.astype("<c16")` of filedata/ildg.py:70 */
#define EDK_LINKS_BIG_ENDIAN 0x100

/* eigenvector input flags of edk_set_eigvecs / edk_calc_host (the `is_c8` argument) */
#define EDK_EIGVECS_C8 1         /* complex64 input (else complex128)                                         */
#define EDK_EIGVECS_BIG_ENDIAN 2 /* big-endian words, as in a QDP timeslice file (filedata/timeslice.py:96)   */

typedef struct edk_handle edk_handle;

/* library ABI version (bumped on any signature change) */
int edk_version(void);
const char* edk_last_error(void);

/*
 * Constructor.  Replaces the numerical part of
 *   ElementalGenerator.__init__            lattice/generator/elemental.py:17-59
 *   DisplacementElementalGenerator.__init__ lattice/generator/displacement_elemental.py:12-51
 * mode/order: EDK_MODE_DERIVATIVE with order = num_nabla (0..3), or
 *             EDK_MODE_DISPLACEMENT with order = distance (>= 0).
 * mom3: nmom integer triples (px,py,pz); the phase is exp(+2 pi i p.x/L)
 *       (lattice/insertion/phase.py:11-13,41-46).
 * Allocates the workspace (derived fields, phase tables, partial sums) on `device`.
 */
int edk_create(int Lx, int Ly, int Lz, int Ne, int mode, int order, int nmom, const int* mom3, int device,
               edk_handle** out);
int edk_destroy(edk_handle* h);

/*
 * MomentumPhase.get for a list of momenta (lattice/insertion/phase.py:41-46):
 * out_dev[ip][z][y][x] = exp(+2 pi i (px x/Lx + py y/Ly + pz z/Lz)), complex128, no handle needed.
 */
int edk_phase_table(int Lx, int Ly, int Lz, int nmom, const int* mom3, void* out_dev, int device, void* stream);

/*
 * Contraction plan for a momentum list, pure host logic (no device needed): out[0] Hermitian pairing
 * G(L,R,p)^dagger = G(R,L,-p) used (sym_request: -1 auto by cost, 0 never, 1 always), out[1] internal momenta
 * (the caller's plus missing negatives), out[2] half set contracted by self pairs L == R, out[3]/out[4] distinct
 * (left, right) pairs without / with the pairing (34 / 19 for num_nabla = 2; the reference runs 43,
 * lattice/generator/elemental.py:309-329), out[5] (pair, momentum) GEMMs per timeslice, out[6] output operators,
 * out[7] self pairs.
 */
int edk_plan(int mode, int order, int nmom, const int* mom3, int sym_request, int out[8]);

/*
 * Mode plan of the plane-wave factorised contraction (edk_debug_algo 2), pure host logic: the xy-part of
 * exp(2 pi i p.x/L) (lattice/insertion/phase.py:41-46) is cos(theta_q) + i sigma sin(theta_q) for the couple
 * {+q, -q} of (px, py).  nmodes real modes; modes3[m] = (qx, qy, kind: 0 cos / 1 sin), room for 2*nmom triples;
 * momode[i] = (cos mode, sin mode or -1, sigma) of momentum i.
 */
int edk_plan_modes(int nmom, const int* mom3, int* nmodes, int* modes3, int* momode);

/*
 * Form of the contraction a handle of this lattice / operator set / momentum list uses, pure host logic (no device
 * needed).  The reference has one form (one einsum per (split, momentum), lattice/generator/elemental.py:322-329); here
 * edk_create picks, from the FP64-pipe work per (e, f, site) of each form, between
 *   1 the GEMM form (3M arithmetic on DMMA; cheapest for one or two momenta),
 *   3 the plane-wave form with folded site pairs (site product once per site, real xy-modes by DMMA; any momentum list,
 *     planes of at least 8 sites),
 *   4 the separable form (x and y transforms of the site product in registers, row by row; Lx even and >= 8 with Lx/2 a
 *     multiple of 4, 6 or 8, |px|, |py| <= 2 and px^2 + py^2 <= 4).
 * out[0] form, out[1] separable form available (0/1), out[2] its max |px|,|py| (1 or 2), out[3] its max px^2 + py^2,
 * out[4] site pairs per stage (8, 6, 4), out[5] separable xy-modes (5, 9, 13).
 */
int edk_plan_form(int Lx, int Ly, int mode, int order, int nmom, const int* mom3, int out[6]);

/*
 * CTA tiles the separable contraction covers the Ne x Ne elements of one (pair, plane) with, pure host logic:
 * tiles4[i] = (first row, first column, rows, columns) of tile i (32 x 32 where whole tiles fit, up to 8 x 128 and
 * 64 x 16 along the ragged edges); every element lies in exactly one tile.  Returns the number of tiles (writes at
 * most max_tiles of them).
 */
int edk_plan_tiles(int Ne, int max_tiles, int* tiles4);

/* number of operators in the output: (3^(num_nabla+1)-1)/2 or distance+1 (elemental.py:48, displacement_elemental.py:45) */
int edk_num_operators(const edk_handle* h);
/* bytes of one timeslice result [Nop][Nmom][Ne][Ne] complex128 */
size_t edk_output_bytes(const edk_handle* h);
size_t edk_workspace_bytes(const edk_handle* h);

/*
 * Inputs of one timeslice, device pointers.  Replace `U[:, t]` (elemental.py:103,320)
 * and the per-eigenvector staging loop `V[e] = eigenvector[t, e]` (elemental.py:297-298,
 * displacement_elemental.py:88-89).  Eigenvectors are [Ne][Lz][Ly][Lx][3]; is_c8 = 1 for
 * complex64 input, 0 for complex128 input, which is value-rounded through complex64 on
 * the device exactly as the reference's complex64 `_V` buffer does (elemental.py:55).
 * `is_c8` is a flag word: EDK_EIGVECS_C8 | EDK_EIGVECS_BIG_ENDIAN; `layout` may carry
 * EDK_LINKS_BIG_ENDIAN.  Big-endian input is the raw file payload, swapped on the device.
 */
int edk_set_links(edk_handle* h, const void* U_dev, int layout, void* stream);
int edk_set_eigvecs(edk_handle* h, const void* V_dev, int is_c8, void* stream);

/*
 * Gauge preprocessing of the reference classes, applied on the device to the links of every
 * timeslice right after edk_set_links / inside edk_calc_host, in the order given:
 *   kinds[i] = 1: stout_smear(nsteps[i], rhos[i])  spatial links only, 3-d staples
 *                 (lattice/generator/elemental.py:175-277, lattice/generator/stout_smear.cu:159-245)
 *   kinds[i] = 2: project_SU3()                    X <- (X + X^-dagger)/2 to 1e-15 (elemental.py:107-117)
 * nops = 0 clears the list.  (Spatial smearing never couples timeslices, so per-timeslice
 * application equals the reference's whole-configuration call after load().)
 */
int edk_set_link_ops(edk_handle* h, int nops, const int* kinds, const int* nsteps, const double* rhos);

/* optional real [Ne][Ne] "blending" matrix multiplied into every output block
 * (stocastic_coeff, elemental.py:61-100,331-337); NULL clears it. Device pointer, copied. */
int edk_set_blending(edk_handle* h, const double* coeff_dev, void* stream);

/*
 * calc(t): elemental.py:290-338 / displacement_elemental.py:78-96.
 * Writes [Nop][Nmom][Ne][Ne] complex128 to out_dev. Asynchronous on `stream`.
 */
int edk_calc(edk_handle* h, void* out_dev, void* stream);

/*
 * Same with HOST buffers: copies links (layout as above) and eigenvectors to the
 * device, runs calc, copies the result back, and synchronises `stream` before
 * returning.  This is the call `bench.py`'s end-to-end figure times.
 */
int edk_calc_host(edk_handle* h, const void* U_host, int layout, const void* V_host, int is_c8, void* out_host,
                  void* stream);

/*
 * The eigensolver's operator on the same links (lattice/generator/eigenvector.py:11-26, `_Laplacian`):
 *   out[v] = 6 F[v] - sum_d [ U_d(x) F[v](x+d) + U_d(x-d)^dagger F[v](x-d) ],  v < nvec,
 * F and out are [nvec][Lz][Ly][Lx][3] complex128 device arrays (vector index slowest; the reference keeps
 * it fastest), out != F.  Uses the links of the last edk_set_links, gauge preprocessing included.
 */
int edk_laplacian(edk_handle* h, const void* F_dev, void* out_dev, int nvec, void* stream);

/* page-locked host memory for the buffers above (cudaHostAlloc / cudaFreeHost) */
int edk_host_alloc(void** p, size_t bytes);
int edk_host_free(void* p);
/* page-lock an existing host range in place (cudaHostRegister) so its timeslices can be uploaded by DMA
 * without a staging copy; fails harmlessly (EDK_ERR_CUDA) for ranges that cannot be locked */
int edk_host_register(void* p, size_t bytes);
int edk_host_unregister(void* p);

/*
 * Measurement hooks.  With profiling on, every kernel class of the next edk_calc is
 * bracketed by CUDA events on the launching stream; edk_get_profile synchronises
 * and returns the milliseconds and launch counts of the LAST calc.
 *   ms[0] prepare (complex64 rounding / link reorder)   ms[1] stencil (nabla or displacement step)
 *   ms[2] contraction (DMMA)                            ms[3] combine / reduce
 * n_launch[i] likewise.
 */
int edk_set_profiling(edk_handle* h, int on);
int edk_get_profile(edk_handle* h, double ms[4], int n_launch[4]);
/* kernels launched by this handle since creation */
long long edk_launch_count(const edk_handle* h);

/*
 * Test hooks (used by tests/ only).
 * edk_debug_field: copy derived field `idx` ([Ne][Lz][Ly][Lx][3] complex128) of the last calc
 *   to dst_dev.  Derivative mode: idx = (3^len-1)/2 + base-3 value of the direction sequence
 *   in application order (0 = W0, 1..3 = nabla_a W0, 4+3a+b = nabla_b nabla_a W0).
 *   Displacement mode: idx = k -> D_k.
 * edk_debug_phase: copy the phase table of momentum ip ([Lz][Ly][Lx] complex128).
 * edk_debug_use_naive_gram: 1 = run the scalar one-thread-per-output contraction instead of
 *   the DMMA kernel (cross-check only; never enabled by the product path).
 * edk_debug_gram_config: force the DMMA tile variant (m-frags per warp) and split-K factor; 0 = auto.
 * edk_debug_symmetry: -1 = auto (default), 0 = contract every (left, right) pair directly,
 *   1 = force the Hermitian pairing G(L,R,p)^dagger = G(R,L,-p) (missing -p are added internally).
 * edk_debug_loader: 0 = TMA producer warp + mbarrier ring (default), 1 = cp.async ring executed by the MMA warps.
 * edk_debug_algo: arithmetic of the TMA kernel, 1 = 3M (default: Re = T1+T2, Im = T3+T1-T2 with
 *   T1 = Lr.Pr, T2 = Li.Pi, T3 = (Lr+Li).(Pi-Pr): three real MMAs per complex block), 0 = 4M (four),
 *   2 = plane-wave factorised form: site products conj(L).R formed once per site, real xy-mode transform by
 *   DMMA per z-plane, z folded by a second kernel (csrc/edk_gram_pw.cu), 3 = form 2 with centre-symmetric site pairs
 *   folded (modes about the centre of the plane are even / odd under s -> A-1-s: the sum of the two site products
 *   feeds the cos modes, their difference the sin modes - half the DMMAs per site), 4 = separable form
 *   (csrc/edk_gram_sep.cu: x transform per pair of sites (x, Lx-1-x) and y transform per row in registers, plain DFMAs;
 *   EDK_ERR_ARG where edk_plan_form says it is not available), -1 = back to the planned form.  A handle starts with the
 *   form edk_plan_form plans; the environment variable EDK_GRAM_ALGO=0..4 asks for one form in every new handle
 *   instead (A/B measurements only; nothing in the package sets it).
 * edk_query: what = 0 pairing in use (0/1), 1 internal momentum count, 2 pair-GEMMs per momentum,
 *   3 split-K factor, 4 m-fragments per tile, 5 contraction jobs, 6 TMA ring depth (0 = cp.async loader),
 *   7 real MMAs per complex block (3 or 4), 8 number of (pair, momentum) GEMMs contracted per timeslice
 *   (self pairs L == R only run one momentum of every +-p couple), 9 size of that half set,
 *   10 contraction form in use (0 / 1 / 2 / 3 as in edk_debug_algo), 11 real xy-modes of forms 2 / 3 (0 = not built),
 *   12 tile shape of forms 2 / 3 as 10 el + fl (24 = 16 x 32 rows, 25 = 16 x 40, 17 = 8 x 56; environment EDK_PW_TILE overrides the pick, EDK_PW_STAGES = 2.. caps the depth of the operand ring: A/B hooks).
 *   With forms 2 / 3, what = 7 answers 2 / 1 (DMMAs per site, real or imaginary part and block of 8 modes), with
 *   form 4 it answers 0 (no MMA), 11 the separable xy-modes and 12 the tile as 100 rows_L + rows_R (1632).
 *   13 the form asked for (-1 = planned), 14 site pairs per stage of form 4.
 */
int edk_debug_field(edk_handle* h, int idx, void* dst_dev, void* stream);
int edk_debug_phase(edk_handle* h, int ip, void* dst_dev, void* stream);
int edk_debug_links(edk_handle* h, void* dst_dev, void* stream); /* processed links [3][Lz][Ly][Lx][3][3] */
int edk_debug_use_naive_gram(edk_handle* h, int on);
int edk_debug_gram_config(edk_handle* h, int mfrag, int ksplit);
int edk_debug_symmetry(edk_handle* h, int mode);
int edk_debug_loader(edk_handle* h, int mode);
int edk_debug_algo(edk_handle* h, int algo);
int edk_query(const edk_handle* h, int what);

/*
 * Stand-alone micro-benchmarks (bench.py's roofline denominators): sustained
 * DMMA.8x8x4 and DFMA throughput of the device in TFLOP/s, measured with CUDA events.
 */
int edk_microbench_fp64(int device, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* EDK_H */
