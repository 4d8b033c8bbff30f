// C ABI (include/edk.h) over the sm_100a kernels: handle, workspace, job lists, calc flows.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include <cuda.h>

#include "edk_common.cuh"

namespace edk {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

enum { PH_PREP = 0, PH_STENCIL = 1, PH_GRAM = 2, PH_COMBINE = 3, PH_COUNT = 4 };

struct EventPair {
    cudaEvent_t a, b;
    int phase;
};

}  // namespace edk

using namespace edk;

struct edk_handle {
    Geom g;
    int Ne, mode, order, nmom, device;  // nmom = momenta the caller asked for (output)
    int nop;
    // internal momentum list: the caller's momenta first, then any missing negatives when the
    // Hermitian pairing G(L,R,p)^dagger = G(R,L,-p) is used to halve the number of contracted pairs
    int nmom_int = 0;
    bool symmetric = false;
    int sym_request = -1;  // -1 auto, 0 off, 1 on (test hook)
    std::vector<int> mom_user, mom_int, negidx, pmap;
    int n_half = 0;  // the first n_half internal momenta hold one of every +-p couple (pairing mode)
    int* negidx_dev = nullptr;
    int* pmap_dev = nullptr;
    int2* cta_map_dev = nullptr;
    int ncta = 0;
    bool cta_dirty = true;
    // TMA-fed contraction (default); loader = 1 selects the cp.async kernel (A/B comparison hook)
    GramTma tma{};
    cplx* phase_tiles = nullptr;
    bool tma_ready = false;
    int loader = 0;
    int algo = 1;  // contraction: 1 = GEMM form, 3M arithmetic (three real MMAs per complex block), 0 = GEMM form, 4M,
                   // 2 = plane-wave factorised form (edk_gram_pw.cu), 3 = the same with centre-symmetric site pairs folded,
                   // 4 = separable form (edk_gram_sep.cu): x and y transforms in registers, row by row
    // plane-wave factorised contraction (algo 2): xy-mode weights, per-plane sums Y, z phases
    PwTma pw_tma{};
    bool pw_ready = false;
    int pw_algo = 0;              // the form (2 or 3) the tables below were built for
    bool pw_fold = false;         // the tables are those of the folded kernel (form 3 on planes of at least 8 sites)
    int pw_npass = 0;             // launches of the plane kernel per timeslice
    int* pw_slotmode = nullptr;   // form 3: [mbtot][8] compact mode index of every weight-tile row (-1 = unused)
    int pw_nmodes = 0, pw_mbtot = 0, pw_kplane = 0;
    int pw_el = 2, pw_fl = 4;     // tile shape: (8 el) rows of L x (8 fl) rows of R per CTA
    double* pw_wtiles = nullptr;  // [kplane][2][mbtot][32]
    cplx* pw_Y = nullptr;         // [njobs][Lz][nmodes][Ne][Ne]
    cplx* pw_zphase = nullptr;    // [nmom_int][Lz]
    int* pw_momode = nullptr;     // [nmom_int][3]: cos mode, sin mode, sigma
    size_t pw_bytes = 0;
    // separable contraction (algo 4): x / y weight tables, per-plane sums of the separable modes, fold tables
    SepTma sep_tma{};
    bool sep_ready = false;
    int sep_qmax = 0, sep_r2 = 0, sep_pairs = 0, sep_nmodes = 0, sep_nclass = 0;
    int sep_variant = 0;  // kernel variant (edk_gram_sep.cu: launch_gram_sep)
    bool sep_per_shape = false;  // gram_sepx: one launch per tile shape instead of one over the whole tile table (A/B)
    SepWeights sep_wx_host{};  // the x weights as a kernel parameter (constant cache)
    SepTmaX sep_tmax{};        // variant 6: tensor maps and ring depth per tile shape
    SepTile* sep_tiles = nullptr;  // grouped by shape
    int sep_ntiles = 0;
    int sep_shape_first[SEP_NSHAPES] = {0, 0, 0}, sep_shape_count[SEP_NSHAPES] = {0, 0, 0};
    double* sep_wx = nullptr;      // [Lx/2][4]
    double* sep_wy = nullptr;      // [Ly][4]
    cplx* sep_Y = nullptr;         // [njobs][Lz][nmodes][Ne][Ne]
    cplx* sep_zphase = nullptr;    // [nmom_int][Lz]
    SepClass* sep_classes = nullptr;
    int* sep_mom = nullptr;
    size_t sep_bytes = 0;
    int algo_request = -1;  // -1 = pick the form from the cost plan (plan_contraction_form), else the form asked for
    size_t field_cplx;  // Ne * V * 3
    // device buffers
    cplx* links = nullptr;    // [3][V][9]
    cplx* links_tmp = nullptr;  // second buffer for the stout ping-pong (allocated on first use)
    struct LinkOp {
        int kind;  // 1 = stout(nstep, rho), 2 = project_SU3
        int nstep;
        double rho;
    };
    std::vector<LinkOp> link_ops;  // applied, in order, to every timeslice's links after upload
    cplx* fields = nullptr;   // nfield fields
    double* fsum = nullptr;   // Re + Im of every field element, [nfield][Ne][3V]: the 3M contraction's third A operand
    int nfield = 0;
    cplx* lines = nullptr;    // displacement: 12 line buffers (ping-pong)
    cplx* phase = nullptr;    // [2][nmom][Vpad]: phase and -i*phase
    cplx* partial = nullptr;  // [ksplit][njobs][nmom][Ne][Ne]
    double* coeff = nullptr;  // [Ne][Ne] or null
    bool have_coeff = false;
    GramJob* jobs_dev = nullptr;
    CombineOp* ops_dev = nullptr;
    int njobs = 0;
    int ksplit = 1, mfrag = 13;
    int force_mfrag = 0, force_ksplit = 0;
    bool naive = false;
    // staging for the host-buffer entry point
    void* stage_U = nullptr;
    size_t stage_U_bytes = 0;
    void* stage_V = nullptr;
    size_t stage_V_bytes = 0;
    cplx* stage_out = nullptr;
    // state
    bool links_set = false, evecs_set = false;
    size_t ws_bytes = 0, cfg_bytes = 0;
    long long launches = 0;
    // profiling
    bool profiling = false;
    std::vector<EventPair> events;
    std::vector<EventPair> pool;
    int n_launch[PH_COUNT] = {0, 0, 0, 0};
    // derivative-mode stencil schedule: (source field, first child field)
    std::vector<std::pair<int, int>> hops;
    std::vector<GramJob> jobs_host;
    std::vector<CombineOp> ops_host;

    cplx* field(int i) const { return fields + (size_t)i * field_cplx; }
    size_t sum_row = 0;       // doubles per eigenvector row of a sum plane: 3V rounded up to even (16-byte TMA strides)
    double* field_sum(int i) const { return fsum + (size_t)i * Ne * sum_row; }
};

namespace {

struct PhaseTimer {
    edk_handle* h;
    cudaStream_t s;
    int phase;
    EventPair ev{};
    bool on;
    PhaseTimer(edk_handle* h_, cudaStream_t s_, int phase_, int launches) : h(h_), s(s_), phase(phase_), on(h_->profiling) {
        h->launches += launches;
        h->n_launch[phase] += launches;
        if (!on) return;
        if (!h->pool.empty()) {
            ev = h->pool.back();
            h->pool.pop_back();
        } else {
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
        }
        ev.phase = phase;
        cudaEventRecord(ev.a, s);
    }
    ~PhaseTimer() {
        if (!on) return;
        cudaEventRecord(ev.b, s);
        h->events.push_back(ev);
    }
};

// Every entry point that launches or allocates runs on the handle's device and leaves the caller's current device as
// it found it (two generators on different GPUs in one process).
struct DeviceGuard {
    int prev = -1, dev;
    bool ok = true;
    explicit DeviceGuard(int dev_) : dev(dev_) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) {
            ok = false;
            cudaGetLastError();
        }
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define EDK_ON_DEVICE(h)                                                  \
    DeviceGuard _guard((h)->device);                                      \
    if (!_guard.ok) {                                                     \
        set_error("cudaSetDevice(%d) failed", (h)->device);               \
        return EDK_ERR_CUDA;                                              \
    }

// field index of a direction sequence in application order (see edk.h, edk_debug_field)
int seq_index(const std::vector<int>& seq) {
    int base = 1, off = 0;
    for (size_t i = 0; i < seq.size(); ++i) {
        off += base;
        base *= 3;
    }
    int v = 0;
    for (int d : seq) v = v * 3 + d;
    return off + v;
}

// the reference's derivative(n): lattice/insertion/derivative.py:23-33
std::vector<int> derivative_tuple(int n) {
    int order = 0, p = 1;
    while (n >= p) {
        n -= p;
        p *= 3;
        ++order;
    }
    std::vector<int> digits(order);
    for (int i = 0; i < order; ++i) {
        digits[order - 1 - i] = n % 3;
        n /= 3;
    }
    return digits;
}

int pow3sum(int n) {  // (3^(n+1)-1)/2
    int s = 0, p = 1;
    for (int i = 0; i <= n; ++i) {
        s += p;
        p *= 3;
    }
    return s;
}

// Build the job list of the derivative elementals (elemental.py:299-329).
// Every output operator is a signed sum over left/right splits of G(L, R).
//  * plain mode: (L, R) pairs that occur in more than one operator (e.g. (nabla_a W0, nabla_b W0)
//    for num_nabla = 2) become shared single-segment jobs, the others are fused into one
//    multi-segment job per operator: 34 instead of the reference's 43 pair-GEMMs.
//  * symmetric mode: G(L,R,p)^dagger = G(R,L,-p), so only pairs with L >= R are contracted (19 for
//    num_nabla = 2, 4 for num_nabla = 1) and the combine step reads the conjugate transpose at -p
//    for the others.  Needs -p in the (internal) momentum list for every p.
void build_derivative_jobs(edk_handle* h) {
    struct Term {
        int L, R, sign;
    };
    std::vector<std::vector<Term>> per_op(h->nop);
    std::map<std::pair<int, int>, int> uses;
    for (int n = 0; n < h->nop; ++n) {
        std::vector<int> dirs = derivative_tuple(n);
        const int len = (int)dirs.size();
        for (int pick = 0; pick < (1 << len); ++pick) {
            std::vector<int> right, left;
            for (int i = 0; i < len; ++i) ((pick >> i) & 1 ? right : left).push_back(dirs[i]);
            std::reverse(left.begin(), left.end());
            Term t{seq_index(left), seq_index(right), (right.size() & 1) ? -1 : 1};
            per_op[n].push_back(t);
            ++uses[{t.L, t.R}];
        }
    }
    h->jobs_host.clear();
    h->ops_host.assign(h->nop, CombineOp{});
    auto add_term = [](CombineOp& o, int jid, double w, int herm, int half) {
        for (int k = 0; k < o.nterm; ++k)
            if (o.job[k] == jid && o.herm[k] == herm) {
                o.weight[k] += w;
                return;
            }
        o.job[o.nterm] = jid;
        o.weight[o.nterm] = w;
        o.herm[o.nterm] = herm;
        o.half[o.nterm] = half;
        ++o.nterm;
    };
    std::map<std::pair<int, int>, int> shared_job;
    auto single_job = [&](int L, int R) {
        auto key = std::make_pair(L, R);
        auto it = shared_job.find(key);
        if (it != shared_job.end()) return it->second;
        GramJob j{};
        j.nseg = 1;
        j.nmom = (h->symmetric && L == R) ? h->n_half : h->nmom_int;
        j.sign[0] = 1;
        j.L[0] = h->field(L);
        j.R[0] = h->field(R);
        j.Lf[0] = L;
        j.Rf[0] = R;
        const int jid = (int)h->jobs_host.size();
        h->jobs_host.push_back(j);
        shared_job[key] = jid;
        return jid;
    };
    if (h->symmetric) {
        for (int n = 0; n < h->nop; ++n)
            for (const Term& t : per_op[n]) {
                if (t.L >= t.R)
                    add_term(h->ops_host[n], single_job(t.L, t.R), t.sign, 0, t.L == t.R);
                else
                    add_term(h->ops_host[n], single_job(t.R, t.L), t.sign, 1, 0);
            }
    } else {
        // private multi-segment jobs first (longest first helps the tail of the grid)
        struct Pending {
            int op;
            GramJob job;
        };
        std::vector<Pending> priv;
        for (int n = 0; n < h->nop; ++n) {
            GramJob j{};
            j.nmom = h->nmom_int;
            for (const Term& t : per_op[n]) {
                if (uses[{t.L, t.R}] > 1) continue;
                j.sign[j.nseg] = t.sign;
                j.L[j.nseg] = h->field(t.L);
                j.R[j.nseg] = h->field(t.R);
                j.Lf[j.nseg] = t.L;
                j.Rf[j.nseg] = t.R;
                ++j.nseg;
            }
            if (j.nseg) priv.push_back({n, j});
        }
        std::stable_sort(priv.begin(), priv.end(),
                         [](const Pending& a, const Pending& b) { return a.job.nseg > b.job.nseg; });
        for (const Pending& p : priv) {
            add_term(h->ops_host[p.op], (int)h->jobs_host.size(), 1.0, 0, 0);
            h->jobs_host.push_back(p.job);
        }
        for (int n = 0; n < h->nop; ++n)
            for (const Term& t : per_op[n])
                if (uses[{t.L, t.R}] > 1) add_term(h->ops_host[n], single_job(t.L, t.R), t.sign, 0, 0);
    }
    // stencil schedule: every field of length < order spawns its three children
    h->hops.clear();
    const int nparents = h->order >= 1 ? pow3sum(h->order - 1) : 0;
    for (int f = 0; f < nparents; ++f) {
        // children of sequence s are s+(a): index = off(len+1) + 3*v(s) + a
        int len = 0, off = 0, p = 1;
        while (f >= off + p) {
            off += p;
            p *= 3;
            ++len;
        }
        const int v = f - off;
        const int child0 = off + p + 3 * v;
        h->hops.push_back({f, child0});
    }
}

// number of pair-GEMMs (segments) the two modes need, to decide which is cheaper
void count_pairs(int order, int& plain, int& sym, int* self_pairs = nullptr) {
    std::map<std::pair<int, int>, int> a, b;
    const int nop = pow3sum(order);
    for (int n = 0; n < nop; ++n) {
        std::vector<int> dirs = derivative_tuple(n);
        const int len = (int)dirs.size();
        for (int pick = 0; pick < (1 << len); ++pick) {
            std::vector<int> right, left;
            for (int i = 0; i < len; ++i) ((pick >> i) & 1 ? right : left).push_back(dirs[i]);
            std::reverse(left.begin(), left.end());
            const int L = seq_index(left), R = seq_index(right);
            a[{L, R}] = 1;
            b[{std::max(L, R), std::min(L, R)}] = 1;
        }
    }
    plain = (int)a.size();
    sym = (int)b.size();
    if (self_pairs) {
        *self_pairs = 0;
        for (const auto& kv : b) *self_pairs += kv.first.first == kv.first.second;
    }
}

void build_displacement_jobs(edk_handle* h) {
    h->jobs_host.clear();
    h->ops_host.assign(h->nop, CombineOp{});
    for (int k = 0; k < h->nop; ++k) {
        GramJob j{};
        j.nseg = 1;
        j.nmom = h->nmom_int;
        j.sign[0] = 1;
        j.L[0] = h->field(0);
        j.R[0] = h->field(k);  // field k = D_k (field 0 = W0 = D_0)
        j.Lf[0] = 0;
        j.Rf[0] = k;
        h->ops_host[k].nterm = 1;
        h->ops_host[k].job[0] = k;
        h->ops_host[k].weight[0] = 1.0;
        h->jobs_host.push_back(j);
    }
}

int effective_algo(const edk_handle* h) { return (h->loader == 0 && !h->naive) ? h->algo : 0; }

int count_n_tiles(const edk_handle* h, int algo, int nmom_job) {
    const int fw = gram_fwidth(algo), nt = gram_nfrag_per_tile(algo);
    const int nfrag_f = (h->Ne + fw - 1) / fw;
    return (nfrag_f * nmom_job + nt - 1) / nt;
}

int row_tiles(const edk_handle* h) {
    const int rows = gram_rows_per_tile(h->mfrag);
    return (h->Ne + rows - 1) / rows;
}

// CTA -> (job, tile inside the job).  Jobs keep their order (longest first), tiles of one job are
// contiguous with nt fastest, so CTAs that run together share the rows of L.
int build_cta_map(edk_handle* h) {
    const int algo = effective_algo(h);
    const int n_mt = row_tiles(h);
    std::vector<int2> map;
    for (int j = 0; j < h->njobs; ++j) {
        const int tiles = n_mt * count_n_tiles(h, algo, h->jobs_host[j].nmom);
        for (int t = 0; t < tiles; ++t) map.push_back(make_int2(j, t));
    }
    h->ncta = (int)map.size();
    cudaFree(h->cta_map_dev);
    h->cta_map_dev = nullptr;
    EDK_CUDA_TRY(cudaMalloc(&h->cta_map_dev, map.size() * sizeof(int2)));
    EDK_CUDA_TRY(cudaMemcpy(h->cta_map_dev, map.data(), map.size() * sizeof(int2), cudaMemcpyHostToDevice));
    return EDK_OK;
}

void pick_gram_config(edk_handle* h) {
    h->cta_dirty = true;
    h->mfrag = h->force_mfrag ? h->force_mfrag : gram_pick_mfrag(h->Ne);
    long long tiles = 0;
    for (const auto& j : h->jobs_host) tiles += (long long)row_tiles(h) * count_n_tiles(h, effective_algo(h), j.nmom);
    const int ksteps = h->g.Vpad / 8;
    int ks = 1;
    if (!h->force_ksplit) {
        // fill the 148 SMs at least ~4 times over, but keep >= 16 stages per CTA
        const long long want = 148LL * 4;
        if (tiles < want) ks = (int)((want + tiles - 1) / tiles);
        // CTAs that share row tiles drift apart over very long K ranges and lose their L2 sharing:
        // measured at 48^3 (13824 stages) 1424 -> 1390 ms with 4 K-segments; 32^3 (4096) is best unsplit
        ks = std::max(ks, (ksteps + 4095) / 4096);
        ks = std::min(ks, std::max(1, ksteps / 16));
        ks = std::min(ks, 64);
        if (h->algo >= 2 && h->loader == 0 && !h->naive) ks = 1;  // the plane-wave forms write split 0 only
    } else {
        ks = std::min(h->force_ksplit, ksteps);
    }
    h->ksplit = std::max(1, ks);
}

int ensure_partial(edk_handle* h) {
    const size_t need = (size_t)h->ksplit * h->njobs * h->nmom_int * h->Ne * h->Ne * sizeof(cplx);
    static_assert(sizeof(cplx) == 16, "complex128");
    if (h->partial) EDK_CUDA_TRY(cudaFree(h->partial));
    h->partial = nullptr;
    EDK_CUDA_TRY(cudaMalloc(&h->partial, need));
    return EDK_OK;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query, so the
// library does not link libcuda.
EncodeFn tensor_map_encoder() {
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return nullptr;
        }
        encode = (EncodeFn)fn;
    }
    return encode;
}

// Tensor maps of the TMA-fed contraction: the field array as [nfield][Ne][2*Kc] doubles, boxes of
// 8 doubles (4 complex k) x rows x 1 field.
int build_tma(edk_handle* h) {
    h->tma_ready = false;
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return EDK_ERR_CUDA;
    int rows = 0, nst = 0, bytes = 0;
    if (gram_tma_plan(h->algo, h->mfrag, h->nmom_int, h->Ne, &rows, &nst, &bytes) != 0) {
        set_error("no shared-memory plan for the TMA contraction (mfrag %d)", h->mfrag);
        return EDK_ERR_ARG;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    const cuuint64_t Kd = (cuuint64_t)2 * 3 * h->g.V;
    const cuuint64_t gdim[3] = {Kd, (cuuint64_t)h->Ne, (cuuint64_t)h->nfield};
    const cuuint64_t gstr[2] = {Kd * 8, Kd * 8 * (cuuint64_t)h->Ne};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t boxA[3] = {8, (cuuint32_t)gram_rows_per_tile(h->mfrag), 1};
    const cuuint32_t boxB[3] = {8, 8, 1};
    CUresult r = encode((CUtensorMap*)h->tma.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxA, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
        r = encode((CUtensorMap*)h->tma.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxB, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
        const cuuint64_t Ks = (cuuint64_t)3 * h->g.V;
        const cuuint64_t sdim[3] = {Ks, (cuuint64_t)h->Ne, (cuuint64_t)(h->mode == EDK_MODE_DERIVATIVE ? h->nfield : 1)};
        const cuuint64_t sstr[2] = {(cuuint64_t)h->sum_row * 8, (cuuint64_t)h->sum_row * 8 * (cuuint64_t)h->Ne};
        const cuuint32_t boxS[3] = {4, (cuuint32_t)gram_rows_per_tile(h->mfrag), 1};
        r = encode((CUtensorMap*)h->tma.mapS, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fsum, sdim, sstr, boxS, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return EDK_ERR_CUDA;
    }
    h->tma.phase_tiles = h->phase_tiles;
    h->tma.brows_alloc = rows;
    h->tma.nstages = nst;
    h->tma_ready = true;
    return EDK_OK;
}

// Plane-wave factorised contraction (algo 2), host side.  The xy-part of every momentum is one of a
// few {+q, -q} couples; couple q contributes the real modes cos(theta_q) and sin(theta_q) (only the
// constant 1 for q = 0).  Pure host logic, unit-tested on CPU through edk_plan_modes.
struct ModePlan {
    std::vector<int> modes3;  // per mode: qx, qy, kind (0 cos, 1 sin)
    std::vector<int> momode;  // per momentum: cos mode, sin mode (-1 if none), sigma with (px,py) = sigma (qx,qy)
};

ModePlan plan_modes(const std::vector<int>& mom) {
    ModePlan M;
    struct Couple {
        int qx, qy, mc, ms;
    };
    std::vector<Couple> couples;
    const int nmom = (int)mom.size() / 3;
    for (int i = 0; i < nmom; ++i) {
        int qx = mom[3 * i], qy = mom[3 * i + 1], sigma = 1;
        if (qx < 0 || (qx == 0 && qy < 0)) {  // representative: first non-zero component positive
            qx = -qx;
            qy = -qy;
            sigma = -1;
        }
        const Couple* hit = nullptr;
        for (const Couple& c : couples)
            if (c.qx == qx && c.qy == qy) hit = &c;
        if (!hit) {
            Couple c{qx, qy, (int)M.modes3.size() / 3, -1};
            M.modes3.insert(M.modes3.end(), {qx, qy, 0});
            if (qx != 0 || qy != 0) {
                c.ms = (int)M.modes3.size() / 3;
                M.modes3.insert(M.modes3.end(), {qx, qy, 1});
            }
            couples.push_back(c);
            hit = &couples.back();
        }
        M.momode.insert(M.momode.end(), {hit->mc, hit->ms, hit->ms < 0 ? 0 : sigma});
    }
    return M;
}

void free_pw(edk_handle* h) {
    cudaFree(h->pw_wtiles);
    cudaFree(h->pw_Y);
    cudaFree(h->pw_zphase);
    cudaFree(h->pw_momode);
    cudaFree(h->pw_slotmode);
    h->pw_slotmode = nullptr;
    h->pw_wtiles = nullptr;
    h->pw_Y = nullptr;
    h->pw_zphase = nullptr;
    h->pw_momode = nullptr;
    h->pw_ready = false;
    h->pw_bytes = 0;
}

// Tables, per-plane buffer and tensor maps of algo 2; needs the job list and the internal momenta.
int build_pw(edk_handle* h) {
    free_pw(h);
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return EDK_ERR_CUDA;
    const ModePlan mp = plan_modes(h->mom_int);
    const int A = h->g.Lx * h->g.Ly;
    // Form 3 on a plane of fewer than 8 sites would start its back run of 8 sites before the plane (for z = 0: before
    // the array); such toy planes simply run the unfolded kernel - same Y, same result.
    const bool fold = h->algo == 3 && A >= 8;
    h->pw_nmodes = (int)mp.modes3.size() / 3;
    // form 3: the {+q, -q} couples, eight per pass; a pass has a block of cos rows and a block of sin rows
    std::vector<int> slotmode;
    if (fold) {
        std::vector<std::pair<int, int>> couples;  // (cos mode, sin mode or -1), in mode order
        for (int m = 0; m < h->pw_nmodes; ++m) {
            if (mp.modes3[3 * m + 2] != 0) continue;
            const bool has_sin = m + 1 < h->pw_nmodes && mp.modes3[3 * (m + 1) + 2] == 1 &&
                                 mp.modes3[3 * (m + 1)] == mp.modes3[3 * m] && mp.modes3[3 * (m + 1) + 1] == mp.modes3[3 * m + 1];
            couples.push_back({m, has_sin ? m + 1 : -1});
        }
        h->pw_npass = ((int)couples.size() + 7) / 8;
        h->pw_mbtot = 2 * h->pw_npass;
        slotmode.assign((size_t)h->pw_mbtot * 8, -1);
        for (size_t c = 0; c < couples.size(); ++c) {
            slotmode[(2 * (c / 8)) * 8 + c % 8] = couples[c].first;
            slotmode[(2 * (c / 8) + 1) * 8 + c % 8] = couples[c].second;
        }
        h->pw_kplane = ((A + 1) / 2 + 7) / 8;  // stages of 8 site pairs
    } else {
        h->pw_mbtot = (h->pw_nmodes + 7) / 8;
        h->pw_npass = (h->pw_mbtot + PW_MAX_MB - 1) / PW_MAX_MB;
        h->pw_kplane = (A + 7) / 8;
    }
    pw_pick_tile(h->Ne, fold, &h->pw_el, &h->pw_fl);
    if (const char* t = getenv("EDK_PW_TILE")) {  // A/B hook: "24" = 16 x 32 tiles, "25" = 16 x 40, "17" = 8 x 56
        const int v = atoi(t);
        if (pw_tile_available(v / 10, v % 10)) h->pw_el = v / 10, h->pw_fl = v % 10;
    }
    const int rows_l = PW_WARPS * h->pw_el, rows_r = 8 * h->pw_fl;
    int smem = 0;
    if ((fold ? pwf_plan_smem(h->pw_el, h->pw_fl, &h->pw_tma.nstages, &smem) : pw_plan_smem(h->pw_el, h->pw_fl, &h->pw_tma.nstages, &smem)) != 0) {
        set_error("no shared-memory plan for the plane-wave contraction");
        return EDK_ERR_ARG;
    }
    if (const char* t = getenv("EDK_PW_STAGES")) {  // A/B hook: a shallower ring than shared memory allows (2 .. planned depth)
        const int v = atoi(t);
        if (v >= 2 && v < h->pw_tma.nstages) h->pw_tma.nstages = v;
    }
    const size_t wt_bytes = (size_t)h->pw_kplane * 2 * h->pw_mbtot * 32 * sizeof(double);
    const size_t y_bytes = (size_t)h->njobs * h->g.Lz * h->pw_nmodes * h->Ne * h->Ne * sizeof(cplx);
    const size_t zp_bytes = (size_t)h->nmom_int * h->g.Lz * sizeof(cplx);
    int* modes_dev = nullptr;
    EDK_CUDA_TRY(cudaMalloc(&h->pw_wtiles, wt_bytes));
    EDK_CUDA_TRY(cudaMalloc(&h->pw_zphase, zp_bytes));
    EDK_CUDA_TRY(cudaMalloc(&h->pw_momode, mp.momode.size() * sizeof(int)));
    if (fold) {
        EDK_CUDA_TRY(cudaMalloc(&h->pw_slotmode, slotmode.size() * sizeof(int)));
        EDK_CUDA_TRY(cudaMemcpy(h->pw_slotmode, slotmode.data(), slotmode.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    {
        const cudaError_t e = cudaMalloc(&h->pw_Y, y_bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc of %zu bytes (per-plane mode sums) failed: %s", y_bytes, cudaGetErrorString(e));
            free_pw(h);
            return e == cudaErrorMemoryAllocation ? EDK_ERR_NOMEM : EDK_ERR_CUDA;
        }
    }
    EDK_CUDA_TRY(cudaMalloc(&modes_dev, mp.modes3.size() * sizeof(int)));
    cudaError_t e = cudaMemcpy(modes_dev, mp.modes3.data(), mp.modes3.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = fold ? launch_pwf_weights(h->pw_wtiles, modes_dev, h->pw_slotmode, h->pw_mbtot, h->pw_kplane, h->g, 0)
                 : launch_pw_weights(h->pw_wtiles, modes_dev, h->pw_nmodes, h->pw_mbtot, h->pw_kplane, h->g, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(modes_dev);
    if (e != cudaSuccess) {
        set_error("plane-wave weight table failed: %s", cudaGetErrorString(e));
        return EDK_ERR_CUDA;
    }
    h->launches += 1;
    // exp(2 pi i pz z / Lz) with pz z reduced mod Lz in integers; form 3 takes its xy-modes about the centre of the
    // plane, which leaves the constant exp(i sigma delta_q), delta_q = pi (qx (Lx-1)/Lx + qy (Ly-1)/Ly), per momentum
    auto unit = [](long long num, long long den, double& c, double& sn) {  // exp(2 pi i num / den), exact on the axes
        const long long r = (num % den + den) % den;
        c = 1.0, sn = 0.0;
        if (4 * r == den) {
            c = 0.0, sn = 1.0;
        } else if (2 * r == den) {
            c = -1.0, sn = 0.0;
        } else if (4 * r == 3 * den) {
            c = 0.0, sn = -1.0;
        } else if (r != 0) {
            const double a = 2.0 * 3.14159265358979323846 * (double)r / (double)den;
            c = cos(a);
            sn = sin(a);
        }
    };
    std::vector<double> zp((size_t)h->nmom_int * h->g.Lz * 2);
    for (int p = 0; p < h->nmom_int; ++p) {
        double kc = 1.0, ks = 0.0;  // the constant of form 3
        if (fold) {
            const int mc = mp.momode[3 * p], sigma = mp.momode[3 * p + 2];
            const long long qx = mp.modes3[3 * mc], qy = mp.modes3[3 * mc + 1];
            double cx, sx, cy, sy;
            unit(sigma * qx * (h->g.Lx - 1), 2LL * h->g.Lx, cx, sx);
            unit(sigma * qy * (h->g.Ly - 1), 2LL * h->g.Ly, cy, sy);
            kc = cx * cy - sx * sy;
            ks = cx * sy + sx * cy;
        }
        for (int z = 0; z < h->g.Lz; ++z) {
            double c, sn;
            unit((long long)h->mom_int[3 * p + 2] * z, h->g.Lz, c, sn);
            zp[2 * ((size_t)p * h->g.Lz + z)] = c * kc - sn * ks;
            zp[2 * ((size_t)p * h->g.Lz + z) + 1] = c * ks + sn * kc;
        }
    }
    EDK_CUDA_TRY(cudaMemcpy(h->pw_zphase, zp.data(), zp_bytes, cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->pw_momode, mp.momode.data(), mp.momode.size() * sizeof(int), cudaMemcpyHostToDevice));
    const cuuint64_t Kd = (cuuint64_t)2 * 3 * h->g.V;
    const cuuint64_t gdim[3] = {Kd, (cuuint64_t)h->Ne, (cuuint64_t)h->nfield};
    const cuuint64_t gstr[2] = {Kd * 8, Kd * 8 * (cuuint64_t)h->Ne};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t boxL[3] = {8, (cuuint32_t)rows_l, 1};
    const cuuint32_t boxR[3] = {8, (cuuint32_t)rows_r, 1};
    CUresult r = encode((CUtensorMap*)h->pw_tma.mapL, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxL, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
        r = encode((CUtensorMap*)h->pw_tma.mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxR, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (plane-wave contraction) failed with CUresult %d", (int)r);
        return EDK_ERR_CUDA;
    }
    h->pw_bytes = wt_bytes + y_bytes + zp_bytes;
    h->pw_algo = h->algo;
    h->pw_fold = fold;
    h->pw_ready = true;
    return EDK_OK;
}

int run_gram_pw(edk_handle* h, cudaStream_t s) {
    PwParams Q{};
    Q.jobs = h->jobs_dev;
    Q.njobs = h->njobs;
    Q.Ne = h->Ne;
    Q.Lz = h->g.Lz;
    Q.A = h->g.Lx * h->g.Ly;
    Q.kplane = h->pw_kplane;
    const int rows_l = PW_WARPS * h->pw_el, rows_r = 8 * h->pw_fl;
    Q.n_et = (h->Ne + rows_l - 1) / rows_l;
    Q.n_ft = (h->Ne + rows_r - 1) / rows_r;
    Q.nmodes = h->pw_nmodes;
    Q.mbtot = h->pw_mbtot;
    Q.wtiles = h->pw_wtiles;
    Q.Y = h->pw_Y;
    Q.slotmode = h->pw_slotmode;
    const int npass = h->pw_npass;
    {
        PhaseTimer t(h, s, PH_GRAM, npass);
        for (int pass = 0; pass < npass; ++pass) {
            if (h->pw_fold) {
                Q.mb0 = 2 * pass;  // the pass's cos block, its sin block follows
                EDK_CUDA_TRY(launch_gram_pwf(Q, h->pw_tma, h->pw_el, h->pw_fl, s));
            } else {
                Q.mb0 = pass * PW_MAX_MB;
                EDK_CUDA_TRY(launch_gram_pw(Q, h->pw_tma, std::min(PW_MAX_MB, h->pw_mbtot - Q.mb0), h->pw_el, h->pw_fl, s));
            }
        }
    }
    // the z fold is a reduction like the combine step and is timed with it
    PhaseTimer t(h, s, PH_COMBINE, 1);
    PwFold F{};
    F.jobs = h->jobs_dev;
    F.njobs = h->njobs;
    F.Ne = h->Ne;
    F.Lz = h->g.Lz;
    F.nmodes = h->pw_nmodes;
    F.nmom_int = h->nmom_int;
    F.rows_l = rows_l;
    F.rows_r = rows_r;
    F.Y = h->pw_Y;
    F.zphase = h->pw_zphase;
    F.momode = h->pw_momode;
    F.partial = h->partial;
    EDK_CUDA_TRY(launch_pw_zfold(F, s));
    return EDK_OK;
}

// exp(2 pi i num / den), exact on the axes
void unit_circle(long long num, long long den, double& c, double& sn) {
    const long long r = (num % den + den) % den;
    c = 1.0, sn = 0.0;
    if (4 * r == den) {
        c = 0.0, sn = 1.0;
    } else if (2 * r == den) {
        c = -1.0, sn = 0.0;
    } else if (4 * r == 3 * den) {
        c = 0.0, sn = -1.0;
    } else if (r != 0) {
        const double a = 2.0 * 3.14159265358979323846 * (double)r / (double)den;
        c = cos(a);
        sn = sin(a);
    }
}

// ---- separable contraction (algo 4), host side ------------------------------------------------------------
// Which lattices / momentum lists the separable kernel covers, pure host logic (edk_plan_form): an even Lx >= 8 whose
// half row is a multiple of 8, 6 or 4 site pairs and fits the weight table, and momenta whose (|px|, |py|) lie in one
// of the instantiated mode structures: max <= 1 with px^2 + py^2 <= 1 (5 modes) or <= 2 (9 modes), max <= 2 with
// px^2 + py^2 <= 4 (13 modes: every list inside |p|^2 <= 4).
struct SepPlan {
    bool ok = false;
    int qmax = 0, r2 = 0, pairs = 0, nmodes = 0;
};
SepPlan plan_sep(int Lx, const std::vector<int>& mom) {
    SepPlan S;
    if (Lx < 8 || (Lx & 1) || Lx / 2 > SEP_MAX_PR) return S;
    const int PR = Lx / 2;
    S.pairs = PR % 8 == 0 ? 8 : (PR % 6 == 0 ? 6 : (PR % 4 == 0 ? 4 : 0));
    if (!S.pairs) return S;
    int qm = 0, r2 = 0;
    for (size_t i = 0; i < mom.size() / 3; ++i) {
        const int ax = abs(mom[3 * i]), ay = abs(mom[3 * i + 1]);
        qm = std::max(qm, std::max(ax, ay));
        r2 = std::max(r2, ax * ax + ay * ay);
    }
    if (qm <= 1)
        S.qmax = 1, S.r2 = r2 <= 1 ? 1 : 2;
    else if (qm == 2 && r2 <= 4)
        S.qmax = 2, S.r2 = 4;
    else
        return S;
    S.nmodes = sep_num_modes(S.qmax, S.r2);
    S.ok = S.nmodes > 0;
    return S;
}

// CTA tiles of gram_sepx_kernel for an Ne x Ne block of elements (pure host logic, edk_plan_tiles): 32 x 32 tiles over the
// part of the square that whole tiles fill, 8 x 128 tiles over the rows left below it (one 8-row unit each, all
// columns), 64 x 16 tiles over the columns left beside it (one 16-column unit each, the rows of the square).  Every
// element lies in exactly one tile; only the last tile of a strip can have idle warps.
std::vector<SepTile> sep_build_tiles(int Ne) {
    std::vector<SepTile> t;
    const int a = Ne / 32, sq = 32 * a;
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < a; ++j) t.push_back(SepTile{32 * i, 32 * j, 0, 32 * i + 32, 32 * j + 32, 0});
    for (int e0 = sq; e0 < Ne; e0 += 8)
        for (int f0 = 0; f0 < Ne; f0 += 128) t.push_back(SepTile{e0, f0, 1, std::min(e0 + 8, Ne), std::min(f0 + 128, Ne), 0});
    for (int f0 = sq; f0 < Ne; f0 += 16)
        for (int e0 = 0; e0 < sq; e0 += 64) t.push_back(SepTile{e0, f0, 2, std::min(e0 + 64, sq), std::min(f0 + 16, Ne), 0});
    return t;
}

void free_sep(edk_handle* h) {
    cudaFree(h->sep_tiles);
    h->sep_tiles = nullptr;
    h->sep_ntiles = 0;
    cudaFree(h->sep_wx);
    cudaFree(h->sep_wy);
    cudaFree(h->sep_Y);
    cudaFree(h->sep_zphase);
    cudaFree(h->sep_classes);
    cudaFree(h->sep_mom);
    h->sep_wx = h->sep_wy = nullptr;
    h->sep_Y = nullptr;
    h->sep_zphase = nullptr;
    h->sep_classes = nullptr;
    h->sep_mom = nullptr;
    h->sep_ready = false;
    h->sep_bytes = 0;
}

// Tables, per-plane buffer and tensor maps of algo 4; needs the job list and the internal momenta.
int build_sep(edk_handle* h) {
    free_sep(h);
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return EDK_ERR_CUDA;
    const SepPlan S = plan_sep(h->g.Lx, h->mom_int);
    if (!S.ok) {
        set_error("the separable contraction does not cover this lattice / momentum list (Lx even, >= 8, Lx/2 a multiple of 4, 6 or 8; "
                  "|px|, |py| <= 2 with px^2 + py^2 <= 4)");
        return EDK_ERR_ARG;
    }
    int smem = 0;
    h->sep_variant = SEP_DEFAULT_VARIANT;
    if (const char* t = getenv("EDK_SEP_VARIANT")) {  // A/B hook: 0 = the first kernel (accumulators in registers), 6 = the product
        const int v = atoi(t);
        if (v == 0 || v == 6) h->sep_variant = v;
    }
    if (h->sep_variant == 6) {
        for (int sh = 0; sh < SEP_NSHAPES; ++sh)
            if (sepx_plan_smem(sh, &h->sep_tmax.nstages[sh], &smem) != 0) {
                set_error("no shared-memory plan for the separable contraction");
                return EDK_ERR_ARG;
            }
        h->sep_tma.nstages = h->sep_tmax.nstages[0];
    } else if (sep_variant_plan(h->sep_variant, &h->sep_tma.nstages, &smem) != 0) {
        set_error("no shared-memory plan for the separable contraction");
        return EDK_ERR_ARG;
    }
    if (const char* t = getenv("EDK_SEP_STAGES")) {  // A/B hook: a shallower operand ring (2 .. planned depth)
        const int v = atoi(t);
        if (v >= 2 && v < h->sep_tma.nstages) h->sep_tma.nstages = v;
        for (int sh = 0; sh < SEP_NSHAPES; ++sh)
            if (v >= 2 && v < h->sep_tmax.nstages[sh]) h->sep_tmax.nstages[sh] = v;
    }
    h->sep_qmax = S.qmax, h->sep_r2 = S.r2, h->sep_pairs = S.pairs, h->sep_nmodes = S.nmodes;
    const int Lx = h->g.Lx, Ly = h->g.Ly, Lz = h->g.Lz, PR = Lx / 2;
    // weights about the centre of the lattice: c_q(x) + i s_q(x) = exp(i pi q (2x - L + 1)/L), q = 1, 2
    std::vector<double> wx((size_t)PR * 4), wy((size_t)Ly * 4);
    for (int x = 0; x < PR; ++x)
        for (int q = 1; q <= 2; ++q) unit_circle((long long)q * (2 * x - Lx + 1), 2LL * Lx, wx[4 * x + 2 * (q - 1)], wx[4 * x + 2 * (q - 1) + 1]);
    for (int y = 0; y < Ly; ++y)
        for (int q = 1; q <= 2; ++q) unit_circle((long long)q * (2 * y - Ly + 1), 2LL * Ly, wy[4 * y + 2 * (q - 1)], wy[4 * y + 2 * (q - 1) + 1]);
    std::copy(wx.begin(), wx.end(), h->sep_wx_host.w);
    // classes of momenta with the same (|px|, |py|), and exp(2 pi i pz z/Lz) times the constant left by the centring
    std::vector<SepClass> classes;
    std::vector<std::pair<int, int>> keys;
    std::vector<std::vector<int>> members;
    for (int p = 0; p < h->nmom_int; ++p) {
        const std::pair<int, int> key{abs(h->mom_int[3 * p]), abs(h->mom_int[3 * p + 1])};
        size_t c = 0;
        while (c < keys.size() && keys[c] != key) ++c;
        if (c == keys.size()) {
            keys.push_back(key);
            members.emplace_back();
        }
        members[c].push_back(p);  // ascending p
    }
    std::vector<int> mom3;
    for (size_t c = 0; c < keys.size(); ++c) {
        SepClass K{};
        const int ax = keys[c].first, ay = keys[c].second;
        K.mode[0] = sep_mode_index(S.qmax, S.r2, ax, 0, ay, 0);
        K.mode[1] = ay ? sep_mode_index(S.qmax, S.r2, ax, 0, ay, 1) : -1;
        K.mode[2] = ax ? sep_mode_index(S.qmax, S.r2, ax, 1, ay, 0) : -1;
        K.mode[3] = (ax && ay) ? sep_mode_index(S.qmax, S.r2, ax, 1, ay, 1) : -1;
        if (K.mode[0] < 0 || (ay && K.mode[1] < 0) || (ax && K.mode[2] < 0) || (ax && ay && K.mode[3] < 0)) {
            set_error("separable contraction: momentum class (%d, %d) is not in the mode structure", ax, ay);
            return EDK_ERR_STATE;
        }
        K.first = (int)mom3.size() / 3;
        K.count = (int)members[c].size();
        for (int p : members[c]) {
            const int px = h->mom_int[3 * p], py = h->mom_int[3 * p + 1];
            mom3.insert(mom3.end(), {p, (px > 0) - (px < 0), (py > 0) - (py < 0)});
        }
        classes.push_back(K);
    }
    h->sep_nclass = (int)classes.size();
    std::vector<double> zp((size_t)h->nmom_int * Lz * 2);
    for (int p = 0; p < h->nmom_int; ++p) {
        double cx, sx, cy, sy;
        unit_circle((long long)h->mom_int[3 * p] * (Lx - 1), 2LL * Lx, cx, sx);
        unit_circle((long long)h->mom_int[3 * p + 1] * (Ly - 1), 2LL * Ly, cy, sy);
        const double kc = cx * cy - sx * sy, ks = cx * sy + sx * cy;
        for (int z = 0; z < Lz; ++z) {
            double c, sn;
            unit_circle((long long)h->mom_int[3 * p + 2] * z, Lz, c, sn);
            zp[2 * ((size_t)p * Lz + z)] = c * kc - sn * ks;
            zp[2 * ((size_t)p * Lz + z) + 1] = c * ks + sn * kc;
        }
    }
    const size_t y_bytes = (size_t)h->njobs * Lz * S.nmodes * h->Ne * h->Ne * sizeof(cplx);
    EDK_CUDA_TRY(cudaMalloc(&h->sep_wx, wx.size() * sizeof(double)));
    EDK_CUDA_TRY(cudaMalloc(&h->sep_wy, wy.size() * sizeof(double)));
    EDK_CUDA_TRY(cudaMalloc(&h->sep_zphase, zp.size() * sizeof(double)));
    EDK_CUDA_TRY(cudaMalloc(&h->sep_classes, classes.size() * sizeof(SepClass)));
    EDK_CUDA_TRY(cudaMalloc(&h->sep_mom, mom3.size() * sizeof(int)));
    {
        const cudaError_t e = cudaMalloc(&h->sep_Y, y_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaMalloc of %zu bytes (per-plane mode sums) failed: %s", y_bytes, cudaGetErrorString(e));
            free_sep(h);
            return e == cudaErrorMemoryAllocation ? EDK_ERR_NOMEM : EDK_ERR_CUDA;
        }
    }
    EDK_CUDA_TRY(cudaMemcpy(h->sep_wx, wx.data(), wx.size() * sizeof(double), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->sep_wy, wy.data(), wy.size() * sizeof(double), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->sep_zphase, zp.data(), zp.size() * sizeof(double), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->sep_classes, classes.data(), classes.size() * sizeof(SepClass), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->sep_mom, mom3.data(), mom3.size() * sizeof(int), cudaMemcpyHostToDevice));
    // the field array as [nfield][Ne][6V doubles]; boxes of 16 doubles (8 complex = 128 bytes) x rows, 128-byte swizzle
    const cuuint64_t Kd = (cuuint64_t)2 * 3 * h->g.V;
    const cuuint64_t gdim[3] = {Kd, (cuuint64_t)h->Ne, (cuuint64_t)h->nfield};
    const cuuint64_t gstr[2] = {Kd * 8, Kd * 8 * (cuuint64_t)h->Ne};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t boxL[3] = {16, (cuuint32_t)sep_variant_rows(h->sep_variant), 1};
    const cuuint32_t boxR[3] = {16, (cuuint32_t)SEP_TF, 1};
    CUresult r = encode((CUtensorMap*)h->sep_tma.mapL, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxL, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
        r = encode((CUtensorMap*)h->sep_tma.mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, boxR, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && h->sep_variant == 6) {
        for (int sh = 0; sh < SEP_NSHAPES && r == CUDA_SUCCESS; ++sh) {
            int we = 0, wf = 0;
            sepx_shape(sh, &we, &wf);
            const cuuint32_t bL[3] = {16, (cuuint32_t)(8 * we), 1}, bR[3] = {16, (cuuint32_t)(16 * wf), 1};
            r = encode((CUtensorMap*)h->sep_tmax.mapL[sh], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, bL, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS)
                r = encode((CUtensorMap*)h->sep_tmax.mapR[sh], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->fields, gdim, gstr, bR, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        const std::vector<SepTile> tiles = sep_build_tiles(h->Ne);  // shape 0 first, then 1, then 2
        h->sep_ntiles = (int)tiles.size();
        for (int sh = 0; sh < SEP_NSHAPES; ++sh) h->sep_shape_first[sh] = h->sep_shape_count[sh] = 0;
        for (size_t i = 0; i < tiles.size(); ++i) {
            if (h->sep_shape_count[tiles[i].shape]++ == 0) h->sep_shape_first[tiles[i].shape] = (int)i;
        }
        // One launch over the whole tile table saves two launches and two tails (config 2: 0.36 -> 0.33 ms); with tens of
        // thousands of CTAs a launch per shape is faster (config 5: 116.0 against 117.7 ms), the strips then run after the
        // 32 x 32 tiles instead of among them.  EDK_SEP_LAUNCHES = 1 | 3 overrides (A/B hook).
        h->sep_per_shape = (long long)h->njobs * Lz * (long long)tiles.size() >= 148LL * 64;
        if (const char* t = getenv("EDK_SEP_LAUNCHES")) h->sep_per_shape = atoi(t) == 3 ? true : (atoi(t) == 1 ? false : h->sep_per_shape);
        EDK_CUDA_TRY(cudaMalloc(&h->sep_tiles, tiles.size() * sizeof(SepTile)));
        EDK_CUDA_TRY(cudaMemcpy(h->sep_tiles, tiles.data(), tiles.size() * sizeof(SepTile), cudaMemcpyHostToDevice));
    }
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (separable contraction) failed with CUresult %d", (int)r);
        return EDK_ERR_CUDA;
    }
    h->sep_bytes = y_bytes + (wx.size() + wy.size() + zp.size()) * sizeof(double);
    h->sep_ready = true;
    return EDK_OK;
}

int run_gram_sep(edk_handle* h, cudaStream_t s) {
    SepParams Q{};
    Q.jobs = h->jobs_dev;
    Q.njobs = h->njobs;
    Q.Ne = h->Ne;
    Q.Lx = h->g.Lx, Q.Ly = h->g.Ly, Q.Lz = h->g.Lz;
    Q.SR = h->g.Lx / 2 / h->sep_pairs;
    const int rows_l = sep_variant_rows(h->sep_variant);
    Q.n_et = (h->Ne + rows_l - 1) / rows_l;
    Q.n_ft = (h->Ne + SEP_TF - 1) / SEP_TF;
    Q.nmodes = h->sep_nmodes;
    Q.wx = h->sep_wx;
    Q.wy = h->sep_wy;
    Q.Y = h->sep_Y;
    if (h->sep_variant == 6 && !h->sep_per_shape) {
        PhaseTimer t(h, s, PH_GRAM, 1);  // one launch over the whole tile table
        Q.tiles = h->sep_tiles;
        Q.ntiles = h->sep_ntiles;
        EDK_CUDA_TRY(launch_gram_sepx(Q, h->sep_tmax, h->sep_wx_host, h->sep_qmax, h->sep_r2, h->sep_pairs, SEP_NSHAPES, s));
    } else if (h->sep_variant == 6) {
        int nlaunch = 0;
        for (int sh = 0; sh < SEP_NSHAPES; ++sh) nlaunch += h->sep_shape_count[sh] > 0;
        PhaseTimer t(h, s, PH_GRAM, nlaunch);
        for (int sh = 0; sh < SEP_NSHAPES; ++sh) {  // A/B reference: one launch per tile shape
            if (!h->sep_shape_count[sh]) continue;
            Q.tiles = h->sep_tiles + h->sep_shape_first[sh];
            Q.ntiles = h->sep_shape_count[sh];
            EDK_CUDA_TRY(launch_gram_sepx(Q, h->sep_tmax, h->sep_wx_host, h->sep_qmax, h->sep_r2, h->sep_pairs, sh, s));
        }
    } else {
        PhaseTimer t(h, s, PH_GRAM, 1);
        EDK_CUDA_TRY(launch_gram_sep(Q, h->sep_tma, h->sep_wx_host, h->sep_qmax, h->sep_r2, h->sep_pairs, h->sep_variant, s));
    }
    // the z fold is a reduction like the combine step and is timed with it
    PhaseTimer t(h, s, PH_COMBINE, 1);
    SepFold F{};
    F.jobs = h->jobs_dev;
    F.njobs = h->njobs;
    F.Ne = h->Ne;
    F.Lz = h->g.Lz;
    F.nmodes = h->sep_nmodes;
    F.nmom_int = h->nmom_int;
    F.rows_l = h->sep_variant == 6 ? 8 : rows_l;  // variant 6 skips 8 x 8 blocks of a self pair below the diagonal
    F.rows_r = h->sep_variant == 6 ? 8 : SEP_TF;
    F.Y = h->sep_Y;
    F.zphase = h->sep_zphase;
    F.nclass = h->sep_nclass;
    F.classes = h->sep_classes;
    F.mom = h->sep_mom;
    F.partial = h->partial;
    EDK_CUDA_TRY(launch_sep_zfold(F, s));
    return EDK_OK;
}

// The contraction form a handle uses unless one is asked for (edk_debug_algo, EDK_GRAM_ALGO), from the FP64-pipe work
// per (e, f, site) of each form and the fraction of the pipe each kernel was measured to sustain on B200:
//   GEMM form (1):       9 DMMA-slot equivalents per (pair, momentum) - three real MMAs over three colours
//   folded plane wave (3): 12 + 2 + 16 per pair and pass of 8 {+q, -q} couples, needs planes of at least 8 sites
//   separable (4):       12 + (4 + 2 + 4 qmax)/2 per pair, needs plan_sep
// Pure host logic; edk_plan_form exposes it to the CPU tests.
int plan_contraction_form(int Lx, int Ly, const std::vector<int>& mom_int, const std::vector<GramJob>& jobs) {
    double w_gemm = 0.0, segs = 0.0;
    for (const auto& j : jobs) {
        w_gemm += 9.0 * j.nseg * j.nmom;
        segs += j.nseg;
    }
    double best = w_gemm / 0.88;
    int form = 1;
    if (Lx * Ly >= 8) {
        const ModePlan mp = plan_modes(mom_int);
        int couples = 0;
        for (size_t m = 0; m < mp.modes3.size() / 3; ++m) couples += mp.modes3[3 * m + 2] == 0;
        const double w = segs * (14.0 + 16.0) * ((couples + 7) / 8) / 0.61;  // measured: FP64 pipe 61 % busy at config 5
        if (w < best) best = w, form = 3;
    }
    const SepPlan S = plan_sep(Lx, mom_int);
    if (S.ok) {
        const double w = segs * (12.0 + 0.5 * (6.0 + 4.0 * S.qmax)) / 0.70;  // measured: 69 - 71 % at configs 4 / 5
        if (w < best) best = w, form = 4;
    }
    return form;
}

// Momentum bookkeeping of the contraction, pure host logic (unit-tested on CPU through edk_plan):
// whether the Hermitian pairing pays, the internal momentum list (the caller's distinct momenta plus
// missing negatives, one representative of every {p, -p} couple first), the index of -p for every
// internal p, the caller's -> internal map and the size of the half set the self pairs contract.
struct MomentumPlan {
    bool symmetric = false;
    std::vector<int> mom_int, negidx, pmap;
    int n_half = 0;
};

MomentumPlan plan_momenta(int mode, int order, int sym_request, const std::vector<int>& mom_user) {
    MomentumPlan P;
    const int nmom = (int)mom_user.size() / 3;
    P.mom_int = mom_user;
    P.pmap.resize(nmom);
    for (int i = 0; i < nmom; ++i) P.pmap[i] = i;
    auto find = [](const std::vector<int>& v, int px, int py, int pz) {
        for (size_t i = 0; i < v.size() / 3; ++i)
            if (v[3 * i] == px && v[3 * i + 1] == py && v[3 * i + 2] == pz) return (int)i;
        return -1;
    };
    if (mode == EDK_MODE_DERIVATIVE && order >= 1 && sym_request != 0) {
        std::vector<int> ext;  // distinct momenta of the caller plus any missing negatives
        for (int i = 0; i < nmom; ++i) {
            const int* m = &mom_user[3 * i];
            if (find(ext, m[0], m[1], m[2]) < 0) ext.insert(ext.end(), m, m + 3);
        }
        const size_t nuser_distinct = ext.size() / 3;
        for (size_t i = 0; i < nuser_distinct; ++i) {
            const int px = -ext[3 * i], py = -ext[3 * i + 1], pz = -ext[3 * i + 2];
            if (find(ext, px, py, pz) < 0) ext.insert(ext.end(), {px, py, pz});
        }
        int plain = 0, sym = 0;
        count_pairs(order, plain, sym);
        const long long cost_sym = (long long)sym * (long long)(ext.size() / 3);
        const long long cost_plain = (long long)plain * nmom;
        if (sym_request == 1 || cost_sym < cost_plain) {
            P.symmetric = true;
            std::vector<int> reps, partners;
            for (size_t i = 0; i < ext.size() / 3; ++i) {
                const int px = ext[3 * i], py = ext[3 * i + 1], pz = ext[3 * i + 2];
                if (find(reps, px, py, pz) >= 0 || find(partners, px, py, pz) >= 0) continue;
                reps.insert(reps.end(), {px, py, pz});
                if (px != 0 || py != 0 || pz != 0) partners.insert(partners.end(), {-px, -py, -pz});
            }
            P.n_half = (int)reps.size() / 3;
            P.mom_int = reps;
            P.mom_int.insert(P.mom_int.end(), partners.begin(), partners.end());
            for (int i = 0; i < nmom; ++i) {
                const int* m = &mom_user[3 * i];
                P.pmap[i] = find(P.mom_int, m[0], m[1], m[2]);
            }
        }
    }
    const int nint = (int)P.mom_int.size() / 3;
    if (!P.symmetric) P.n_half = nint;
    P.negidx.assign(nint, -1);
    for (int i = 0; i < nint; ++i) {
        if (!P.symmetric) {
            P.negidx[i] = i;  // never used without the pairing
            continue;
        }
        P.negidx[i] = find(P.mom_int, -P.mom_int[3 * i], -P.mom_int[3 * i + 1], -P.mom_int[3 * i + 2]);
    }
    return P;
}

// (Re)build everything that depends on the momentum list and on the pairing mode:
// internal momentum list, phase tables, contraction jobs, combine recipe, partial-sum buffer.
int configure(edk_handle* h) {
    const MomentumPlan plan = plan_momenta(h->mode, h->order, h->sym_request, h->mom_user);
    h->symmetric = plan.symmetric;
    h->mom_int = plan.mom_int;
    h->negidx = plan.negidx;
    h->pmap = plan.pmap;
    h->n_half = plan.n_half;
    h->nmom_int = (int)h->mom_int.size() / 3;
    if (h->mode == EDK_MODE_DERIVATIVE)
        build_derivative_jobs(h);
    else
        build_displacement_jobs(h);
    h->njobs = (int)h->jobs_host.size();
    h->algo = h->algo_request >= 0 ? h->algo_request : plan_contraction_form(h->g.Lx, h->g.Ly, h->mom_int, h->jobs_host);

    free_pw(h);
    free_sep(h);
    cudaFree(h->phase);
    cudaFree(h->phase_tiles);
    h->phase_tiles = nullptr;
    cudaFree(h->jobs_dev);
    cudaFree(h->ops_dev);
    cudaFree(h->negidx_dev);
    cudaFree(h->pmap_dev);
    cudaFree(h->partial);
    h->pmap_dev = nullptr;
    h->phase = nullptr;
    h->jobs_dev = nullptr;
    h->ops_dev = nullptr;
    h->negidx_dev = nullptr;
    h->partial = nullptr;
    const size_t nm = (size_t)h->nmom_int;
    int* mom_dev = nullptr;
    EDK_CUDA_TRY(cudaMalloc(&h->phase, 2 * nm * h->g.Vpad * sizeof(cplx)));
    EDK_CUDA_TRY(cudaMalloc(&mom_dev, nm * 3 * sizeof(int)));
    cudaError_t e = cudaMemcpy(mom_dev, h->mom_int.data(), nm * 3 * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_phase_table(h->phase, h->phase + nm * h->g.Vpad, mom_dev, h->nmom_int, h->g, 0);
    if (e == cudaSuccess) e = cudaMalloc(&h->phase_tiles, 2 * nm * h->g.Vpad * sizeof(cplx));
    if (e == cudaSuccess) e = launch_phase_tiles(h->phase, h->phase_tiles, h->nmom_int, h->g.Vpad, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(mom_dev);
    if (e != cudaSuccess) {
        set_error("phase table failed: %s", cudaGetErrorString(e));
        return EDK_ERR_CUDA;
    }
    h->launches += 2;
    EDK_CUDA_TRY(cudaMalloc(&h->jobs_dev, h->jobs_host.size() * sizeof(GramJob)));
    EDK_CUDA_TRY(cudaMalloc(&h->ops_dev, h->ops_host.size() * sizeof(CombineOp)));
    EDK_CUDA_TRY(cudaMalloc(&h->negidx_dev, nm * sizeof(int)));
    EDK_CUDA_TRY(cudaMemcpy(h->jobs_dev, h->jobs_host.data(), h->jobs_host.size() * sizeof(GramJob), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->ops_dev, h->ops_host.data(), h->ops_host.size() * sizeof(CombineOp), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMemcpy(h->negidx_dev, h->negidx.data(), nm * sizeof(int), cudaMemcpyHostToDevice));
    EDK_CUDA_TRY(cudaMalloc(&h->pmap_dev, (size_t)h->nmom * sizeof(int)));
    EDK_CUDA_TRY(cudaMemcpy(h->pmap_dev, h->pmap.data(), (size_t)h->nmom * sizeof(int), cudaMemcpyHostToDevice));
    pick_gram_config(h);
    {
        const int rc = build_tma(h);
        if (rc != EDK_OK) return rc;
    }
    const size_t partial_bytes = (size_t)h->ksplit * h->njobs * nm * h->Ne * h->Ne * sizeof(cplx);
    cudaError_t pe = cudaMalloc(&h->partial, partial_bytes);
    if (pe != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes (partial sums) failed: %s", partial_bytes, cudaGetErrorString(pe));
        return pe == cudaErrorMemoryAllocation ? EDK_ERR_NOMEM : EDK_ERR_CUDA;
    }
    h->cfg_bytes = 4 * nm * h->g.Vpad * sizeof(cplx) + partial_bytes;
    if (h->algo >= 2) {
        int rc = h->algo == 4 ? build_sep(h) : build_pw(h);
        if (rc == EDK_ERR_NOMEM && h->algo_request < 0) {
            // the per-plane sums of the planned form do not fit: the GEMM form needs no such buffer
            free_sep(h);
            free_pw(h);
            h->algo = 1;
            pick_gram_config(h);
            rc = build_tma(h);
            if (rc == EDK_OK) rc = ensure_partial(h);
            if (rc == EDK_OK) h->cfg_bytes = 4 * nm * h->g.Vpad * sizeof(cplx) + (size_t)h->ksplit * h->njobs * nm * h->Ne * h->Ne * sizeof(cplx);
        }
        if (rc != EDK_OK) return rc;
    }
    return EDK_OK;
}

int run_gram_and_combine(edk_handle* h, cplx* out, cudaStream_t s) {
    GramParams P{};
    P.jobs = h->jobs_dev;
    P.njobs = h->njobs;
    P.Ne = h->Ne;
    P.nmom = h->nmom_int;
    P.Kc = 3 * h->g.V;
    P.ksteps = h->g.Vpad / 8;
    P.Vpad = h->g.Vpad;
    P.ksplit = h->naive ? 1 : h->ksplit;
    P.n_mt = row_tiles(h);
    const bool use_pw = !h->naive && h->loader == 0 && h->algo >= 2;
    if (use_pw) {
        if (h->algo == 4 ? !h->sep_ready : (!h->pw_ready || h->pw_algo != h->algo)) {
            set_error("plane-wave / separable contraction selected but its tables are not built");
            return EDK_ERR_STATE;
        }
        const int rc = h->algo == 4 ? run_gram_sep(h, s) : run_gram_pw(h, s);
        if (rc != EDK_OK) return rc;
        PhaseTimer t(h, s, PH_COMBINE, 1);
        EDK_CUDA_TRY(launch_combine(h->ops_dev, h->nop, h->partial, h->njobs, 1, h->nmom_int, h->nmom, h->pmap_dev,
                                    h->negidx_dev, h->n_half, h->Ne, h->have_coeff ? h->coeff : nullptr, out, s));
        return EDK_OK;
    }
    const bool use_tma = !h->naive && h->loader == 0 && h->tma_ready;
    if (h->cta_dirty) {
        const int rc = build_cta_map(h);
        if (rc != EDK_OK) return rc;
        h->cta_dirty = false;
    }
    P.cta_map = h->cta_map_dev;
    P.ncta = h->ncta;
    P.phase = h->phase;
    P.partial = h->partial;
    {
        PhaseTimer t(h, s, PH_GRAM, 1);
        if (h->naive)
            EDK_CUDA_TRY(launch_gram_naive(P, s));
        else if (use_tma)
            EDK_CUDA_TRY(launch_gram_tma(P, h->tma, h->mfrag, h->algo, s));
        else
            EDK_CUDA_TRY(launch_gram_dmma(P, h->mfrag, s));
    }
    {
        PhaseTimer t(h, s, PH_COMBINE, 1);
        EDK_CUDA_TRY(launch_combine(h->ops_dev, h->nop, h->partial, h->njobs, P.ksplit, h->nmom_int, h->nmom, h->pmap_dev,
                                    h->negidx_dev, h->n_half, h->Ne, h->have_coeff ? h->coeff : nullptr, out, s));
    }
    return EDK_OK;
}

void recycle_events(edk_handle* h) {
    for (auto& e : h->events) h->pool.push_back(e);
    h->events.clear();
    for (int i = 0; i < PH_COUNT; ++i) h->n_launch[i] = 0;
}

}  // namespace

extern "C" {

int edk_version(void) { return 1; }
const char* edk_last_error(void) { return g_err; }

int edk_create(int Lx, int Ly, int Lz, int Ne, int mode, int order, int nmom, const int* mom3, int device,
               edk_handle** out) {
    if (!out) {
        set_error("edk_create: out is NULL");
        return EDK_ERR_ARG;
    }
    *out = nullptr;
    if (Lx < 1 || Ly < 1 || Lz < 1 || Ne < 1 || nmom < 1 || !mom3) {
        set_error("edk_create: lattice extents, Ne and nmom must be positive (got %d %d %d, Ne=%d, nmom=%d)", Lx, Ly, Lz, Ne,
                  nmom);
        return EDK_ERR_ARG;
    }
    if (mode != EDK_MODE_DERIVATIVE && mode != EDK_MODE_DISPLACEMENT) {
        set_error("edk_create: unknown mode %d", mode);
        return EDK_ERR_ARG;
    }
    if (order < 0 || (mode == EDK_MODE_DERIVATIVE && order > 3)) {
        set_error("edk_create: order %d out of range (num_nabla 0..3, distance >= 0)", order);
        return EDK_ERR_ARG;
    }
    if ((long long)Lx * Ly * Lz * 3 >= (1LL << 31) / 2) {
        set_error("edk_create: spatial volume too large for 32-bit k indices");
        return EDK_ERR_ARG;
    }
    DeviceGuard guard(device);  // the caller's current device is restored on return
    if (!guard.ok) {
        set_error("edk_create: cudaSetDevice(%d) failed", device);
        return EDK_ERR_CUDA;
    }
    edk_handle* h = new edk_handle();
    h->g.Lx = Lx;
    h->g.Ly = Ly;
    h->g.Lz = Lz;
    h->g.V = Lx * Ly * Lz;
    h->g.Vpad = (h->g.V + 7) / 8 * 8;
    h->Ne = Ne;
    h->mode = mode;
    h->order = order;
    h->nmom = nmom;
    h->device = device;
    h->field_cplx = (size_t)Ne * h->g.V * 3;
    if (mode == EDK_MODE_DERIVATIVE) {
        h->nop = pow3sum(order);
        h->nfield = h->nop;
    } else {
        h->nop = order + 1;
        h->nfield = order + 1;
    }
    size_t ws = 0;
    auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
        ws += bytes;
        return cudaMalloc(p, bytes);
    };
#define EDK_ALLOC(ptr, bytes)                                                                      \
    do {                                                                                           \
        cudaError_t _e = alloc((void**)&(ptr), (bytes));                                           \
        if (_e != cudaSuccess) {                                                                   \
            set_error("edk_create: cudaMalloc of %zu bytes failed: %s", (size_t)(bytes), cudaGetErrorString(_e)); \
            edk_destroy(h);                                                                        \
            return _e == cudaErrorMemoryAllocation ? EDK_ERR_NOMEM : EDK_ERR_CUDA;                 \
        }                                                                                          \
    } while (0)
    EDK_ALLOC(h->links, (size_t)3 * h->g.V * 9 * sizeof(cplx));
    EDK_ALLOC(h->fields, (size_t)h->nfield * h->field_cplx * sizeof(cplx));
    // displacement mode only ever has W0 on the left of a pair, so only its plane is needed
    h->sum_row = ((size_t)3 * h->g.V + 1) & ~(size_t)1;
    EDK_ALLOC(h->fsum, (size_t)(mode == EDK_MODE_DERIVATIVE ? h->nfield : 1) * Ne * h->sum_row * sizeof(double));
    if (mode == EDK_MODE_DISPLACEMENT && order >= 1) EDK_ALLOC(h->lines, (size_t)12 * h->field_cplx * sizeof(cplx));
    EDK_ALLOC(h->coeff, (size_t)Ne * Ne * sizeof(double));
    h->mom_user.assign(mom3, mom3 + 3 * (size_t)nmom);
    // The form of the contraction is planned per handle (plan_contraction_form).  A/B hook for measurements only:
    // EDK_GRAM_ALGO = 0 (4M GEMM), 1 (3M GEMM), 2 / 3 (plane-wave forms), 4 (separable form) asks for one form.
    if (const char* a = getenv("EDK_GRAM_ALGO")) {
        const int v = atoi(a);
        if (v >= 0 && v <= 4 && a[0] >= '0' && a[0] <= '9') h->algo_request = v;
    }
    {
        const int rc = configure(h);
        if (rc != EDK_OK) {
            edk_destroy(h);
            return rc;
        }
    }
#undef EDK_ALLOC
    h->ws_bytes = ws;
    *out = h;
    return EDK_OK;
}

int edk_destroy(edk_handle* h) {
    if (!h) return EDK_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->links);
    cudaFree(h->links_tmp);
    cudaFree(h->fields);
    cudaFree(h->fsum);
    cudaFree(h->lines);
    cudaFree(h->phase);
    cudaFree(h->partial);
    cudaFree(h->coeff);
    cudaFree(h->jobs_dev);
    cudaFree(h->ops_dev);
    cudaFree(h->negidx_dev);
    cudaFree(h->pmap_dev);
    cudaFree(h->cta_map_dev);
    cudaFree(h->phase_tiles);
    free_pw(h);
    free_sep(h);
    cudaFree(h->stage_U);
    cudaFree(h->stage_V);
    cudaFree(h->stage_out);
    for (auto& e : h->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto& e : h->pool) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    delete h;
    return EDK_OK;
}

int edk_phase_table(int Lx, int Ly, int Lz, int nmom, const int* mom3, void* out_dev, int device, void* stream) {
    if (Lx < 1 || Ly < 1 || Lz < 1 || nmom < 1 || !mom3 || !out_dev) {
        set_error("edk_phase_table: bad argument");
        return EDK_ERR_ARG;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("edk_phase_table: cudaSetDevice(%d) failed", device);
        return EDK_ERR_CUDA;
    }
    Geom g{Lx, Ly, Lz, Lx * Ly * Lz, Lx * Ly * Lz};  // unpadded rows: the caller's buffer is [nmom][V]
    int* mom_dev = nullptr;
    EDK_CUDA_TRY(cudaMalloc(&mom_dev, (size_t)nmom * 3 * sizeof(int)));
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(mom_dev, mom3, (size_t)nmom * 3 * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_phase_table((cplx*)out_dev, nullptr, mom_dev, nmom, g, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(mom_dev);
    if (e != cudaSuccess) {
        set_error("edk_phase_table: %s", cudaGetErrorString(e));
        return EDK_ERR_CUDA;
    }
    return EDK_OK;
}

int edk_plan(int mode, int order, int nmom, const int* mom3, int sym_request, int out[8]) {
    if (nmom < 1 || !mom3 || !out || order < 0 || (mode != EDK_MODE_DERIVATIVE && mode != EDK_MODE_DISPLACEMENT) ||
        (mode == EDK_MODE_DERIVATIVE && order > 3)) {
        set_error("edk_plan: bad argument");
        return EDK_ERR_ARG;
    }
    const std::vector<int> user(mom3, mom3 + 3 * (size_t)nmom);
    const MomentumPlan P = plan_momenta(mode, order, sym_request, user);
    const int nint = (int)P.mom_int.size() / 3;
    int plain = 1, sym = 1, self = 0;
    if (mode == EDK_MODE_DERIVATIVE)
        count_pairs(order, plain, sym, &self);
    else
        plain = sym = order + 1;
    out[0] = P.symmetric ? 1 : 0;
    out[1] = nint;
    out[2] = P.n_half;
    out[3] = plain;
    out[4] = sym;
    out[5] = P.symmetric ? (sym - self) * nint + self * P.n_half : plain * nint;  // (pair, momentum) GEMMs per timeslice
    out[6] = mode == EDK_MODE_DERIVATIVE ? pow3sum(order) : order + 1;            // operators in the output
    out[7] = self;
    return EDK_OK;
}

int edk_plan_modes(int nmom, const int* mom3, int* nmodes, int* modes3, int* momode) {
    if (nmom < 1 || !mom3 || !nmodes || !modes3 || !momode) {
        set_error("edk_plan_modes: bad argument");
        return EDK_ERR_ARG;
    }
    const ModePlan M = plan_modes(std::vector<int>(mom3, mom3 + 3 * (size_t)nmom));
    *nmodes = (int)M.modes3.size() / 3;
    std::copy(M.modes3.begin(), M.modes3.end(), modes3);
    std::copy(M.momode.begin(), M.momode.end(), momode);
    return EDK_OK;
}

int edk_plan_form(int Lx, int Ly, int mode, int order, int nmom, const int* mom3, int out[6]) {
    if (Lx < 1 || Ly < 1 || nmom < 1 || !mom3 || !out || order < 0 || (mode != EDK_MODE_DERIVATIVE && mode != EDK_MODE_DISPLACEMENT) ||
        (mode == EDK_MODE_DERIVATIVE && order > 3)) {
        set_error("edk_plan_form: bad argument");
        return EDK_ERR_ARG;
    }
    // the job list of a handle, built on the host only (field pointers are never dereferenced here)
    edk_handle tmp;
    tmp.mode = mode;
    tmp.order = order;
    tmp.nmom = nmom;
    tmp.Ne = 1;
    tmp.field_cplx = 0;
    tmp.nop = mode == EDK_MODE_DERIVATIVE ? pow3sum(order) : order + 1;
    tmp.mom_user.assign(mom3, mom3 + 3 * (size_t)nmom);
    const MomentumPlan plan = plan_momenta(mode, order, -1, tmp.mom_user);
    tmp.symmetric = plan.symmetric;
    tmp.mom_int = plan.mom_int;
    tmp.n_half = plan.n_half;
    tmp.nmom_int = (int)plan.mom_int.size() / 3;
    if (mode == EDK_MODE_DERIVATIVE)
        build_derivative_jobs(&tmp);
    else
        build_displacement_jobs(&tmp);
    const SepPlan S = plan_sep(Lx, tmp.mom_int);
    out[0] = plan_contraction_form(Lx, Ly, tmp.mom_int, tmp.jobs_host);
    out[1] = S.ok ? 1 : 0;
    out[2] = S.qmax;
    out[3] = S.r2;
    out[4] = S.pairs;
    out[5] = S.nmodes;
    return EDK_OK;
}

int edk_plan_tiles(int Ne, int max_tiles, int* tiles4) {
    if (Ne < 1 || max_tiles < 0 || (max_tiles > 0 && !tiles4)) {
        set_error("edk_plan_tiles: bad argument");
        return EDK_ERR_ARG;
    }
    const std::vector<SepTile> t = sep_build_tiles(Ne);
    for (size_t i = 0; i < t.size() && (int)i < max_tiles; ++i) {
        int we = 0, wf = 0;
        sepx_shape(t[i].shape, &we, &wf);
        tiles4[4 * i] = t[i].e0;
        tiles4[4 * i + 1] = t[i].f0;
        tiles4[4 * i + 2] = t[i].e1 - t[i].e0;
        tiles4[4 * i + 3] = t[i].f1 - t[i].f0;
    }
    return (int)t.size();
}

int edk_num_operators(const edk_handle* h) { return h ? h->nop : EDK_ERR_ARG; }
size_t edk_output_bytes(const edk_handle* h) {
    return h ? (size_t)h->nop * h->nmom * h->Ne * h->Ne * sizeof(cplx) : 0;
}
size_t edk_workspace_bytes(const edk_handle* h) { return h ? h->ws_bytes + h->cfg_bytes + h->pw_bytes + h->sep_bytes : 0; }

int edk_set_links(edk_handle* h, const void* U_dev, int layout, void* stream) {
    const int big_endian = (layout & EDK_LINKS_BIG_ENDIAN) ? 1 : 0;
    layout &= ~EDK_LINKS_BIG_ENDIAN;
    if (!h || !U_dev || (layout != EDK_LINKS_DIR_MAJOR && layout != EDK_LINKS_FILE_T)) {
        set_error("edk_set_links: bad argument");
        return EDK_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EDK_ON_DEVICE(h);
    {
        PhaseTimer t(h, s, PH_PREP, 1);
        EDK_CUDA_TRY(launch_reorder_links((const cplx*)U_dev, layout, big_endian, h->links, h->g, s));
    }
    for (const auto& op : h->link_ops) {
        if (op.kind == 2) {
            PhaseTimer t(h, s, PH_PREP, 1);
            EDK_CUDA_TRY(launch_project_su3(h->links, h->g, s));
        } else {
            if (!h->links_tmp) EDK_CUDA_TRY(cudaMalloc(&h->links_tmp, (size_t)3 * h->g.V * 9 * sizeof(cplx)));
            for (int i = 0; i < op.nstep; ++i) {
                PhaseTimer t(h, s, PH_PREP, 1);
                EDK_CUDA_TRY(launch_stout_step(h->links, h->links_tmp, op.rho, h->g, s));
                std::swap(h->links, h->links_tmp);
            }
        }
    }
    h->links_set = true;
    return EDK_OK;
}

int edk_set_link_ops(edk_handle* h, int nops, const int* kinds, const int* nsteps, const double* rhos) {
    if (!h || nops < 0 || (nops > 0 && (!kinds || !nsteps || !rhos))) {
        set_error("edk_set_link_ops: bad argument");
        return EDK_ERR_ARG;
    }
    std::vector<edk_handle::LinkOp> ops;
    for (int i = 0; i < nops; ++i) {
        if ((kinds[i] != 1 && kinds[i] != 2) || (kinds[i] == 1 && nsteps[i] < 0)) {
            set_error("edk_set_link_ops: op %d is neither stout (1, nstep >= 0) nor project (2)", i);
            return EDK_ERR_ARG;
        }
        ops.push_back({kinds[i], nsteps[i], rhos[i]});
    }
    h->link_ops = ops;
    h->links_set = false;  // links already on the device were processed with the old list
    return EDK_OK;
}

int edk_debug_links(edk_handle* h, void* dst_dev, void* stream) {
    if (!h || !dst_dev) return EDK_ERR_ARG;
    EDK_ON_DEVICE(h);
    EDK_CUDA_TRY(cudaMemcpyAsync(dst_dev, h->links, (size_t)3 * h->g.V * 9 * sizeof(cplx), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return EDK_OK;
}

int edk_set_eigvecs(edk_handle* h, const void* V_dev, int is_c8, void* stream) {
    if (!h || !V_dev || (is_c8 & ~(EDK_EIGVECS_C8 | EDK_EIGVECS_BIG_ENDIAN))) {
        set_error("edk_set_eigvecs: bad argument");
        return EDK_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EDK_ON_DEVICE(h);
    PhaseTimer t(h, s, PH_PREP, 1);
    EDK_CUDA_TRY(launch_round_eigvecs(V_dev, is_c8, h->field(0), h->field_sum(0), h->field_cplx, (size_t)3 * h->g.V,
                                      h->sum_row, s));
    h->evecs_set = true;
    return EDK_OK;
}

int edk_set_blending(edk_handle* h, const double* coeff_dev, void* stream) {
    if (!h) {
        set_error("edk_set_blending: NULL handle");
        return EDK_ERR_ARG;
    }
    if (!coeff_dev) {
        h->have_coeff = false;
        return EDK_OK;
    }
    EDK_ON_DEVICE(h);
    EDK_CUDA_TRY(cudaMemcpyAsync(h->coeff, coeff_dev, (size_t)h->Ne * h->Ne * sizeof(double), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    h->have_coeff = true;
    return EDK_OK;
}

int edk_calc(edk_handle* h, void* out_dev, void* stream) {
    if (!h || !out_dev) {
        set_error("edk_calc: bad argument");
        return EDK_ERR_ARG;
    }
    if (!h->links_set || !h->evecs_set) {
        set_error("edk_calc: links and eigenvectors of the timeslice must be set first");
        return EDK_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EDK_ON_DEVICE(h);
    if (h->mode == EDK_MODE_DERIVATIVE) {
        // only the GEMM form's 3M arithmetic reads the Re + Im planes of the derived fields; the fields are rebuilt by
        // every call, so the choice follows the contraction form in use right now (W0's plane is always written)
        const bool planes = effective_algo(h) < 2;
        for (const auto& hop : h->hops) {
            PhaseTimer t(h, s, PH_STENCIL, 1);
            EDK_CUDA_TRY(launch_nabla3(h->field(hop.first), h->field(hop.second), h->field(hop.second + 1),
                                       h->field(hop.second + 2), planes ? h->field_sum(hop.second) : nullptr,
                                       planes ? h->field_sum(hop.second + 1) : nullptr,
                                       planes ? h->field_sum(hop.second + 2) : nullptr, h->sum_row, h->links, h->g, h->Ne, s));
        }
    } else {
        for (int k = 1; k <= h->order; ++k) {
            Ptr6 p;
            for (int l = 0; l < 6; ++l) {
                p.src[l] = (k == 1) ? h->field(0) : h->lines + (size_t)(((k - 1) & 1) * 6 + l) * h->field_cplx;
                p.dst[l] = h->lines + (size_t)((k & 1) * 6 + l) * h->field_cplx;
            }
            PhaseTimer t(h, s, PH_STENCIL, 1);
            EDK_CUDA_TRY(launch_displace_step6(p, h->field(k), h->links, h->g, h->Ne, s));
        }
    }
    return run_gram_and_combine(h, (cplx*)out_dev, s);
}

int edk_laplacian(edk_handle* h, const void* F_dev, void* out_dev, int nvec, void* stream) {
    if (!h || !F_dev || !out_dev || nvec < 1 || F_dev == out_dev) {
        set_error("edk_laplacian: bad argument (in-place application is not supported)");
        return EDK_ERR_ARG;
    }
    if (!h->links_set) {
        set_error("edk_laplacian: the links of the timeslice must be set first");
        return EDK_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EDK_ON_DEVICE(h);
    PhaseTimer t(h, s, PH_STENCIL, 1);
    EDK_CUDA_TRY(launch_laplacian((const cplx*)F_dev, (cplx*)out_dev, h->links, h->g, nvec, s));
    return EDK_OK;
}

int edk_calc_host(edk_handle* h, const void* U_host, int layout, const void* V_host, int is_c8, void* out_host,
                  void* stream) {
    if (!h || !U_host || !V_host || !out_host) {
        set_error("edk_calc_host: bad argument");
        return EDK_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EDK_ON_DEVICE(h);
    const size_t ub = (size_t)((layout & ~EDK_LINKS_BIG_ENDIAN) == EDK_LINKS_FILE_T ? 4 : 3) * h->g.V * 9 * sizeof(cplx);
    const size_t vb = h->field_cplx * ((is_c8 & EDK_EIGVECS_C8) ? 8 : 16);
    if (h->stage_U_bytes < ub) {
        cudaFree(h->stage_U);
        h->stage_U = nullptr;
        h->stage_U_bytes = 0;
        EDK_CUDA_TRY(cudaMalloc(&h->stage_U, ub));
        h->stage_U_bytes = ub;
    }
    if (h->stage_V_bytes < vb) {
        cudaFree(h->stage_V);
        h->stage_V = nullptr;
        h->stage_V_bytes = 0;
        EDK_CUDA_TRY(cudaMalloc(&h->stage_V, vb));
        h->stage_V_bytes = vb;
    }
    if (!h->stage_out) EDK_CUDA_TRY(cudaMalloc(&h->stage_out, edk_output_bytes(h)));
    EDK_CUDA_TRY(cudaMemcpyAsync(h->stage_U, U_host, ub, cudaMemcpyHostToDevice, s));
    EDK_CUDA_TRY(cudaMemcpyAsync(h->stage_V, V_host, vb, cudaMemcpyHostToDevice, s));
    int rc = edk_set_links(h, h->stage_U, layout, stream);
    if (rc) return rc;
    rc = edk_set_eigvecs(h, h->stage_V, is_c8, stream);
    if (rc) return rc;
    rc = edk_calc(h, h->stage_out, stream);
    if (rc) return rc;
    EDK_CUDA_TRY(cudaMemcpyAsync(out_host, h->stage_out, edk_output_bytes(h), cudaMemcpyDeviceToHost, s));
    EDK_CUDA_TRY(cudaStreamSynchronize(s));
    return EDK_OK;
}

int edk_host_alloc(void** p, size_t bytes) {
    if (!p) return EDK_ERR_ARG;
    EDK_CUDA_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return EDK_OK;
}
int edk_host_free(void* p) {
    EDK_CUDA_TRY(cudaFreeHost(p));
    return EDK_OK;
}

int edk_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return EDK_ERR_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();  // not sticky: e.g. file-backed or already registered ranges; the caller stages instead
        set_error("cudaHostRegister failed: %s", cudaGetErrorString(e));
        return EDK_ERR_CUDA;
    }
    return EDK_OK;
}
int edk_host_unregister(void* p) {
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostUnregister failed: %s", cudaGetErrorString(e));
        return EDK_ERR_CUDA;
    }
    return EDK_OK;
}

int edk_set_profiling(edk_handle* h, int on) {
    if (!h) return EDK_ERR_ARG;
    h->profiling = on != 0;
    recycle_events(h);
    return EDK_OK;
}

int edk_get_profile(edk_handle* h, double ms[4], int n_launch[4]) {
    if (!h || !ms || !n_launch) return EDK_ERR_ARG;
    for (int i = 0; i < PH_COUNT; ++i) {
        ms[i] = 0.0;
        n_launch[i] = h->n_launch[i];
    }
    for (auto& e : h->events) {
        EDK_CUDA_TRY(cudaEventSynchronize(e.b));
        float t = 0.f;
        EDK_CUDA_TRY(cudaEventElapsedTime(&t, e.a, e.b));
        ms[e.phase] += t;
    }
    recycle_events(h);
    return EDK_OK;
}

long long edk_launch_count(const edk_handle* h) { return h ? h->launches : 0; }

int edk_debug_field(edk_handle* h, int idx, void* dst_dev, void* stream) {
    if (!h || !dst_dev || idx < 0 || idx >= h->nfield) {
        set_error("edk_debug_field: bad argument");
        return EDK_ERR_ARG;
    }
    EDK_ON_DEVICE(h);
    EDK_CUDA_TRY(cudaMemcpyAsync(dst_dev, h->field(idx), h->field_cplx * sizeof(cplx), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return EDK_OK;
}

int edk_debug_phase(edk_handle* h, int ip, void* dst_dev, void* stream) {
    if (!h || !dst_dev || ip < 0 || ip >= h->nmom) {
        set_error("edk_debug_phase: bad argument");
        return EDK_ERR_ARG;
    }
    EDK_ON_DEVICE(h);
    EDK_CUDA_TRY(cudaMemcpyAsync(dst_dev, h->phase + (size_t)ip * h->g.Vpad, (size_t)h->g.V * sizeof(cplx),
                                 cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return EDK_OK;
}

int edk_debug_use_naive_gram(edk_handle* h, int on) {
    if (!h) return EDK_ERR_ARG;
    h->naive = on != 0;
    h->cta_dirty = true;
    return EDK_OK;
}

int edk_debug_gram_config(edk_handle* h, int mfrag, int ksplit) {
    if (!h || mfrag < 0 || ksplit < 0) return EDK_ERR_ARG;
    if (mfrag && !gram_mfrag_available(mfrag)) {
        set_error("edk_debug_gram_config: mfrag %d not instantiated (2..13)", mfrag);
        return EDK_ERR_ARG;
    }
    EDK_ON_DEVICE(h);
    h->force_mfrag = mfrag;
    h->force_ksplit = ksplit;
    pick_gram_config(h);
    {
        const int rc = build_tma(h);
        if (rc != EDK_OK) return rc;
    }
    return ensure_partial(h);
}

int edk_debug_loader(edk_handle* h, int mode) {
    if (!h || mode < 0 || mode > 1) return EDK_ERR_ARG;
    EDK_ON_DEVICE(h);
    h->loader = mode;
    pick_gram_config(h);
    return ensure_partial(h);
}

int edk_debug_algo(edk_handle* h, int algo) {
    if (!h || algo < -1 || algo > 4) return EDK_ERR_ARG;
    EDK_ON_DEVICE(h);
    if (algo == 4 && !plan_sep(h->g.Lx, h->mom_int).ok) {
        set_error("edk_debug_algo: the separable form does not cover this lattice / momentum list");
        return EDK_ERR_ARG;
    }
    EDK_CUDA_TRY(cudaDeviceSynchronize());
    h->algo_request = algo;
    h->algo = algo >= 0 ? algo : plan_contraction_form(h->g.Lx, h->g.Ly, h->mom_int, h->jobs_host);
    pick_gram_config(h);
    int rc = build_tma(h);
    if (rc != EDK_OK) return rc;
    // only the buffers of the form in use are kept (the per-plane sums are the largest part of the workspace)
    if (h->algo == 4) {
        free_pw(h);
        if (!h->sep_ready) rc = build_sep(h);
    } else if (h->algo >= 2) {
        free_sep(h);
        if (!h->pw_ready || h->pw_algo != h->algo) rc = build_pw(h);  // the two plane-wave forms have different tables
    }
    if (rc != EDK_OK) return rc;
    return ensure_partial(h);
}

int edk_debug_symmetry(edk_handle* h, int mode) {
    if (!h || mode < -1 || mode > 1) return EDK_ERR_ARG;
    EDK_ON_DEVICE(h);
    EDK_CUDA_TRY(cudaDeviceSynchronize());
    h->sym_request = mode;
    return configure(h);
}

int edk_query(const edk_handle* h, int what) {
    if (!h) return EDK_ERR_ARG;
    switch (what) {
        case 0: return h->symmetric ? 1 : 0;
        case 1: return h->nmom_int;
        case 2: {
            int segs = 0;
            for (const auto& j : h->jobs_host) segs += j.nseg;
            return segs;
        }
        case 3: return h->ksplit;
        case 4: return h->mfrag;
        case 5: return h->njobs;
        case 6: return (h->loader == 0 && h->tma_ready) ? h->tma.nstages : 0;
        case 7: return effective_algo(h) == 4 ? 0 : (effective_algo(h) >= 2 ? 4 - effective_algo(h) : (effective_algo(h) ? 3 : 4));
        case 8: {  // (pair, momentum) GEMMs actually contracted
            int n = 0;
            for (const auto& j : h->jobs_host) n += j.nseg * j.nmom;
            return n;
        }
        case 9: return h->n_half;
        case 10: return effective_algo(h);
        case 11: return effective_algo(h) == 4 ? (h->sep_ready ? h->sep_nmodes : 0) : (h->pw_ready ? h->pw_nmodes : 0);
        case 12: return effective_algo(h) == 4 ? (h->sep_ready ? (h->sep_variant == 6 ? 3232 : 100 * sep_variant_rows(h->sep_variant) + SEP_TF) : 0) : (h->pw_ready ? 10 * h->pw_el + h->pw_fl : 0);
        case 13: return h->algo_request;
        case 14: return effective_algo(h) == 4 && h->sep_ready ? h->sep_pairs : 0;
        case 15: return effective_algo(h) == 4 && h->sep_ready ? h->sep_variant : -1;
        default: return EDK_ERR_ARG;
    }
}

int edk_microbench_fp64(int device, double* dmma_tflops, double* dfma_tflops) {
    if (!dmma_tflops || !dfma_tflops) return EDK_ERR_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("edk_microbench_fp64: cudaSetDevice(%d) failed", device);
        return EDK_ERR_CUDA;
    }
    EDK_CUDA_TRY(microbench_fp64(dmma_tflops, dfma_tflops));
    return EDK_OK;
}

}  // extern "C"
