// Calibration of the host emulator (tests/emu/edk_emu.h) against kernels that are validated on the GPU: runs
// the REAL sources of csrc/edk_gram.cu - phase_tiles_kernel, gram_tma_kernel<2, ALGO> (TMA producer warp +
// mbarrier ring; ALGO 1 = 3M with the Re+Im plane, 0 = 4M) or gram_dmma_kernel<2> (cp.async ring) - on the
// host.  If these reproduce the direct contraction, the emulator's TMA-box, mbarrier and m8n8k4 semantics are the
// ones the hardware has.  TEST INFRASTRUCTURE ONLY - driven by tests/test_pw_model.py.
//
//   gram_emu <input.bin> <output.bin>
// input : int32 header {Lx, Ly, Lz, Ne, nfield, njobs, nmom, kernel (0 tma 4M, 1 tma 3M, 2 cp.async 4M), ksplit}
//         int32 jobs[njobs][26] = nseg, nmom, Lf[8], Rf[8], sign[8]
//         f64 phase[2][nmom][Vpad][2] (phase, then -i*phase), f64 fields[nfield][Ne][3V][2]
// output: f64 partial[ksplit][njobs][nmom][Ne][Ne][2]
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define EDK_EMU_NO_LAUNCHERS  // this harness drives the kernels itself
#include "edk_emu.h"

thread_local EmuIdx threadIdx, blockIdx, blockDim;
namespace edk {
alignas(1024) unsigned char smem[EMU_SMEM_BYTES];
EmuCta g_cta;
}  // namespace edk

#include "edk_gram.cu"

using namespace edk;

template <class Body>
static void run_cta(int nthreads, unsigned bx, unsigned by, Body body) {
    g_cta.nthreads = nthreads;
    for (auto& b : g_cta.mbar) b = EmuMbar{};
    std::memset(smem, 0xff, sizeof(smem));
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=] {
            threadIdx.x = (unsigned)t;
            blockIdx.x = bx;
            blockIdx.y = by;
            blockDim.x = (unsigned)nthreads;
            try {
                body();
            } catch (const std::exception& e) {
                std::fprintf(stderr, "gram_emu: CTA (%u,%u) thread %d: %s\n", bx, by, t, e.what());
                std::_Exit(2);
            }
        });
    for (auto& t : th) t.join();
}

template <class T>
static std::vector<T> read_vec(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) std::exit(1);
    return v;
}

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    const auto hd = read_vec<int>(f, 9);
    const int Lx = hd[0], Ly = hd[1], Lz = hd[2], Ne = hd[3], nfield = hd[4], njobs = hd[5], nmom = hd[6], kernel = hd[7],
              ksplit = hd[8];
    const int V = Lx * Ly * Lz, Kc = 3 * V, Vpad = (V + 7) / 8 * 8;
    const auto jraw = read_vec<int>(f, (size_t)njobs * 26);
    const auto phase = read_vec<double>(f, (size_t)2 * nmom * Vpad * 2);
    auto fields = read_vec<double>(f, (size_t)nfield * Ne * Kc * 2);
    std::fclose(f);
    const cplx* fld = reinterpret_cast<const cplx*>(fields.data());

    const int mfrag = 2, algo = kernel == 1 ? 1 : 0;
    std::vector<GramJob> jobs(njobs);
    for (int j = 0; j < njobs; ++j) {
        const int* r = &jraw[(size_t)j * 26];
        GramJob J{};
        J.nseg = r[0];
        J.nmom = r[1];
        for (int s = 0; s < 8; ++s) {
            J.Lf[s] = r[2 + s];
            J.Rf[s] = r[10 + s];
            J.sign[s] = r[18 + s];
            J.L[s] = fld + (size_t)J.Lf[s] * Ne * Kc;
            J.R[s] = fld + (size_t)J.Rf[s] * Ne * Kc;
        }
        jobs[j] = J;
    }
    // Re + Im planes, rows padded to an even number of doubles (edk_api.cu: sum_row)
    const size_t sum_row = ((size_t)Kc + 1) & ~(size_t)1;
    std::vector<double> fsum((size_t)nfield * Ne * sum_row, 0.0);
    for (int a = 0; a < nfield; ++a)
        for (int e = 0; e < Ne; ++e)
            for (int k = 0; k < Kc; ++k) {
                const cplx v = fld[((size_t)a * Ne + e) * Kc + k];
                fsum[((size_t)a * Ne + e) * sum_row + k] = v.x + v.y;
            }
    // phase tiles [kstep][2][nmom][8]
    std::vector<cplx> tiles((size_t)2 * nmom * Vpad);
    {
        const size_t n = (size_t)2 * nmom * Vpad;
        for (unsigned b = 0; b < (n + 255) / 256; ++b)
            for (int t = 0; t < 256; ++t) {
                threadIdx.x = t;
                blockIdx.x = b;
                blockDim.x = 256;
                phase_tiles_kernel(reinterpret_cast<const cplx*>(phase.data()), tiles.data(), nmom, Vpad);
            }
    }
    // CTA map as edk_api.cu builds it: tiles of one job contiguous, column tile fastest
    const int rows = gram_rows_per_tile(mfrag);
    const int n_mt = (Ne + rows - 1) / rows;
    const int fw = gram_fwidth(algo), ntile = gram_nfrag_per_tile(algo);
    std::vector<int2> cta_map;
    for (int j = 0; j < njobs; ++j) {
        const int nfrag_f = (Ne + fw - 1) / fw;
        const int n_nt = (nfrag_f * jobs[j].nmom + ntile - 1) / ntile;
        for (int t = 0; t < n_mt * n_nt; ++t) cta_map.push_back(make_int2(j, t));
    }
    std::vector<cplx> partial((size_t)ksplit * njobs * nmom * Ne * Ne, make_double2(0.0, 0.0));
    GramParams P{};
    P.jobs = jobs.data();
    P.njobs = njobs;
    P.Ne = Ne;
    P.nmom = nmom;
    P.Kc = Kc;
    P.ksteps = Vpad / 8;
    P.Vpad = Vpad;
    P.ksplit = ksplit;
    P.n_mt = n_mt;
    P.ncta = (int)cta_map.size();
    P.cta_map = cta_map.data();
    P.phase = reinterpret_cast<const cplx*>(phase.data());
    P.partial = partial.data();

    if (kernel == 2) {
        for (int by = 0; by < ksplit; ++by)
            for (int bx = 0; bx < P.ncta; ++bx) run_cta(GRAM_NTHREADS, bx, by, [&] { gram_dmma_kernel<2>(P); });
    } else {
        GramTma T{};
        int brows = 0, nst = 0, bytes = 0;
        if (gram_tma_plan(algo, mfrag, nmom, Ne, &brows, &nst, &bytes) != 0 || bytes > EMU_SMEM_BYTES) return 1;
        T.brows_alloc = brows;
        T.nstages = nst;
        T.phase_tiles = tiles.data();
        EmuTensorMap M{};
        M.base = fields.data();
        M.dim[0] = 2LL * Kc;
        M.dim[1] = Ne;
        M.dim[2] = nfield;
        M.stride_bytes[0] = 2LL * Kc * 8;
        M.stride_bytes[1] = 2LL * Kc * 8 * Ne;
        M.box[0] = 8;
        M.box[1] = rows;
        M.box[2] = 1;
        std::memcpy(T.mapA, &M, sizeof(M));
        M.box[1] = 8;
        std::memcpy(T.mapB, &M, sizeof(M));
        EmuTensorMap S{};
        S.base = fsum.data();
        S.dim[0] = Kc;
        S.dim[1] = Ne;
        S.dim[2] = nfield;
        S.stride_bytes[0] = (long long)sum_row * 8;
        S.stride_bytes[1] = (long long)sum_row * 8 * Ne;
        S.box[0] = 4;
        S.box[1] = rows;
        S.box[2] = 1;
        std::memcpy(T.mapS, &S, sizeof(S));
        for (int by = 0; by < ksplit; ++by)
            for (int bx = 0; bx < P.ncta; ++bx) {
                if (algo)
                    run_cta(GT_THREADS, bx, by, [&] { gram_tma_kernel<2, 1>(P, T); });
                else
                    run_cta(GT_THREADS, bx, by, [&] { gram_tma_kernel<2, 0>(P, T); });
            }
    }
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 1;
    std::fwrite(partial.data(), sizeof(cplx), partial.size(), o);
    std::fclose(o);
    return 0;
}
