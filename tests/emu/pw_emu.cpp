// Runs the REAL kernel sources of csrc/edk_gram_pw.cu on the host (tests/emu/edk_emu.h), one thread per CUDA
// thread: pw_weights_kernel, gram_pw_kernel<MB> (TMA producer warp, mbarrier ring, DMMA fragments, epilogue)
// and pw_zfold_kernel.  TEST INFRASTRUCTURE ONLY - driven by tests/test_pw_model.py.
//
//   pw_emu <input.bin> <output.bin>
// input : int32 header {Lx, Ly, Lz, Ne, nfield, njobs, nmom_int, nmodes, max_mb, nstages, el, fl}
//         int32 jobs[njobs][2 + 3*8]  = nseg, nmom, Lf[8], Rf[8], sign[8]
//         int32 modes3[nmodes][3], int32 momode[nmom_int][3]
//         f64   zphase[nmom_int][Lz][2], f64 fields[nfield][Ne][3V][2]
// output: f64 partial[njobs][nmom_int][Ne][Ne][2]
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

#include "edk_gram_pw.cu"

using namespace edk;

template <class Body>
static void run_cta(int nthreads, unsigned bx, Body body) {
    g_cta.nthreads = nthreads;
    for (auto& b : g_cta.mbar) b = EmuMbar{};
    std::memset(smem, 0xff, sizeof(smem));  // NaN pattern: reading what was never written shows up in the result
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=] {
            threadIdx.x = (unsigned)t;
            blockIdx.x = bx;
            blockDim.x = (unsigned)nthreads;
            try {
                body();
            } catch (const std::exception& e) {
                std::fprintf(stderr, "pw_emu: CTA %u thread %d: %s\n", bx, t, e.what());
                std::_Exit(2);
            }
        });
    for (auto& t : th) t.join();
}

// data-parallel kernels without barriers: run the threads of a block one after another
template <class Body>
static void run_simple(unsigned nblocks, int nthreads, Body body) {
    for (unsigned b = 0; b < nblocks; ++b)
        for (int t = 0; t < nthreads; ++t) {
            threadIdx.x = (unsigned)t;
            blockIdx.x = b;
            blockDim.x = (unsigned)nthreads;
            body();
        }
}

template <class T>
static std::vector<T> read_vec(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
        std::fprintf(stderr, "pw_emu: short read\n");
        std::exit(1);
    }
    return v;
}

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    const auto hd = read_vec<int>(f, 12);
    const int Lx = hd[0], Ly = hd[1], Lz = hd[2], Ne = hd[3], nfield = hd[4], njobs = hd[5], nmom = hd[6], nmodes = hd[7],
              max_mb = hd[8], nstages = hd[9], el = hd[10], fl = hd[11];
    const int V = Lx * Ly * Lz, Kc = 3 * V;
    const auto jraw = read_vec<int>(f, (size_t)njobs * 26);
    const auto modes3 = read_vec<int>(f, (size_t)nmodes * 3);
    const auto momode = read_vec<int>(f, (size_t)nmom * 3);
    const auto zphase = read_vec<double>(f, (size_t)nmom * Lz * 2);
    const auto fields = read_vec<double>(f, (size_t)nfield * Ne * Kc * 2);
    std::fclose(f);

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
        }
        jobs[j] = J;
    }
    Geom g{Lx, Ly, Lz, V, (V + 7) / 8 * 8};
    const int mbtot = (nmodes + 7) / 8, kplane = (Lx * Ly + 7) / 8;

    // ---- pw_weights_kernel ------------------------------------------------------------------------
    std::vector<double> wt_raw((size_t)kplane * 2 * mbtot * 32 + 2, -1.0);
    double* wtiles = wt_raw.data();
    if (reinterpret_cast<uintptr_t>(wtiles) % 16) ++wtiles;  // cp.async.bulk wants 16-byte aligned sources
    {
        const size_t total = (size_t)kplane * 2 * mbtot * 32;
        run_simple((unsigned)((total + 255) / 256), 256, [&] { pw_weights_kernel(wtiles, modes3.data(), nmodes, mbtot, kplane, g); });
    }

    // ---- gram_pw_kernel ---------------------------------------------------------------------------
    const size_t mat = (size_t)Ne * Ne;
    std::vector<cplx> Y((size_t)njobs * Lz * nmodes * mat, make_double2(NAN, NAN));
    PwParams P{};
    P.jobs = jobs.data();
    P.njobs = njobs;
    P.Ne = Ne;
    P.Lz = Lz;
    P.A = Lx * Ly;
    P.kplane = kplane;
    const int rows_l = PW_WARPS * el, rows_r = 8 * fl;
    P.n_et = (Ne + rows_l - 1) / rows_l;
    P.n_ft = (Ne + rows_r - 1) / rows_r;
    P.nmodes = nmodes;
    P.mbtot = mbtot;
    P.wtiles = wtiles;
    P.Y = Y.data();
    PwTma T{};
    int plan_nst = 0, smem_bytes = 0;
    if (pw_plan_smem(el, fl, &plan_nst, &smem_bytes) != 0 || smem_bytes > EMU_SMEM_BYTES) return 1;
    T.nstages = nstages > 0 ? nstages : plan_nst;
    EmuTensorMap M{};
    M.base = fields.data();
    M.dim[0] = 2LL * Kc;
    M.dim[1] = Ne;
    M.dim[2] = nfield;
    M.stride_bytes[0] = 2LL * Kc * 8;
    M.stride_bytes[1] = 2LL * Kc * 8 * Ne;
    M.box[0] = 8;
    M.box[2] = 1;
    M.box[1] = rows_l;
    std::memcpy(T.mapL, &M, sizeof(M));
    M.box[1] = rows_r;
    std::memcpy(T.mapR, &M, sizeof(M));
    static_assert(sizeof(EmuTensorMap) <= 128, "descriptor fits the CUtensorMap slot");
    const unsigned items = (unsigned)(njobs * Lz * P.n_et * P.n_ft);
    for (int mb0 = 0; mb0 < mbtot; mb0 += max_mb) {
        P.mb0 = mb0;
        const int MB = std::min(max_mb, mbtot - mb0);
        for (unsigned b = 0; b < items; ++b) {
            if (el == 2 && fl == 4 && MB == 1)
                run_cta(PW_THREADS, b, [&] { gram_pw_kernel<1, 2, 4>(P, T); });
            else if (el == 2 && fl == 4)
                run_cta(PW_THREADS, b, [&] { gram_pw_kernel<2, 2, 4>(P, T); });
            else if (el == 2 && fl == 5 && MB == 1)
                run_cta(PW_THREADS, b, [&] { gram_pw_kernel<1, 2, 5>(P, T); });
            else if (el == 2 && fl == 5)
                run_cta(PW_THREADS, b, [&] { gram_pw_kernel<2, 2, 5>(P, T); });
            else
                return 1;
        }
    }

    // ---- pw_zfold_kernel --------------------------------------------------------------------------
    std::vector<cplx> partial((size_t)njobs * nmom * mat, make_double2(NAN, NAN));
    PwFold F{};
    F.jobs = jobs.data();
    F.njobs = njobs;
    F.Ne = Ne;
    F.Lz = Lz;
    F.nmodes = nmodes;
    F.nmom_int = nmom;
    F.rows_l = rows_l;
    F.rows_r = rows_r;
    F.Y = Y.data();
    F.zphase = reinterpret_cast<const cplx*>(zphase.data());
    F.momode = momode.data();
    F.partial = partial.data();
    const unsigned nblk = (unsigned)((mat + PW_FOLD_THREADS - 1) / PW_FOLD_THREADS);
    run_simple(nblk * njobs * nmodes, PW_FOLD_THREADS, [&] { pw_zfold_kernel(F); });  // one block per (job, 256 elements, mode)

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 1;
    std::fwrite(partial.data(), sizeof(cplx), partial.size(), o);
    std::fclose(o);
    return 0;
}
