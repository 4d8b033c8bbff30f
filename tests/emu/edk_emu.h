// Host emulation of the device helpers in csrc/edk_pipe.cuh (TEST INFRASTRUCTURE ONLY).
//
// The contraction kernels are compiled by g++ with -DEDK_HOST_EMU and executed with one host thread per
// CUDA thread of a CTA, CTAs one after another:
//   * shared memory  = one global buffer, "shared addresses" are byte offsets into it
//   * mbarrier       = arrival count + transaction bytes + phase counter (try_wait.parity semantics of PTX:
//                      waiting on parity P succeeds once the phase with that parity has completed; a fresh
//                      barrier is in phase 0, so waiting on parity 1 succeeds at once)
//   * TMA            = cp.async.bulk.tensor.3d over a tiled descriptor {base, dims, byte strides, box}: box
//                      written dimension-0 fastest, out-of-range elements zero-filled, the full box size is
//                      signalled to the barrier (complete_tx) whether or not parts were out of range
//   * DMMA           = mma.sync.m8n8k4 row.col f64 by exchanging the operands of the 32 lanes of the calling
//                      warp: A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
//                      C/D[row = lane/4][col = 2 (lane%4) + {0,1}]
// These are the semantics the GPU-validated gram_tma_kernel already relies on.
#pragma once
#include <math.h>

#include <algorithm>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <string>
#include <mutex>
#include <stdexcept>

#include "edk_common.cuh"

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

struct EmuIdx {
    unsigned x = 0, y = 0, z = 0;
};
extern thread_local EmuIdx threadIdx, blockIdx, blockDim;
#ifndef EDK_EMU_NO_LAUNCHERS
extern thread_local EmuIdx gridDim;  // only the whole-library build (emu_runtime.cpp) defines it
#endif
// linear thread index inside the CTA (x fastest), which is what warps are made of; harnesses that only set
// threadIdx.x and blockDim.x get threadIdx.x
inline int emu_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }

namespace edk {

constexpr int EMU_SMEM_BYTES = 232448;
extern unsigned char smem[EMU_SMEM_BYTES];  // `extern __shared__ unsigned char smem[]` of the kernels binds to this

// ---- CTA-wide state (reset by the driver before every CTA) --------------------------------------------
struct EmuBarrier {  // reusable counting barrier; threads that have left the kernel no longer count
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0, generation = 0, gone = 0;
    void release_locked() {
        waiting = 0;
        ++generation;
        cv.notify_all();
    }
    void wait(int n) {
        std::unique_lock<std::mutex> lk(m);
        const int gen = generation;
        if (++waiting >= n - gone) {
            release_locked();
        } else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen; })) {
            throw std::runtime_error("barrier timed out (divergent __syncthreads / __syncwarp?)");
        }
    }
    void leave(int n) {  // the calling thread returned from the kernel
        std::lock_guard<std::mutex> lk(m);
        ++gone;
        if (waiting > 0 && waiting >= n - gone) release_locked();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(m);
        waiting = 0;
        gone = 0;
    }
};
struct EmuMbar {
    bool live = false;
    int init = 0, pending = 0;
    long long tx = 0;
    unsigned phase = 0;
};
struct EmuCta {
    int nthreads = 0;
    EmuBarrier cta_barrier;
    EmuBarrier warp_barrier[32];
    double xa[2][32][32], xb[2][32][32];
    std::mutex mb_mutex;
    std::condition_variable mb_cv;
    EmuMbar mbar[EMU_SMEM_BYTES / 8];
    uint32_t tmem[128][512];
    int tmem_cols = 0;  // columns allocated by tcgen05.alloc (0 = none); a CTA that exits with columns allocated is a bug
    bool failed = false;
};
extern EmuCta g_cta;

struct EmuTensorMap {  // what the driver stores in the 128 bytes of a CUtensorMap
    const double* base;
    long long dim[3];
    long long stride_bytes[2];  // of dimensions 1 and 2
    int box[3];
    int swizzle128;  // CU_TENSOR_MAP_SWIZZLE_128B: 16-byte slot s of box row r lands at slot s ^ (r % 8) of its 128-byte row
};

inline uint32_t __cvta_generic_to_shared(const void* p) {
    const long long off = (const unsigned char*)p - smem;
    if (off < 0 || off >= EMU_SMEM_BYTES) throw std::runtime_error("pointer outside shared memory");
    return (uint32_t)off;
}
inline void __syncthreads() { g_cta.cta_barrier.wait(g_cta.nthreads); }
inline int emu_warp_size(int warp) { return std::min(32, g_cta.nthreads - 32 * warp); }  // the last warp may be partial
inline void __syncwarp() {
    const int warp = emu_tid() >> 5;
    g_cta.warp_barrier[warp].wait(emu_warp_size(warp));
}

inline void dmma884(double& c0, double& c1, const double a, const double b) {
    // operands are exchanged through one of two buffers, alternating per call, so that one warp barrier per DMMA is
    // enough: a lane can only overwrite buffer k after every lane has passed the barrier of call k+1, i.e. finished
    // reading buffer k.  The lanes of a warp execute the same sequence of DMMAs (mma.sync.aligned), so their private
    // call counters stay in step.
    static thread_local unsigned turn = 0;
    const int warp = emu_tid() >> 5, lane = emu_tid() & 31;
    double(*xa)[32] = g_cta.xa[turn & 1], (*xb)[32] = g_cta.xb[turn & 1];
    ++turn;
    xa[warp][lane] = a;
    xb[warp][lane] = b;
    g_cta.warp_barrier[warp].wait(emu_warp_size(warp));
    const int row = lane >> 2;
    for (int q = 0; q < 2; ++q) {
        const int col = 2 * (lane & 3) + q;
        double acc = q ? c1 : c0;
        for (int k = 0; k < 4; ++k) acc = fma(xa[warp][row * 4 + k], xb[warp][col * 4 + k], acc);
        (q ? c1 : c0) = acc;
    }
}

inline double flip_sign(double x) { return -x; }
inline double2 lds128_again(const void* p) { return *reinterpret_cast<const double2*>(p); }

using std::max;
using std::min;

// cp.async (16 bytes, zero-fill when !valid), executed synchronously: the kernels wait and barrier before use
inline void cp_async16(uint32_t dst, const void* src, bool valid) {
    if (dst % 16 != 0 || (long long)dst + 16 > EMU_SMEM_BYTES) throw std::runtime_error("bad cp.async destination");
    if (valid)
        std::memcpy(smem + dst, src, 16);
    else
        std::memset(smem + dst, 0, 16);
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}

inline void mbar_init_fence() {}
template <int N>
inline void warpgroup_reg_dealloc() {
    static_assert(N % 8 == 0 && N >= 24 && N <= 256, "setmaxnreg operand");
}
template <int N>
inline void warpgroup_reg_alloc() {
    static_assert(N % 8 == 0 && N >= 24 && N <= 256, "setmaxnreg operand");
}

inline EmuMbar& emu_bar(uint32_t bar) {
    if (bar % 8 != 0 || bar >= (uint32_t)EMU_SMEM_BYTES) throw std::runtime_error("bad mbarrier address");
    return g_cta.mbar[bar / 8];
}
inline void emu_bar_check_complete(EmuMbar& b) {  // caller holds mb_mutex
    if (b.pending == 0 && b.tx == 0) {
        ++b.phase;
        b.pending = b.init;
        g_cta.mb_cv.notify_all();
    }
}
inline void mbar_init(uint32_t bar, uint32_t count) {
    std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
    EmuMbar& b = emu_bar(bar);
    b.live = true;
    b.init = b.pending = (int)count;
    b.tx = 0;
    b.phase = 0;
}
inline void mbar_arrive(uint32_t bar) {
    std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
    EmuMbar& b = emu_bar(bar);
    if (!b.live || b.pending <= 0) {
        g_cta.failed = true;
        throw std::runtime_error("arrive on an uninitialised or over-arrived mbarrier");
    }
    --b.pending;
    emu_bar_check_complete(b);
}
inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
    EmuMbar& b = emu_bar(bar);
    if (!b.live || b.pending <= 0) {
        g_cta.failed = true;
        throw std::runtime_error("arrive.expect_tx on an uninitialised or over-arrived mbarrier");
    }
    b.tx += bytes;
    --b.pending;
    emu_bar_check_complete(b);
}
inline void emu_complete_tx(uint32_t bar, long long bytes) {
    std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
    EmuMbar& b = emu_bar(bar);
    if (!b.live) throw std::runtime_error("complete_tx on an uninitialised mbarrier");
    b.tx -= bytes;
    emu_bar_check_complete(b);
}
inline void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef EDK_EMU_BREAK_PROTOCOL  // self-test of the race detection: waits that do not wait
    (void)bar;
    (void)parity;
    return;
#endif
    std::unique_lock<std::mutex> lk(g_cta.mb_mutex);
    EmuMbar& b = emu_bar(bar);
    if (!b.live) throw std::runtime_error("wait on an uninitialised mbarrier");
    // a hang in the kernel's protocol becomes a test failure instead of a stuck test
    if (!g_cta.mb_cv.wait_for(lk, std::chrono::seconds(20), [&] { return (b.phase & 1u) != parity; })) {
        g_cta.failed = true;
        throw std::runtime_error("mbarrier wait timed out (pipeline protocol bug)");
    }
}

inline void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    EmuTensorMap T;
    std::memcpy(&T, tmap, sizeof(T));
    if (dst % 128 != 0) throw std::runtime_error("TMA destination not 128-byte aligned");
    const long long bytes = 8LL * T.box[0] * T.box[1] * T.box[2];
    if ((long long)dst + bytes > EMU_SMEM_BYTES) throw std::runtime_error("TMA box past the end of shared memory");
    double* out = reinterpret_cast<double*>(smem + dst);
    if (T.swizzle128) {
        // the hardware XORs address bits [4:6] with bits [7:9] of the shared-memory address; with 128-byte box rows
        // that is slot ^ (row % 8) provided the tile starts on a 1024-byte boundary
        if (T.box[0] != 16) throw std::runtime_error("128-byte swizzle needs an inner box of 128 bytes");
        if (dst % 1024 != 0) throw std::runtime_error("swizzled TMA destination not 1024-byte aligned");
    }
    for (int k = 0; k < T.box[2]; ++k)
        for (int j = 0; j < T.box[1]; ++j)
            for (int i = 0; i < T.box[0]; ++i) {
                const long long x0 = (long long)c0 + i, x1 = (long long)c1 + j, x2 = (long long)c2 + k;
                double v = 0.0;
                if (x0 >= 0 && x0 < T.dim[0] && x1 >= 0 && x1 < T.dim[1] && x2 >= 0 && x2 < T.dim[2])
                    v = *reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(T.base) + x0 * 8 +
                                                         x1 * T.stride_bytes[0] + x2 * T.stride_bytes[1]);
                if (T.swizzle128) {
                    const long long r = (long long)k * T.box[1] + j;
                    out[r * 16 + (((i >> 1) ^ (int)(r & 7)) << 1) + (i & 1)] = v;
                } else {
                    *out++ = v;
                }
            }
    emu_complete_tx(bar, bytes);
}
inline void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if (dst % 16 != 0 || bytes % 16 != 0 || bytes == 0 || reinterpret_cast<uintptr_t>(src) % 16 != 0)
        throw std::runtime_error("cp.async.bulk needs 16-byte aligned addresses and size");
    if ((long long)dst + bytes > EMU_SMEM_BYTES) throw std::runtime_error("bulk copy past the end of shared memory");
    std::memcpy(smem + dst, src, bytes);
    emu_complete_tx(bar, bytes);
}

// ---- tensor memory: 128 lanes x 512 columns of 32-bit words per CTA; warp w reaches lanes 32 (w % 4) .. + 31 ----------
inline void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) throw std::runtime_error("tcgen05.alloc: columns must be a power of two in [32, 512]");
    if (smem_dst % 4 != 0 || smem_dst + 4 > (uint32_t)EMU_SMEM_BYTES) throw std::runtime_error("tcgen05.alloc: bad shared-memory address");
    if ((emu_tid() & 31) == 0) {
        std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
        if (g_cta.tmem_cols) throw std::runtime_error("tcgen05.alloc: tensor memory already allocated by this CTA");
        g_cta.tmem_cols = (int)ncols;
        const uint32_t base = 0;
        std::memcpy(smem + smem_dst, &base, 4);
    }
    g_cta.warp_barrier[emu_tid() >> 5].wait(emu_warp_size(emu_tid() >> 5));  // .sync.aligned
}
inline void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if ((emu_tid() & 31) == 0) {
        std::lock_guard<std::mutex> lk(g_cta.mb_mutex);
        if (taddr != 0 || (int)ncols != g_cta.tmem_cols) throw std::runtime_error("tcgen05.dealloc does not match the allocation");
        g_cta.tmem_cols = 0;
    }
    g_cta.warp_barrier[emu_tid() >> 5].wait(emu_warp_size(emu_tid() >> 5));
}
inline void tmem_fence_before_sync() {}
inline void tmem_fence_after_sync() {}
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline uint32_t* emu_tmem_row(uint32_t taddr, int ncols) {
    const int warp = emu_tid() >> 5, lane = emu_tid() & 31;
    const int lane0 = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    if (lane0 != 32 * (warp & 3)) throw std::runtime_error("TMEM access outside the warp's lane quadrant");
    if (col < 0 || col + ncols > g_cta.tmem_cols) throw std::runtime_error("TMEM access outside the allocated columns");
    return &g_cta.tmem[lane0 + lane][col];
}
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) { std::memcpy(r, emu_tmem_row(taddr, 32), sizeof(r)); }
inline void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) { std::memcpy(emu_tmem_row(taddr, 32), r, sizeof(r)); }
inline void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) { std::memcpy(r, emu_tmem_row(taddr, 16), sizeof(r)); }
inline void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) { std::memcpy(emu_tmem_row(taddr, 16), r, sizeof(r)); }
template <int N>
inline double tmem_get_f64(const uint32_t (&r)[N], int k) {
    double v;
    std::memcpy(&v, &r[2 * k], 8);
    return v;
}
template <int N>
inline void tmem_put_f64(uint32_t (&r)[N], int k, double v) { std::memcpy(&r[2 * k], &v, 8); }

inline void sincospi(double x, double* s, double* c) {
    // exact on multiples of 1/2, like the device function
    const double r = x - 2.0 * floor(x / 2.0);  // [0, 2)
    if (r == 0.0) { *s = 0.0; *c = 1.0; }
    else if (r == 0.5) { *s = 1.0; *c = 0.0; }
    else if (r == 1.0) { *s = 0.0; *c = -1.0; }
    else if (r == 1.5) { *s = -1.0; *c = 0.0; }
    else { *s = sin(3.14159265358979323846 * r); *c = cos(3.14159265358979323846 * r); }
}


// ---- device intrinsics of the stencil / gauge / preparation kernels ---------------------------------------
template <class T>
inline T __ldg(const T* p) { return *p; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned n = (sel >> (4 * i)) & 0x7;  // selectors 8..15 (sign replication) are not used by the kernels
        r |= (unsigned)((src >> (8 * n)) & 0xff) << (8 * i);
    }
    return r;
}
inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __double2float_rn(double x) { return (float)x; }  // round-to-nearest-even is the host default

#ifndef EDK_EMU_NO_LAUNCHERS
// ---- kernel launch: one host thread per CUDA thread, CTAs one after another ----------------------------------
// Used by EDK_LAUNCH (csrc/edk_common.cuh) when the whole library is built for the host (tests/emu/emu_runtime.cpp).
// The threads are created once per launch and walk the grid together; between two CTAs they meet at a barrier and
// thread 0 resets the per-CTA state (mbarriers, participation counts, shared memory poisoned with a NaN pattern).
cudaError_t emu_launch_error(cudaError_t e, const char* what);  // records cudaGetLastError + message (emu_runtime.cpp)
void emu_run_threads(int nthreads, const std::function<void(int)>& fn);

template <class Body>
cudaError_t emu_launch(dim3 grid, dim3 block, size_t smem_bytes, Body body) {
    const long long nthreads = (long long)block.x * block.y * block.z;
    const long long nctas = (long long)grid.x * grid.y * grid.z;
    if (nthreads < 1 || nthreads > 1024 || nctas < 1 || grid.y > 65535 || grid.z > 65535)
        return emu_launch_error(cudaErrorInvalidConfiguration, "launch configuration");
    if (smem_bytes > (size_t)EMU_SMEM_BYTES) return emu_launch_error(cudaErrorInvalidValue, "dynamic shared memory above 227 KB");
    g_cta.nthreads = (int)nthreads;
    g_cta.failed = false;
    EmuBarrier between;  // all threads of the launch, between two CTAs
    std::mutex err_m;
    std::string err;
    emu_run_threads((int)nthreads, [&](int t) {
        for (long long c = 0; c < nctas; ++c) {
            if (t == 0) {
                for (auto& b : g_cta.mbar) b.live = false;
                if (g_cta.tmem_cols != 0 && !g_cta.failed) {
                    std::lock_guard<std::mutex> lk(err_m);
                    if (err.empty()) err = "a CTA exited without tcgen05.dealloc";
                    g_cta.failed = true;
                }
                g_cta.tmem_cols = 0;
                std::memset(g_cta.tmem, 0xff, sizeof(g_cta.tmem));
                g_cta.cta_barrier.reset();
                for (auto& w : g_cta.warp_barrier) w.reset();
                std::memset(smem, 0xff, smem_bytes ? smem_bytes : 0);
            }
            between.wait((int)nthreads);
            threadIdx.x = (unsigned)(t % block.x);
            threadIdx.y = (unsigned)((t / block.x) % block.y);
            threadIdx.z = (unsigned)(t / (block.x * block.y));
            blockIdx.x = (unsigned)(c % grid.x);
            blockIdx.y = (unsigned)((c / grid.x) % grid.y);
            blockIdx.z = (unsigned)(c / ((long long)grid.x * grid.y));
            blockDim.x = block.x, blockDim.y = block.y, blockDim.z = block.z;
            gridDim.x = grid.x, gridDim.y = grid.y, gridDim.z = grid.z;
            try {
                body();
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(err_m);
                if (err.empty()) err = e.what();
                g_cta.failed = true;
            }
            // a thread that has returned no longer takes part in the CTA's barriers
            g_cta.cta_barrier.leave((int)nthreads);
            g_cta.warp_barrier[t >> 5].leave(emu_warp_size(t >> 5));
            between.wait((int)nthreads);
            if (g_cta.failed) return;
        }
    });
    if (!err.empty()) return emu_launch_error(cudaErrorLaunchFailure, err.c_str());
    return cudaSuccess;
}
#endif  // EDK_EMU_NO_LAUNCHERS

}  // namespace edk
