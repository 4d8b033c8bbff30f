// Device helpers shared by the contraction kernels: DMMA.8x8x4, mbarrier ring, TMA / bulk copies (sm_100a).
#pragma once
#include "edk_common.cuh"

#ifdef EDK_HOST_EMU
// tests/emu: the same helpers executed by host threads (one per CUDA thread), so that the kernel sources
// themselves can be run on a machine without a GPU.  Test infrastructure only, never part of the library.
#include "edk_emu.h"
#else
namespace edk {

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> 16 zero bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// A 16-byte shared-memory load the compiler neither merges with an identical earlier one nor hoists out of an
// unrolled loop: re-reading a broadcast operand costs one wavefront, keeping it costs four registers.
__device__ __forceinline__ double2 lds128_again(const void* p) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// sign flip on the integer pipe: the FP64 pipe is the one the DMMAs need
__device__ __forceinline__ double flip_sign(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// make the initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// setmaxnreg: a producer warpgroup hands registers to the MMA warpgroups (all warps of a warpgroup call it)
template <int N>
__device__ __forceinline__ void warpgroup_reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void warpgroup_reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// A wait that cannot hang the GPU: a pipeline bug traps after ~1 s instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spin >= 64) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > (1LL << 31)) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace edk
#endif  // EDK_HOST_EMU
