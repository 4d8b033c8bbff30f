// Host stand-in for the CUDA runtime, for the whole-library emulator build (TEST INFRASTRUCTURE ONLY).
//
// tests/test_emu_library.py compiles csrc/*.cu with g++ -DEDK_HOST_EMU together with this file into
// libedk_emu.so: the C ABI of include/edk.h with "device memory" = host memory, kernel launches = host threads
// (tests/emu/edk_emu.h, emu_launch) and cuTensorMapEncodeTiled = the emulator's own box descriptor.  What this
// buys: the host glue of csrc/edk_api.cu (job lists, buffer sizes, tensor maps, launch configurations, the
// switching between contraction forms) and every launcher run on a machine without a GPU, end to end against the
// oracle.  It is never loaded by the package (easydistillation_b200/_capi.py refuses a library that exports the
// marker edk_host_emulator_build below).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "edk_common.cuh"

thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
namespace edk {
alignas(1024) unsigned char smem[EMU_SMEM_BYTES];
EmuCta g_cta;

static thread_local cudaError_t g_last = cudaSuccess;
static thread_local char g_launch_msg[256] = "";

cudaError_t emu_launch_error(cudaError_t e, const char* what) {
    g_last = e;
    std::snprintf(g_launch_msg, sizeof(g_launch_msg), "%s", what);
    std::fprintf(stderr, "edk emulator: launch failed: %s\n", what);
    return e;
}

void emu_run_threads(int nthreads, const std::function<void(int)>& fn) {
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&fn, t] { fn(t); });
    for (auto& t : th) t.join();
}
}  // namespace edk

namespace {

struct FakeEvent {
    std::chrono::steady_clock::time_point t;
};

// cuTensorMapEncodeTiled for the emulator: rank-3 FLOAT64 maps only, which is all the library builds
CUresult fake_encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim,
                           const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride, CUtensorMapInterleave il,
                           CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill fill) {
    // the constraints of the real encoder that the library could violate
    if (!map || !base || rank != 3 || dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || il != CU_TENSOR_MAP_INTERLEAVE_NONE ||
        (sw != CU_TENSOR_MAP_SWIZZLE_NONE && sw != CU_TENSOR_MAP_SWIZZLE_128B) || fill != CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        return CUDA_ERROR_INVALID_VALUE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * 8 > 128) return CUDA_ERROR_INVALID_VALUE;  // inner extent <= swizzle span
    if (reinterpret_cast<uintptr_t>(map) % 64 || reinterpret_cast<uintptr_t>(base) % 16) return CUDA_ERROR_INVALID_VALUE;
    for (int i = 0; i < 3; ++i)
        if (gdim[i] < 1 || gdim[i] > (1ULL << 32) || box[i] < 1 || box[i] > 256 || estride[i] != 1) return CUDA_ERROR_INVALID_VALUE;
    if ((box[0] * 8) % 16) return CUDA_ERROR_INVALID_VALUE;  // inner box extent is a multiple of 16 bytes
    for (int i = 0; i < 2; ++i)
        if (gstride[i] % 16 || gstride[i] >= (1ULL << 40)) return CUDA_ERROR_INVALID_VALUE;
    edk::EmuTensorMap M{};
    M.base = static_cast<const double*>(base);
    for (int i = 0; i < 3; ++i) {
        M.dim[i] = (long long)gdim[i];
        M.box[i] = (int)box[i];
    }
    M.stride_bytes[0] = (long long)gstride[0];
    M.stride_bytes[1] = (long long)gstride[1];
    M.swizzle128 = sw == CU_TENSOR_MAP_SWIZZLE_128B;
    static_assert(sizeof(edk::EmuTensorMap) <= sizeof(CUtensorMap), "descriptor fits");
    std::memset(map, 0, sizeof(CUtensorMap));
    std::memcpy(map, &M, sizeof(M));
    return CUDA_SUCCESS;
}

}  // namespace

extern "C" {

// marker by which easydistillation_b200/_capi.py refuses to load this build: it is test infrastructure, not a CPU path
int edk_host_emulator_build(void) { return 1; }

cudaError_t cudaSetDevice(int dev) { return dev == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDevice(int* dev) {
    *dev = 0;
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) {
    const cudaError_t e = edk::g_last;
    edk::g_last = cudaSuccess;
    return e;
}
const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        case cudaErrorLaunchFailure: return edk::g_launch_msg[0] ? edk::g_launch_msg : "unspecified launch failure";
        case cudaErrorNotSupported: return "operation not supported";
        default: return "emulated CUDA error";
    }
}
cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (!p) return cudaErrorInvalidValue;
    // cudaMalloc returns at least 256-byte aligned memory; fill with a NaN pattern so that reads of memory the
    // library never wrote show up in the results
    void* q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) return cudaErrorMemoryAllocation;
    std::memset(q, 0xff, bytes);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) {
    std::memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = reinterpret_cast<cudaEvent_t>(new FakeEvent());
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete reinterpret_cast<FakeEvent*>(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    reinterpret_cast<FakeEvent*>(e)->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(reinterpret_cast<FakeEvent*>(b)->t - reinterpret_cast<FakeEvent*>(a)->t).count();
    return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void* fn, cudaFuncAttribute attr, int value) {
    if (!fn) return cudaErrorInvalidValue;
    if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize && value > 227 * 1024) return cudaErrorInvalidValue;
    return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res) {
    if (std::strcmp(symbol, "cuTensorMapEncodeTiled") != 0) {
        if (res) *res = cudaDriverEntryPointSymbolNotFound;
        return cudaErrorInvalidValue;
    }
    *fn = reinterpret_cast<void*>(&fake_encode_tiled);
    if (res) *res = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}

}  // extern "C"
