#include <cstdint>
#include <cstdio>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
#define R8(a, o) "%" #a
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// every warp keeps 16 doubles per lane in TMEM, adds to them a few times, reads back
__global__ void __launch_bounds__(256) tk(double* out, int iters) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc((uint32_t)__cvta_generic_to_shared(&base_s), 512);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t taddr = base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 32;
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0;
    tmem_st32(taddr, r);
    tmem_wait_st();
    for (int it = 0; it < iters; ++it) {
        tmem_ld32(taddr, r);
        tmem_wait_ld();
        for (int k = 0; k < 16; ++k) {
            double v = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
            v += (double)(threadIdx.x * 16 + k) + 0.25;
            r[2 * k] = (uint32_t)__double2loint(v);
            r[2 * k + 1] = (uint32_t)__double2hiint(v);
        }
        tmem_st32(taddr, r);
        tmem_wait_st();
    }
    tmem_ld32(taddr, r);
    tmem_wait_ld();
    for (int k = 0; k < 16; ++k) out[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16 + k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(base_s, 512);
}
int main() {
    const int nb = 300, nt = 256, iters = 5;
    double* d;
    cudaMalloc(&d, sizeof(double) * nb * nt * 16);
    tk<<<nb, nt>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    double* h = (double*)malloc(sizeof(double) * nb * nt * 16);
    cudaMemcpy(h, d, sizeof(double) * nb * nt * 16, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int b = 0; b < nb; ++b)
        for (int t = 0; t < nt; ++t)
            for (int k = 0; k < 16; ++k)
                if (h[((size_t)b * nt + t) * 16 + k] != iters * ((double)(t * 16 + k) + 0.25)) ++bad;
    printf("TMEM_TEST bad=%ld\n", bad);
    return bad != 0;
}
