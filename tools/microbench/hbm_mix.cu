// HBM ceilings for different read:write mixes on B200: copy (1:1), the stencil's mix (1 read : 3 writes),
// write-only and read-only streams.  All coalesced 16-byte accesses, grid-stride, 1 GiB per array.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_mix hbm_mix.cu && ./hbm_mix
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_1r3w(const double2* __restrict__ a, double2* __restrict__ b, double2* __restrict__ c, double2* __restrict__ d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = a[i];
        b[i] = v;
        c[i] = make_double2(2.0 * v.x, 2.0 * v.y);
        d[i] = make_double2(3.0 * v.x, 3.0 * v.y);
    }
}
__global__ void k_write(double2* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_double2(1.0, 2.0);
}
__global__ void k_read(const double2* __restrict__ a, double* sink, size_t n) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = a[i];
        s += v.x + v.y;
    }
    if (s == 1.2345) *sink = s;
}

template <class F>
static double best_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t n = (size_t)1 << 26;  // 64 Mi double2 = 1 GiB per array
    double2 *a, *b, *c, *d; double* sink;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16); cudaMalloc(&d, n * 16); cudaMalloc(&sink, 8);
    cudaMemset(a, 0, n * 16);
    const int grid = 148 * 16, block = 256;
    const double gb = n * 16 / 1e9;
    printf("copy  (1r:1w): %.0f GB/s\n", 2 * gb / (best_ms([&] { k_copy<<<grid, block>>>(a, b, n); }) * 1e-3));
    printf("1r:3w (stencil mix): %.0f GB/s\n", 4 * gb / (best_ms([&] { k_1r3w<<<grid, block>>>(a, b, c, d, n); }) * 1e-3));
    printf("write only: %.0f GB/s\n", gb / (best_ms([&] { k_write<<<grid, block>>>(b, n); }) * 1e-3));
    printf("read only: %.0f GB/s\n", gb / (best_ms([&] { k_read<<<grid, block>>>(a, sink, n); }) * 1e-3));
    return 0;
}
