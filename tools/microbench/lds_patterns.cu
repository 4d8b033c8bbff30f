// LDS.128 / LDS.64 cost per lane->address pattern on B200: the shared-memory data pipe retires one "wavefront" per
// cycle and SM; how many wavefronts a warp-wide load costs depends on how its 32 addresses group.  Every pattern
// below is bank-conflict free in the classical sense (distinct addresses fall into distinct 16-byte bank groups,
// equal addresses are broadcasts).  Output: cycles per warp-load at saturation (16 warps per SM, loads independent).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ int slot_of(int pattern, int lane) {
    switch (pattern) {
        case 0: return 0;                                   // uniform
        case 1: return lane >> 3;                           // 4 distinct: uniform inside each quarter-warp
        case 2: return lane & 7;                            // 8 distinct, the same 8 in every quarter
        case 3: return lane;                                // 32 distinct
        case 4: return (lane >> 3) * 2 + (lane & 1);        // 8 distinct: 2 per quarter
        case 5: return (lane >> 1) & 3;                     // 4 distinct, the same 4 in every quarter
        case 6: return lane >> 1;                           // 16 distinct: 4 per quarter
        case 7: return lane >> 2;                           // 8 distinct: 2 per quarter, in blocks of 4 lanes
        case 8: return lane >> 4;                           // 2 distinct: uniform per half-warp
        case 9: return lane & 1;                            // 2 distinct, alternating
        case 10: return lane & 3;                           // 4 distinct, the same in every group of 4
        case 11: return (lane & 1) + 2 * (lane >> 4);       // 4 distinct: 2 per half-warp
        case 12: return lane & 15;                          // 16 distinct, repeated in both half-warps
        case 13: return (lane >> 3) + 4 * (lane & 1);       // 8 distinct: 2 per quarter, far apart
        default: return lane;
    }
}

template <int BYTES>
__global__ void __launch_bounds__(512) lds_kernel(int pattern, int iters, unsigned long long* cycles, double* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = (double)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + slot_of(pattern, lane) * 16;
    unsigned long long acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const uint32_t a = base + ((u * 512 + (it & 3) * 8192) & 16383);  // another 512-byte row each time, same bank picture
            if (BYTES == 16) {
                unsigned long long x, y;
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
                if (u & 1) { acc0 ^= x; acc1 ^= y; } else { acc2 ^= x; acc3 ^= y; }
            } else {
                unsigned long long x;
                asm volatile("ld.shared.u64 %0, [%1];" : "=l"(x) : "r"(a));
                if (u & 1) acc0 ^= x; else acc2 ^= x;
            }
        }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if ((acc0 ^ acc1 ^ acc2 ^ acc3) == 0x123456789abcdefULL) sink[0] = (double)acc0;
}

int main() {
    unsigned long long* cyc;
    double* sink;
    cudaMalloc(&cyc, 148 * sizeof(unsigned long long));
    cudaMalloc(&sink, 8);
    const int iters = 2048, warps = 16;
    printf("# bytes pattern cycles_per_warp_load (SM data pipe, %d warps resident; the loads are consumed by integer XORs)\n", warps);
    for (int bytes : {16, 8})
        for (int p = 0; p <= 13; ++p) {
            std::vector<unsigned long long> h(148);
            for (int rep = 0; rep < 2; ++rep) {
                if (bytes == 16)
                    lds_kernel<16><<<148, warps * 32, 16384>>>(p, iters, cyc, sink);
                else
                    lds_kernel<8><<<148, warps * 32, 16384>>>(p, iters, cyc, sink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h.data(), cyc, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            double mean = 0;
            for (auto v : h) mean += (double)v;
            mean /= 148.0;
            printf("%2d %2d %.3f\n", bytes, p, mean / ((double)iters * 16 * warps));
        }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
