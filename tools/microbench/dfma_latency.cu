// DFMA issue / latency picture of one B200 SM sub-partition: throughput of CH independent DFMA chains per warp with
// W warps per sub-partition, operands shared between the chains or distinct registers per chain (register-bank and
// operand-collector effects).  Peak = 16 lanes per cycle and sub-partition = one warp-wide DFMA every 2 cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cuda_runtime.h>
#include <stdio.h>

template <int CH, bool DISTINCT>
__global__ void k_dfma(double* sink, int iters, unsigned long long* cyc) {
    double c[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        c[i] = threadIdx.x * 1e-3 + i;
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1e-9 * (i + 1);
    }
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) c[i] = fma(c[i], DISTINCT ? a[i] : a[0], DISTINCT ? b[i] : b[0]);
    }
    const unsigned long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH, bool DISTINCT>
static void run(int warps_per_smsp, double* sink, unsigned long long* cyc) {
    const int iters = 4096;
    k_dfma<CH, DISTINCT><<<148, 128 * warps_per_smsp>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    k_dfma<CH, DISTINCT><<<148, 128 * warps_per_smsp>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_instr = (double)h / ((double)iters * 8 * CH);          // cycles between two DFMAs of one warp
    const double pipe = 2.0 * warps_per_smsp / per_warp_instr;                    // fraction of the 16-lane pipe in use
    printf("chains %2d  %s  warps/SMSP %d : %.2f cycles per DFMA of a warp, pipe %.0f %%\n", CH, DISTINCT ? "distinct operands" : "shared operands  ",
           warps_per_smsp, per_warp_instr, 100.0 * pipe);
}

int main() {
    double* sink;
    unsigned long long* cyc;
    cudaMalloc(&sink, 8);
    cudaMalloc(&cyc, 8);
    for (int w = 1; w <= 4; ++w) {
        run<1, false>(w, sink, cyc);
        run<2, false>(w, sink, cyc);
        run<4, false>(w, sink, cyc);
        run<8, false>(w, sink, cyc);
        run<16, false>(w, sink, cyc);
        run<4, true>(w, sink, cyc);
        run<8, true>(w, sink, cyc);
        run<16, true>(w, sink, cyc);
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
