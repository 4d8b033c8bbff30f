// FP64 pipe micro-benchmarks for B200 (sm_100a): what is the real ceiling of DMMA.8x8x4,
// does it depend on occupancy / chain count / PTX shape, and do DMMA and DFMA overlap?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH, bool DISTINCT>
__global__ void k_dmma884(double* sink, int iters) {
    double c[CH][2];
    double a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = c[i][1] = 0.0; a[i] = 1.0 + (threadIdx.x + i) * 1e-9; b[i] = 1.0 - (threadIdx.x + i) * 1e-9; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], DISTINCT ? a[i] : a[0], DISTINCT ? b[i] : b[0]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}
template <int CH>
__global__ void k_dmma1688(double* sink, int iters) {
    double c[CH][4];
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
    b[0] = 0.5; b[1] = 0.25 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma1688(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) sink[0] = s;
}
template <int CH>
__global__ void k_dmma16816(double* sink, int iters) {
    double c[CH][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 0.5 + (threadIdx.x + i) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma16816(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) sink[0] = s;
}
template <int CH>
__global__ void k_dfma(double* sink, int iters) {
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}
// hybrid: per warp, every DMMA is followed by NF independent DFMAs (same instruction stream)
template <int NF>
__global__ void k_hybrid_interleaved(double* sink, int iters) {
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = threadIdx.x * 1e-3 + i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    const double fa = 1.0000001, fb = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < NF; ++j) f[(i + j) & 7] = fma(f[(i + j) & 7], fa, fb);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) sink[0] = s;
}
// hybrid: warps with (warp % MOD) < NDMMA run DMMA chains, the others DFMA chains
__global__ void k_hybrid_warps(double* sink, int iters_mma, int iters_fma, int mod, int ndmma) {
    const int warp = threadIdx.x >> 5;
    if ((warp % mod) < ndmma) {
        double c[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
        double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
        for (int it = 0; it < iters_mma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
        if (s == 123.456) sink[0] = s;
    } else {
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
        const double a = 1.0000001, b = 1e-9;
        for (int it = 0; it < iters_fma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i];
        if (s == 123.456) sink[0] = s;
    }
}

static double* g_sink;
template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main() {
    cudaMalloc(&g_sink, 8);
    const int SM = 148;
    const int IT = 20000;
    printf("# kernel, blocks/SM, warps/block, chains, TFLOP/s\n");
    for (int wpb : {4, 8}) for (int bps : {1, 2, 4}) {
        const int blocks = SM * bps, thr = wpb * 32;
        double warps = (double)blocks * wpb;
#define RUN884(CH, D) { double ms = time_ms([&] { k_dmma884<CH, D><<<blocks, thr>>>(g_sink, IT); }); \
            printf("dmma884 distinct=%d, %d, %d, %d, %.2f\n", (int)D, bps, wpb, CH, warps * IT * CH * 512.0 / (ms * 1e-3) / 1e12); }
        RUN884(4, false) RUN884(8, false) RUN884(16, false) RUN884(26, false) RUN884(16, true)
#define RUN1688(CH) { double ms = time_ms([&] { k_dmma1688<CH><<<blocks, thr>>>(g_sink, IT); }); \
            printf("dmma1688, %d, %d, %d, %.2f\n", bps, wpb, CH, warps * IT * CH * 2048.0 / (ms * 1e-3) / 1e12); }
        RUN1688(4) RUN1688(8)
#define RUN16816(CH) { double ms = time_ms([&] { k_dmma16816<CH><<<blocks, thr>>>(g_sink, IT / 2); }); \
            printf("dmma16816, %d, %d, %d, %.2f\n", bps, wpb, CH, warps * (IT / 2) * CH * 4096.0 / (ms * 1e-3) / 1e12); }
        RUN16816(4) RUN16816(8)
        { double ms = time_ms([&] { k_dfma<16><<<blocks, thr>>>(g_sink, 2 * IT); });
          printf("dfma, %d, %d, 16, %.2f\n", bps, wpb, warps * 32 * 2.0 * IT * 16 * 2.0 / (ms * 1e-3) / 1e12); }
    }
    // interleaved hybrid: DMMA flops + DFMA flops
    {
        const int blocks = SM * 2, thr = 256; double warps = blocks * 8.0;
#define RUNHY(NF) { double ms = time_ms([&] { k_hybrid_interleaved<NF><<<blocks, thr>>>(g_sink, IT); }); \
            double mma = warps * IT * 8 * 512.0, fm = warps * 32.0 * IT * 8 * NF * 2.0; \
            printf("hybrid_interleaved NF=%d: total %.2f TFLOP/s (dmma part %.2f, dfma part %.2f)\n", NF, (mma + fm) / (ms * 1e-3) / 1e12, mma / (ms * 1e-3) / 1e12, fm / (ms * 1e-3) / 1e12); }
        RUNHY(0) RUNHY(1) RUNHY(2) RUNHY(4) RUNHY(8)
    }
    // warp-specialised hybrid: find balance by equalising run time roughly (iters tuned by rate guess)
    for (int nd : {1, 2, 3}) {
        const int mod = 4, blocks = SM * 2, thr = 512;  // 16 warps/block, 32 warps/SM -> 8 per SMSP
        // warps (w % 4) < nd do DMMA.  Choose iteration counts so both finish together if rates add up.
        for (double ratio : {0.1, 0.25, 0.5}) {
            int it_mma = IT, it_fma = (int)(IT * ratio * 8);
            double ms = time_ms([&] { k_hybrid_warps<<<blocks, thr>>>(g_sink, it_mma, it_fma, mod, nd); });
            double wd = blocks * 16.0 * nd / mod, wf = blocks * 16.0 * (mod - nd) / mod;
            double mma = wd * it_mma * 16 * 512.0, fm = wf * 32.0 * it_fma * 16 * 2.0;
            printf("hybrid_warps dmma_warps=%d/4 fma_iters_ratio=%.2f: total %.2f TFLOP/s (dmma %.2f + dfma %.2f) in %.2f ms\n", nd, ratio,
                   (mma + fm) / (ms * 1e-3) / 1e12, mma / (ms * 1e-3) / 1e12, fm / (ms * 1e-3) / 1e12, ms);
        }
    }
    cudaFree(g_sink);
    return 0;
}
