// Do DMMA (FP64 tensor sub-pipe) and DFMA (FP64 vector pipe) overlap on sm_100a?  Development tool.
// Every warp runs ITER iterations of ND independent m8n8k4 DMMAs + NF independent DFMAs; the grid fills every SM.
// If the two pipes are separate hardware, time(ND, NF) ~ max(time(ND, 0), time(0, NF)); if DMMA is executed by the
// vector FP64 units, time(ND, NF) ~ the sum.
#include <stdio.h>
#include <cuda_runtime.h>

template <int ND, int NF>
__global__ void __launch_bounds__(512) mix(double *out, int iters, double seed) {
    double c[ND ? ND : 1][2];
    double f[NF ? NF : 1];
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-3 - threadIdx.x * 1e-12, y = 0.999999 + seed;
#pragma unroll
    for (int i = 0; i < (ND ? ND : 1); ++i) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
    for (int i = 0; i < (NF ? NF : 1); ++i) f[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
        // interleave: one DMMA after every NF/ND DFMAs
#pragma unroll
        for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
            if (ND && i % ((NF > ND ? NF : ND) / ND) == 0 && i / ((NF > ND ? NF : ND) / ND) < ND) {
                const int k = i / ((NF > ND ? NF : ND) / ND);
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
            }
            if (i < NF) f[i] = fma(f[i], y, 1e-9);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (ND ? ND : 1); ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NF ? NF : 1); ++i) s += f[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ND, int NF>
static void run(double *out, int threads, int ctas_per_sm) {
    const int iters = 4096, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<ND, NF><<<grid, threads>>>(out, iters, 1e-7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mix<ND, NF><<<grid, threads>>>(out, iters, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)grid * threads / 32;
    const double dmma_fma = warps * iters * ND * 256.0, dfma_fma = warps * iters * NF * 32.0;
    printf("ND=%2d NF=%2d threads=%d ctas/SM=%d: %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  sum %.2f  (%s)\n", ND, NF, threads,
           ctas_per_sm, ms, 2e-9 * dmma_fma / ms, 2e-9 * dfma_fma / ms, 2e-9 * (dmma_fma + dfma_fma) / ms,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *out;
    cudaMalloc(&out, (size_t)148 * 4 * 512 * 8);
    for (int threads : {256, 512}) {
        run<8, 0>(out, threads, 1);
        run<0, 16>(out, threads, 1);
        run<0, 56>(out, threads, 1);
        run<4, 56>(out, threads, 1);
        run<8, 56>(out, threads, 1);
        run<4, 28>(out, threads, 1);
        run<8, 16>(out, threads, 1);
        run<8, 8>(out, threads, 1);
        run<2, 56>(out, threads, 1);
    }
    run<4, 56>(out, 256, 2);
    run<8, 56>(out, 256, 2);
    return 0;
}
