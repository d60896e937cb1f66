// Would the RBF pair loop gain from computing s = |x - c|^2 on the tensor sub-pipe?  Development tool.
// Both kernels do, per lane and iteration, the work of 8 (proposal, center) pairs with the instruction mix of the chain kernel's
// pair loop: 7 dependent-ish FP64 operations + 1 MUFU.RSQ64H + 1 integer op per pair for the kernel value and the accumulation,
// and for s either 4 DFMA per pair (VEC) or 1 DMMA per 2 pairs (MMA: an m8n8k4 yields 64 values = 2 per lane).
// DMMA and DFMA share the FP64 datapath (mb_mix.cu), but a DFMA holds the issue port for 2 cycles per warp while a DMMA does not.
#include <stdio.h>
#include <cuda_runtime.h>

template <bool MMA>
__global__ void __launch_bounds__(448) pairs(double *out, int iters, double seed) {
    const double x0 = seed + threadIdx.x * 1e-3, x1 = seed - threadIdx.x * 2e-3, w = 1.0 + seed;
    double acc[4] = {0, 0, 0, 0};
    double c0 = 0.3 + seed, c1 = 0.7 - seed;
    int guard = 0;
    for (int it = 0; it < iters; ++it) {
        double s[8];
        if (MMA) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                s[2 * t] = x0; s[2 * t + 1] = x1;      // C preloaded with |x|^2
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(s[2 * t]), "+d"(s[2 * t + 1]) : "d"(x0 + t), "d"(c0));
            }
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const double d0 = c0 - (x0 + p), d1 = c1 - (x1 + p);
                s[p] = fma(d1, d1, fma(d0, d0, 1e-300));
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s[p]));
            const double t = s[p] * y;
            const double e = fma(-t, y, 1.0);
            const double q = e * fma(0.375, e, 0.5);
            const double a = s[p] * t;
            acc[p & 3] = fma(fma(a, q, a), w, acc[p & 3]);
            guard += __double2hiint(a) & 1;
        }
        c0 += 1e-3; c1 -= 1e-3;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3] + guard;
}

template <bool MMA>
static float run(double *out, int threads) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pairs<MMA><<<148, threads>>>(out, iters, 1e-7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pairs<MMA><<<148, threads>>>(out, iters, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double pairs_n = 148.0 * threads * iters * 8;
    printf("%s threads=%d: %.3f ms  %.3e pairs/s  %.2f cycles per 32 pairs per scheduler (1.965 GHz)  (%s)\n", MMA ? "s by DMMA" : "s by DFMA",
           threads, ms, pairs_n / (ms * 1e-3), ms * 1e-3 * 1.965e9 / (threads / 32.0 / 4.0 * iters * 8), cudaGetErrorString(cudaGetLastError()));
    return ms;
}

int main() {
    double *out;
    cudaMalloc(&out, (size_t)148 * 512 * 8);
    for (int threads : {256, 448}) {
        const float a = run<false>(out, threads), b = run<true>(out, threads);
        printf("  -> speed-up %.3f\n", a / b);
    }
    return 0;
}
