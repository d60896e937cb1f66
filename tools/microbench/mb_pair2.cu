// What does each part of the RBF pair loop cost?  Development tool (round 2).  148 CTAs x 512 threads (4 warps per sub-partition, the chain
// kernel's residency), every lane runs the 8-point pass of sdb_rbf_partialn_kt against centers in shared memory; variants drop or alter one
// ingredient at a time.  Prints cycles per pair and sub-partition (the FP64 pipe holds an instruction for 2 cycles: 11 instructions = 22).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_pair2 mb_pair2.cu && ./mb_pair2
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(512, 1) pass8(double *out, int iters, int n) {
    extern __shared__ double sm[];
    double *cs = sm, *ws = sm + 2 * n;
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) cs[i] = 5.0 + 1e-3 * i;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ws[i] = 1.0 + 1e-6 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double xs[8][2], acc[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) { xs[p][0] = 1.0 + 0.1 * p + 1e-3 * lane; xs[p][1] = 2.0 - 0.1 * p; acc[p] = 0.0; }
    const double wc = 1.000001;
    for (int it = 0; it < iters; ++it) {
        const double *c = cs + 2 * lane, *w = ws + lane;
        for (int i = lane; i < n; i += 32, c += 64, w += 32) {
            const double c0 = c[0], c1 = c[1], wk = (V == 2 || V == 3) ? wc : w[0];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                double r2 = 1e-300;
                const double d0 = c0 - xs[p][0], d1 = c1 - xs[p][1];
                r2 = fma(d0, d0, r2);
                r2 = fma(d1, d1, r2);
                if (V == 4) { acc[p] = fma(r2, wk, acc[p]); continue; }
                double y;
                if (V == 1 || V == 3) y = r2 * 0.03125;
                else asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
                const double t = r2 * y;
                const double e = fma(-t, y, 1.0);
                if (V == 5) { const double a = r2 * t; acc[p] = fma(fma(a, 0.5 * e, a), wk, acc[p]); continue; }   // second-order: 2 FP64 fewer
                if (V == 6) {   // the e^2 term of the correction in FP32 (two conversions + two FP32 operations), 10 FP64
                    const float ef = (float)e;
                    const double q2 = (double)(0.375f * ef * ef);
                    const double qq = fma(0.5, e, q2);
                    const double a = r2 * t;
                    acc[p] = fma(fma(a, qq, a), wk, acc[p]);
                    continue;
                }
                if (V == 7) {   // the same with the e^2 term from the HIGH WORD of e through integer / FP32 operations only (no F2F)
                    const int hi = __double2hiint(e);                       // sign, 11-bit exponent, 20 mantissa bits
                    // float with the same value as the truncated e: exponent re-biased (1023 -> 127), mantissa shifted by 3
                    const int fb = ((hi & 0x7fffffff) >> 3 << 0) - ((1023 - 127) << 20 >> 0 << 0) * 0;   // placeholder, replaced below
                    (void)fb;
                    const unsigned mag = (unsigned)(hi & 0x7fffffff);
                    const float ef = __uint_as_float(mag > (896u << 20) ? ((mag - (896u << 20)) << 3) : 0u);   // |e| as a float (sign irrelevant for e^2)
                    const float q2f = 0.375f * ef * ef;
                    const unsigned qb = __float_as_uint(q2f);
                    const double q2 = __hiloint2double((int)(qb ? ((qb >> 3) + (896u << 20)) : 0u), (int)(qb << 29));
                    const double qq = fma(0.5, e, q2);
                    const double a = r2 * t;
                    acc[p] = fma(fma(a, qq, a), wk, acc[p]);
                    continue;
                }
                const double q = e * fma(0.375, e, 0.5);
                const double a = r2 * t;
                acc[p] = fma(fma(a, q, a), wk, acc[p]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) s += acc[p];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
static void run(const char *what, int fp64_per_pair, double *out) {
    const int iters = 40, n = 4096, smem = 3 * n * 8;
    cudaFuncSetAttribute((const void *)pass8<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pass8<V><<<148, 512, smem>>>(out, iters, n);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pass8<V><<<148, 512, smem>>>(out, iters, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // pairs per warp: iters * (n / 32) * 8; 4 warps per sub-partition
    const double pairs_per_smsp = 4.0 * iters * (n / 32) * 8;
    const double cyc = ms * 1e-3 * 1.965e9 / pairs_per_smsp;
    printf("%-64s %7.3f ms  %6.2f cycles/pair/sub-partition  (%d FP64 -> %d pipe cycles, %.0f %%)  %s\n", what, ms, cyc, fp64_per_pair,
           2 * fp64_per_pair, 100.0 * 2 * fp64_per_pair / cyc, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *out;
    cudaMalloc(&out, 148 * 512 * 8);
    run<0>("V0 the kernel's pass (11 FP64 + MUFU + zeroing move + loads)", 11, out);
    run<1>("V1 seed by a multiply instead of MUFU.RSQ64H (12 FP64, no MUFU)", 12, out);
    run<2>("V2 weight a constant (the accumulating FMA reads 2 registers)", 11, out);
    run<3>("V3 both", 12, out);
    run<4>("V4 distance + accumulate only (5 FP64)", 5, out);
    run<5>("V5 second-order correction (10 FP64 + MUFU)", 10, out);
    run<6>("V6 e^2 term in FP32 via two F2F conversions (10 FP64 + MUFU)", 10, out);
    run<7>("V7 e^2 term in FP32 via integer re-biasing, no F2F (10 FP64)", 10, out);
    return 0;
}
