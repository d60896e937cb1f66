// Latency probes for the sequential factorisation kernels (development tool): dependent DFMA / DMUL chains, MUFU.RSQ64H,
// shared-memory round trip, __syncthreads, all timed with clock64 inside one warp / one CTA.
#include <stdio.h>
#include <cuda_runtime.h>

__global__ void probe(double *out, long long *cyc, double seed) {
    __shared__ double sm[256];
    const int t = threadIdx.x;
    double x = seed + t * 1e-9, y = 1.0000001;
    long long t0, t1;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    const long long c0 = clock64();
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (t == 0) cyc[0] = t1 - t0;
    // 2. dependent DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = x * y;
    t1 = clock64();
    if (t == 0) cyc[1] = t1 - t0;
    // 3. 4 independent DFMA chains
    double a = x, b = x + 1, c = x + 2, d = x + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { a = fma(a, y, 1e-9); b = fma(b, y, 1e-9); c = fma(c, y, 1e-9); d = fma(d, y, 1e-9); }
    t1 = clock64();
    if (t == 0) cyc[2] = t1 - t0;
    x = a + b + c + d;
    // 4. dependent rsqrt chain (MUFU.RSQ64H + one DMUL to rebuild a double)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        int hi = __double2hiint(x), r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
        (void)hi; (void)r;
        x = x + 1.5;
    }
    t1 = clock64();
    if (t == 0) cyc[3] = t1 - t0;
    // 5. smem store -> barrier -> load round trip, 64 times
    t0 = clock64();
    for (int i = 0; i < 64; ++i) {
        sm[t] = x;
        __syncthreads();
        x = sm[(t + 1) & 255] + 1.0;
    }
    t1 = clock64();
    if (t == 0) cyc[4] = t1 - t0;
    // 6. barrier only
    t0 = clock64();
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64();
    if (t == 0) cyc[5] = t1 - t0;
    // 7. 16 independent DFMA per iteration, dependent across iterations (the rank-1 update pattern)
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = x + k;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], y, 1e-9);
    }
    t1 = clock64();
    if (t == 0) cyc[6] = t1 - t0;
#pragma unroll
    for (int k = 0; k < 16; ++k) x += acc[k];
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (t == 0) { cyc[7] = (long long)((clock64() - c0) * 1000 / (long long)(g1 - g0 + 1)); }
    out[blockIdx.x * blockDim.x + t] = x;
}

__global__ void dmma_probe(double *out, long long *cyc, int nchain) {
    double c[8][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nchain)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[threadIdx.x] = s;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 256 * 8); cudaMalloc(&cyc, 8 * 8);
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 200; ++rep) probe<<<1, threads>>>(out, cyc, 1.25);
        long long h[8];
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%d: dep DFMA %.1f cyc  dep DMUL %.1f cyc  4-indep DFMA %.1f cyc/instr  rsqrt+add %.1f cyc  st/bar/ld %.1f cyc  bar %.1f cyc  16-indep DFMA block %.1f cyc/iter  clock %lld MHz (%s)\n",
               threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7], cudaGetErrorString(cudaGetLastError()));
    }
    for (int threads : {32, 128, 256, 512}) for (int nchain : {1, 8}) {
        dmma_probe<<<1, threads>>>(out, cyc, nchain);
        dmma_probe<<<1, threads>>>(out, cyc, nchain);
        long long h[1];
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA m8n8k4: threads=%d chains=%d: %.1f cycles per DMMA per warp (%.2f FMA/clk/SM)\n", threads, nchain, h[0] / (64.0 * nchain),
               256.0 * nchain * 64 * (threads / 32) / h[0]);
    }
    return 0;
}
