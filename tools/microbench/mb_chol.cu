// Micro-benchmark of the null-space Cholesky building blocks (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Isurrdamh_b200/csrc tools/microbench/mb_chol.cu \
//        surrdamh_b200/csrc/sdb_util.cu surrdamh_b200/csrc/sdb_fit.cu -o build/mb_chol
#include <stdio.h>
#include <functional>
#include <vector>
#define SDB_PI_PROF_ARG , long long *prof
#define SDB_PI_PROF_PASS , nullptr
#define SDB_PI_PROF_INIT long long pt[5] = {0, 0, 0, 0, 0}, plast = clock64();
#define SDB_PI_STAMP(i) { const long long now = clock64(); pt[i] += now - plast; plast = now; }
#define SDB_PI_PROF_END if (prof && (threadIdx.x & 31) == 0) for (int i = 0; i < 5; ++i) prof[(threadIdx.x >> 5) * 5 + i] = pt[i];
#include "../../surrdamh_b200/csrc/sdb_chol.cu"

static float time_loop(cudaStream_t st, int reps, const std::function<void(int)> &f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(61);
    cudaStreamSynchronize(st);
    cudaEventRecord(a, st);
    for (int i = 0; i < reps; ++i) f(i);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1000.f / reps;
}

__global__ void fill_spd(double *A, int64_t ld, int n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)n * n; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % n), j = (int)(e / n);
        const int di = i > j ? i - j : j - i;
        A[(int64_t)j * ld + i] = (i == j) ? 200.0 : 1.0 / (1.0 + di);
    }
}

int main() {
    const int n = 4096 + 64;
    const int64_t ld = n + 3;
    double *A, *xinv;
    int *info;
    long long *prof;
    cudaMalloc(&prof, 8 * 64);
    cudaMalloc(&A, sizeof(double) * ld * n);
    cudaMalloc(&xinv, sizeof(double) * 4096 * 80);
    cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    cudaStream_t st;
    cudaStreamCreate(&st);
    const int reps = 60;
    for (int upd = 0; upd < 2; ++upd) {
        fill_spd<<<1024, 256, 0, st>>>(A, ld, n);
        float us = time_loop(st, reps, [&](int i) { chol_potrf_inv64_kernel<<<1, PI_T, 0, st>>>(A, ld, 64 + 64 * (i % 60), 64, xinv + 4096 * (i % 60), info, upd, prof); });
        printf("potrf_inv64 upd=%d: %.2f us per launch (back-to-back, includes launch gap)\n", upd, us);
        long long hp[40];
        cudaMemcpy(hp, prof, 8 * 40, cudaMemcpyDeviceToHost);
        for (int w = 0; w < 8; ++w)
            printf("   warp %d cycles: prologue %lld | loop-back %lld  wait-panel %lld  priority+extract %lld  rest+stores %lld  (sums over 64 columns)\n", w, hp[w * 5], hp[w * 5 + 1],
                   hp[w * 5 + 2], hp[w * 5 + 3], hp[w * 5 + 4]);
    }
    {
        float us = time_loop(st, reps, [&](int i) { fill_spd<<<1, 32, 0, st>>>(A, ld, 1); });
        printf("empty-ish kernel back-to-back: %.2f us\n", us);
    }
    for (int M : {4096, 2048, 512, 64}) {
        float us = time_loop(st, reps, [&](int i) { sdb_gemm_k64_inplace(st, M, 64, A + 64, ld, xinv, 64); });
        printf("trsm-as-gemm M=%d: %.2f us\n", M, us);
        us = time_loop(st, reps, [&](int i) { sdb_gemm_k64_skinny_update(st, M, 64, 64, A + 64, A + 64, A + 64 * ld + 64, ld, ld); });
        printf("next-columns gemm M=%d: %.2f us\n", M, us);
        us = time_loop(st, reps, [&](int i) { sdb_syrk_k64_update(st, M, M, 64, A + 64, A + 64, A + 64 * ld + 64, ld); });
        printf("syrk M=%d: %.2f us  (%.2f TF/s)\n", M, us, (double)M * M * 64 / (us * 1e-6) / 1e12);
    }
    int hinfo;
    cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    printf("info %d  err %s\n", hinfo, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
