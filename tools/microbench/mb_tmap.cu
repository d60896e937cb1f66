// Which 2-D tensor maps over a column-major FP64 matrix does the TMA unit accept?  Development tool (round 2).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_tmap mb_tmap.cu && ./mb_tmap
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BI, int BK>
__global__ void load_box(const __grid_constant__ CUtensorMap tm, int i0, int k0, double *out) {
    extern __shared__ __align__(128) double sm[];
    uint64_t *bar = (uint64_t *)(sm + BI * BK);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(BI * BK * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                     "l"(&tm), "r"(i0), "r"(k0), "r"(s32(bar))
                     : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(bar)) : "memory");
    for (int e = threadIdx.x; e < BI * BK; e += blockDim.x) out[e] = sm[e];
}

typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BI, int BK>
static void test(enc_t enc, double *A, int64_t ld, int cols, int i0, int k0, double *out, double *hout) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)cols};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {BI, BK}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %3d x %2d  ld %lld  at (%d, %d): encode %d", BI, BK, (long long)ld, i0, k0, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return; }
    const int smem = BI * BK * 8 + 128;
    cudaFuncSetAttribute((const void *)load_box<BI, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    load_box<BI, BK><<<1, 128, smem>>>(tm, i0, k0, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  run: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(hout, out, (size_t)BI * BK * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int k = 0; k < BK; ++k)
        for (int i = 0; i < BI; ++i) {
            const bool in = (i0 + i) < ld && (k0 + k) < cols && (i0 + i) >= 0 && (k0 + k) >= 0;
            const double want = in ? (double)((int64_t)(k0 + k) * ld + i0 + i) : 0.0;
            if (hout[k * BI + i] != want) ++bad;
        }
    printf("  run ok, %d wrong of %d\n", bad, BI * BK);
}

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc_t enc = (enc_t)p;
    const int64_t ld = 4102;
    const int cols = 300;
    double *A, *out;
    cudaMalloc(&A, ld * cols * 8);
    cudaMalloc(&out, 256 * 64 * 8);
    double *h = (double *)malloc(ld * cols * 8), *hout = (double *)malloc(256 * 64 * 8);
    for (int64_t e = 0; e < ld * cols; ++e) h[e] = (double)e;
    cudaMemcpy(A, h, ld * cols * 8, cudaMemcpyHostToDevice);
    test<128, 16>(enc, A, ld, cols, 0, 0, out, hout);
    test<128, 16>(enc, A, ld, cols, 2, 5, out, hout);
    test<132, 16>(enc, A, ld, cols, 0, 0, out, hout);
    test<132, 16>(enc, A, ld, cols, 2, 5, out, hout);
    test<132, 16>(enc, A, ld, cols, 4000, 290, out, hout);     // runs past the last row and the last column: zero fill
    test<130, 16>(enc, A, ld, cols, 2, 5, out, hout);
    test<16, 16>(enc, A, ld, cols, 2, 5, out, hout);
    test<128, 16>(enc, A, ld, cols, 3, 5, out, hout);          // odd inner coordinate = box start not 16-byte aligned (LAST: the fault is sticky)
    return 0;
}
