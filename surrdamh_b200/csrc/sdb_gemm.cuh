// Internal FP64 GEMM on the tensor pipe (mma.sync m8n8k4 f64 -> DMMA), generic strides.
//
//   C[M x N] (column-major, ldc)  =  beta * C  +  alpha * A[M x K] * B[K x N]
//   A(i,k) = A[i*sa_m + k*sa_k],  B(k,j) = B[k*sb_k + j*sb_n]      (any transposition via strides)
//
// Split-K: gridDim.z slabs of the K range; slab z writes alpha * (partial product) to
// C + z*c_slab (beta ignored) and a second kernel (sdb_gemm_reduce) sums the slabs in a fixed order
// -- deterministic, no atomics.  gridDim.z == 1 writes C directly.
//
// Used by: LU trailing updates / triangular solves (sdb_lu.cu), normal equations of the polynomial
// fit (sdb_fit.cu).  Tile 128 x 64 x 8, 256 threads = 8 warps (4 x 2), warp tile 32 x 32,
// two-stage shared-memory pipeline, register prefetch of the next k-slab.
#pragma once
#include <string.h>

#include <cuda.h>      // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include "sdb_common.cuh"

#define GM_BM 128
#define GM_BN 64
#define GM_BK 8
#define GM_PAD 8  // row stride == 8 (mod 16) doubles: the 4 k-rows of a fragment load hit disjoint bank halves

__device__ __forceinline__ void sdb_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct sdb_gemm_args {
    const double *A, *B;
    double *C;
    int64_t sa_m, sa_k, sb_k, sb_n, ldc, c_slab;
    int M, N, K;
    int k_per_slab;  // multiple of GM_BK
    double alpha, beta;
};

static __global__ void __launch_bounds__(256) sdb_gemm_kernel(const sdb_gemm_args g) {
    __shared__ double As[2][GM_BK][GM_BM + GM_PAD];
    __shared__ double Bs[2][GM_BK][GM_BN + GM_PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * GM_BN;
    const int kbeg = blockIdx.z * g.k_per_slab;
    const int kend = min(g.K, kbeg + g.k_per_slab);
    const int gid = lane >> 2, tig = lane & 3;

    // global -> register staging: A tile 128 x 8 = 1024 elements = 4 per thread, B tile 64 x 8 = 2 per thread
    const bool a_m_fast = g.sa_m <= g.sa_k;  // which index runs fastest over consecutive threads
    const bool b_n_fast = g.sb_n <= g.sb_k;
    double ra[4], rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            const int mm = a_m_fast ? (idx % GM_BM) : (idx / GM_BK);
            const int kk = a_m_fast ? (idx / GM_BM) : (idx % GM_BK);
            const int gm = m0 + mm, gk = k0 + kk;
            ra[e] = (gm < g.M && gk < kend) ? g.A[(int64_t)gm * g.sa_m + (int64_t)gk * g.sa_k] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid + e * 256;
            const int nn = b_n_fast ? (idx % GM_BN) : (idx / GM_BK);
            const int kk = b_n_fast ? (idx / GM_BN) : (idx % GM_BK);
            const int gn = n0 + nn, gk = k0 + kk;
            rb[e] = (gn < g.N && gk < kend) ? g.B[(int64_t)gk * g.sb_k + (int64_t)gn * g.sb_n] : 0.0;
        }
    };
    auto store_tiles = [&](int st) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            const int mm = a_m_fast ? (idx % GM_BM) : (idx / GM_BK);
            const int kk = a_m_fast ? (idx / GM_BM) : (idx % GM_BK);
            As[st][kk][mm] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid + e * 256;
            const int nn = b_n_fast ? (idx % GM_BN) : (idx / GM_BK);
            const int kk = b_n_fast ? (idx / GM_BN) : (idx % GM_BK);
            Bs[st][kk][nn] = rb[e];
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int st = 0;
    if (kbeg < kend) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
        const bool more = k0 + GM_BK < kend;
        if (more) load_tiles(k0 + GM_BK);
#pragma unroll
        for (int ks = 0; ks < GM_BK; ks += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) fa[i] = As[st][ks + tig][wm + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) fb[j] = Bs[st][ks + tig][wn + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sdb_dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
        if (more) store_tiles(st ^ 1);
        __syncthreads();
        st ^= 1;
    }
    // epilogue: accumulator (row gid, cols 2*tig, 2*tig+1) of each 8x8 tile
    double *Cz = g.C + (int64_t)blockIdx.z * g.c_slab;
    const bool direct = gridDim.z == 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wm + i * 8 + gid, gn = n0 + wn + j * 8 + tig * 2 + e;
                if (gm < g.M && gn < g.N) {
                    double *p = Cz + (int64_t)gn * g.ldc + gm;
                    const double v = g.alpha * acc[i][j][e];
                    *p = (direct && g.beta != 0.0) ? fma(g.beta, *p, v) : v;
                }
            }
}

// C = beta*C + sum_z slab_z   (fixed order z = 0, 1, ...)
static __global__ void __launch_bounds__(256) sdb_gemm_reduce(double *C, int64_t ldc, const double *slabs, int64_t c_slab,
                                                       int64_t ld_slab, int M, int N, int nslab, double beta) {
    const int64_t total = (int64_t)M * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % M), j = (int)(e / M);
        double s = 0.0;
        for (int z = 0; z < nslab; ++z) s += slabs[(int64_t)z * c_slab + (int64_t)j * ld_slab + i];
        double *p = C + (int64_t)j * ldc + i;
        *p = (beta != 0.0) ? fma(beta, *p, s) : s;
    }
}

// Host launcher.  `slab_ws` (doubles, >= nslab * M * N) is needed only when nslab > 1.
static inline int sdb_gemm(cudaStream_t st, int M, int N, int K, double alpha, const double *A, int64_t sa_m, int64_t sa_k,
                           const double *B, int64_t sb_k, int64_t sb_n, double beta, double *C, int64_t ldc, int nslab = 1,
                           double *slab_ws = nullptr) {
    if (M <= 0 || N <= 0) return SDB_OK;
    sdb_gemm_args g;
    g.A = A; g.B = B; g.sa_m = sa_m; g.sa_k = sa_k; g.sb_k = sb_k; g.sb_n = sb_n;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    if (nslab < 1) nslab = 1;
    int kps = ((K + nslab - 1) / nslab + GM_BK - 1) / GM_BK * GM_BK;
    if (kps < GM_BK) kps = GM_BK;
    nslab = (K + kps - 1) / kps;
    if (nslab < 1) nslab = 1;
    g.k_per_slab = kps;
    dim3 grid((M + GM_BM - 1) / GM_BM, (N + GM_BN - 1) / GM_BN, nslab);
    if (nslab == 1) {
        g.C = C; g.ldc = ldc; g.c_slab = 0;
        sdb_count_launch();
        sdb_gemm_kernel<<<grid, 256, 0, st>>>(g);
    } else {
        g.C = slab_ws; g.ldc = M; g.c_slab = (int64_t)M * N;
        sdb_count_launch();
        sdb_gemm_kernel<<<grid, 256, 0, st>>>(g);
        int64_t blocks = ((int64_t)M * N + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        sdb_count_launch();
        sdb_gemm_reduce<<<(unsigned)blocks, 256, 0, st>>>(C, ldc, slab_ws, (int64_t)M * N, M, M, N, nslab, beta);
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}


// ---------------------------------------------------------------------------------------------------
// Rank-K update (K = 64: the LU trailing update; K = 128: the paired Cholesky update):  C[M x N] -= A[M x K] * B[K x N],
// all column-major with a common leading dimension (A m-contiguous, B k-contiguous).  The K extent of a 128 x 64
// tile passes through shared memory by cp.async (8-byte pieces: lda may be odd, so 16-byte alignment is
// not guaranteed) in slices of 32, two resident, so the next slice loads while the current one is multiplied.
// ---------------------------------------------------------------------------------------------------
#define GK_K 64
#define GK_APAD 4   // As / Bt row stride == 4 (mod 16) doubles: the four k-rows of a fragment load (tig = 0..3) start 8 banks apart
#define GK_BPAD 4   // Bs is [n][k]: row stride 68 doubles == 4 (mod 16): the 8 n-rows of a fragment load spread over all banks

__device__ __forceinline__ void sdb_cp_async8(void *smem, const void *gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 8 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

// BT = false: B(k, j) = B[j*ld + k]  (k-contiguous: the U12 block of the LU);  shared layout Bs[n][k]
// BT = true : B(k, j) = B[k*ld + j]  (j-contiguous: B = L21^T of the Cholesky SYRK); shared layout Bs[k][n]
// LOWER: tiles strictly above the diagonal are skipped (symmetric rank-k update of the lower triangle)
// ASSIGN: C = A B instead of C -= A B (C may alias A when N <= 64: a CTA reads its whole A tile before it stores)
// NH: number of 32-wide K slices the kernel is unrolled for (2: K <= 64, 4: K <= 128, 8: K <= 256)
template <bool BT, bool LOWER, bool ASSIGN, int NH>
static __global__ void __launch_bounds__(256, 2) sdb_gemm_k64_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                                     double *__restrict__ C, int64_t ld, int64_t ldb, int M, int N,
                                                                     int K) {
    const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * GM_BN;
    if (LOWER && n0 > m0 + GM_BM - 1) return;
    extern __shared__ __align__(16) double gk_sm[];
    double (*As)[GM_BM + GK_APAD] = (double (*)[GM_BM + GK_APAD])gk_sm;                               // [64 k][136]
    double *Bsm = gk_sm + GK_K * (GM_BM + GK_APAD);
    double (*Bs)[GK_K + GK_BPAD] = (double (*)[GK_K + GK_BPAD])Bsm;                                   // [64 n][68]   (!BT)
    double (*Bt)[GM_BN + GK_APAD] = (double (*)[GM_BN + GK_APAD])Bsm;                                 // [64 k][72]   (BT)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int gid = lane >> 2, tig = lane & 3;
    // K is walked in slices of 32: slice h lives in rows [32 (h & 1), 32 (h & 1) + 32) of As / Bs, so two slices are
    // resident and slice h + 2 is fetched into the rows slice h has just left.  K <= 64 is the one-shot case (both
    // slices requested up front, no refill); a larger K (the rank-128 update of the Cholesky) re-uses the C tile in the
    // accumulators, which halves the read-modify-write traffic of C per flop.
    constexpr int H = NH;
    auto fetch = [&](int h) {
        const int kb = (h & 1) * 32, kg = h * 32;       // rows in shared memory / first k of the slice
        // A: 32 k-rows x 128 m  (4096 elements, 16 per thread), m fastest over threads
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int idx = tid + e * 256;
            const int mm = idx & (GM_BM - 1), kk = idx >> 7;
            const bool ok = (m0 + mm < M) && (kg + kk < K);
            sdb_cp_async8(&As[kb + kk][mm], A + (int64_t)(kg + kk) * ld + (m0 + mm), ok);
        }
        // B: 32 k x 64 n (2048 elements, 8 per thread), the memory-contiguous index fastest over threads
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * 256;
            if (!BT) {
                const int kk = idx & 31, nn = idx >> 5;
                const bool ok = (n0 + nn < N) && (kg + kk < K);
                sdb_cp_async8(&Bs[nn][kb + kk], B + (int64_t)(n0 + nn) * ldb + (kg + kk), ok);
            } else {
                const int nn = idx & 63, kk = idx >> 6;
                const bool ok = (n0 + nn < N) && (kg + kk < K);
                sdb_cp_async8(&Bt[kb + kk][nn], B + (int64_t)(kg + kk) * ldb + (n0 + nn), ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0);
    fetch(1);            // beyond K: zero fill (keeps the group count uniform)
    // C - A B is accumulated as (-A) B + C: the C tile is loaded into the accumulators NOW, while the cp.async copies
    // are in flight, so the epilogue is a plain store (no exposed read-modify-write latency at the end)
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wm + i * 8 + gid, gn = n0 + wn + j * 8 + tig * 2 + e;
                acc[i][j][e] = (!ASSIGN && gm < M && gn < N) ? C[(int64_t)gn * ld + gm] : 0.0;
            }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        // groups committed so far: min(h + 2, H); slice h is complete when at most the one after it is pending
        if (h + 1 < H) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int kb = (h & 1) * 32;
        if (h * 32 < K) {
#pragma unroll
        for (int ks = 0; ks < 32; ks += 4) {
            const int k = kb + ks;
            double fa[4], fb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) fa[i] = ASSIGN ? As[k + tig][wm + i * 8 + gid] : -As[k + tig][wm + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) fb[j] = BT ? Bt[k + tig][wn + j * 8 + gid] : Bs[wn + j * 8 + gid][k + tig];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sdb_dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
        }
        if (h + 2 < H) {
            __syncthreads();          // every warp has left slice h
            fetch(h + 2);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wm + i * 8 + gid, gn = n0 + wn + j * 8 + tig * 2 + e;
                if (gm < M && gn < N) {
                    double *p = C + (int64_t)gn * ld + gm;
                    *p = acc[i][j][e];
                }
            }
}

template <bool BT, bool LOWER, bool ASSIGN, int NH>
static inline int sdb_gemm_k64_launch_nh(cudaStream_t st, int M, int N, int K, const double *A, const double *B, double *C,
                                         int64_t ld, int64_t ldb) {
    const int smem = (GK_K * (GM_BM + GK_APAD) + GK_K * (GM_BN + GK_APAD)) * 8;
    static bool attr_set = false;
    if (!attr_set) {
        SDB_CUDA(cudaFuncSetAttribute((const void *)sdb_gemm_k64_kernel<BT, LOWER, ASSIGN, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid((M + GM_BM - 1) / GM_BM, (N + GM_BN - 1) / GM_BN);
    sdb_count_launch();
    sdb_gemm_k64_kernel<BT, LOWER, ASSIGN, NH><<<grid, 256, smem, st>>>(A, B, C, ld, ldb, M, N, K);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

template <bool BT, bool LOWER, bool ASSIGN>
static inline int sdb_gemm_k64_launch(cudaStream_t st, int M, int N, int K, const double *A, const double *B, double *C,
                                      int64_t ld, int64_t ldb) {
    if (M <= 0 || N <= 0 || K <= 0) return SDB_OK;
    if (K <= 64) return sdb_gemm_k64_launch_nh<BT, LOWER, ASSIGN, 2>(st, M, N, K, A, B, C, ld, ldb);
    if (K <= 128) return sdb_gemm_k64_launch_nh<BT, LOWER, ASSIGN, 4>(st, M, N, K, A, B, C, ld, ldb);
    if (K <= 256) return sdb_gemm_k64_launch_nh<BT, LOWER, ASSIGN, 8>(st, M, N, K, A, B, C, ld, ldb);
    sdb_set_error("sdb_gemm_k64: K = %d exceeds 256", K);
    return SDB_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------------
// The same products for a SKINNY result (N, K <= 64 and a tall M): the panel solve and the next-panel column update
// of the Cholesky sit on its critical path, and with 128 x 64 tiles a 4096 x 64 result keeps only 32 SMs busy for the
// ~4 us a tile's 2048 DMMAs take.  Here a CTA owns 32 rows (4 warps x 8 rows): the A fragments go straight from
// global memory into registers (a warp reads its 8 x 64 slice once), only B passes through shared memory, and each
// warp runs 8 independent accumulator chains (a dependent DMMA has ~150 cycles of latency).  In place (C == A) is
// safe: a warp reads exactly the rows it later writes.  B(k, j) = B[k*ldb + j].
// ---------------------------------------------------------------------------------------------------
#define GS_BS 68     // row stride of B in shared memory: == 4 (mod 16) doubles -> conflict-free fragment loads

template <bool ASSIGN>
static __global__ void __launch_bounds__(128) sdb_gemm_skinny64_kernel(const double *A, const double *__restrict__ B, double *C,
                                                                       int64_t ld, int64_t ldb, int M, int N, int K) {
    __shared__ __align__(16) double Bt[64][GS_BS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int row = blockIdx.x * 32 + warp * 8 + gid;
    const bool rok = row < M;
    double breg[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) {                           // 64 x 64 elements of B, 32 per thread, all loads in flight
        const int e = tid + u * 128, n = e & 63, k = e >> 6;
        breg[u] = (n < N && k < K) ? B[(int64_t)k * ldb + n] : 0.0;
    }
    double a[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int k = 4 * s + tig;
        const double v = (rok && k < K) ? A[(int64_t)k * ld + row] : 0.0;
        a[s] = ASSIGN ? v : -v;
    }
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = j * 8 + tig * 2 + e;
            acc[j][e] = (!ASSIGN && rok && n < N) ? C[(int64_t)n * ld + row] : 0.0;
        }
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        const int e = tid + u * 128;
        Bt[e >> 6][e & 63] = breg[u];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 16; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) sdb_dmma(acc[j][0], acc[j][1], a[s], Bt[4 * s + tig][j * 8 + gid]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = j * 8 + tig * 2 + e;
            if (rok && n < N) C[(int64_t)n * ld + row] = acc[j][e];
        }
}

template <bool ASSIGN>
static inline int sdb_gemm_skinny64_launch(cudaStream_t st, int M, int N, int K, const double *A, const double *B, double *C,
                                           int64_t ld, int64_t ldb) {
    if (M <= 0 || N <= 0 || K <= 0) return SDB_OK;
    sdb_count_launch();
    sdb_gemm_skinny64_kernel<ASSIGN><<<(M + 31) / 32, 128, 0, st>>>(A, B, C, ld, ldb, M, N, K);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Symmetric rank-K update of a LARGE trailing matrix (the paired Cholesky update, K = 128), a TMA-fed DMMA kernel:
//   lower(C[M x N]) -= A[M x K] A2[N x K]^T,   A(i, k) = A[k*ld + i],  A2(j, k) = A2[k*ld + j],  C column-major (ld).
// The 128 x 64 one-shot kernel above moves 0.16 B through L2 per flop at K = 128 and runs two short-lived CTAs per SM
// whose load -> multiply -> store phases overlap only with each other (ncu: DMMA sub-pipe active 31 %, top stall long
// scoreboard).  Here one CTA per SM owns a 128 x 128 tile (8 warps, warp tile 64 x 32: 32 independent DMMA accumulator
// chains per warp, 0.09 B per flop), the C tile is loaded straight into the accumulators while the first slices are in
// flight, and the operands stream through a 5-stage ring of 16-wide K slices:
//  * lane l of warp 0 issues ONE 1-D bulk copy (cp.async.bulk, completion by transaction count on the stage's `full`
//    mbarrier) per k-row of a slice -- 32 copies of 1040 bytes.  (A first version used 4096 eight-byte cp.async per slice:
//    their 64-bit address arithmetic and the full load/store queue cost 17 % of the stall samples, lg_throttle.)
//  * a stage is handed back through its `empty` mbarrier, one arrival per warp; only warp 0 waits on it.  There is no CTA-wide
//    barrier in the loop (with a __syncthreads per slice the kernel was 15 % slower: barrier stall 14 %).
//  * the fragments of k-step ks+4 are requested before the 32 DMMAs of k-step ks are issued (two register sets).
// Bulk copies need 16-byte aligned addresses and the saddle matrix has an odd leading dimension, so every row is fetched
// from the aligned address at or one element below its first element (130 elements: the tile row plus one on either side)
// and lands with that shift in shared memory; the shift of the row a lane reads, ((address of row 0) / 8 + tig * ld) & 1,
// is a per-lane constant.
// PRECONDITION (met by the one caller, the factor inside the Cholesky work buffer): one element before and two after each
// 128-element row segment of A / A2 are readable; rows beyond M / N may hold anything (they only reach unused accumulators).
// Measured: 24.8 TFLOP/s = 0.67 of the FP64 peak on a 15 872^2 update (profiles/r01_syrk_big_ncu.md).
// ---------------------------------------------------------------------------------------------------
#define SB_T 128
// slice width and ring depth (compile-time; -DSB_KS=32 -DSB_ST=3, the other split of the same shared memory, was 18 % slower
// at n = 32768: three stages are too shallow a ring; six stages of 16 were 3 % slower than five)
#ifndef SB_KS
#define SB_KS 16
#endif
#ifndef SB_ST
#define SB_ST 5
#endif
#define SB_LD (SB_T + 4)    // row stride == 4 (mod 16) doubles: the k-rows tig = 0..3 of a fragment load start 8 banks apart

// TM (round 2): the operands are described by ONE 2-D tensor map over the whole work buffer (inner dimension = a column of `ld`
// doubles, ld padded to an even number so that the column stride is a multiple of 16 bytes) and a ring stage is filled by TWO
// cp.async.bulk.tensor copies of a 132 x 16 box -- the tile's 128 rows plus the 4 columns of padding that keep the fragment
// loads conflict-free (box rows land 132 doubles apart, exactly the layout the row-by-row producer builds); rows / columns beyond the
// buffer are zero-filled by the TMA unit.
// (ia, ka) / (ib, kb): element coordinates of A(m0 = 0, k = 0) and A2(n0 = 0, k = 0) inside the mapped buffer.
template <bool LOWER, bool TM>
static __global__ void __launch_bounds__(256, 1) sdb_syrk_tma_kernel(const double *__restrict__ A, const double *__restrict__ A2,
                                                                     double *__restrict__ C, int64_t ld, int M, int N, int K,
                                                                     const __grid_constant__ CUtensorMap tmap, int ia, int ka, int ib,
                                                                     int kb) {
    constexpr int SLICE = 2 * SB_KS * SB_LD;                     // doubles per ring stage (A rows, then A2 rows)
    constexpr uint32_t ROW_BYTES = (SB_T + 2) * 8;
    const int m0 = blockIdx.x * SB_T, n0 = blockIdx.y * SB_T;
    if (LOWER && n0 > m0 + SB_T - 1) return;
    // the ring comes first in dynamic shared memory and the kernel has NO static shared memory, so every stage starts 128-byte
    // aligned (a tensor-map copy faults on a destination that is not: "illegal instruction"); the barriers follow the ring
    extern __shared__ __align__(128) double sb_sm[];
    uint64_t *full = (uint64_t *)(sb_sm + (size_t)SB_ST * SLICE), *empty = full + SB_ST;
    typedef double (*slice_t)[SB_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int gid = lane >> 2, tig = lane & 3;
    const int S = K / SB_KS;                                     // K is a multiple of 16 (checked by the launcher)
    const void *tmap_ptr = &tmap;                                // address of the __grid_constant__ parameter itself (param space)
    const double *Arow = A + m0, *Brow = A2 + n0;
    // TM: the box must START 16-byte aligned (an odd inner coordinate of an FP64 tensor faults: tools/microbench/mb_tmap.cu), so it
    // starts at the even coordinate at or one below the tile's first row and the data lands with that shift, like the rows of the
    // row-by-row producer; ld is even, so the shift is the same for every k-row.
    const int pa = TM ? (ia & 1) : (int)(((uintptr_t)Arow >> 3) & 1), pb = TM ? (ib & 1) : (int)(((uintptr_t)Brow >> 3) & 1),
              pl = TM ? 0 : (int)(ld & 1);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < SB_ST; ++i) { sdb_mbar_init(&full[i], 1); sdb_mbar_init(&empty[i], 8); }
        sdb_mbar_fence_init();
    }
    __syncthreads();
    // warp 0 is also the producer: one bulk copy per row of a stage,
    // row r = lane + 32 c of a stage (c = 0 .. SB_KS/16 - 1): rows 0 .. SB_KS-1 are k-rows of A, the rest k-rows of A2
    constexpr int PC = 2 * SB_KS / 32;
    const double *psrc[PC];
    double *pdst[PC];
#pragma unroll
    for (int c = 0; c < PC; ++c) {
        const int r = lane + 32 * c, pkk = r & (SB_KS - 1);
        const bool pisb = r >= SB_KS;
        psrc[c] = (pisb ? Brow : Arow) + (int64_t)pkk * ld - (((pisb ? pb : pa) + pkk * pl) & 1);   // s * SB_KS * ld is even
        pdst[c] = sb_sm + (size_t)r * SB_LD;
    }
    auto issue = [&](int s) {
        uint64_t *bar = &full[s % SB_ST];
        if (TM) {
            if (lane == 0) {
                double *dst = sb_sm + (size_t)(s % SB_ST) * SLICE;
                sdb_mbar_expect_tx(bar, 2 * SB_KS * SB_LD * 8);
                sdb_mbar_arrive(bar);
                sdb_tensor_g2s_2d(dst, tmap_ptr, ia + m0 - pa, ka + s * SB_KS, bar);
                sdb_tensor_g2s_2d(dst + SB_KS * SB_LD, tmap_ptr, ib + n0 - pb, kb + s * SB_KS, bar);
            }
            __syncwarp();
            return;
        }
        if (lane == 0) {
            sdb_mbar_expect_tx(bar, 2 * SB_KS * ROW_BYTES);
            sdb_mbar_arrive(bar);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < PC; ++c)
            sdb_bulk_g2s(pdst[c] + (size_t)(s % SB_ST) * SLICE, psrc[c] + (int64_t)s * SB_KS * ld, ROW_BYTES, bar);
    };
    if (warp == 0) {
        for (int s = 0; s < SB_ST - 1 && s < S; ++s) issue(s);
    }
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wm + i * 8 + gid, gn = n0 + wn + j * 8 + tig * 2 + e;
                acc[i][j][e] = (gm < M && gn < N) ? C[(int64_t)gn * ld + gm] : 0.0;
            }
    const int ca = wm + gid + ((pa + tig * pl) & 1), cb = wn + gid + ((pb + tig * pl) & 1);    // per-lane column offsets
    for (int s = 0; s < S; ++s) {
        if (warp == 0 && s + SB_ST - 1 < S) {
            // the stage slice s-1 lived in is refilled once all eight warps have left it: per-stage `empty` barriers, no
            // CTA-wide barrier in the loop (with __syncthreads here the kernel was 15 % slower: ncu barrier stall 14 %)
            if (s > 0) sdb_mbar_wait(&empty[(s - 1) % SB_ST], (uint32_t)(((s - 1) / SB_ST) & 1));
            issue(s + SB_ST - 1);
        }
        sdb_mbar_wait(&full[s % SB_ST], (uint32_t)((s / SB_ST) & 1));
        slice_t As = (slice_t)(sb_sm + (size_t)(s % SB_ST) * SLICE);
        slice_t Bs = As + SB_KS;
        // fragments of k-step ks+4 are requested before the 32 DMMAs of k-step ks are issued (two register sets)
        double fa[2][8], fb[2][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) fa[0][i] = As[tig][ca + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) fb[0][j] = Bs[tig][cb + j * 8];
#pragma unroll
        for (int ks = 0; ks < SB_KS; ks += 4) {
            const int cur = (ks >> 2) & 1, nxt = cur ^ 1;
            if (ks + 4 < SB_KS) {
#pragma unroll
                for (int i = 0; i < 8; ++i) fa[nxt][i] = As[ks + 4 + tig][ca + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) fb[nxt][j] = Bs[ks + 4 + tig][cb + j * 8];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sdb_dmma(acc[i][j][0], acc[i][j][1], -fa[cur][i], fb[cur][j]);
        }
        __syncwarp();
        if (lane == 0) sdb_mbar_arrive(&empty[s % SB_ST]);      // this warp's loads of slice s have all returned (their values were used)
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wm + i * 8 + gid, gn = n0 + wn + j * 8 + tig * 2 + e;
                if (gm < M && gn < N) C[(int64_t)gn * ld + gm] = acc[i][j][e];
            }
}

typedef CUresult (*sdb_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline sdb_tmap_encode_fn sdb_tmap_encoder() {
    static sdb_tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (sdb_tmap_encode_fn)p;
        cudaGetLastError();
    }
    return fn;
}

// base / base_cols: the column-major buffer (leading dimension ld) the operands live in, for the tensor map; nullptr = row-by-row producer
static inline int sdb_syrk_tma_launch(cudaStream_t st, int M, int N, int K, const double *A, const double *A2, double *C, int64_t ld,
                                      const double *base = nullptr, int64_t base_cols = 0) {
    const int smem = SB_ST * 2 * SB_KS * SB_LD * 8 + 128;      // ring + the full / empty barriers
    static bool attr_set = false;
    if (!attr_set) {
        SDB_CUDA(cudaFuncSetAttribute((const void *)sdb_syrk_tma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SDB_CUDA(cudaFuncSetAttribute((const void *)sdb_syrk_tma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid((M + SB_T - 1) / SB_T, (N + SB_T - 1) / SB_T);
    static const int use_tm = getenv("SDB_SYRK_TENSORMAP") ? atoi(getenv("SDB_SYRK_TENSORMAP")) : 1;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    sdb_tmap_encode_fn enc = sdb_tmap_encoder();
    if (use_tm && enc && base && (ld & 1) == 0 && (((uintptr_t)base) & 15) == 0 && A >= base && A2 >= base && ld < (1ll << 31)) {
        const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)base_cols};
        const cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
        const cuuint32_t box[2] = {SB_LD, SB_KS};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            const int64_t oa = A - base, ob = A2 - base;
            sdb_count_launch();
            sdb_syrk_tma_kernel<true, true><<<grid, 256, smem, st>>>(A, A2, C, ld, M, N, K, tm, (int)(oa % ld), (int)(oa / ld), (int)(ob % ld),
                                                                     (int)(ob / ld));
            SDB_CUDA(cudaGetLastError());
            return SDB_OK;
        }
    }
    sdb_count_launch();
    sdb_syrk_tma_kernel<true, false><<<grid, 256, smem, st>>>(A, A2, C, ld, M, N, K, tm, 0, 0, 0, 0);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// lower(C) -= A A2^T for a large trailing matrix, K a multiple of 16 (else, or with SDB_SYRK_TMA=0, the rank-K kernel above)
static inline int sdb_syrk_big_update(cudaStream_t st, int M, int N, int K, const double *A, const double *A2, double *C, int64_t ld,
                                      const double *base = nullptr, int64_t base_cols = 0) {
    if (M <= 0 || N <= 0 || K <= 0) return SDB_OK;
    static const int tma = getenv("SDB_SYRK_TMA") ? atoi(getenv("SDB_SYRK_TMA")) : 1;
    if (tma && K % SB_KS == 0) return sdb_syrk_tma_launch(st, M, N, K, A, A2, C, ld, base, base_cols);
    return sdb_gemm_k64_launch<true, true, false>(st, M, N, K, A, A2, C, ld, ld);
}

// C -= A B, B = U12 (k-contiguous): the LU trailing update
static inline int sdb_gemm_k64_update(cudaStream_t st, int M, int N, int K, const double *A, const double *B, double *C,
                                      int64_t ld) {
    return sdb_gemm_k64_launch<false, false, false>(st, M, N, K, A, B, C, ld, ld);
}
// lower(C) -= A A2^T with A2 given as rows (B(k, j) = A2[k*ld + j]): the Cholesky trailing update
static inline int sdb_syrk_k64_update(cudaStream_t st, int M, int N, int K, const double *A, const double *A2, double *C,
                                      int64_t ld) {
    return sdb_gemm_k64_launch<true, true, false>(st, M, N, K, A, A2, C, ld, ld);
}
// A (M x K, leading dimension ld) := A * Xt with Xt(k, j) = X[k*ldx + j], K, N <= 64: the triangular solve of the
// Cholesky panel as a multiplication by the inverse of the diagonal block, in place
static inline int sdb_gemm_k64_inplace(cudaStream_t st, int M, int K, double *A, int64_t ld, const double *X, int64_t ldx) {
    return sdb_gemm_skinny64_launch<true>(st, M, K, K, A, X, A, ld, ldx);
}
// C (M x N, N <= 64) -= A (M x K) B with B(k, j) = B[k*ldb + j], K <= 64: the next panel's columns of the Cholesky
static inline int sdb_gemm_k64_skinny_update(cudaStream_t st, int M, int N, int K, const double *A, const double *B, double *C,
                                             int64_t ld, int64_t ldb) {
    return sdb_gemm_skinny64_launch<false>(st, M, N, K, A, B, C, ld, ldb);
}
