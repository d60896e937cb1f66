// Dense LU with partial pivoting + triangular solves, FP64, on the device (sm_100a).
// Replaces np.linalg.solve (LAPACK dgesv) at surrogate_rbf.py:129-130 (solver_type='direct').
//
// Column-major N x N matrix A (leading dimension lda), factored in place as P A = L U with the
// LAPACK conventions (unit lower L, pivot row indices piv[k] >= k applied in order).
//
//   * blocked right-looking outer loop, panel width LU_NB
//   * inside a panel: recursive halving down to 8 columns; the 8-column base case keeps the whole
//     (N - j) x 8 sub-panel in REGISTERS, spread over one thread-block cluster of 8 CTAs (8 SMs);
//     per column one block-level arg-max, one exchange of the candidate rows through distributed
//     shared memory and ONE cluster barrier
//   * everything else (row interchanges, triangular solves with the 8..64-wide diagonal blocks,
//     rank-k updates) runs on the whole GPU; the updates use the DMMA GEMM of sdb_gemm.cuh
//   * solve phase: warp-per-right-hand-side triangular solves on 64 x 64 diagonal blocks + GEMM
#include <cooperative_groups.h>
#include <limits.h>

#include "sdb_gemm.cuh"

namespace cg = cooperative_groups;

#define LU_CL 8    // CTAs per cluster (portable maximum)
#define LU_T 512   // threads per CTA of the base case
#define LU_W 8     // base-case width
#define LU_NB 64   // outer panel width

struct lu_cand {
    double val;
    int row;
    int pad;
    double vals[LU_W];
};

// ---------------------------------------------------------------------------------------------------
// Base case: factor the (N - j0) x w sub-panel (w <= 8) starting at A(j0, j0); row interchanges are
// applied to these w columns only.  Row j0 + slot + q * 4096 lives in registers a[q][*] of thread `slot`.
// ---------------------------------------------------------------------------------------------------
template <int R>
__global__ void __cluster_dims__(LU_CL, 1, 1) __launch_bounds__(LU_T, 1)
    lu_base_kernel(double *__restrict__ A, int64_t lda, int N, int j0, int w, int *__restrict__ piv, int *__restrict__ info) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ lu_cand cand[2][LU_CL];
    __shared__ double diag[2][LU_W];
    __shared__ double red_val[LU_T / 32];
    __shared__ int red_row[LU_T / 32];
    __shared__ double win_val;
    __shared__ int win_row;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nslots = LU_CL * LU_T;
    const int slot = rank * LU_T + tid;

    double a[R][LU_W];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int row = j0 + slot + q * nslots;
#pragma unroll
        for (int c = 0; c < LU_W; ++c) a[q][c] = (row < N && c < w) ? A[(int64_t)(j0 + c) * lda + row] : 0.0;
    }

#pragma unroll
    for (int c = 0; c < LU_W; ++c) {
        if (c < w) {  // uniform over the whole cluster
            const int buf = c & 1;
            const int drow = j0 + c;
            // ---- arg-max of |a(:, c)| over my rows >= drow (first maximum wins, as idamax) ------------
            double best = -1.0;
            int brow = INT_MAX;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = j0 + slot + q * nslots;
                const double v = fabs(a[q][c]);
                if (row < N && row >= drow && v > best) { best = v; brow = row; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
                if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
            }
            if (lane == 0) { red_val[warp] = best; red_row[warp] = brow; }
            __syncthreads();
            if (warp == 0) {
                double v = lane < LU_T / 32 ? red_val[lane] : -1.0;
                int r = lane < LU_T / 32 ? red_row[lane] : INT_MAX;
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
                    const int orow = __shfl_xor_sync(0xffffffffu, r, off);
                    if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
                }
                if (lane == 0) { win_val = v; win_row = r; }
            }
            __syncthreads();
            // ---- publish this CTA's candidate row (and the diagonal row) to every CTA ----------------------
            const int wrow = win_row;
            if (wrow == INT_MAX) {
                if (tid == 0)
                    for (int r = 0; r < LU_CL; ++r) {
                        lu_cand *dst = cluster.map_shared_rank(&cand[buf][rank], r);
                        dst->val = -1.0;
                        dst->row = INT_MAX;
                    }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = j0 + slot + q * nslots;
                if (row == wrow) {
                    for (int r = 0; r < LU_CL; ++r) {
                        lu_cand *dst = cluster.map_shared_rank(&cand[buf][rank], r);
                        dst->val = win_val;
                        dst->row = row;
#pragma unroll
                        for (int k = 0; k < LU_W; ++k) dst->vals[k] = a[q][k];
                    }
                }
                if (row == drow) {
                    for (int r = 0; r < LU_CL; ++r) {
                        double *dst = cluster.map_shared_rank(&diag[buf][0], r);
#pragma unroll
                        for (int k = 0; k < LU_W; ++k) dst[k] = a[q][k];
                    }
                }
            }
            cluster.sync();
            // ---- every thread picks the global winner ------------------------------------------------------
            double pv = -1.0;
            int prow = INT_MAX, pr = 0;
#pragma unroll
            for (int r = 0; r < LU_CL; ++r) {
                const double v = cand[buf][r].val;
                const int row = cand[buf][r].row;
                if (v > pv || (v == pv && row < prow)) { pv = v; prow = row; pr = r; }
            }
            double pvals[LU_W], dvals[LU_W];
#pragma unroll
            for (int k = 0; k < LU_W; ++k) { pvals[k] = cand[buf][pr].vals[k]; dvals[k] = diag[buf][k]; }
            const bool singular = !(pv > 0.0);  // exactly zero (or NaN) column: LAPACK records info and moves on
            if (singular) prow = drow;
            if (slot == 0) {
                piv[drow] = prow;
                if (singular && *info == 0) *info = drow + 1;
            }
            if (!singular) {
                if (prow != drow) {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int row = j0 + slot + q * nslots;
                        if (row == prow) {
#pragma unroll
                            for (int k = 0; k < LU_W; ++k) a[q][k] = dvals[k];
                        } else if (row == drow) {
#pragma unroll
                            for (int k = 0; k < LU_W; ++k) a[q][k] = pvals[k];
                        }
                    }
                }
                const double rcp = 1.0 / pvals[c];  // dgetf2 scales by the reciprocal
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int row = j0 + slot + q * nslots;
                    if (row > drow && row < N) {
                        const double l = a[q][c] * rcp;
                        a[q][c] = l;
#pragma unroll
                        for (int c2 = c + 1; c2 < LU_W; ++c2) a[q][c2] = fma(-l, pvals[c2], a[q][c2]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int row = j0 + slot + q * nslots;
#pragma unroll
        for (int c = 0; c < LU_W; ++c)
            if (row < N && c < w) A[(int64_t)(j0 + c) * lda + row] = a[q][c];
    }
}

// ---------------------------------------------------------------------------------------------------
// Row interchanges k = k0..k1-1 (row k <-> piv[k], in order) applied to columns [c0, c1).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lu_laswp_kernel(double *__restrict__ A, int64_t lda, int c0, int c1,
                                                       const int *__restrict__ piv, int k0, int k1) {
    const int col = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= c1) return;
    double *a = A + (int64_t)col * lda;
    for (int k = k0; k < k1; ++k) {
        const int p = piv[k];
        if (p != k) {
            const double t = a[k];
            a[k] = a[p];
            a[p] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// B := L^{-1} B with L = unit lower triangle of the h x h block at Lp (h <= W), B is h x ncols,
// one thread per column, the column in registers, L broadcast from shared memory.
// ---------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(128) lu_trsm_lunit_kernel(const double *__restrict__ Lp, int64_t ldl, double *__restrict__ B,
                                                            int64_t ldb, int h, int ncols) {
    __shared__ double Ls[W][W + 1];
    for (int e = threadIdx.x; e < W * W; e += blockDim.x) {
        const int i = e % W, k = e / W;
        Ls[i][k] = (i < h && k < i) ? Lp[(int64_t)k * ldl + i] : 0.0;
    }
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    double *bp = B + (int64_t)col * ldb;
    double b[W];
#pragma unroll
    for (int i = 0; i < W; ++i) b[i] = i < h ? bp[i] : 0.0;
#pragma unroll
    for (int i = 1; i < W; ++i) {
        double t = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) t = fma(-Ls[i][k], b[k], t);
        b[i] = t;
    }
#pragma unroll
    for (int i = 0; i < W; ++i)
        if (i < h) bp[i] = b[i];
}

static int lu_trsm_lunit(cudaStream_t st, const double *Lp, int64_t ldl, double *B, int64_t ldb, int h, int ncols) {
    if (ncols <= 0 || h <= 0) return SDB_OK;
    const int blocks = (ncols + 127) / 128;
    if (h <= 8) { sdb_count_launch(); lu_trsm_lunit_kernel<8><<<blocks, 128, 0, st>>>(Lp, ldl, B, ldb, h, ncols); }
    else if (h <= 16) { sdb_count_launch(); lu_trsm_lunit_kernel<16><<<blocks, 128, 0, st>>>(Lp, ldl, B, ldb, h, ncols); }
    else if (h <= 32) { sdb_count_launch(); lu_trsm_lunit_kernel<32><<<blocks, 128, 0, st>>>(Lp, ldl, B, ldb, h, ncols); }
    else { sdb_count_launch(); lu_trsm_lunit_kernel<64><<<blocks, 128, 0, st>>>(Lp, ldl, B, ldb, h, ncols); }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

static int lu_base(cudaStream_t st, double *A, int64_t lda, int N, int j0, int w, int *piv, int *info) {
    const int rows = N - j0;
    const int per = (rows + LU_CL * LU_T - 1) / (LU_CL * LU_T);
    if (per <= 1) { sdb_count_launch(); lu_base_kernel<1><<<LU_CL, LU_T, 0, st>>>(A, lda, N, j0, w, piv, info); }
    else if (per <= 2) { sdb_count_launch(); lu_base_kernel<2><<<LU_CL, LU_T, 0, st>>>(A, lda, N, j0, w, piv, info); }
    else if (per <= 4) { sdb_count_launch(); lu_base_kernel<4><<<LU_CL, LU_T, 0, st>>>(A, lda, N, j0, w, piv, info); }
    else if (per <= 8) { sdb_count_launch(); lu_base_kernel<8><<<LU_CL, LU_T, 0, st>>>(A, lda, N, j0, w, piv, info); }
    else if (per <= 12) { sdb_count_launch(); lu_base_kernel<12><<<LU_CL, LU_T, 0, st>>>(A, lda, N, j0, w, piv, info); }
    else {
        sdb_set_error("sdb_lu: %d rows exceed the compiled limit of %d", rows, 12 * LU_CL * LU_T);
        return SDB_ERR_LIMIT;
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// Recursive panel: columns [j0, j0 + w), rows [j0, N); interchanges applied inside the panel only.
static int lu_panel(cudaStream_t st, double *A, int64_t lda, int N, int j0, int w, int *piv, int *info) {
    if (w <= LU_W) return lu_base(st, A, lda, N, j0, w, piv, info);
    const int h = ((w / 2 + LU_W - 1) / LU_W) * LU_W;  // left half, a multiple of the base width
    int rc = lu_panel(st, A, lda, N, j0, h, piv, info);
    if (rc) return rc;
    const int rcols = w - h;
    sdb_count_launch();
    lu_laswp_kernel<<<(rcols + 127) / 128, 128, 0, st>>>(A, lda, j0 + h, j0 + w, piv, j0, j0 + h);
    double *A11 = A + (int64_t)j0 * lda + j0;
    double *A12 = A + (int64_t)(j0 + h) * lda + j0;
    double *A21 = A11 + h;
    double *A22 = A12 + h;
    rc = lu_trsm_lunit(st, A11, lda, A12, lda, h, rcols);
    if (rc) return rc;
    rc = sdb_gemm(st, N - j0 - h, rcols, h, -1.0, A21, 1, lda, A12, 1, lda, 1.0, A22, lda);
    if (rc) return rc;
    rc = lu_panel(st, A, lda, N, j0 + h, rcols, piv, info);
    if (rc) return rc;
    sdb_count_launch();
    lu_laswp_kernel<<<(h + 127) / 128, 128, 0, st>>>(A, lda, j0, j0 + h, piv, j0 + h, j0 + w);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

static int lu_factor(cudaStream_t st, double *A, int64_t lda, int N, int *piv, int *info) {
    for (int j0 = 0; j0 < N; j0 += LU_NB) {
        const int w = (N - j0) < LU_NB ? (N - j0) : LU_NB;
        int rc = lu_panel(st, A, lda, N, j0, w, piv, info);
        if (rc) return rc;
        if (j0 > 0) { sdb_count_launch(); lu_laswp_kernel<<<(j0 + 127) / 128, 128, 0, st>>>(A, lda, 0, j0, piv, j0, j0 + w); }
        const int rest = N - j0 - w;
        if (rest > 0) {
            sdb_count_launch();
            lu_laswp_kernel<<<(rest + 127) / 128, 128, 0, st>>>(A, lda, j0 + w, N, piv, j0, j0 + w);
            double *A11 = A + (int64_t)j0 * lda + j0;
            double *A12 = A + (int64_t)(j0 + w) * lda + j0;
            rc = lu_trsm_lunit(st, A11, lda, A12, lda, w, rest);
            if (rc) return rc;
            rc = sdb_gemm(st, rest, rest, w, -1.0, A11 + w, 1, lda, A12, 1, lda, 1.0, A12 + w, lda);
            if (rc) return rc;
        }
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Solve phase.  B is column-major N x nrhs (ldb).  One warp per right-hand side works on a diagonal
// block of up to 64 rows: lane l owns rows l and l + 32.
// ---------------------------------------------------------------------------------------------------
template <bool UPPER>
__global__ void __launch_bounds__(256) lu_trsv_block_kernel(const double *__restrict__ A, int64_t lda, double *__restrict__ B,
                                                            int64_t ldb, int k0, int h, int nrhs) {
    __shared__ double T[64][65];
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int i = e % 64, k = e / 64;
        T[i][k] = (i < h && k < h) ? A[(int64_t)(k0 + k) * lda + k0 + i] : (i == k ? 1.0 : 0.0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int rhs = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rhs < nrhs; rhs += gridDim.x * (blockDim.x >> 5)) {
        double *b = B + (int64_t)rhs * ldb + k0;
        double x0 = lane < h ? b[lane] : 0.0, x1 = lane + 32 < h ? b[lane + 32] : 0.0;
        if (!UPPER) {
            for (int i = 0; i < h; ++i) {
                const double xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);  // unit diagonal
                if (lane > i) x0 = fma(-T[lane][i], xi, x0);
                if (lane + 32 > i) x1 = fma(-T[lane + 32][i], xi, x1);
            }
        } else {
            for (int i = h - 1; i >= 0; --i) {
                double xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);
                xi = xi / T[i][i];
                if (lane == (i & 31)) { if (i < 32) x0 = xi; else x1 = xi; }
                if (lane < i) x0 = fma(-T[lane][i], xi, x0);
                if (lane + 32 < i) x1 = fma(-T[lane + 32][i], xi, x1);
            }
        }
        if (lane < h) b[lane] = x0;
        if (lane + 32 < h) b[lane + 32] = x1;
    }
}

__global__ void __launch_bounds__(128) lu_apply_piv_rhs_kernel(double *__restrict__ B, int64_t ldb, int nrhs,
                                                               const int *__restrict__ piv, int N) {
    const int rhs = blockIdx.x * blockDim.x + threadIdx.x;
    if (rhs >= nrhs) return;
    double *b = B + (int64_t)rhs * ldb;
    for (int k = 0; k < N; ++k) {
        const int p = piv[k];
        if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
    }
}

static int lu_solve_factored(cudaStream_t st, const double *A, int64_t lda, int N, const int *piv, double *B, int64_t ldb,
                             int nrhs) {
    sdb_count_launch();
    lu_apply_piv_rhs_kernel<<<(nrhs + 127) / 128, 128, 0, st>>>(B, ldb, nrhs, piv, N);
    const int tb = (nrhs + 7) / 8;
    for (int k0 = 0; k0 < N; k0 += 64) {  // L y = P b
        const int h = (N - k0) < 64 ? (N - k0) : 64;
        sdb_count_launch();
        lu_trsv_block_kernel<false><<<tb, 256, 0, st>>>(A, lda, B, ldb, k0, h, nrhs);
        const int rest = N - k0 - h;
        if (rest > 0) {
            int rc = sdb_gemm(st, rest, nrhs, h, -1.0, A + (int64_t)k0 * lda + k0 + h, 1, lda, B + k0, 1, ldb, 1.0, B + k0 + h, ldb);
            if (rc) return rc;
        }
    }
    for (int k1 = N; k1 > 0;) {  // U x = y
        const int h = (k1 % 64) ? (k1 % 64) : 64;
        const int k0 = k1 - h;
        sdb_count_launch();
        lu_trsv_block_kernel<true><<<tb, 256, 0, st>>>(A, lda, B, ldb, k0, h, nrhs);
        if (k0 > 0) {
            int rc = sdb_gemm(st, k0, nrhs, h, -1.0, A + (int64_t)k0 * lda, 1, lda, B + k0, 1, ldb, 1.0, B, ldb);
            if (rc) return rc;
        }
        k1 = k0;
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// Factor A (column-major N x N, overwritten by L and U) and solve A X = B for nrhs columns
// (B column-major N x nrhs, leading dimension N, overwritten by X).  d_piv: N ints.  d_info: 0, or
// k+1 when U(k,k) is exactly zero (LAPACK's info).
extern "C" int sdb_lu_solve(double *d_A, int64_t N, double *d_B, int32_t nrhs, int32_t *d_piv, int32_t *d_info,
                            void *stream) {
    SDB_CHECK_ARG(d_A && d_piv && d_info, "sdb_lu_solve: NULL pointer");
    SDB_CHECK_ARG(N >= 1 && N < (1 << 30), "sdb_lu_solve: N = %lld out of range", (long long)N);
    SDB_CHECK_ARG(nrhs >= 0 && (nrhs == 0 || d_B), "sdb_lu_solve: bad right-hand sides");
    cudaStream_t st = (cudaStream_t)stream;
    SDB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    int rc = lu_factor(st, d_A, N, (int)N, d_piv, d_info);
    if (rc) return rc;
    if (nrhs > 0) rc = lu_solve_factored(st, d_A, N, (int)N, d_piv, d_B, N, nrhs);
    return rc;
}

// C (column-major, ldc) = beta C + alpha A B with generic strides (see sdb_gemm.cuh); nslab > 1 splits K
// deterministically and needs d_ws >= nslab * M * N doubles.  Exposed for tests and tooling.
extern "C" int sdb_dgemm(int32_t M, int32_t N, int32_t K, double alpha, const double *d_A, int64_t sa_m, int64_t sa_k,
                         const double *d_B, int64_t sb_k, int64_t sb_n, double beta, double *d_C, int64_t ldc, int32_t nslab,
                         double *d_ws, void *stream) {
    SDB_CHECK_ARG(d_A && d_B && d_C && M >= 0 && N >= 0 && K >= 0, "sdb_dgemm: bad arguments");
    SDB_CHECK_ARG(nslab <= 1 || d_ws, "sdb_dgemm: split-K needs a workspace");
    return sdb_gemm((cudaStream_t)stream, M, N, K, alpha, d_A, sa_m, sa_k, d_B, sb_k, sb_n, beta, d_C, ldc, nslab, d_ws);
}
