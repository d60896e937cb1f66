// RBF interpolation system solved by the NULL-SPACE method + Cholesky (sm_100a, FP64).
//
// The saddle-point system  [[Phi, P], [P^T, 0]] [lambda; c] = [F; 0]  (surrogate_rbf.py:121-130) is
// symmetric INDEFINITE, so the reference's 'direct' solver (and sdb_lu_solve) is an LU with partial pivoting
// whose pivot search serialises the whole GPU once per column.  For the kernels that are conditionally
// definite on the null space of P^T -- r (type 0, 6: negative definite there), r^3 (type 1), exp(-r^2)
// (type 4) and (r^2+1)^(-1/2) (type 5) -- the same solution is obtained without any pivoting:
//     P = Q [R; 0]                      Householder QR, q = d + 1 reflectors
//     B = Q^T Phi Q,  g = Q^T F         two-sided reflector application, O(q n^2)
//     y1 = 0,  (s B22) y2 = s g2        B22 = B[q:, q:] is s-definite (s = -1 for r, +1 otherwise): CHOLESKY
//     R c = g1 - B12 y2,  lambda = Q [0; y2]
// One third of the LU's flops, and the factorisation's critical path per 64 columns is a 64 x 64 POTRF in
// one CTA instead of 64 cluster-wide pivot searches.  Opt-in (solver_type 'nullspace' of Surrogate_update);
// a non-positive pivot is reported in info (the caller falls back to sdb_lu_solve).
#include <float.h>

#include "sdb_gemm.cuh"
#include "sdb_graph.cuh"

#define NS_T 1024

__device__ __forceinline__ double ns_block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    return t;
}

// ---------------------------------------------------------------------------------------------------
// Householder QR of P (n x q, column-major, leading dimension ldp), LAPACK dgeqr2 conventions: v_j below the
// diagonal (v_j(j) = 1 implicit), R on and above it, tau[q].  Q^T is applied to the m columns of F (ldf).
// One CTA: the norm of column j is a block reduction; the remaining columns are handled one per warp.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NS_T, 1) ns_house_qr_kernel(double *__restrict__ P, int64_t ldp, int n, int q,
                                                              double *__restrict__ tau, double *__restrict__ F, int64_t ldf, int m) {
    __shared__ double scratch[32];
    __shared__ double s_tau, s_scale, s_beta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int j = 0; j < q && j < n; ++j) {
        double *x = P + (int64_t)j * ldp;
        double ss = 0.0;
        for (int i = j + 1 + tid; i < n; i += blockDim.x) ss = fma(x[i], x[i], ss);
        ss = ns_block_sum(ss, scratch);
        if (tid == 0) {
            const double alpha = x[j];
            if (ss == 0.0) { s_tau = 0.0; s_scale = 0.0; s_beta = alpha; }
            else {
                const double beta = -copysign(sqrt(alpha * alpha + ss), alpha);
                s_tau = (beta - alpha) / beta;
                s_scale = 1.0 / (alpha - beta);
                s_beta = beta;
            }
        }
        __syncthreads();
        const double tj = s_tau, sc = s_scale;
        for (int i = j + 1 + tid; i < n; i += blockDim.x) x[i] *= sc;
        if (tid == 0) { x[j] = s_beta; tau[j] = tj; }
        __syncthreads();
        // apply H_j = I - tau v v^T to the later columns of P and to F: one warp per column
        const int ncol = (q - 1 - j) + m;
        for (int cc = warp; cc < ncol; cc += nwarp) {
            double *col = cc < q - 1 - j ? P + (int64_t)(j + 1 + cc) * ldp : F + (int64_t)(cc - (q - 1 - j)) * ldf;
            double w = 0.0;
            for (int i = j + 1 + lane; i < n; i += 32) w = fma(x[i], col[i], w);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
            w += col[j];                      // v_j(j) = 1
            const double tw = tj * w;
            for (int i = j + 1 + lane; i < n; i += 32) col[i] = fma(-tw, x[i], col[i]);
            if (lane == 0) col[j] -= tw;
        }
        __syncthreads();
    }
}

// explicit reflector vector v_j (length n): 0 above j, 1 at j, P(i, j) below
__global__ void __launch_bounds__(256) ns_vec_kernel(const double *__restrict__ P, int64_t ldp, int n, int j, double *__restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i < j ? 0.0 : (i == j ? 1.0 : P[(int64_t)j * ldp + i]);
}

// w = Phi v, Phi symmetric n x n with leading dimension ld; one warp per row
__global__ void __launch_bounds__(256) ns_symv_kernel(const double *__restrict__ Phi, int64_t ld, int n, const double *__restrict__ v,
                                                      double *__restrict__ w) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *a = Phi + (int64_t)row * ld;
    double s = 0.0;
    for (int c = lane; c < n; c += 32) s = fma(a[c], v[c], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) w[row] = s;
}

// z = tau w - (tau^2 (v.w) / 2) v
__global__ void __launch_bounds__(NS_T, 1) ns_z_kernel(int n, const double *__restrict__ v, const double *__restrict__ w,
                                                       const double *__restrict__ tau, int j, double *__restrict__ z) {
    __shared__ double scratch[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fma(v[i], w[i], a);
    a = ns_block_sum(a, scratch);
    const double t = tau[j], h = 0.5 * t * t * a;
    for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = t * w[i] - h * v[i];
}

// Phi -= v z^T + z v^T; entries with r >= q0 and c >= q0 are additionally multiplied by `scale` (the sign that
// makes B22 positive definite, applied with the last reflector)
__global__ void __launch_bounds__(256) ns_syr2_kernel(double *__restrict__ Phi, int64_t ld, int n, const double *__restrict__ v,
                                                      const double *__restrict__ z, int q0, double scale) {
    const int r = blockIdx.x * 32 + (threadIdx.x & 31);
    const int c0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
    if (r >= n) return;
    const double vr = v[r], zr = z[r];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        if (c < n) {
            double *p = Phi + (int64_t)c * ld + r;
            double val = *p - (vr * z[c] + zr * v[c]);
            if (r >= q0 && c >= q0) val *= scale;
            *p = val;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Cholesky (lower, column-major, leading dimension ld) of the nb x nb matrix A, 64-column blocks
// ---------------------------------------------------------------------------------------------------
// 64 x 64 diagonal block in ONE CTA of 256 threads arranged 16 x 16; thread (ty, tx) keeps the 4 x 4 tile
// rows 4ty.., columns 4tx.. in registers.  Per column: the owners of column j publish it unscaled (double-buffered),
// every thread forms 1/sqrt(a_jj) itself and applies the rank-1 update to its tile: ONE barrier per column.
__global__ void __launch_bounds__(256, 1) chol_potrf64_kernel(double *__restrict__ A, int64_t ld, int k0, int h, int *__restrict__ info) {
    __shared__ double col[2][64];     // the UNSCALED column j, double-buffered: one barrier per column
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = ty * 4 + r, cc = tx * 4 + c;
            a[r][c] = (i < h && cc <= i) ? A[(int64_t)(k0 + cc) * ld + k0 + i] : (i == cc ? 1.0 : 0.0);
        }
    for (int jb = 0; jb < 16; ++jb) {            // column block; inside, the 4 columns are unrolled (static register indices)
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
            const int j = jb * 4 + jc;
            if (j < h) {                          // uniform
                double *cj = col[j & 1];
                if (tx == jb) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) cj[ty * 4 + r] = (ty * 4 + r >= j) ? a[r][jc] : 0.0;
                }
                __syncthreads();
                const double djj = cj[j];
                const bool pd = djj > 0.0;
                if (!pd && tid == 0 && *info == 0) *info = k0 + j + 1;      // not positive definite (or NaN)
                const double rd = pd ? sdb_fast_rsqrt(djj) : 1.0;           // 1 / L(j, j)
                if (tx == jb) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = ty * 4 + r;
                        a[r][jc] = i > j ? a[r][jc] * rd : (i == j ? (pd ? djj * rd : 1.0) : 0.0);
                    }
                }
                if (tx >= jb) {
                    double lr[4], lc[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) lr[r] = cj[ty * 4 + r] * rd;
#pragma unroll
                    for (int c = 0; c < 4; ++c) lc[c] = cj[tx * 4 + c] * rd;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (tx * 4 + c > j) a[r][c] = fma(-lr[r], lc[c], a[r][c]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = ty * 4 + r, cc = tx * 4 + c;
            if (i < h && cc <= i) A[(int64_t)(k0 + cc) * ld + k0 + i] = a[r][c];
        }
}

// X = L11^{-1} (64 x 64 lower triangular) -> Xout column-major with leading dimension 64.  Sixteen lanes share a
// column: row i of X needs sum_k L(i, k) X(k, c), split over the lanes and combined with four shuffles; only those
// sixteen lanes ever read column c, so a warp-level sync per row suffices.  The panel's triangular solve then
// becomes one DMMA multiplication L21 = A21 X^T (sdb_gemm_k64_inplace).
__global__ void __launch_bounds__(1024, 1) chol_trinv64_kernel(const double *__restrict__ A, int64_t ld, int k0, int h,
                                                               double *__restrict__ Xout) {
    // one 64 x 65 array holds both triangles: L(i, c) at Ls[i][c] (c <= i), X(k, c) at Ls[c][k + 1] (k >= c)
    __shared__ double Ls[64][65];
    __shared__ double rdiag[64];
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * 64; e += 1024) {
        const int i = e & 63, c = e >> 6;
        if (c <= i) Ls[i][c] = (i < h) ? A[(int64_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid < 64) rdiag[tid] = 1.0 / Ls[tid][tid];
    __syncthreads();
    const int c = tid >> 4, part = tid & 15;     // sixteen lanes share a column: at most 4 terms of the row sum each
    for (int i = 0; i < 64; ++i) {      // uniform trip count: the full-mask shuffles need every lane of the warp
        double s = 0.0;
        for (int k = c + part; k < i; k += 16) s = fma(Ls[i][k], Ls[c][k + 1], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        if (part == 0 && i >= c) Ls[c][i + 1] = ((i == c ? 1.0 : 0.0) - s) * rdiag[i];
        __syncwarp();
    }
    __syncthreads();
    for (int e = tid; e < 64 * 64; e += 1024) {
        const int i = e & 63, cc = e >> 6;
        Xout[cc * 64 + i] = (i < h && cc < h && cc <= i) ? Ls[cc][i + 1] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------
// Triangular solves with the inverses of the diagonal blocks (kept from the factorisation), one kernel per
// 64-block and direction: every CTA forms the block's solution redundantly (a 64 x 64 mat-vec) and then
// updates its own rows of the remaining right-hand side.
//   forward : x_blk = Linv_k b_blk              -> X;   B[r] -= sum_k L(r, k0+k) x_k   for r >= k0 + h
//   backward: y_blk = Linv_k^T w_blk            -> Y;   W[i] -= sum_k L(k0+k, i) y_k   for i <  k0
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chol_fwd_kernel(const double *__restrict__ A, int64_t ld, const double *__restrict__ Linv,
                                                       double *__restrict__ B, double *__restrict__ X, int64_t ldb, int k0, int h,
                                                       int nb, int m) {
    __shared__ double Ls[64][65];      // Ls[t][k] = Linv(t, k): staged with coalesced, independent loads
    __shared__ double bb[64], xs[64];
    const int t = threadIdx.x;
    for (int e = t; e < 64 * 64; e += 256) Ls[e & 63][e >> 6] = Linv[e];      // Linv[k*64 + t]
    for (int rhs = 0; rhs < m; ++rhs) {
        double *b = B + (int64_t)rhs * ldb;
        if (t < 64) bb[t] = t < h ? b[k0 + t] : 0.0;
        __syncthreads();
        {
            const int o = t >> 2, part = t & 3;               // four lanes per entry of the block solution
            double s = 0.0;
            for (int k = part; k <= o && k < h; k += 4) s = fma(Ls[o][k], bb[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0) {
                xs[o] = o < h ? s : 0.0;
                if (blockIdx.x == 0 && o < h) X[(int64_t)rhs * ldb + k0 + o] = s;
            }
        }
        __syncthreads();
        const int r = k0 + h + blockIdx.x * 256 + t;
        if (r < nb) {
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
            for (int k = 0; k + 1 < h; k += 2) {
                acc0 = fma(A[(int64_t)(k0 + k) * ld + r], xs[k], acc0);
                acc1 = fma(A[(int64_t)(k0 + k + 1) * ld + r], xs[k + 1], acc1);
            }
            if (h & 1) acc0 = fma(A[(int64_t)(k0 + h - 1) * ld + r], xs[h - 1], acc0);
            b[r] -= acc0 + acc1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) chol_bwd_kernel(const double *__restrict__ A, int64_t ld, const double *__restrict__ Linv,
                                                       double *__restrict__ W, double *__restrict__ Y, int64_t ldb, int k0, int h,
                                                       int m) {
    __shared__ double Ls[64][65];      // Ls[k][t] = Linv(k, t), read as Linv^T(t, k)
    __shared__ double wb[64], ys[64];
    const int t = threadIdx.x;
    for (int e = t; e < 64 * 64; e += 256) Ls[e & 63][e >> 6] = Linv[e];      // Linv[c*64 + i] -> Ls[i][c]
    for (int rhs = 0; rhs < m; ++rhs) {
        double *w = W + (int64_t)rhs * ldb;
        if (t < 64) wb[t] = t < h ? w[k0 + t] : 0.0;
        __syncthreads();
        {
            const int o = t >> 2, part = t & 3;
            double s = 0.0;
            for (int k = o + part; k < h; k += 4) s = fma(Ls[k][o], wb[k], s);   // Linv^T(o, k) = Linv(k, o)
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0) {
                ys[o] = o < h ? s : 0.0;
                if (blockIdx.x == 0 && o < h) Y[(int64_t)rhs * ldb + k0 + o] = s;
            }
        }
        __syncthreads();
        const int i = blockIdx.x * 256 + t;
        if (i < k0) {
            const double *a = A + (int64_t)i * ld + k0;        // L(k0 + k, i), contiguous in k
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
            for (int k = 0; k + 1 < h; k += 2) {
                acc0 = fma(a[k], ys[k], acc0);
                acc1 = fma(a[k + 1], ys[k + 1], acc1);
            }
            if (h & 1) acc0 = fma(a[h - 1], ys[h - 1], acc0);
            w[i] -= acc0 + acc1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) ns_scale_rhs_kernel(double *__restrict__ F, int64_t ldf, int q, int n, int m, double scale) {
    const int64_t total = (int64_t)(n - q) * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / (n - q)), i = q + (int)(e % (n - q));
        F[(int64_t)k * ldf + i] *= scale;
    }
}

// c = R^{-1} (g1 - B12 y2),  lambda = Q [0; y2]  ->  COEFS [(n + q) x m] row-major.  One CTA per right-hand side.
__global__ void __launch_bounds__(NS_T, 1) ns_finish_kernel(const double *__restrict__ S, int64_t ld, int n, int q,
                                                            const double *__restrict__ tau, double *__restrict__ F, int64_t ldf,
                                                            int m, double *__restrict__ COEFS) {
    __shared__ double scratch[32];
    __shared__ double t1[SDB_MAX_D + 1], cc[SDB_MAX_D + 1];
    const int rhs = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    double *f = F + (int64_t)rhs * ldf;
    const double *P = S + (int64_t)n * ld;          // columns n .. n+q-1: reflectors below, R on / above the diagonal
    // t1_i = g1_i - sum_c B12(i, c) y2_c,  B12(i, c) = S[c*ld + i] (unscaled rows i < q)
    for (int i = warp; i < q; i += nwarp) {
        double s = 0.0;
        for (int c = q + lane; c < n; c += 32) s = fma(S[(int64_t)c * ld + i], f[c], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) t1[i] = f[i] - s;
    }
    __syncthreads();
    if (tid == 0) {                                   // R c = t1, R upper triangular
        for (int i = q - 1; i >= 0; --i) {
            double s = t1[i];
            for (int k = i + 1; k < q; ++k) s -= P[(int64_t)k * ld + i] * cc[k];
            cc[i] = s / P[(int64_t)i * ld + i];
        }
    }
    __syncthreads();
    for (int i = tid; i < q; i += blockDim.x) f[i] = 0.0;      // [0; y2]
    __syncthreads();
    for (int j = q - 1; j >= 0; --j) {                          // lambda = H_0 ... H_{q-1} [0; y2]
        const double *v = P + (int64_t)j * ld;
        double w = 0.0;
        for (int i = j + 1 + tid; i < n; i += blockDim.x) w = fma(v[i], f[i], w);
        w = ns_block_sum(w, scratch);
        w += f[j];
        const double tw = tau[j] * w;
        __syncthreads();
        for (int i = j + 1 + tid; i < n; i += blockDim.x) f[i] = fma(-tw, v[i], f[i]);
        if (tid == 0) f[j] -= tw;
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x) COEFS[(int64_t)i * m + rhs] = f[i];
    for (int i = tid; i < q; i += blockDim.x) COEFS[(int64_t)(n + i) * m + rhs] = cc[i];
}

static bool ns_kernel_supported(int kt) { return kt == 0 || kt == 1 || kt == 4 || kt == 5 || kt == 6; }

static int ns_enqueue(cudaStream_t st, double *S, int n, int d, int m, int kernel_type, double *F, double *COEFS,
                      double *vec, int32_t *d_info) {
    const int q = d + 1, N = n + q, nb = n - q;
    const int64_t ld = N;
    double *P = S + (int64_t)n * ld;
    double *tau = vec, *v = vec + 64, *w = v + n, *z = w + n;
    double *xf = z + n;                                   // forward-solve result, m columns of length ld
    double *xinv_all = xf + (int64_t)m * ld;              // inverse of every 64 x 64 diagonal block
    const double sgn = (kernel_type == 0 || kernel_type == 6) ? -1.0 : 1.0;
    SDB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    sdb_count_launch();
    ns_house_qr_kernel<<<1, NS_T, 0, st>>>(P, ld, n, q, tau, F, ld, m);
    for (int j = 0; j < q; ++j) {
        sdb_count_launch();
        ns_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, ld, n, j, v);
        sdb_count_launch();
        ns_symv_kernel<<<(n + 7) / 8, 256, 0, st>>>(S, ld, n, v, w);
        sdb_count_launch();
        ns_z_kernel<<<1, NS_T, 0, st>>>(n, v, w, tau, j, z);
        sdb_count_launch();
        ns_syr2_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), 256, 0, st>>>(S, ld, n, v, z, q, j == q - 1 ? sgn : 1.0);
    }
    if (sgn != 1.0) {
        sdb_count_launch();
        ns_scale_rhs_kernel<<<148, 256, 0, st>>>(F, ld, q, n, m, sgn);
    }
    // Cholesky of B22 = S[q:n, q:n], right-looking with a look-ahead of one panel: after the triangular solve
    // of panel k the trailing update is split into the 64 columns of panel k+1 (main stream) and the rest (side
    // stream); panel k+1 is factored on the main stream while the rest of update k is still running.
    double *A = S + (int64_t)q * ld + q;
    static cudaStream_t aux = nullptr;
    static cudaEvent_t ev_trsm = nullptr, ev_rest = nullptr;
    if (!aux) {
        SDB_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        SDB_CUDA(cudaEventCreateWithFlags(&ev_trsm, cudaEventDisableTiming));
        SDB_CUDA(cudaEventCreateWithFlags(&ev_rest, cudaEventDisableTiming));
    }
    bool rest_pending = false;   // a side-stream update whose completion the main stream has not yet waited for
    for (int k0 = 0; k0 < nb; k0 += 64) {
        const int h = (nb - k0) < 64 ? (nb - k0) : 64;
        sdb_count_launch();
        chol_potrf64_kernel<<<1, 256, 0, st>>>(A, ld, k0, h, d_info);
        const int rest = nb - k0 - h;
        double *xinv = xinv_all + (int64_t)(k0 / 64) * 4096;
        sdb_count_launch();
        chol_trinv64_kernel<<<1, 1024, 0, st>>>(A, ld, k0, h, xinv);
        if (rest > 0) {
            double *L21 = A + (int64_t)k0 * ld + k0 + h;
            // L21 = A21 L11^{-T}:  (A21 X^T)(r, j) = sum_k A21(r, k) X(j, k),  X(j, k) = xinv[k*64 + j]
            int rc2 = sdb_gemm_k64_inplace(st, rest, h, L21, ld, xinv, 64);
            if (rc2) return rc2;
            double *A22 = A + (int64_t)(k0 + h) * ld + k0 + h;
            const int nxt = rest < 64 ? rest : 64;      // columns of the next panel
            // the previous side-stream update also wrote into the columns touched from here on
            if (rest_pending) { SDB_CUDA(cudaStreamWaitEvent(st, ev_rest, 0)); rest_pending = false; }
            SDB_CUDA(cudaEventRecord(ev_trsm, st));
            int rc = sdb_gemm_k64_launch<true, false, false>(st, rest, nxt, h, L21, L21, A22, ld, ld);     // next panel's columns
            if (rc) return rc;
            if (rest > 64) {
                SDB_CUDA(cudaStreamWaitEvent(aux, ev_trsm, 0));
                rc = sdb_syrk_k64_update(aux, rest - 64, rest - 64, h, L21 + 64, L21 + 64, A22 + (int64_t)64 * ld + 64, ld);
                if (rc) return rc;
                SDB_CUDA(cudaEventRecord(ev_rest, aux));
                rest_pending = true;
            }
        }
    }
    if (rest_pending) SDB_CUDA(cudaStreamWaitEvent(st, ev_rest, 0));
    // (s B22) y2 = s g2: forward with L, backward with L^T; right-hand sides in F[q:n], result back in F[q:n]
    double *G2 = F + q;
    for (int k0 = 0; k0 < nb; k0 += 64) {
        const int h = (nb - k0) < 64 ? (nb - k0) : 64;
        const int rest = nb - k0 - h;
        sdb_count_launch();
        chol_fwd_kernel<<<rest > 0 ? (rest + 255) / 256 : 1, 256, 0, st>>>(A, ld, xinv_all + (int64_t)(k0 / 64) * 4096, G2, xf, ld, k0,
                                                                           h, nb, m);
    }
    for (int k0 = ((nb - 1) / 64) * 64; k0 >= 0; k0 -= 64) {
        const int h = (nb - k0) < 64 ? (nb - k0) : 64;
        sdb_count_launch();
        chol_bwd_kernel<<<k0 > 0 ? (k0 + 255) / 256 : 1, 256, 0, st>>>(A, ld, xinv_all + (int64_t)(k0 / 64) * 4096, xf, G2, ld, k0, h,
                                                                       m);
    }
    sdb_count_launch();
    ns_finish_kernel<<<m, NS_T, 0, st>>>(S, ld, n, q, tau, F, ld, m, COEFS);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

extern "C" int64_t sdb_rbf_fit_nullspace_work(int64_t n, int32_t d, int32_t m) {
    const int64_t N = n + d + 1;
    // system | right-hand sides | tau, v, w, z | forward-solve result | inverses of the diagonal blocks
    return N * N + N * m + 3 * n + 128 + N * m + (n / 64 + 2) * 4096;
}

// Same contract as sdb_rbf_fit (X [n x d], Y [n x m] -> d_COEFS [(n+d+1) x m] row-major) for kernel types
// 0, 1, 4, 5, 6.  d_info: 0, or k+1 if the k-th pivot of the Cholesky factorisation was not positive (the
// projected kernel matrix is numerically not definite: use sdb_rbf_fit).  SDB_ERR_ARG for other kernel types.
extern "C" int sdb_rbf_fit_nullspace(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                                     double *d_COEFS, double *d_work, int32_t *d_info, void *stream) {
    SDB_CHECK_ARG(d_X && d_Y && d_COEFS && d_work && d_info, "sdb_rbf_fit_nullspace: NULL pointer");
    SDB_CHECK_ARG(ns_kernel_supported(kernel_type),
                  "sdb_rbf_fit_nullspace: kernel_type %d is not conditionally definite with a linear tail (use sdb_rbf_fit)", kernel_type);
    SDB_CHECK_ARG(n > d + 1, "sdb_rbf_fit_nullspace: needs more than d + 1 points");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D && m >= 1 && m <= SDB_MAX_M, "sdb_rbf_fit_nullspace: d=%d m=%d outside limits", d, m);
    const int64_t N = n + d + 1;
    double *S = d_work;
    double *F = S + N * N;
    double *vec = F + N * m;
    cudaStream_t st = (cudaStream_t)stream;
    auto enqueue = [&]() -> int {
        int rc = sdb_rbf_system(d_X, d_Y, n, d, m, kernel_type, S, F, stream);
        if (rc) return rc;
        return ns_enqueue(st, S, (int)n, d, m, kernel_type, F, d_COEFS, vec, d_info);
    };
    if (n < 256) return enqueue();
    sdb_graph_key key = {{d_X, d_Y, d_COEFS, d_work, d_info}, {n, d, m, kernel_type}, 2, 0};
    return sdb_graph_run(st, key, enqueue);
}
