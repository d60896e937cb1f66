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

// 512 threads x <= 64 registers: these single-CTA kernels must fit beside a resident chain CTA (448 threads x 64
// registers) when the refit runs concurrently with the next block of chain steps
#define NS_T 512

__device__ __forceinline__ int ns_ld_acquire_early(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ns_st_release_early(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
__global__ void __launch_bounds__(NS_T, 2) ns_house_qr_kernel(double *__restrict__ P, int64_t ldp, int n, int q,
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
            // four independent partial sums / eight loads in flight per lane: one warp walks a whole column, so the loop is
            // bound by the latency of its loads unless several are outstanding
            double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
            int i = j + 1 + lane;
            for (; i + 96 < n; i += 128) {
                w0 = fma(x[i], col[i], w0);
                w1 = fma(x[i + 32], col[i + 32], w1);
                w2 = fma(x[i + 64], col[i + 64], w2);
                w3 = fma(x[i + 96], col[i + 96], w3);
            }
            for (; i < n; i += 32) w0 = fma(x[i], col[i], w0);
            double w = (w0 + w1) + (w2 + w3);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
            w += col[j];                      // v_j(j) = 1
            const double tw = tj * w;
#pragma unroll 4
            for (int i2 = j + 1 + lane; i2 < n; i2 += 32) col[i2] = fma(-tw, x[i2], col[i2]);
            if (lane == 0) col[j] -= tw;
        }
        __syncthreads();
    }
}

// The same factorisation for many reflectors (q > 4), one CTA per column of [P | F], cooperative launch: the CTA of column
// j generates H_j as soon as H_0..H_{j-1} have been applied to its column and publishes it (release flag); every CTA
// to the right waits for that flag (acquire) and applies H_j to its own column with all its threads.  One flag
// hand-over and two short passes over a column per reflector, instead of one warp walking whole columns.
__global__ void __launch_bounds__(NS_T, 2) ns_house_qr_grid_kernel(double *P, int64_t ldp, int n, int q, double *tau, double *F,
                                                                   int64_t ldf, int m, int *flag) {
    __shared__ double scratch[32];
    __shared__ double s_tau, s_scale, s_beta;
    const int tid = threadIdx.x, cc = blockIdx.x;
    double *col = cc < q ? P + (int64_t)cc * ldp : F + (int64_t)(cc - q) * ldf;
    const int nj = cc < q ? cc : q;                      // reflectors to apply to this column
    for (int j = 0; j < nj && j < n; ++j) {
        if (tid == 0)
            while (ns_ld_acquire_early(flag + j) == 0) { }
        __syncthreads();
        const double *x = P + (int64_t)j * ldp;          // v_j below the diagonal, written by the CTA of column j
        const double tj = __ldcg(tau + j);
        double w = 0.0;
        for (int i = j + 1 + tid; i < n; i += NS_T) w = fma(__ldcg(x + i), col[i], w);
        w = ns_block_sum(w, scratch);
        w += col[j];                                     // v_j(j) = 1
        const double tw = tj * w;
        __syncthreads();                                 // col[j] read by everyone before thread 0 changes it
        for (int i = j + 1 + tid; i < n; i += NS_T) col[i] = fma(-tw, __ldcg(x + i), col[i]);
        if (tid == 0) col[j] -= tw;
        __syncthreads();
    }
    if (cc < q && cc < n) {                              // generate H_cc from the updated column
        const int j = cc;
        double ss = 0.0;
        for (int i = j + 1 + tid; i < n; i += NS_T) ss = fma(col[i], col[i], ss);
        ss = ns_block_sum(ss, scratch);
        if (tid == 0) {
            const double alpha = col[j];
            if (ss == 0.0) { s_tau = 0.0; s_scale = 0.0; s_beta = alpha; }
            else {
                const double beta = -copysign(sqrt(alpha * alpha + ss), alpha);
                s_tau = (beta - alpha) / beta;
                s_scale = 1.0 / (alpha - beta);
                s_beta = beta;
            }
        }
        __syncthreads();
        const double sc = s_scale;
        for (int i = j + 1 + tid; i < n; i += NS_T) col[i] *= sc;
        if (tid == 0) { col[j] = s_beta; tau[j] = s_tau; }
        __threadfence();
        __syncthreads();
        if (tid == 0) ns_st_release_early(flag + j, 1);
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
__global__ void __launch_bounds__(NS_T, 2) ns_z_kernel(int n, const double *__restrict__ v, const double *__restrict__ w,
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
// Blocked two-sided application of the q reflectors (compact WY, LAPACK dlarft / dsytrd-style):
//     Q = H_0 ... H_{q-1} = I - V T V^T,      B = Q^T Phi Q = Phi - V Y^T - Y V^T,
//     W = Phi V,  M = V^T W,  Y = W T - 1/2 V (T^T M T)
// i.e. ONE read of Phi (a tall-skinny DMMA GEMM) and ONE read-modify-write (a rank-2q DMMA update) instead of the
// three sweeps per reflector of the one-at-a-time form: 3 n^2 words of traffic instead of 3 q n^2.
// ---------------------------------------------------------------------------------------------------
// explicit V (n x q, unit lower trapezoidal) into both operand panels: VY = [V | Y], YV = [Y | V] (leading dimension n)
__global__ void __launch_bounds__(256) ns_vmat_kernel(const double *__restrict__ P, int64_t ldp, int n, int q,
                                                      double *__restrict__ VY, double *__restrict__ YV) {
    const int64_t total = (int64_t)n * q;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % n), j = (int)(e / n);
        const double v = i < j ? 0.0 : (i == j ? 1.0 : P[(int64_t)j * ldp + i]);
        VY[(int64_t)j * n + i] = v;
        YV[(int64_t)(q + j) * n + i] = v;
    }
}

// T (q x q upper triangular, dlarft forward / columnwise) from tau and G = V^T V, then Z = T^T M T.  One CTA; q <= 33.
__global__ void __launch_bounds__(256) ns_wy_small_kernel(int q, const double *__restrict__ tau, const double *__restrict__ G,
                                                          const double *__restrict__ Mq, double *__restrict__ Tout,
                                                          double *__restrict__ Zout) {
    __shared__ double T[SDB_MAX_D + 1][SDB_MAX_D + 2], U[SDB_MAX_D + 1][SDB_MAX_D + 2];
    const int tid = threadIdx.x;
    for (int e = tid; e < q * q; e += blockDim.x) T[e / q][e % q] = 0.0;
    __syncthreads();
    for (int j = 0; j < q; ++j) {
        // T(0:j, j) = -tau_j T(0:j, 0:j) (V(:, 0:j)^T v_j),  V(:, i)^T v_j = G(i, j)
        if (tid < j) {
            double s = 0.0;
            for (int l = tid; l < j; ++l) s = fma(T[tid][l], G[l * q + j], s);
            T[tid][j] = -tau[j] * s;
        }
        if (tid == j) T[j][j] = tau[j];
        __syncthreads();
    }
    // U = M T,  Z = T^T U
    for (int e = tid; e < q * q; e += blockDim.x) {
        const int r = e / q, c = e % q;
        double s = 0.0;
        for (int l = 0; l <= c; ++l) s = fma(Mq[r * q + l], T[l][c], s);       // M symmetric: row-major == column-major
        U[r][c] = s;
    }
    __syncthreads();
    for (int e = tid; e < q * q; e += blockDim.x) {
        const int r = e / q, c = e % q;
        double s = 0.0;
        for (int l = 0; l <= r; ++l) s = fma(T[l][r], U[l][c], s);
        Zout[r * q + c] = s;
        Tout[r * q + c] = T[r][c];
    }
}

// Y = W T - 1/2 V Z  ->  VY[:, q:2q] and YV[:, 0:q].  One thread per (row, column); W, V with leading dimension n.
__global__ void __launch_bounds__(256) ns_wy_y_kernel(int n, int q, const double *__restrict__ W, const double *__restrict__ T,
                                                      const double *__restrict__ Z, double *__restrict__ VY,
                                                      double *__restrict__ YV) {
    __shared__ double Ts[(SDB_MAX_D + 1) * (SDB_MAX_D + 1)], Zs[(SDB_MAX_D + 1) * (SDB_MAX_D + 1)];
    for (int e = threadIdx.x; e < q * q; e += blockDim.x) { Ts[e] = T[e]; Zs[e] = Z[e]; }
    __syncthreads();
    const int64_t total = (int64_t)n * q;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % n), c = (int)(e / n);
        double s = 0.0, t = 0.0;
        for (int l = 0; l <= c; ++l) s = fma(W[(int64_t)l * n + i], Ts[l * q + c], s);          // T upper triangular
        for (int l = 0; l < q; ++l) t = fma(VY[(int64_t)l * n + i], Zs[l * q + c], t);
        const double y = fma(-0.5, t, s);
        VY[(int64_t)(q + c) * n + i] = y;
        YV[(int64_t)c * n + i] = y;
    }
}

// B22 *= -1 (kernels that are NEGATIVE definite on the null space of P^T)
__global__ void __launch_bounds__(256) ns_negate_kernel(double *__restrict__ S, int64_t ld, int q, int n) {
    const int r = q + blockIdx.x * 32 + (threadIdx.x & 31);
    const int c0 = q + blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
    if (r >= n) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c0 + k < n) S[(int64_t)(c0 + k) * ld + r] = -S[(int64_t)(c0 + k) * ld + r];
}

// ---------------------------------------------------------------------------------------------------
// Cholesky (lower, column-major, leading dimension ld) of the nb x nb matrix A, 64-column blocks
// ---------------------------------------------------------------------------------------------------
// POTRF + inverse of the 64 x 64 diagonal block, blocked by FOUR columns, as a producer / consumer pipeline.
//   workers (warps 0..7): the trailing matrix and the growing inverse live in their DMMA accumulator fragments
//       (8 x 8 blocks, lower block triangle: every warp owns one block row of A and the complementary block row of X);
//       the rank-4 updates  A -= Lp Lp^T,  X -= Lp Xr  are m8n8k4 DMMAs whose operands are two small k-major panels
//       in shared memory (double-buffered);
//   panel warps (8: L, 9: X): factor the 64 x 4 panel / the 4 x 64 rows of X.  Every lane factors the 4 x 4 pivot
//       block redundantly in registers (broadcast loads, no shuffles): the serial chain is four reciprocal roots.
// Look-ahead inside the kernel: after panel s the workers FIRST update and hand over the four columns / rows of
// panel s+1 (named barrier 2), then finish the rest of update s while the panel warps already factor panel s+1
// (named barrier 1 hands the factored panel back).  A dependent DMMA has ~150 cycles of latency, a dependent DFMA 8:
// the serial chain therefore contains one DMMA per panel, everything else of the update overlaps with it.
// upd != 0: the block first receives the previous panel's contribution  D -= T T^T,  T = A[k0.., k0-64..k0-1]
// (the rows of the previous panel solve that face this block), also as DMMAs -- the look-ahead update of the
// diagonal block is done HERE, so the factorisation's critical path never waits for the panel-wide update kernel.
// ---------------------------------------------------------------------------------------------------
#define PI_S 68      // row stride of the k-major panels: == 4 (mod 16) doubles -> conflict-free fragment loads
#define PI_XS 65     // column stride of the staged inverse: row-wise stores by 32 lanes hit 32 distinct banks
#define PI_T 320     // 8 worker warps + 2 panel warps

#ifndef SDB_PI_PROF_ARG
#define SDB_PI_PROF_ARG
#define SDB_PI_PROF_PASS
#define SDB_PI_PROF_INIT
#define SDB_PI_STAMP(i)
#define SDB_PI_PROF_END
#endif

__device__ __forceinline__ void pi_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void pi_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(PI_T, 1) chol_potrf_inv64_kernel(double *__restrict__ A, int64_t ld, int k0, int h,
                                                                   double *__restrict__ Xout, int *__restrict__ info, int upd SDB_PI_PROF_ARG) {
    SDB_PI_PROF_INIT
    __shared__ __align__(16) double sm[(64 + 24) * PI_S];
    double (*T)[PI_S] = (double (*)[PI_S])sm;              // prologue: T[k][i]; afterwards the staged inverse Xs[c*PI_XS + i]
    double *Xs = sm;
    double (*LpB)[PI_S] = T + 64;                           // [2][4]: final L(:, j0..j0+3), k-major, double-buffered
    double (*XrB)[PI_S] = LpB + 8;                          // [2][4]: final X(j0..j0+3, :)
    double (*Pn)[PI_S] = XrB + 8;                           // the panel as extracted from the accumulators (not yet factored)
    double (*Xn)[PI_S] = Pn + 4;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int nsteps = (h + 3) >> 2;
    for (int e = tid; e < 4 * PI_S; e += PI_T) (&Xn[0][0])[e] = 0.0;      // columns right of the diagonal block are never written

    if (w >= 8) {
        // ---------------- panel warps ----------------
        __syncthreads();
        int bad = 0;
        for (int s = 0; s < nsteps; ++s) {
            const int j0 = 4 * s;
            double (*Lp)[PI_S] = LpB + 4 * (s & 1), (*Xr)[PI_S] = XrB + 4 * (s & 1);
            pi_bar_sync(2, PI_T);                            // panel s has left the accumulators
            double v[4][2], g[4][4], rd[4];
            double (*src)[PI_S] = w == 8 ? Pn : Xn;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int e = 0; e < 2; ++e) v[k][e] = src[k][lane + 32 * e];
#pragma unroll
                for (int i = k; i < 4; ++i) g[i][k] = Pn[k][j0 + i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double d = g[k][k];
                const bool pd = d > 0.0;
                if (!pd && !bad) bad = j0 + k + 1;
                const double rs = sdb_fast_rsqrt(pd ? d : 1.0);
                rd[k] = pd ? rs : 1.0;                                      // 1 / L(j, j)
#pragma unroll
                for (int i = k + 1; i < 4; ++i) g[i][k] *= rd[k];           // L(j0 + i, j0 + k)
#pragma unroll
                for (int k2 = k + 1; k2 < 4; ++k2)
#pragma unroll
                    for (int i = k2; i < 4; ++i) g[i][k2] = fma(-g[i][k], g[k2][k], g[i][k2]);
            }
            // rows of the panel (warp 8) / columns of the X rows (warp 9): the same forward substitution with the 4 x 4 factor.
            // No masks: rows above the pivot and columns right of it only ever feed entries that are already final.
            double (*dst)[PI_S] = w == 8 ? Lp : Xr;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) o[e] = v[k][e] * rd[k];          // row j itself: d * rd = L(j, j)
#pragma unroll
                for (int k2 = k + 1; k2 < 4; ++k2)
#pragma unroll
                    for (int e = 0; e < 2; ++e) v[k2][e] = fma(-o[e], g[k2][k], v[k2][e]);
#pragma unroll
                for (int e = 0; e < 2; ++e) dst[k][lane + 32 * e] = o[e];
            }
            pi_bar_arrive(1, PI_T);                          // panel s is factored
        }
        if (w == 8 && bad && lane == 0 && *info == 0) *info = k0 + bad;     // not positive definite (or NaN)
        __syncthreads();
    } else {
        // ---------------- workers ----------------
        // block rows: warps w and w + 4 share a scheduler, so their rows pair up as (0,7) (1,6) (2,5) (3,4)
        const int ra = w < 4 ? w : 11 - w, bx = 7 - ra;
        double ca[8][2], cx[8][2];
        if (upd) {
            double tv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {                      // all sixteen loads in flight before the first store
                const int e = tid + u * 256, i = e & 63, c = e >> 6;
                tv[u] = i < h ? A[(int64_t)(k0 - 64 + c) * ld + k0 + i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = tid + u * 256;
                T[e >> 6][e & 63] = tv[u];
            }
        }
#pragma unroll
        for (int bj = 0; bj < 8; ++bj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = ra * 8 + gid, c = bj * 8 + tig * 2 + e;
                ca[bj][e] = (bj <= ra) ? ((i < h && c < h) ? A[(int64_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0)) : 0.0;
                cx[bj][e] = (bx * 8 + gid == c) ? 1.0 : 0.0;
            }
        __syncthreads();
        if (upd) {
#pragma unroll 2
            for (int kk = 0; kk < 64; kk += 4) {
                const double fa = -T[kk + tig][ra * 8 + gid];
#pragma unroll
                for (int bj = 0; bj < 8; ++bj)
                    if (bj <= ra) sdb_dmma(ca[bj][0], ca[bj][1], fa, T[kk + tig][bj * 8 + gid]);
            }
            pi_bar_sync(3, 256);                                // T is reused for the staged inverse below
        }
        SDB_PI_STAMP(0)
        // hand the four columns of A / rows of X of the panel starting at column jn to the panel warps
        auto extract = [&](int jn) {
            const int bjs = jn >> 3, hi4 = (jn >> 2) & 1;       // block column / row of the panel, and which half of the block
            if (ra >= bjs && (tig >> 1) == hi4) {
#pragma unroll
                for (int bj = 0; bj < 8; ++bj)
                    if (bj == bjs) {
                        Pn[(tig & 1) * 2 + 0][ra * 8 + gid] = ca[bj][0];
                        Pn[(tig & 1) * 2 + 1][ra * 8 + gid] = ca[bj][1];
                    }
            }
            if (bx == bjs && (gid >> 2) == hi4) {
#pragma unroll
                for (int bj = 0; bj < 8; ++bj)
                    if (bj <= bx) {
                        Xn[gid & 3][bj * 8 + tig * 2 + 0] = cx[bj][0];
                        Xn[gid & 3][bj * 8 + tig * 2 + 1] = cx[bj][1];
                    }
            }
        };
        extract(0);
        pi_bar_arrive(2, PI_T);
        for (int s = 0; s < nsteps; ++s) {
            const int j0 = 4 * s;
            double (*Lp)[PI_S] = LpB + 4 * (s & 1), (*Xr)[PI_S] = XrB + 4 * (s & 1);
            SDB_PI_STAMP(1)
            pi_bar_sync(1, PI_T);                               // panel s is factored
            SDB_PI_STAMP(2)
            const bool more = s + 1 < nsteps;
            // rank-4 updates of what is still live: rows / columns beyond the panel of A, rows beyond it of X
            const int bj_lo = (j0 + 4) >> 3, bj_hi = (j0 + 3) >> 3;     // bj_lo is also the block column / row of panel s+1
            const double fa_a = -Lp[tig][ra * 8 + gid], fa_x = -Lp[tig][bx * 8 + gid];
            if (more) {
                // first what panel s+1 needs: block column bj_lo of A and (for its owner) block row bj_lo of X
                if (ra >= bj_lo) {
#pragma unroll
                    for (int bj = 0; bj < 8; ++bj)
                        if (bj == bj_lo) sdb_dmma(ca[bj][0], ca[bj][1], fa_a, Lp[tig][bj * 8 + gid]);
                }
                if (bx == bj_lo) {
#pragma unroll
                    for (int bj = 0; bj < 8; ++bj)
                        if (bj <= bj_hi && bj <= bx) sdb_dmma(cx[bj][0], cx[bj][1], fa_x, Xr[tig][bj * 8 + gid]);
                }
                extract(j0 + 4);
                pi_bar_arrive(2, PI_T);
                SDB_PI_STAMP(3)
                // then the rest, while the panel warps are already at work
                if (ra > bj_lo) {
#pragma unroll
                    for (int bj = 0; bj < 8; ++bj)
                        if (bj > bj_lo && bj <= ra) sdb_dmma(ca[bj][0], ca[bj][1], fa_a, Lp[tig][bj * 8 + gid]);
                }
                if (bx > bj_lo) {
#pragma unroll
                    for (int bj = 0; bj < 8; ++bj)
                        if (bj <= bj_hi && bj <= bx) sdb_dmma(cx[bj][0], cx[bj][1], fa_x, Xr[tig][bj * 8 + gid]);
                }
            }
            {   // the factored panel goes out: L to global memory, the X rows to the staging area (one entry per thread each)
                const int k = tid >> 6, r = tid & 63, j = j0 + k;
                if (r >= j && r < h) A[(int64_t)(k0 + j) * ld + k0 + r] = Lp[k][r];
                Xs[r * PI_XS + j] = r <= j ? Xr[k][r] : 0.0;
            }
            SDB_PI_STAMP(4)
        }
        SDB_PI_PROF_END
        __syncthreads();
    }
    for (int e = tid; e < 64 * 64; e += PI_T) {
        const int i = e & 63, c = e >> 6;
        Xout[e] = (i < h && c <= i) ? Xs[c * PI_XS + i] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------
// The whole backward solve  L^T x = y  in ONE cooperative kernel.  CTA g owns the 64-blocks nblk-1-g, nblk-1-g-G, ...
// (descending); for its block kb it accumulates  s = sum_{jb > kb} L(jb, kb)^T x_jb  as the x_jb are published by
// their owners (release / acquire flags in global memory, L tiles requested before the wait), then
// x_kb = Linv_kb^T (y_kb - s).  The critical path per block is one flag hand-over, a 64 x 64 tile product and a
// 64 x 64 mat-vec instead of a kernel launch.  All CTAs are co-resident (cooperative launch) and every CTA takes its
// blocks in dependency order, so the waits cannot deadlock.  y: row nb + rhs of A (left there by the factorisation).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ns_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ns_st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) chol_bwd_flags_kernel(const double *__restrict__ A, int64_t ld,
                                                             const double *__restrict__ LinvAll, double *Xv, int64_t ldb, int nb,
                                                             int m, int *ready) {
    __shared__ double Ls[64][65];      // Ls[k][o] = Linv(k, o)
    __shared__ double sv[64], vv[64];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nblk = (nb + 63) / 64;
    for (int p = blockIdx.x; p < nblk; p += gridDim.x) {
        const int kb = nblk - 1 - p, k0 = kb * 64;
        const int h = (nb - k0) < 64 ? (nb - k0) : 64;
        const double *Linv = LinvAll + (int64_t)kb * 4096;
        for (int e = t; e < 64 * 64; e += 256) Ls[e & 63][e >> 6] = Linv[e];      // Linv[c*64 + i] -> Ls[i][c]
        const int c0 = warp * 8;                                                   // this warp's 8 columns of the block
        for (int rhs = 0; rhs < m; ++rhs) {
            double *xg = Xv + (int64_t)rhs * ldb;
            double acc[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.0;
            for (int jb = nblk - 1; jb > kb; --jb) {
                const int r0 = jb * 64 + lane, r1 = r0 + 32;
                const bool ok0 = r0 < nb, ok1 = r1 < nb;
                double l0[8], l1[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const bool okc = c0 + cc < h;
                    const double *colp = A + (int64_t)(k0 + c0 + cc) * ld;
                    l0[cc] = (ok0 && okc) ? colp[r0] : 0.0;
                    l1[cc] = (ok1 && okc) ? colp[r1] : 0.0;
                }
                if (lane == 0)
                    while (ns_ld_acquire(ready + jb) <= rhs) { }
                __syncwarp();
                const double x0 = ok0 ? __ldcg(xg + r0) : 0.0, x1 = ok1 ? __ldcg(xg + r1) : 0.0;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(l0[cc], x0, fma(l1[cc], x1, acc[cc]));
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[cc] += __shfl_xor_sync(0xffffffffu, acc[cc], off);
            if (lane == 0) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) sv[c0 + cc] = acc[cc];
            }
            __syncthreads();
            if (t < 64) vv[t] = t < h ? A[(int64_t)(k0 + t) * ld + nb + rhs] - sv[t] : 0.0;
            __syncthreads();
            {
                const int o = t >> 2, part = t & 3;               // four lanes per entry of the block solution
                double s = 0.0;
                for (int k = o + part; k < h; k += 4) s = fma(Ls[k][o], vv[k], s);   // Linv^T(o, k) = Linv(k, o)
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (part == 0 && o < h) { xg[k0 + o] = s; __threadfence(); }
            }
            __syncthreads();
            if (t == 0) ns_st_release(ready + kb, rhs + 1);
        }
        __syncthreads();
    }
}

// right-hand sides (m columns of F, rows q..n-1, after Q^T) -> m extra ROWS below B22, times the definiteness sign:
// the factorisation's panel solve and trailing update then carry out the forward substitution on them for free
__global__ void __launch_bounds__(256) ns_rhs_rows_kernel(const double *__restrict__ F, int64_t ldf, double *__restrict__ A, int64_t ld,
                                                          int q, int nb, int m, double scale) {
    const int64_t total = (int64_t)nb * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / nb), c = (int)(e % nb);
        A[(int64_t)c * ld + nb + k] = scale * F[(int64_t)k * ldf + q + c];
    }
}

// c = R^{-1} (g1 - B12 y2),  lambda = Q [0; y2]  ->  COEFS [(n + q) x m] row-major.  One CTA per right-hand side.
__global__ void __launch_bounds__(NS_T, 2) ns_finish_kernel(const double *__restrict__ S, int64_t ld, int n, int q,
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

// leading dimension of the system in the work buffer: room for the m right-hand-side rows below row n
// even: the column stride of the work matrix is then a multiple of 16 bytes, which a TMA tensor map requires (sdb_syrk_tma_kernel)
static inline int64_t ns_ld(int64_t n, int d, int m) { return (n + (m > d + 1 ? m : d + 1) + 1) & ~(int64_t)1; }

static int ns_enqueue(cudaStream_t st, double *S, int64_t ld, int n, int d, int m, int kernel_type, double *F, double *COEFS,
                      double *vec, int32_t *d_info) {
    const int q = d + 1, nb = n - q;
    double *P = S + (int64_t)n * ld;
    double *tau = vec, *v = vec + 64, *w = v + n, *z = w + n;
    double *xinv_all = z + n;                              // inverse of every 64 x 64 diagonal block
    int *ready = (int *)(xinv_all + (int64_t)(nb / 64 + 1) * 4096);
    const double sgn = (kernel_type == 0 || kernel_type == 6) ? -1.0 : 1.0;
    SDB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    SDB_CUDA(cudaMemsetAsync(ready, 0, sizeof(int) * (nb / 64 + 1), st));
    if (q <= 4) {
        sdb_count_launch();
        ns_house_qr_kernel<<<1, NS_T, 0, st>>>(P, ld, n, q, tau, F, ld, m);
    } else {
        int *qflag = ready + (nb / 64 + 1);
        SDB_CUDA(cudaMemsetAsync(qflag, 0, sizeof(int) * (SDB_MAX_D + 1), st));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(q + m); cfg.blockDim = dim3(NS_T); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeCooperative; at.val.cooperative = 1;     // the flag waits need every CTA resident
        cfg.attrs = &at; cfg.numAttrs = 1;
        sdb_count_launch();
        int64_t ld64 = ld;
        SDB_CUDA(cudaLaunchKernelEx(&cfg, ns_house_qr_grid_kernel, P, ld64, n, q, tau, F, ld64, m, qflag));
    }
    if (q <= 4) {
        // few reflectors (d <= 3): one at a time, three sweeps over the matrix each -- cheaper than the fixed cost of the
        // blocked form's nine launches
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
    } else {   // B = Q^T Phi Q by the compact-WY form: W = Phi V, M = V^T W, G = V^T V, T, Y, then Phi -= V Y^T + Y V^T
        double *VY = (double *)(ready + 2 * (nb / 64 + 2) + 2 * ((SDB_MAX_D + 2) / 2 + 1));
        double *YV = VY + (int64_t)n * 2 * q;
        double *W = YV + (int64_t)n * 2 * q;
        double *Gq = W + (int64_t)n * q, *Mq = Gq + q * q, *Tq = Mq + q * q, *Zq = Tq + q * q;
        double *slab = Zq + q * q;
        const int tiles_m = (n + GM_BM - 1) / GM_BM;
        int nslab_w = (296 + tiles_m - 1) / tiles_m;
        if (nslab_w > 8) nslab_w = 8;
        int nslab_s = n / 512;
        if (nslab_s > 64) nslab_s = 64;
        if (nslab_s < 1) nslab_s = 1;
        sdb_count_launch();
        ns_vmat_kernel<<<592, 256, 0, st>>>(P, ld, n, q, VY, YV);
        int rc = sdb_gemm(st, n, q, n, 1.0, S, 1, ld, VY, 1, n, 0.0, W, n, nslab_w, slab);                 // W = Phi V
        if (rc) return rc;
        rc = sdb_gemm(st, q, q, n, 1.0, VY, n, 1, W, 1, n, 0.0, Mq, q, nslab_s, slab);                     // M = V^T W
        if (rc) return rc;
        rc = sdb_gemm(st, q, q, n, 1.0, VY, n, 1, VY, 1, n, 0.0, Gq, q, nslab_s, slab);                    // G = V^T V
        if (rc) return rc;
        sdb_count_launch();
        ns_wy_small_kernel<<<1, 256, 0, st>>>(q, tau, Gq, Mq, Tq, Zq);
        sdb_count_launch();
        ns_wy_y_kernel<<<592, 256, 0, st>>>(n, q, W, Tq, Zq, VY, YV);
        rc = sdb_gemm(st, n, n, 2 * q, -1.0, VY, 1, n, YV, n, 1, 1.0, S, ld);                                // Phi -= V Y^T + Y V^T
        if (rc) return rc;
        if (sgn != 1.0) {
            sdb_count_launch();
            ns_negate_kernel<<<dim3((nb + 31) / 32, (nb + 31) / 32), 256, 0, st>>>(S, ld, q, n);
        }
    }
    // Cholesky of B22 = S[q:n, q:n] with the right-hand sides as m extra rows (the forward substitution rides on the
    // panel solve and the trailing update).  Right-looking with a look-ahead of one panel: after the triangular solve
    // of panel k the trailing update is split into the 64 columns of panel k+1 (main stream) and the rest (side
    // stream); panel k+1 is factored on the main stream while the rest of update k is still running.
    double *A = S + (int64_t)q * ld + q;
    sdb_count_launch();
    ns_rhs_rows_kernel<<<148, 256, 0, st>>>(F, ld, A, ld, q, nb, m, sgn);
    // three streams: st carries the critical path  POTRF+inverse(k) -> panel solve(k) -> POTRF+inverse(k+1) ...;
    // `nxt` updates the columns of panel k+1 below its diagonal block, `aux` the rest of the trailing matrix.
    static cudaStream_t aux = nullptr, nxs = nullptr;
    static cudaEvent_t ev_trsm[4], ev_rest[4], ev_next[4];
    if (!aux) {
        // the bulk updates yield to the critical path when SMs are scarce (a refit confined to the few SMs an asynchronous
        // collector gets beside the chain kernel): lowest priority for `aux`, highest for the next-panel update
        int prio_lo = 0, prio_hi = 0;
        SDB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        SDB_CUDA(cudaStreamCreateWithPriority(&aux, cudaStreamNonBlocking, prio_lo));
        SDB_CUDA(cudaStreamCreateWithPriority(&nxs, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < 4; ++i) {
            SDB_CUDA(cudaEventCreateWithFlags(&ev_trsm[i], cudaEventDisableTiming));
            SDB_CUDA(cudaEventCreateWithFlags(&ev_rest[i], cudaEventDisableTiming));
            SDB_CUDA(cudaEventCreateWithFlags(&ev_next[i], cudaEventDisableTiming));
        }
    }
    // Who writes what after the panel solve of panel k (T = solved panel, T1 = its first `nxt` rows):
    //   diagonal block of panel k+1      -= T1 T1^T         prologue of POTRF+inverse(k+1)           (st)
    //   rest of the columns of panel k+1 -= T[nxt:] T1^T    `nxs`, awaited by the panel solve of k+1
    //   columns of panels >= k+2         -= lower(T' T'^T)  `aux`, awaited by POTRF+inverse(k+2) and by nxs(k+1)
    // nb >= SDB_CHOL_PAIR_MIN (default 3072; no gain measured below): the `aux` update is PAIRED.  After an even panel k only
    // the 64 columns of panel k+2 receive its contribution (rank 64); everything beyond them waits for the odd panel k+1 and
    // then receives both contributions in ONE rank-128 pass, which reads and writes each C tile once per 128 columns of L
    // instead of once per 64 -- the update is bound by that traffic and by its per-tile prologue, not by the DMMA pipe.
    static const int pair_min = getenv("SDB_CHOL_PAIR_MIN") ? atoi(getenv("SDB_CHOL_PAIR_MIN")) : 3072;
    const bool paired = nb >= pair_min;
    // group size: G consecutive panels share one rank-64G pass over everything beyond the group (G = 2 is the pairing just
    // described).  After panel i of a group only the `window` -- the columns of the panels up to the first one of the next
    // group -- receives its contribution at rank 64; the C tiles of the bulk are read and written once per 64 G columns of L,
    // and a tile's exposed load / store is amortised over G x 64 k-steps.  Measured (d = 4, m = 3, whole fit, ms):
    //   n = 4096 (d = 2, m = 1): G = 1 / 2 / 3 -> 2.88 / 2.70 / 2.75, (d = 4, m = 3) 2 / 4 -> 2.96 / 3.07;
    //   8192: G = 2 / 4 / 8 -> 12.1 / 11.6 / 11.4;  16384: 69.8 / 62.8 / 61.1;  32768: 487 / 427 / 412.
    static const int group_env = getenv("SDB_CHOL_GROUP") ? atoi(getenv("SDB_CHOL_GROUP")) : 0;
    const int G = group_env > 1 ? group_env : (nb >= 12288 ? 8 : nb >= 6144 ? 4 : 2);
    // trailing matrices of at least this many columns use the TMA-fed 128 x 128 ring kernel (fewer leave SMs without a tile)
    static const int big_min = getenv("SDB_CHOL_BIG_MIN") ? atoi(getenv("SDB_CHOL_BIG_MIN")) : 1024;
    int last_rest = -1, prev_rest = -1, last_next = -1;      // last panels after whose solve an aux / nxs update was issued
    int kp = 0;
    for (int k0 = 0; k0 < nb; k0 += 64, ++kp) {
        const int h = (nb - k0) < 64 ? (nb - k0) : 64;
        const int rest = nb - k0 - h, restx = rest + m;     // rows below the diagonal block, with the right-hand-side rows
        double *xinv = xinv_all + (int64_t)kp * 4096;
        {   // aux updates issued after the solve of a panel <= kp-2 wrote this panel's columns
            const int q = (last_rest <= kp - 2) ? last_rest : prev_rest;
            if (q >= 0) SDB_CUDA(cudaStreamWaitEvent(st, ev_rest[q & 3], 0));
        }
        sdb_count_launch();
        chol_potrf_inv64_kernel<<<1, PI_T, 0, st>>>(A, ld, k0, h, xinv, d_info, kp > 0 ? 1 : 0 SDB_PI_PROF_PASS);
        if (last_next == kp - 1 && kp > 0) SDB_CUDA(cudaStreamWaitEvent(st, ev_next[(kp - 1) & 3], 0));
        double *L21 = A + (int64_t)k0 * ld + k0 + h;
        // L21 = A21 L11^{-T}:  (A21 X^T)(r, j) = sum_k A21(r, k) X(j, k),  X(j, k) = xinv[k*64 + j]
        int rc = sdb_gemm_k64_inplace(st, restx, h, L21, ld, xinv, 64);
        if (rc) return rc;
        if (rest > 0) {
            SDB_CUDA(cudaEventRecord(ev_trsm[kp & 3], st));
            double *A22 = A + (int64_t)(k0 + h) * ld + k0 + h;
            const int nxt = rest < 64 ? rest : 64;      // columns of the next panel = rows of its diagonal block
            SDB_CUDA(cudaStreamWaitEvent(nxs, ev_trsm[kp & 3], 0));
            if (last_rest >= 0) SDB_CUDA(cudaStreamWaitEvent(nxs, ev_rest[last_rest & 3], 0));   // earlier aux updates wrote these columns too
            rc = sdb_gemm_k64_skinny_update(nxs, restx - nxt, nxt, h, L21 + nxt, L21, A22 + nxt, ld, ld);
            if (rc) return rc;
            SDB_CUDA(cudaEventRecord(ev_next[kp & 3], nxs));
            last_next = kp;
            bool issued = false;
            if (!paired) {
                if (rest > 64) {
                    SDB_CUDA(cudaStreamWaitEvent(aux, ev_trsm[kp & 3], 0));
                    rc = sdb_syrk_k64_update(aux, restx - 64, rest - 64, h, L21 + 64, L21 + 64, A22 + (int64_t)64 * ld + 64, ld);
                    if (rc) return rc;
                    issued = true;
                }
            } else if (kp % G != G - 1) {
                if (rest > 64) {        // window: the columns of the panels kp+2 .. first panel of the next group
                    const int nw = 64 * (G - 1 - kp % G), n2 = (rest - 64) < nw ? (rest - 64) : nw;
                    SDB_CUDA(cudaStreamWaitEvent(aux, ev_trsm[kp & 3], 0));
                    rc = sdb_syrk_k64_update(aux, restx - 64, n2, h, L21 + 64, L21 + 64, A22 + (int64_t)64 * ld + 64, ld);
                    if (rc) return rc;
                    issued = true;
                }
            } else {
                if (rest > 64) {        // the G panels of the group together: everything beyond panel kp+1
                    const int kb = 64 * (G - 1);
                    const double *Lp = A + (int64_t)(k0 - kb) * ld + (k0 + 128);      // rows from k0+128, columns k0-kb .. k0+63
                    SDB_CUDA(cudaStreamWaitEvent(aux, ev_trsm[kp & 3], 0));
                    if (rest - 64 >= big_min || kb + h > 256) rc = sdb_syrk_big_update(aux, restx - 64, rest - 64, kb + h, Lp, Lp, A22 + (int64_t)64 * ld + 64, ld, S, (int64_t)n + (d + 1) + m);
                    else rc = sdb_syrk_k64_update(aux, restx - 64, rest - 64, kb + h, Lp, Lp, A22 + (int64_t)64 * ld + 64, ld);
                    if (rc) return rc;
                    issued = true;
                }
            }
            if (issued) {
                SDB_CUDA(cudaEventRecord(ev_rest[kp & 3], aux));
                prev_rest = last_rest;
                last_rest = kp;
            }
        }
    }
    // join the side streams (the last panel solve already waited for the last nxs update)
    if (last_rest >= 0) SDB_CUDA(cudaStreamWaitEvent(st, ev_rest[last_rest & 3], 0));
    if (last_next >= 0) SDB_CUDA(cudaStreamWaitEvent(st, ev_next[last_next & 3], 0));
    // L^T x = y (y = rows nb.. of A), x -> F[q:n]
    {
        static int per_sm = 0;
        if (!per_sm) SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_bwd_flags_kernel, 256, 0));
        sdb_device_props dp;
        int rc = sdb_get_device_props(&dp);
        if (rc) return rc;
        const int nblk = (nb + 63) / 64;
        int grid = dp.sm_count * (per_sm > 2 ? 2 : per_sm);
        if (grid > nblk) grid = nblk;
        // a cooperative grid starts only when ALL its CTAs fit at once: 64 CTAs could not start on the SMs an asynchronous
        // collector has beside the chain kernel and the round waited for the chain block to end.  16 CTAs carry the same
        // critical path (measured, whole round on 148 SMs: 3.75 ms with 64 or 32 CTAs, 3.77 with 16, 3.85 with 8).
        // With a rotating solve duty on several GPUs the collector gets as few as 6 SMs (two CTAs of this kernel each): 8 CTAs up
        // to 64 blocks (n <= 4096), 16 beyond.
        static const int bwd_env = getenv("SDB_BWD_GRID") ? atoi(getenv("SDB_BWD_GRID")) : 0;
        const int bwd_cap = bwd_env > 0 ? bwd_env : (nblk <= 64 ? 8 : 16);
        if (grid > bwd_cap) grid = bwd_cap;
        if (grid < 1) { sdb_set_error("sdb_rbf_fit_nullspace: the backward-solve kernel does not fit on the device"); return SDB_ERR_CUDA; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeCooperative; at.val.cooperative = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        sdb_count_launch();
        SDB_CUDA(cudaLaunchKernelEx(&cfg, chol_bwd_flags_kernel, (const double *)A, ld, (const double *)xinv_all, F + q, ld, nb, m, ready));
    }
    sdb_count_launch();
    ns_finish_kernel<<<m, NS_T, 0, st>>>(S, ld, n, q, tau, F, ld, m, COEFS);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

extern "C" int64_t sdb_rbf_fit_nullspace_work(int64_t n, int32_t d, int32_t m) {
    const int64_t N = n + d + 1, ld = ns_ld(n, d, m), q = d + 1;
    // system (N columns of ld) | right-hand sides | tau, v, w, z | inverses of the diagonal blocks | block flags |
    // WY panels [V | Y], [Y | V], W | G, M, T, Z | split-K slabs of the tall-skinny products
    return ld * N + ld * m + 3 * n + 128 + (n / 64 + 2) * 4096 + (n / 64 + 4) + 8 + SDB_MAX_D + 8 + 5 * n * q + 4 * q * q + 8 * n * q + 64 * q * q;
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
    const int64_t N = n + d + 1, ld = ns_ld(n, d, m);
    double *S = d_work;
    double *F = S + ld * N;
    double *vec = F + ld * m;
    cudaStream_t st = (cudaStream_t)stream;
    auto enqueue = [&]() -> int {
        int rc = sdb_rbf_system_ld(d_X, d_Y, n, d, m, kernel_type, S, F, ld, stream);
        if (rc) return rc;
        return ns_enqueue(st, S, ld, (int)n, d, m, kernel_type, F, d_COEFS, vec, d_info);
    };
    if (n < 256) return enqueue();
    sdb_graph_key key = {{d_X, d_Y, d_COEFS, d_work, d_info}, {n, d, m, kernel_type}, 2, 0};
    return sdb_graph_run(st, key, enqueue);
}
