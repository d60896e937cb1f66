// Surrogate update kernels (the collector's refit), FP64, sm_100a.
//   sdb_poly_basis / sdb_poly_fit   replace surrogate_poly.py:68-97  (PHI, normal equations, pinv solve)
//   sdb_rbf_thin                    replaces surrogate_rbf.py:97-120 (greedy thinning to no_keep)
//   sdb_rbf_system / sdb_rbf_fit    replace surrogate_rbf.py:91-96,121-130 (distance + kernel matrix,
//                                   saddle-point system, direct solve through sdb_lu_solve)
#include <float.h>

#include "sdb_gemm.cuh"
#include "sdb_models.cuh"

// ===================================================================================================
// polynomial surrogate
// ===================================================================================================
template <int D>
__global__ void __launch_bounds__(128) poly_basis_kernel(const double *__restrict__ pts, int64_t N, int d, const sdb_model mo,
                                                         double *__restrict__ PHI) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int HS = SDB_MAX_DEGREE + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *p = smem + 16;
    if (threadIdx.x == 0) {
        sdb_mbar_init(bar, 1);
        sdb_mbar_fence_init();
    }
    __syncthreads();
    // only alpha and herm are needed; stage them like the evaluation kernels do (coefs may be NULL)
    const size_t hb = (size_t)(mo.degree + 1) * (mo.degree + 1) * 8, ab = (size_t)mo.n_terms * d * 4;
    const double *herm = (const double *)p; sdb_stage_table(p, mo.herm, hb, bar); p += sdb_align16(hb);
    const int *alpha = (const int *)p;      sdb_stage_table(p, mo.alpha, ab, bar); p += sdb_align16(ab);
    if (threadIdx.x == 0) sdb_mbar_arrive(bar);
    sdb_mbar_wait(bar, 0);
    __syncthreads();
    const int deg = mo.degree, nh = deg + 1, P = mo.n_terms;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double h[DD * HS];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) {
                const double x = pts[i * d + j];
                double pw[HS];
                pw[0] = 1.0;
                for (int q = 1; q <= deg; ++q) pw[q] = __dmul_rn(pw[q - 1], x);
                for (int k = 0; k <= deg; ++k) {
                    double v = herm[k * nh];
                    for (int q = 1; q <= k; ++q) v = __dadd_rn(v, __dmul_rn(herm[k * nh + q], pw[q]));
                    h[j * HS + k] = v;
                }
            }
        for (int t = 0; t < P; ++t) {
            const int *a = alpha + t * d;
            double phi = h[a[0]];
#pragma unroll UD
            for (int j = 1; j < DD; ++j)
                if (j < d) phi = __dmul_rn(phi, h[j * HS + a[j]]);
            PHI[i * P + t] = phi;
        }
    }
}

extern "C" int sdb_poly_basis(const double *d_points, int64_t N, int32_t d, const sdb_model *model, double *d_PHI,
                              void *stream) {
    SDB_CHECK_ARG(model && d_points && d_PHI && model->alpha && model->herm, "sdb_poly_basis: NULL pointer");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D, "sdb_poly_basis: d=%d outside limits", d);
    SDB_CHECK_LIMIT(model->degree >= 0 && model->degree <= SDB_MAX_DEGREE, "sdb_poly_basis: degree %d outside [0,%d]",
                    model->degree, SDB_MAX_DEGREE);
    if (N <= 0) return SDB_OK;
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    const size_t smem = 16 + sdb_align16((size_t)(model->degree + 1) * (model->degree + 1) * 8) +
                        sdb_align16((size_t)model->n_terms * d * 4);
    SDB_CHECK_LIMIT(smem <= (size_t)dp.smem_optin, "sdb_poly_basis: tables need %zu B of shared memory", smem);
    int64_t blocks = (N + 127) / 128;
    if (blocks > (int64_t)dp.sm_count * 16) blocks = (int64_t)dp.sm_count * 16;
    auto launch = [&](auto kern) -> int {
        SDB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sdb_count_launch();
        kern<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(d_points, N, d, *model, d_PHI);
        SDB_CUDA(cudaGetLastError());
        return SDB_OK;
    };
    if (d == 2) return launch(poly_basis_kernel<2>);
    if (d == 4) return launch(poly_basis_kernel<4>);
    return launch(poly_basis_kernel<0>);
}

// ---------------------------------------------------------------------------------------------------
// MAT = pinv(A) RHS for the symmetric P x P matrix A (PHI^T PHI) by one-sided Jacobi (Hestenes):
// columns of G = A are rotated pairwise until mutually orthogonal; then G = U Sigma, and for a
// symmetric A the right singular vectors are v_j = sign(u_j^T A u_j) u_j, so
//   pinv(A) RHS = sum_{sigma_j > cutoff} sign_j / sigma_j * u_j (u_j^T RHS),  cutoff = 1e-15 sigma_max
// (numpy.linalg.pinv's default rcond).  One CTA; G lives in shared memory when it fits, else in
// global memory.  Round-robin ordering: P-1 (P even) rounds of P/2 disjoint pairs, one warp per pair.
// info[0] = number of directions kept, info[1] = sweeps, info[2] = 1 if converged.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) poly_pinv_solve_kernel(const double *__restrict__ A, const double *__restrict__ RHS,
                                                                  int P, int m, double *__restrict__ MAT,
                                                                  double *__restrict__ Gglob, int use_smem, int *info) {
    extern __shared__ __align__(16) double sm[];
    double *G = use_smem ? sm : Gglob;         // column-major P x P, column j at G + j*P
    double *sig = use_smem ? sm + (size_t)P * P : Gglob + (size_t)P * P;  // [P] singular values, then signs
    __shared__ int rotated;
    __shared__ double s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int e = tid; e < P * P; e += blockDim.x) G[e] = A[e];
    __syncthreads();
    const int Pe = (P + 1) & ~1;  // players (one dummy when P is odd)
    const double tol = 1e-15;
    int sweeps = 0, converged = 0;
    for (; sweeps < 60 && !converged; ++sweeps) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int round = 0; round < Pe - 1; ++round) {
            for (int pair = warp; pair < Pe / 2; pair += nwarp) {
                // circle method: player Pe-1 fixed, the others rotate
                int a = (pair == 0) ? Pe - 1 : (round + pair) % (Pe - 1);
                int b = (round + Pe - 1 - pair) % (Pe - 1);
                if (a >= P || b >= P) continue;
                if (a > b) { const int t = a; a = b; b = t; }
                double *ga = G + (size_t)a * P, *gb = G + (size_t)b * P;
                double aa = 0.0, bb = 0.0, ab = 0.0;
                for (int i = lane; i < P; i += 32) {
                    const double x = ga[i], y = gb[i];
                    aa = fma(x, x, aa); bb = fma(y, y, bb); ab = fma(x, y, ab);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    aa += __shfl_xor_sync(0xffffffffu, aa, off);
                    bb += __shfl_xor_sync(0xffffffffu, bb, off);
                    ab += __shfl_xor_sync(0xffffffffu, ab, off);
                }
                if (fabs(ab) > tol * sqrt(aa * bb) && ab != 0.0) {
                    const double zeta = (bb - aa) / (2.0 * ab);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int i = lane; i < P; i += 32) {
                        const double x = ga[i], y = gb[i];
                        ga[i] = c * x - s * y;
                        gb[i] = s * x + c * y;
                    }
                    if (lane == 0) rotated = 1;
                }
            }
            __syncthreads();
        }
        converged = !rotated;
        __syncthreads();
    }
    // singular values and sign of the eigenvalue
    if (tid == 0) s_max = 0.0;
    __syncthreads();
    for (int j = warp; j < P; j += nwarp) {
        const double *g = G + (size_t)j * P;
        double nn = 0.0;
        for (int i = lane; i < P; i += 32) nn = fma(g[i], g[i], nn);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
        // u^T A u has the sign of g^T A g
        double q = 0.0;
        for (int i = lane; i < P; i += 32) {
            double r = 0.0;
            for (int k = 0; k < P; ++k) r = fma(A[(size_t)k * P + i], g[k], r);
            q = fma(g[i], r, q);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
        if (lane == 0) sig[j] = copysign(sqrt(nn), q);
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < P; ++j) mx = fmax(mx, fabs(sig[j]));
        s_max = mx;
        int kept = 0;
        for (int j = 0; j < P; ++j) kept += fabs(sig[j]) > 1e-15 * mx;
        info[0] = kept; info[1] = sweeps; info[2] = converged;
    }
    __syncthreads();
    const double cutoff = 1e-15 * s_max;
    // MAT[:, k] = sum_j sign_j / sigma_j^3 * g_j (g_j^T RHS[:, k])   (u_j = g_j / sigma_j)
    for (int e = tid; e < P * m; e += blockDim.x) MAT[e] = 0.0;
    __syncthreads();
    // one warp per observation column; directions summed in index order (deterministic)
    for (int k = warp; k < m; k += nwarp) {
        for (int j = 0; j < P; ++j) {
            const double sj = fabs(sig[j]);
            if (!(sj > cutoff)) continue;
            const double *g = G + (size_t)j * P;
            double dot = 0.0;
            for (int i = lane; i < P; i += 32) dot = fma(g[i], RHS[(size_t)i * m + k], dot);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
            const double w = copysign(1.0, sig[j]) * ((dot / sj) / sj) / sj;
            for (int i = lane; i < P; i += 32) MAT[(size_t)i * m + k] = fma(w, g[i], MAT[(size_t)i * m + k]);
        }
    }
}

static int poly_gram_slabs(int64_t n, int P) {
    // enough K-slabs to fill the GPU for small P; bounded workspace
    const int tiles = ((P + GM_BM - 1) / GM_BM) * ((P + GM_BN - 1) / GM_BN);
    int64_t s = (148 * 4 + tiles - 1) / tiles;
    const int64_t by_k = (n + 255) / 256;
    if (s > by_k) s = by_k;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return (int)s;
}

extern "C" int64_t sdb_poly_fit_work(int64_t n, int32_t P, int32_t m) {
    const int64_t slabs = poly_gram_slabs(n, P);
    // PHI [n x P] | A [P x P] | RHS [P x m] | G + sigma (P*P + P) | slab workspace for the larger of the two products
    const int64_t cmax = (int64_t)P * (P > m ? P : m);
    return n * P + (int64_t)P * P + (int64_t)P * m + (int64_t)P * P + P + slabs * cmax + 64;
}

extern "C" int sdb_poly_fit(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, const sdb_model *model,
                            double *d_MAT, double *d_work, int32_t *d_info, void *stream) {
    SDB_CHECK_ARG(d_X && d_Y && model && d_MAT && d_work && d_info, "sdb_poly_fit: NULL pointer");
    SDB_CHECK_ARG(n >= 1 && n < ((int64_t)1 << 31), "sdb_poly_fit: n = %lld out of range", (long long)n);
    SDB_CHECK_LIMIT(m >= 1 && m <= SDB_MAX_M, "sdb_poly_fit: m=%d outside limits", m);
    const int P = model->n_terms;
    SDB_CHECK_ARG(P >= 1, "sdb_poly_fit: n_terms must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    double *PHI = d_work;
    double *A = PHI + n * P;
    double *RHS = A + (int64_t)P * P;
    double *Gg = RHS + (int64_t)P * m;
    double *slab = Gg + (int64_t)P * P + P;
    rc = sdb_poly_basis(d_X, n, d, model, PHI, stream);
    if (rc) return rc;
    const int slabs = poly_gram_slabs(n, P);
    // A = PHI^T PHI  (column-major P x P; symmetric).  PHI is row-major n x P: PHI^T(i,k) = PHI[k*P+i].
    rc = sdb_gemm(st, P, P, (int)n, 1.0, PHI, 1, P, PHI, P, 1, 0.0, A, P, slabs, slab);
    if (rc) return rc;
    // RHS = PHI^T Y, stored row-major [P x m] == column-major m x P: compute its transpose Y^T PHI
    // (M = m, N = P): A-operand Y^T(i,k) = Y[k*m+i], B-operand PHI(k,j) = PHI[k*P+j].
    rc = sdb_gemm(st, m, P, (int)n, 1.0, d_Y, 1, m, PHI, P, 1, 0.0, RHS, m, slabs, slab);
    if (rc) return rc;
    const size_t need = ((size_t)P * P + P) * 8;
    const int use_smem = need + 1024 <= (size_t)dp.smem_optin;
    const size_t smem = use_smem ? need : 0;
    SDB_CUDA(cudaFuncSetAttribute((const void *)poly_pinv_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(use_smem ? need : 0)));
    sdb_count_launch();
    poly_pinv_solve_kernel<<<1, 1024, smem, st>>>(A, RHS, P, m, d_MAT, Gg, use_smem, d_info);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ===================================================================================================
// RBF surrogate
// ===================================================================================================
// distance exactly as the reference forms it: sum_k (x_jk - x_ik)^2 with separately rounded
// square and add (np.power(Q-T, 2) accumulated k = 0..d-1), then an IEEE square root.
__device__ __forceinline__ double ref_sumsq(const double *__restrict__ xi, const double *__restrict__ xj, int d) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
        const double df = __dadd_rn(xj[k], -xi[k]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}
__device__ __forceinline__ double ref_distance(const double *__restrict__ xi, const double *__restrict__ xj, int d) {
    return __dsqrt_rn(ref_sumsq(xi, xj, d));
}

// Running "smallest distance, ties -> the earlier index" over a sequence visited in ascending index order, WITHOUT a square
// root per candidate: squared distances are compared, and only when a new one is smaller by less than 1e-15 (relative) --
// i.e. when the two correctly rounded roots could coincide -- are the roots themselves compared.  sqrt is monotone, so
// the winner is exactly the one the reference's arg-min over rounded distances (surrogate_rbf.py:98-103) picks.
// The FP64 pipe of one SM is what bounds a rescan: 5 instructions per candidate instead of ~20 with the root.
struct sq_best {
    double sq;
    int j;
    __device__ __forceinline__ void init() { sq = DBL_MAX; j = INT_MAX; }
    __device__ __forceinline__ void offer(double s, int idx) {
        if (s < sq) {       // rarely true after the first few candidates: the band test costs nothing on the common path
            if (s >= sq * (1.0 - 1e-15) && !(__dsqrt_rn(s) < __dsqrt_rn(sq))) return;   // same rounded distance: the earlier index stays
            sq = s; j = idx;
        }
    }
    __device__ __forceinline__ double dist() const { return j == INT_MAX ? DBL_MAX : __dsqrt_rn(sq); }
};

__device__ __forceinline__ double ref_phi(double r, int kt) {
    switch (kt) {
        case 1: return r * r * r;
        case 2: { const double r2 = r * r; return r2 * r2 * r; }
        case 3: { const double r2 = r * r; return r2 * r2 * r2 * r; }
        case 7: { const double r2 = r * r, r4 = r2 * r2; return r4 * r4 * r; }
        case 4: return exp(-(r * r));
        case 5: return 1.0 / sqrt(r * r + 1.0);
        default: return r;  // 0, and 6 (a no-op in the reference)
    }
}

// S [(n+d+1) columns of ld >= n+d+1] (symmetric, so row- and column-major coincide) and RHS as m columns of ld.
__global__ void __launch_bounds__(256) rbf_system_kernel(const double *__restrict__ X, const double *__restrict__ Y, int n, int d,
                                                         int m, int kt, double *__restrict__ S, double *__restrict__ RHS,
                                                         int64_t ld) {
    const int N = n + d + 1;
    const int i = blockIdx.x * 16 + (threadIdx.x & 15);  // fast index: contiguous in memory
    const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (i < N && j < N) {
        double v;
        if (i < n && j < n) v = ref_phi(ref_distance(X + (size_t)i * d, X + (size_t)j * d, d), kt);
        else if (i < n) v = (j - n < d) ? X[(size_t)i * d + (j - n)] : 1.0;
        else if (j < n) v = (i - n < d) ? X[(size_t)j * d + (i - n)] : 1.0;
        else v = 0.0;
        S[(size_t)j * ld + i] = v;
    }
    if (RHS && blockIdx.y == 0 && (threadIdx.x >> 4) == 0 && i < N)
        for (int k = 0; k < m; ++k) RHS[(size_t)k * ld + i] = i < n ? Y[(size_t)i * m + k] : 0.0;
}

// The same for larger d: a 32 x 32 tile of entries per CTA, the 2 x 32 rows of X it needs staged (transposed) in shared
// memory, every thread a 2 x 2 micro-tile -- per coordinate four loads from shared memory feed four differences instead of
// eight loads from global memory.  Same per-entry arithmetic and order as rbf_system_kernel.
__global__ void __launch_bounds__(256) rbf_system_tiled_kernel(const double *__restrict__ X, const double *__restrict__ Y, int n, int d,
                                                               int m, int kt, double *__restrict__ S, double *__restrict__ RHS,
                                                               int64_t ld) {
    __shared__ double Xi[SDB_MAX_D][33], Xj[SDB_MAX_D][33];
    const int N = n + d + 1;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int e = tid; e < 32 * d; e += 256) {
        const int r = e / d, k = e % d;
        Xi[k][r] = (i0 + r < n) ? X[(size_t)(i0 + r) * d + k] : 0.0;
        Xj[k][r] = (j0 + r < n) ? X[(size_t)(j0 + r) * d + k] : 0.0;
    }
    __syncthreads();
    double s[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int k = 0; k < d; ++k) {
        const double a0 = Xi[k][tx], a1 = Xi[k][tx + 16], b0 = Xj[k][ty], b1 = Xj[k][ty + 16];
        const double d00 = __dadd_rn(b0, -a0), d10 = __dadd_rn(b0, -a1), d01 = __dadd_rn(b1, -a0), d11 = __dadd_rn(b1, -a1);
        s[0][0] = __dadd_rn(s[0][0], __dmul_rn(d00, d00));
        s[1][0] = __dadd_rn(s[1][0], __dmul_rn(d10, d10));
        s[0][1] = __dadd_rn(s[0][1], __dmul_rn(d01, d01));
        s[1][1] = __dadd_rn(s[1][1], __dmul_rn(d11, d11));
    }
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
            if (i < N && j < N) {
                double v;
                if (i < n && j < n) v = ref_phi(__dsqrt_rn(s[a][b]), kt);
                else if (i < n) v = (j - n < d) ? X[(size_t)i * d + (j - n)] : 1.0;
                else if (j < n) v = (i - n < d) ? X[(size_t)j * d + (i - n)] : 1.0;
                else v = 0.0;
                S[(size_t)j * ld + i] = v;
            }
        }
    if (RHS && blockIdx.y == 0 && ty == 0)
        for (int a = 0; a < 2; ++a) {
            const int i = i0 + tx + 16 * a;
            if (i < N)
                for (int k = 0; k < m; ++k) RHS[(size_t)k * ld + i] = i < n ? Y[(size_t)i * m + k] : 0.0;
        }
}

// the system with a leading dimension ld >= n + d + 1 (rows beyond n + d + 1 are left untouched)
int sdb_rbf_system_ld(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type, double *d_S,
                      double *d_RHS, int64_t ld, void *stream) {
    SDB_CHECK_ARG(d_X && d_S && (d_Y || !d_RHS), "sdb_rbf_system: NULL pointer");
    SDB_CHECK_ARG(n >= 1 && n < (1 << 30), "sdb_rbf_system: n out of range");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D && m >= 1 && m <= SDB_MAX_M, "sdb_rbf_system: d=%d m=%d outside limits", d, m);
    SDB_CHECK_ARG(kernel_type >= 0 && kernel_type <= 7, "sdb_rbf_system: kernel_type %d outside [0,7]", kernel_type);
    const int N = (int)n + d + 1;
    SDB_CHECK_ARG(ld >= N, "sdb_rbf_system: leading dimension below n + d + 1");
    sdb_count_launch();
    if (d > 4) {
        dim3 grid((N + 31) / 32, (N + 31) / 32);
        rbf_system_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, d_Y, (int)n, d, m, kernel_type, d_S, d_RHS, ld);
    } else {
        dim3 grid((N + 15) / 16, (N + 15) / 16);
        rbf_system_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, d_Y, (int)n, d, m, kernel_type, d_S, d_RHS, ld);
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

extern "C" int sdb_rbf_system(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                              double *d_S, double *d_RHS, void *stream) {
    return sdb_rbf_system_ld(d_X, d_Y, n, d, m, kernel_type, d_S, d_RHS, n + d + 1, stream);
}

__global__ void __launch_bounds__(256) transpose_out_kernel(const double *__restrict__ Xc, int N, int m, double *__restrict__ out) {
    // Xc: m columns of length N  ->  out [N x m] row-major
    const int64_t total = (int64_t)N * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / m), k = (int)(e % m);
        out[e] = Xc[(size_t)k * N + i];
    }
}

extern "C" int64_t sdb_rbf_fit_work(int64_t n, int32_t d, int32_t m) {
    const int64_t N = n + d + 1;
    return N * N + N * m + N / 2 + 64;  // system | right-hand sides | pivots (int32)
}

extern "C" int sdb_rbf_fit(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                           double *d_COEFS, double *d_work, int32_t *d_info, void *stream) {
    SDB_CHECK_ARG(d_X && d_Y && d_COEFS && d_work && d_info, "sdb_rbf_fit: NULL pointer");
    const int64_t N = n + d + 1;
    double *S = d_work;
    double *RHS = S + N * N;
    int32_t *piv = (int32_t *)(RHS + N * m);
    int rc = sdb_rbf_system(d_X, d_Y, n, d, m, kernel_type, S, RHS, stream);
    if (rc) return rc;
    rc = sdb_lu_solve(S, N, RHS, m, piv, d_info, stream);
    if (rc) return rc;
    int64_t blocks = (N * m + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    sdb_count_launch();
    transpose_out_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(RHS, (int)N, m, d_COEFS);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Greedy thinning (surrogate_rbf.py:97-120): repeatedly find the globally closest pair -- the first
// flat arg-min of the symmetric distance matrix, i.e. the smallest row index among the minima, then the
// smallest column index -- and drop the point owning that ROW, until no_keep points remain.
// The n x n matrix is never stored: each point keeps (distance, index) of its nearest alive neighbour
// (ties -> smallest index); removing a point only invalidates the rows that pointed at it.
//   kernel 1 (grid): nearest neighbour of every point.   kernel 2 (one CTA): the removal loop.
// work: nn_dist [n] doubles | nn_idx [n] int32
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) thin_nearest_kernel(const double *__restrict__ X, int n, int d, double *__restrict__ nn_dist,
                                                           int *__restrict__ nn_idx, const uint8_t *__restrict__ alive) {
    // one warp per row; rows that are not alive on entry (alive != NULL) neither get nor are a neighbour
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    if (alive && !alive[row]) {
        if (lane == 0) { nn_dist[row] = DBL_MAX; nn_idx[row] = INT_MAX; }
        return;
    }
    sq_best sb;
    sb.init();
    for (int j = lane; j < n; j += 32) {
        if (j == row || (alive && !alive[j])) continue;
        sb.offer(ref_sumsq(X + (size_t)row * d, X + (size_t)j * d, d), j);   // j ascending per lane: first minimum kept
    }
    double best = sb.dist();
    int bj = sb.j;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
    }
    if (lane == 0) { nn_dist[row] = best; nn_idx[row] = bj; }
}

// (distance, index) arg-min over the block, ties -> smallest index; result broadcast to all threads.
__device__ __forceinline__ void block_argmin(double &best, int &bidx, double *rv, int *rr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
        if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    __syncthreads();  // previous readers of rv / rr are done
    if (lane == 0) { rv[warp] = best; rr[warp] = bidx; }
    __syncthreads();
    best = rv[lane]; bidx = rr[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
        if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
}

// The removal loop, one CTA.  When they fit, the points and the three per-point arrays are staged in
// shared memory (use_smem), so an iteration costs a handful of barriers instead of L2 round trips.
__global__ void __launch_bounds__(1024, 1) thin_remove_kernel(const double *__restrict__ Xg, int n, int d, int no_keep,
                                                              double *__restrict__ nn_dist_g, int *__restrict__ nn_idx_g,
                                                              uint8_t *__restrict__ keep_g, int use_smem,
                                                              const uint8_t *__restrict__ alive_g, int dedupe) {
    extern __shared__ __align__(16) unsigned char thin_sm[];
    __shared__ double rv[32];
    __shared__ int rr[32];
    __shared__ int todo[1024];
    __shared__ int ntodo;
    __shared__ int nalive;
    const int tid = threadIdx.x;
    const double *X = Xg;
    double *nn_dist = nn_dist_g;
    int *nn_idx = nn_idx_g;
    uint8_t *keep = keep_g;
    if (use_smem) {
        double *Xs = (double *)thin_sm;
        nn_dist = Xs + (size_t)n * d;
        nn_idx = (int *)(nn_dist + n);
        keep = (uint8_t *)(nn_idx + n);
        for (int e = tid; e < n * d; e += blockDim.x) Xs[e] = Xg[e];
        for (int i = tid; i < n; i += blockDim.x) { nn_dist[i] = nn_dist_g[i]; nn_idx[i] = nn_idx_g[i]; }
        X = Xs;
    }
    if (tid == 0) { ntodo = 0; nalive = 0; }
    __syncthreads();
    {
        int mine = 0;
        for (int i = tid; i < n; i += blockDim.x) {
            const uint8_t a = alive_g ? (alive_g[i] != 0) : 1;
            keep[i] = a;
            mine += a;
        }
        if (mine) atomicAdd(&nalive, mine);
    }
    __syncthreads();
    // rows dead on entry do not count (the collector's fixed-shape candidate block); dedupe: past the quota the loop goes on
    // while the closest pair coincides (distance exactly 0), so that no two identical rows reach the saddle matrix
    const int n_remove = nalive - no_keep;
    for (int it = 0;; ++it) {
        // global minimum over alive rows; ties -> smallest row (= the first flat arg-min of the reference)
        double best = DBL_MAX;
        int brow = INT_MAX;
        for (int i = tid; i < n; i += blockDim.x)
            if (keep[i] && nn_idx[i] != INT_MAX && nn_dist[i] < best) { best = nn_dist[i]; brow = i; }
        block_argmin(best, brow, rv, rr);
        const int v = brow;
        if (v == INT_MAX) break;  // fewer than two alive points
        if (it >= n_remove && !(dedupe && best == 0.0)) break;
        if (tid == 0) keep[v] = 0;
        // rows whose nearest neighbour was the victim must be rescanned (usually one or two rows); collected in one
        // pass, repeated only if more than 1024 rows pointed at the victim
        bool again;
        do {
            for (int i = tid; i < n; i += blockDim.x)
                if (i != v && keep[i] && nn_idx[i] == v) {
                    const int slot = atomicAdd(&ntodo, 1);
                    if (slot < 1024) todo[slot] = i;
                }
            __syncthreads();
            const int found = ntodo;
            again = found > 1024;
            const int cnt = again ? 1024 : found;
            for (int t = 0; t < cnt; ++t) {   // the whole CTA scans one row at a time
                const int row = todo[t];
                sq_best sb;
                sb.init();
                for (int j = tid; j < n; j += blockDim.x) {
                    if (j == row || !keep[j]) continue;
                    sb.offer(ref_sumsq(X + (size_t)row * d, X + (size_t)j * d, d), j);
                }
                double b = sb.dist();
                int bj = sb.j;
                block_argmin(b, bj, rv, rr);
                if (tid == 0) { nn_dist[row] = b; nn_idx[row] = bj; }
            }
            __syncthreads();
            if (tid == 0) ntodo = 0;
            __syncthreads();
        } while (again);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) keep_g[i] = keep[i];
    }
}

// ---------------------------------------------------------------------------------------------------
// The removal loop when everything fits in shared memory -- the normal case (n up to ~8000 points for d = 2).
// The loop is a chain of reductions over n, so it is bound by INSTRUCTIONS ISSUED per point, and this version is
// built to issue few of them:
//   * every thread keeps the nearest-neighbour entries of its own points (tid, tid + 512, ...) in registers; shared
//     memory only holds the coordinates (70 KB for 4352 points in 2-D: the kernel fits beside a resident chain CTA);
//   * a removed point is "poisoned" in the CTA's private copy (coordinate 0 := 1e150 -> every distance to it is
//     ~1e150, finite so that the square root stays on its fast path; nn_dist := +inf), so no scan tests an alive flag;
//   * D is a template parameter (vector loads, unrolled differences);
//   * (distance, index) arg-min: distances are non-negative, so their bit patterns order like unsigned integers
//     and a warp reduces with three REDUX instructions (high word, low word, index) instead of five shuffle
//     rounds of a (double, int) pair; the 16 per-warp results are combined by every warp redundantly (no second
//     barrier for a broadcast).
// Same decisions as thin_remove_kernel: distance = correctly rounded square root of the reference's sum of
// squares, ties -> smallest index.
// ---------------------------------------------------------------------------------------------------
#define THF_T 512
#define THF_POISON 1e150

struct thf_red { unsigned hi[2][16], lo[2][16]; int idx[2][16]; };

__device__ __forceinline__ void thf_argmin(double &best, int &bidx, thf_red &R, int buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned hi = (unsigned)__double2hiint(best), lo = (unsigned)__double2loint(best);
    {
        const unsigned hm = __reduce_min_sync(0xffffffffu, hi);
        const unsigned lm = __reduce_min_sync(0xffffffffu, hi == hm ? lo : 0xffffffffu);
        const int im = __reduce_min_sync(0xffffffffu, (hi == hm && lo == lm) ? bidx : INT_MAX);
        if (lane == 0) { R.hi[buf][warp] = hm; R.lo[buf][warp] = lm; R.idx[buf][warp] = im; }
    }
    __syncthreads();
    const bool in = lane < (THF_T >> 5);
    hi = in ? R.hi[buf][lane] : 0xffffffffu;
    lo = in ? R.lo[buf][lane] : 0xffffffffu;
    const int id = in ? R.idx[buf][lane] : INT_MAX;
    const unsigned hm = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lm = __reduce_min_sync(0xffffffffu, hi == hm ? lo : 0xffffffffu);
    bidx = __reduce_min_sync(0xffffffffu, (hi == hm && lo == lm) ? id : INT_MAX);
    best = __hiloint2double((int)hm, (int)lm);
}

#define THF_K 16      // points per thread (registers): n <= THF_T * THF_K
static_assert(THF_T == 512, "the owner of point i is thread i & 511, slot i >> 9");

// KT: table slots per thread.  The 9-slot instantiation (n <= 4608) is capped at 72 registers so that the CTA fits beside
// a resident chain CTA (concurrent refit); the 16-slot one takes what it needs.
template <int D, int KT>
__global__ void __maxnreg__(KT <= 9 ? 72 : 128) thin_remove_fast_kernel(const double *__restrict__ Xg, int n, int dd, int no_keep,
                                                                    const double *__restrict__ nn_dist_g,
                                                                    const int *__restrict__ nn_idx_g, uint8_t *__restrict__ keep_g,
                                                                    const uint8_t *__restrict__ alive_g, int dedupe) {
    extern __shared__ __align__(16) unsigned char thin_sm[];
    __shared__ thf_red R;
    __shared__ int todo[1024];
    __shared__ int ntodo;
    __shared__ int nalive;
    const int d = D ? D : dd;
    const int tid = threadIdx.x;
    double *Xs = (double *)thin_sm;                    // the only per-point array in shared memory: the coordinates
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    // nearest-neighbour table of the points tid, tid + THF_T, ... in registers
    const int kmax = (n + THF_T - 1) / THF_T;
    double nd[KT];
    int ni[KT];
    if (tid == 0) { ntodo = 0; nalive = 0; }
    for (int e = tid; e < n * d; e += THF_T) Xs[e] = Xg[e];
    __syncthreads();
    {
        int mine = 0;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int i = tid + k * THF_T;
            const bool a = i < n && (!alive_g || alive_g[i]);
            ni[k] = a ? nn_idx_g[i] : INT_MAX;           // rows dead on entry carry INT_MAX / DBL_MAX (thin_nearest_kernel)
            nd[k] = (a && ni[k] != INT_MAX) ? nn_dist_g[i] : INF;
            if (i < n) keep_g[i] = a;
            if (i < n && !a) Xs[(size_t)i * d] = THF_POISON;   // like a removed point: no scan has to test a flag
            mine += a;
        }
        if (mine) atomicAdd(&nalive, mine);
    }
    __syncthreads();
    const int n_remove = nalive - no_keep;
    int buf = 0;
    for (int it = 0;; ++it) {
        // global minimum of the nearest-neighbour distances; ties -> smallest row (the first flat arg-min of the reference)
        double best = INF;
        int brow = INT_MAX;
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < kmax) {
                const bool lt = nd[k] < best;       // index ascending with k: the first minimum is kept
                best = lt ? nd[k] : best;
                brow = lt ? tid + k * THF_T : brow;
            }
        thf_argmin(best, brow, R, buf);
        buf ^= 1;
        const int v = brow;
        if (v == INT_MAX) break;                // fewer than two alive points (uniform)
        if (it >= n_remove && !(dedupe && best == 0.0)) break;      // uniform: thf_argmin broadcasts (best, brow)
        {
            const bool owner = tid == (v & (THF_T - 1));
            if (owner) { keep_g[v] = 0; Xs[(size_t)v * d] = THF_POISON; }
            const int slot = owner ? (v >> 9) : -1;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                nd[k] = k == slot ? INF : nd[k];
                ni[k] = k == slot ? INT_MAX : ni[k];
            }
        }
        // rows whose nearest neighbour was the victim must be rescanned (usually one or two)
        int pending = 0;                         // bit k: own point k still waits for its rescan (more than 1024 rows at once)
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < kmax && ni[k] == v) pending |= 1 << k;
        bool again;
        do {
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (pending & (1 << k)) {
                    const int slot = atomicAdd(&ntodo, 1);
                    if (slot < 1024) { todo[slot] = tid + k * THF_T; pending &= ~(1 << k); }
                }
            __syncthreads();                     // also publishes the poisoned coordinates of v
            const int found = ntodo;
            again = found > 1024;
            const int cnt = again ? 1024 : found;
            for (int t = 0; t < cnt; ++t) {      // the whole CTA scans one row at a time
                const int row = todo[t];
                double xr[D ? D : 1];
                if (D) {
#pragma unroll
                    for (int k = 0; k < D; ++k) xr[k] = Xs[(size_t)row * D + k];
                }
                sq_best sb;
                sb.init();
#pragma unroll 4
                for (int j = tid; j < n; j += THF_T) {
                    double sq = 0.0;
                    if (D == 2) {
                        const double2 p = *(const double2 *)(Xs + (size_t)j * 2);
                        const double d0 = __dadd_rn(p.x, -xr[0]), d1 = __dadd_rn(p.y, -xr[1]);
                        sq = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));      // 0 + d0^2 == d0^2 exactly
                    } else if (D) {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const double df = __dadd_rn(Xs[(size_t)j * D + k], -xr[k]);
                            sq = __dadd_rn(sq, __dmul_rn(df, df));
                        }
                    } else {
                        sq = ref_sumsq(Xs + (size_t)row * d, Xs + (size_t)j * d, d);
                    }
                    if (j != row) sb.offer(sq, j);
                }
                double b = sb.j == INT_MAX ? INF : __dsqrt_rn(sb.sq);
                int bj = sb.j;
                thf_argmin(b, bj, R, buf);
                buf ^= 1;
                {
                    const bool none = !(b < 0.5 * THF_POISON);       // only poisoned points left: nobody else is alive
                    const int slot = tid == (row & (THF_T - 1)) ? (row >> 9) : -1;
                    const double nb = none ? INF : b;
                    const int nj = none ? INT_MAX : bj;
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        nd[k] = k == slot ? nb : nd[k];
                        ni[k] = k == slot ? nj : ni[k];
                    }
                }
            }
            __syncthreads();
            if (tid == 0) ntodo = 0;
            __syncthreads();
        } while (again);
    }
}

extern "C" int64_t sdb_rbf_thin_work(int64_t n) { return n + n / 2 + 64; }

static int thin_impl(const double *d_X, int64_t n, int32_t d, int64_t no_keep, const uint8_t *d_alive, int dedupe,
                     uint8_t *d_keep, double *d_work, void *stream) {
    SDB_CHECK_ARG(d_X && d_keep && d_work, "sdb_rbf_thin: NULL pointer");
    SDB_CHECK_ARG(n >= 1 && n < (1 << 30) && no_keep >= 0, "sdb_rbf_thin: bad sizes");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D, "sdb_rbf_thin: d=%d outside limits", d);
    cudaStream_t st = (cudaStream_t)stream;
    double *nn_dist = d_work;
    int *nn_idx = (int *)(d_work + n);
    if (no_keep >= n && !dedupe) {
        if (d_alive) SDB_CUDA(cudaMemcpyAsync(d_keep, d_alive, (size_t)n, cudaMemcpyDeviceToDevice, st));
        else SDB_CUDA(cudaMemsetAsync(d_keep, 1, (size_t)n, st));
        return SDB_OK;
    }
    sdb_count_launch();
    thin_nearest_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(d_X, (int)n, d, nn_dist, nn_idx, d_alive);
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    const size_t need_fast = (size_t)n * (size_t)d * 8 + 16;
    if (n <= THF_T * THF_K && need_fast + 12288 <= (size_t)dp.smem_optin) {
        void (*fn)(const double *, int, int, int, const double *, const int *, uint8_t *, const uint8_t *, int);
        if (n <= THF_T * 9)
            fn = d == 1 ? thin_remove_fast_kernel<1, 9> : d == 2 ? thin_remove_fast_kernel<2, 9> : d == 3 ? thin_remove_fast_kernel<3, 9>
                 : d == 4 ? thin_remove_fast_kernel<4, 9> : thin_remove_fast_kernel<0, 9>;
        else
            fn = d == 1 ? thin_remove_fast_kernel<1, THF_K> : d == 2 ? thin_remove_fast_kernel<2, THF_K>
                 : d == 3 ? thin_remove_fast_kernel<3, THF_K> : d == 4 ? thin_remove_fast_kernel<4, THF_K> : thin_remove_fast_kernel<0, THF_K>;
        SDB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need_fast));
        sdb_count_launch();
        fn<<<1, THF_T, need_fast, st>>>(d_X, (int)n, d, (int)no_keep, nn_dist, nn_idx, d_keep, d_alive, dedupe);
        SDB_CUDA(cudaGetLastError());
        return SDB_OK;
    }
    const size_t need = (size_t)n * ((size_t)d * 8 + 8 + 4 + 1) + 16;
    const int use_smem = need + 8192 <= (size_t)dp.smem_optin;
    SDB_CUDA(cudaFuncSetAttribute((const void *)thin_remove_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(use_smem ? need : 0)));
    sdb_count_launch();
    thin_remove_kernel<<<1, 1024, use_smem ? need : 0, st>>>(d_X, (int)n, d, (int)no_keep, nn_dist, nn_idx, d_keep, use_smem,
                                                             d_alive, dedupe);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

extern "C" int sdb_rbf_thin(const double *d_X, int64_t n, int32_t d, int64_t no_keep, uint8_t *d_keep, double *d_work,
                            void *stream) {
    return thin_impl(d_X, n, d, no_keep, nullptr, 0, d_keep, d_work, stream);
}

extern "C" int sdb_rbf_thin_masked(const double *d_X, int64_t n, int32_t d, int64_t no_keep, const uint8_t *d_alive,
                                   int32_t drop_coincident, uint8_t *d_keep, double *d_work, void *stream) {
    return thin_impl(d_X, n, d, no_keep, d_alive, drop_coincident != 0, d_keep, d_work, stream);
}
