// Device functions: forward models (poly / RBF / toy), Gaussian log-posterior.
// Everything here works on models whose tables already sit in shared memory.
#pragma once
#include <type_traits>

#include "sdb_common.cuh"

// A model whose tables have been staged into shared memory.
struct sdb_smodel {
    int kind, kernel_type, degree, n_terms;
    int n_centers;
    const double *coefs;    // smem
    const double *centers;  // smem
    const int *alpha;       // smem
    const double *herm;     // smem
    double toy_f, toy_L, toy_M;
};

// bytes of shared memory a model needs (each table rounded up to 16 B)
__host__ __device__ inline size_t sdb_align16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ inline size_t sdb_model_smem_bytes(const sdb_model &mo, int d, int m) {
    if (mo.kind == SDB_MODEL_RBF)
        return sdb_align16((size_t)mo.n_centers * d * 8) + sdb_align16((size_t)(mo.n_centers + d + 1) * m * 8);
    if (mo.kind == SDB_MODEL_POLY)
        return sdb_align16((size_t)mo.n_terms * m * 8) + sdb_align16((size_t)(mo.degree + 1) * (mo.degree + 1) * 8) +
               sdb_align16((size_t)mo.n_terms * d * 4);
    return 0;
}

// Stage one table: thread 0 issues the 16-byte-granular part as TMA bulk copies on `bar`,
// the (<16 B) tail goes through ordinary 4-byte loads.  Call from all threads of the CTA.
__device__ __forceinline__ void sdb_stage_table(void *dst, const void *src, size_t bytes, uint64_t *bar) {
    const size_t bulk = bytes & ~(size_t)15;
    if (threadIdx.x == 0 && bulk) sdb_bulk_stage(dst, src, bulk, bar);
    const size_t tail_words = (bytes - bulk) >> 2;
    if (threadIdx.x < tail_words)
        ((uint32_t *)((char *)dst + bulk))[threadIdx.x] = ((const uint32_t *)((const char *)src + bulk))[threadIdx.x];
}

// Carve `base` for model `mo`, start the copies, return the smem view.  All threads call it.
__device__ inline sdb_smodel sdb_stage_model(const sdb_model &mo, int d, int m, unsigned char *&base, uint64_t *bar) {
    sdb_smodel s;
    s.kind = mo.kind; s.kernel_type = mo.kernel_type; s.degree = mo.degree; s.n_terms = mo.n_terms;
    s.n_centers = (int)mo.n_centers;
    s.coefs = nullptr; s.centers = nullptr; s.alpha = nullptr; s.herm = nullptr;
    s.toy_f = mo.toy_f; s.toy_L = mo.toy_L; s.toy_M = mo.toy_M;
    if (mo.kind == SDB_MODEL_RBF) {
        const size_t cb = (size_t)mo.n_centers * d * 8, wb = (size_t)(mo.n_centers + d + 1) * m * 8;
        s.centers = (const double *)base; sdb_stage_table(base, mo.centers, cb, bar); base += sdb_align16(cb);
        s.coefs = (const double *)base;   sdb_stage_table(base, mo.coefs, wb, bar);   base += sdb_align16(wb);
    } else if (mo.kind == SDB_MODEL_POLY) {
        const size_t mb = (size_t)mo.n_terms * m * 8, hb = (size_t)(mo.degree + 1) * (mo.degree + 1) * 8,
                     ab = (size_t)mo.n_terms * d * 4;
        s.coefs = (const double *)base; sdb_stage_table(base, mo.coefs, mb, bar); base += sdb_align16(mb);
        s.herm = (const double *)base;  sdb_stage_table(base, mo.herm, hb, bar);  base += sdb_align16(hb);
        s.alpha = (const int *)base;    sdb_stage_table(base, mo.alpha, ab, bar); base += sdb_align16(ab);
    }
    return s;
}

// ---------------------------------------------------------------------------------------
// Polynomial surrogate, one point per thread (surrogate_poly.py:18-30,162-175).
// Univariate values by monomial accumulation in the reference's order, products left to
// right, both with separately rounded multiply/add so PHI is bit-identical to numpy's;
// only the final PHI.MAT contraction uses fma (BLAS order is not reproducible anyway).
// ---------------------------------------------------------------------------------------
template <int D, int M>
__device__ __forceinline__ void sdb_poly_point(const sdb_smodel &s, int d, int m, const double *x, double *out) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int HS = SDB_MAX_DEGREE + 1;
    const int deg = s.degree, nh = deg + 1;
    double h[DD * HS];
#pragma unroll UD
    for (int j = 0; j < DD; ++j) {
        if (j < d) {
            double pw[HS];
            pw[0] = 1.0;
            for (int i = 1; i <= deg; ++i) pw[i] = __dmul_rn(pw[i - 1], x[j]);
            for (int k = 0; k <= deg; ++k) {
                double v = s.herm[k * nh];
                for (int i = 1; i <= k; ++i) {   // coefficients above the diagonal are zero: adding 0*x^i is exact
                    v = __dadd_rn(v, __dmul_rn(s.herm[k * nh + i], pw[i]));
                }
                h[j * HS + k] = v;
            }
        }
    }
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) out[k] = 0.0;
    for (int t = 0; t < s.n_terms; ++t) {
        const int *a = s.alpha + t * d;
        double phi = h[a[0]];                     // 1 * h = h exactly
#pragma unroll UD
        for (int j = 1; j < DD; ++j)
            if (j < d) phi = __dmul_rn(phi, h[j * HS + a[j]]);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[k] = fma(phi, s.coefs[t * m + k], out[k]);
    }
}

// ---------------------------------------------------------------------------------------
// RBF surrogate (surrogate_rbf.py:20-40): partial sums over centers i = first, first+stride, ...
// ---------------------------------------------------------------------------------------
// ST: the lane stride when it is 1 or 2 (the plans of the named shapes), 0 = run-time value
template <int D, int M, int KT, int ST>
__device__ __forceinline__ void sdb_rbf_partial_kt(const sdb_smodel &s, int d, int m, const double *x, double *acc,
                                                   int first, int stride_rt) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int n = s.n_centers;
    const int stride = ST ? ST : stride_rt;
    // running pointers: with a compile-time stride the unrolled body addresses center k of an iteration as pointer +
    // immediate; indexing by i = first + k * stride costs four ALU instructions per pair, a run-time stride two -- in
    // a loop whose limit is the issue port (11 FP64 instructions x 2 issue cycles + everything else)
    const double *c = s.centers + (size_t)first * d;
    const double *w = s.coefs + (size_t)first * m;
    const int cs = stride * d, ws = stride * m;
#pragma unroll 8
    for (int i = first; i < n; i += stride, c += cs, w += ws) {
        double r2 = SDB_R2_FLOOR;          // keeps the reciprocal square root on normal numbers: no zero / subnormal guard
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) {
                const double df = c[j] - x[j];
                r2 = fma(df, df, r2);
            }
        const double phi = sdb_rbf_phi(r2, KT);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[k] = fma(phi, w[k], acc[k]);
    }
}

template <int D, int M, int ST>
__device__ __forceinline__ void sdb_rbf_partial_st(const sdb_smodel &s, int d, int m, const double *x, double *acc,
                                                   int first, int stride) {
    switch (s.kernel_type) {
        case 1: sdb_rbf_partial_kt<D, M, 1, ST>(s, d, m, x, acc, first, stride); break;
        case 2: sdb_rbf_partial_kt<D, M, 2, ST>(s, d, m, x, acc, first, stride); break;
        case 3: sdb_rbf_partial_kt<D, M, 3, ST>(s, d, m, x, acc, first, stride); break;
        case 4: sdb_rbf_partial_kt<D, M, 4, ST>(s, d, m, x, acc, first, stride); break;
        case 5: sdb_rbf_partial_kt<D, M, 5, ST>(s, d, m, x, acc, first, stride); break;
        case 7: sdb_rbf_partial_kt<D, M, 7, ST>(s, d, m, x, acc, first, stride); break;
        default: sdb_rbf_partial_kt<D, M, 0, ST>(s, d, m, x, acc, first, stride); break;
    }
}

template <int D, int M>
__device__ __forceinline__ void sdb_rbf_partial(const sdb_smodel &s, int d, int m, const double *x, double *acc,
                                                int first, int stride) {
    if (stride == 1) sdb_rbf_partial_st<D, M, 1>(s, d, m, x, acc, first, 1);
    else if (stride == 2) sdb_rbf_partial_st<D, M, 2>(s, d, m, x, acc, first, 2);
    else sdb_rbf_partial_st<D, M, 0>(s, d, m, x, acc, first, stride);
}

// NP points against the same centers: every center / weight load serves NP pairs (the loads are two of the ~4.5
// non-FP64 issue slots a pair costs).  Same per-pair arithmetic as sdb_rbf_partial_kt.  xs[p][j], acc[p][k].
template <int D, int M, int KT, int ST, int NP>
__device__ __forceinline__ void sdb_rbf_partialn_kt(const sdb_smodel &s, int d, int m, const double (*xs)[D ? D : SDB_MAX_D],
                                                    double (*acc)[M ? M : SDB_MAX_M], int first, int stride_rt) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int n = s.n_centers;
    const int stride = ST ? ST : stride_rt;
    const double *c = s.centers + (size_t)first * d;
    const double *w = s.coefs + (size_t)first * m;
    const int cs = stride * d, ws = stride * m;
#pragma unroll(8 / NP)
    for (int i = first; i < n; i += stride, c += cs, w += ws) {
        double r2[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) r2[p] = SDB_R2_FLOOR;
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) {
                const double cj = c[j];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double df = cj - xs[p][j];
                    r2[p] = fma(df, df, r2[p]);
                }
            }
        double phi[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) phi[p] = sdb_rbf_phi(r2[p], KT);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) {
                const double wk = w[k];
#pragma unroll
                for (int p = 0; p < NP; ++p) acc[p][k] = fma(phi[p], wk, acc[p][k]);
            }
    }
}

template <int D, int M, int ST, int NP>
__device__ __forceinline__ void sdb_rbf_partialn_st(const sdb_smodel &s, int d, int m, const double (*xs)[D ? D : SDB_MAX_D],
                                                    double (*acc)[M ? M : SDB_MAX_M], int first, int stride) {
    switch (s.kernel_type) {
        case 1: sdb_rbf_partialn_kt<D, M, 1, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 2: sdb_rbf_partialn_kt<D, M, 2, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 3: sdb_rbf_partialn_kt<D, M, 3, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 4: sdb_rbf_partialn_kt<D, M, 4, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 5: sdb_rbf_partialn_kt<D, M, 5, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 7: sdb_rbf_partialn_kt<D, M, 7, ST, NP>(s, d, m, xs, acc, first, stride); break;
        default: sdb_rbf_partialn_kt<D, M, 0, ST, NP>(s, d, m, xs, acc, first, stride); break;
    }
}

// linear tail: COEFS[n+d] + sum_j x_j * COEFS[n+j]   (surrogate_rbf.py:38)
template <int D, int M>
__device__ __forceinline__ void sdb_rbf_add_tail(const sdb_smodel &s, int d, int m, const double *x, double *acc) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const double *lin = s.coefs + (size_t)s.n_centers * m;
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = 0.0;
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) t = fma(x[j], lin[j * m + k], t);
            acc[k] += lin[d * m + k] + t;
        }
}

// L lanes (a power of two, aligned inside a warp) share one point: lane `sub` takes centers
// sub, sub+L, ...; an xor butterfly leaves the identical total on all L lanes.
template <int D, int M>
__device__ __forceinline__ void sdb_rbf_point_split(const sdb_smodel &s, int d, int m, const double *x, double *out,
                                                    int sub, int L) {
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) out[k] = 0.0;
    sdb_rbf_partial<D, M>(s, d, m, x, out, sub, L);
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[k] += sdb_shfl_xor(out[k], off);
    }
    sdb_rbf_add_tail<D, M>(s, d, m, x, out);
}

// The same for a call every lane of the warp makes (the proposal of every chain in a lockstep step): NP neighbouring
// L-lane groups -- NP chains -- form one team of NP*L lanes; lane t of the team takes centers t, t + NP*L, ... and evaluates
// ALL the team's points against them (point p of a lane is the one of the group at xor-distance p), then the groups
// exchange the partial sums that belong to each other.  Must be called by the full warp, converged; NP * L <= 32.
template <int D, int M, int NP>
__device__ __forceinline__ void sdb_rbf_point_team(const sdb_smodel &s, int d, int m, const double *x, double *out, int L) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int team = (threadIdx.x & 31) & (NP * L - 1);
    double xs[NP][DD], acc[NP][MM];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) xs[p][j] = p ? sdb_shfl_xor(x[j], p * L) : x[j];
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[p][k] = 0.0;
    }
    if (NP * L == 2) sdb_rbf_partialn_st<D, M, 2, NP>(s, d, m, xs, acc, team, 2);
    else if (NP * L == 4) sdb_rbf_partialn_st<D, M, 4, NP>(s, d, m, xs, acc, team, 4);
    else sdb_rbf_partialn_st<D, M, 0, NP>(s, d, m, xs, acc, team, NP * L);
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = acc[0][k];
#pragma unroll
            for (int p = 1; p < NP; ++p) t += sdb_shfl_xor(acc[p][k], p * L);   // group (mine xor p) holds its partial sum for MY point in slot p
            out[k] = t;
        }
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[k] += sdb_shfl_xor(out[k], off);
    }
    sdb_rbf_add_tail<D, M>(s, d, m, x, out);
}

// Warp-cooperative evaluation for a data-dependent subset of chains (the "true model" after a
// stage-1 pass): every lane whose `need` is set and that is the first lane of its L-group gets
// served in turn; all 32 lanes split the centers of that one point.  Must be called by the full
// warp, converged.  Lanes with need set receive the result in out[].
template <int D, int M>
__device__ __forceinline__ void sdb_rbf_point_warp(const sdb_smodel &s, int d, int m, const double *x, double *out,
                                                   bool need, int L) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, need && ((lane & (L - 1)) == 0));
    // two points per pass: every center / weight load serves two pairs (same per-point arithmetic and summation order as one
    // point per pass, so the results do not depend on how the survivors pair up); a last odd point rides with itself
    while (todo) {
        const int src0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const int src1 = todo ? __ffs(todo) - 1 : src0;
        todo &= todo - 1;                       // no-op when todo is already 0
        double xs[2][DD], acc[2][MM];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) { xs[0][j] = sdb_shfl(x[j], src0); xs[1][j] = sdb_shfl(x[j], src1); }
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[0][k] = acc[1][k] = 0.0;
        sdb_rbf_partialn_st<D, M, 32, 2>(s, d, m, xs, acc, lane, 32);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) { acc[0][k] += sdb_shfl_xor(acc[0][k], off); acc[1][k] += sdb_shfl_xor(acc[1][k], off); }
        }
        sdb_rbf_add_tail<D, M>(s, d, m, xs[0], acc[0]);
        sdb_rbf_add_tail<D, M>(s, d, m, xs[1], acc[1]);
        if ((lane / L) == (src0 / L)) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) out[k] = acc[0][k];
        }
        if (src1 != src0 && (lane / L) == (src1 / L)) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) out[k] = acc[1][k];
        }
    }
}

// ---------------------------------------------------------------------------------------
// FP32 evaluation of the RBF surrogate (additive option surrogate_precision = "f32"; the reference evaluates in
// float64, surrogate_rbf.py:20-40).  Centers are stored relative to the first center (so that parameters far from the
// origin do not lose the digits that matter: the difference), weights in FP32; distance, kernel and the sum over a
// lane's centers run on the FP32 pipe with MUFU.RSQ / MUFU.EX2; the per-lane partial sums are combined and the
// linear tail is added in FP64.  Per pair at d = 2, m = 1, cubic kernel: 7 FP32 instructions + 1 MUFU (the FP64
// path: 11 FP64 instructions, each holding the issue port two cycles).  The result carries FP32 rounding of every
// term: |error| <~ 1e-7 * sum_j |w_j phi_j|, i.e. 1e-4-accurate only while the expansion is not badly cancelling.
// In a DAMH chain the surrogate only screens proposals (stage 1); stage 2 corrects with the true posterior, so the
// chain targets the exact posterior whatever the surrogate's precision.
// ---------------------------------------------------------------------------------------
struct sdb_smodel32 {
    int kernel_type, n_centers;
    const float *centers;   // smem [n][d], relative to `shift`
    const float *coefs;     // smem [n][m]
    const double *tail;     // smem [(d+1)][m]: linear rows + constant row (FP64)
    const double *shift;    // smem [d]
};

__host__ __device__ inline size_t sdb_model32_smem_bytes(const sdb_model &mo, int d, int m) {
    return sdb_align16((size_t)mo.n_centers * d * 4) + sdb_align16((size_t)mo.n_centers * m * 4) +
           sdb_align16((size_t)(d + 1) * m * 8) + sdb_align16((size_t)d * 8);
}

// All threads: convert the FP64 tables in global memory into the FP32 shared-memory layout (ends with __syncthreads).
__device__ inline sdb_smodel32 sdb_stage_model32(const sdb_model &mo, int d, int m, unsigned char *&base) {
    sdb_smodel32 s;
    const int n = (int)mo.n_centers;
    s.kernel_type = mo.kernel_type; s.n_centers = n;
    float *c = (float *)base; base += sdb_align16((size_t)n * d * 4);
    float *w = (float *)base; base += sdb_align16((size_t)n * m * 4);
    double *t = (double *)base; base += sdb_align16((size_t)(d + 1) * m * 8);
    double *sh = (double *)base; base += sdb_align16((size_t)d * 8);
    for (int i = threadIdx.x; i < n * d; i += blockDim.x) c[i] = (float)(mo.centers[i] - mo.centers[i % d]);
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) w[i] = (float)mo.coefs[i];
    for (int i = threadIdx.x; i < (d + 1) * m; i += blockDim.x) t[i] = mo.coefs[(size_t)n * m + i];
    for (int i = threadIdx.x; i < d; i += blockDim.x) sh[i] = mo.centers[i];
    __syncthreads();
    s.centers = c; s.coefs = w; s.tail = t; s.shift = sh;
    return s;
}

__device__ __forceinline__ float sdb_rsqrt32(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sdb_ex2_32(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#define SDB_R2_FLOOR32 1e-30f   // keeps rsqrt away from 0 (r = 1e-15 for a coincident point)

template <int KT>
__device__ __forceinline__ float sdb_rbf_phi32(float s) {
    if (KT == 4) return sdb_ex2_32(-1.4426950408889634f * s);
    if (KT == 5) return sdb_rsqrt32(s + 1.0f);
    const float r = s * sdb_rsqrt32(s);
    if (KT == 1) return r * s;
    if (KT == 2) return r * s * s;
    if (KT == 3) return (r * s) * (s * s);
    if (KT == 7) return (r * s * s) * (s * s);
    return r;
}

template <int D, int M, int KT, int ST, int NP>
__device__ __forceinline__ void sdb_rbf32_partialn_kt(const sdb_smodel32 &s, int d, int m, const float (*xs)[D ? D : SDB_MAX_D],
                                                      float (*acc)[M ? M : SDB_MAX_M], int first, int stride_rt) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int n = s.n_centers;
    const int stride = ST ? ST : stride_rt;
    const float *c = s.centers + (size_t)first * d;
    const float *w = s.coefs + (size_t)first * m;
    const int cs = stride * d, ws = stride * m;
#pragma unroll(8 / NP)
    for (int i = first; i < n; i += stride, c += cs, w += ws) {
        float r2[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) r2[p] = SDB_R2_FLOOR32;
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) {
                const float cj = c[j];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float df = cj - xs[p][j];
                    r2[p] = fmaf(df, df, r2[p]);
                }
            }
        float phi[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) phi[p] = sdb_rbf_phi32<KT>(r2[p]);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) {
                const float wk = w[k];
#pragma unroll
                for (int p = 0; p < NP; ++p) acc[p][k] = fmaf(phi[p], wk, acc[p][k]);
            }
    }
}

template <int D, int M, int ST, int NP>
__device__ __forceinline__ void sdb_rbf32_partialn_st(const sdb_smodel32 &s, int d, int m, const float (*xs)[D ? D : SDB_MAX_D],
                                                      float (*acc)[M ? M : SDB_MAX_M], int first, int stride) {
    switch (s.kernel_type) {
        case 1: sdb_rbf32_partialn_kt<D, M, 1, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 2: sdb_rbf32_partialn_kt<D, M, 2, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 3: sdb_rbf32_partialn_kt<D, M, 3, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 4: sdb_rbf32_partialn_kt<D, M, 4, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 5: sdb_rbf32_partialn_kt<D, M, 5, ST, NP>(s, d, m, xs, acc, first, stride); break;
        case 7: sdb_rbf32_partialn_kt<D, M, 7, ST, NP>(s, d, m, xs, acc, first, stride); break;
        default: sdb_rbf32_partialn_kt<D, M, 0, ST, NP>(s, d, m, xs, acc, first, stride); break;
    }
}

// The FP32 counterpart of sdb_rbf_point_team: NP neighbouring L-lane groups share their centers' loads.  Full warp, converged.
template <int D, int M, int NP>
__device__ __forceinline__ void sdb_rbf32_point_team(const sdb_smodel32 &s, int d, int m, const double *x, double *out, int L) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int team = (threadIdx.x & 31) & (NP * L - 1);
    float xs[NP][DD], acc[NP][MM];
#pragma unroll UD
    for (int j = 0; j < DD; ++j)
        if (j < d) {
            const float xj = (float)(x[j] - s.shift[j]);
#pragma unroll
            for (int p = 0; p < NP; ++p) xs[p][j] = p ? __shfl_xor_sync(0xffffffffu, xj, p * L) : xj;
        }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[p][k] = 0.0f;
    }
    if (NP * L == 2) sdb_rbf32_partialn_st<D, M, 2, NP>(s, d, m, xs, acc, team, 2);
    else if (NP * L == 4) sdb_rbf32_partialn_st<D, M, 4, NP>(s, d, m, xs, acc, team, 4);
    else sdb_rbf32_partialn_st<D, M, 0, NP>(s, d, m, xs, acc, team, NP * L);
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = (double)acc[0][k];
#pragma unroll
            for (int p = 1; p < NP; ++p) t += (double)__shfl_xor_sync(0xffffffffu, acc[p][k], p * L);
            out[k] = t;
        }
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[k] += sdb_shfl_xor(out[k], off);
    }
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = 0.0;
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) t = fma(x[j], s.tail[j * m + k], t);
            out[k] += s.tail[d * m + k] + t;
        }
}

// ---------------------------------------------------------------------------------------
// Streamed RBF surrogate: the tables do not fit in shared memory (no_keep = None in the reference lets the center set
// grow without bound, surrogate_rbf.py:43,97), so every evaluation streams them from L2 through a two-stage ring of
// chunks filled by TMA bulk copies (full mbarriers; the stage is handed back by the CTA-wide barrier that follows its
// use).  The ring simply keeps cycling over the table: the first chunks of the NEXT evaluation are already in flight
// when an evaluation ends.  Chunks hold a multiple of 32 centers, so lane t of a team visits the same centers in the
// same order as with resident tables: results are bit-identical to the resident path.
// ---------------------------------------------------------------------------------------
struct sdb_ring {
    uint64_t *full;            // smem: full[0], full[1]
    unsigned char *base;       // smem: 2 x (cbytes + wbytes)
    size_t cbytes, wbytes;
    int chunk, nchunks, n, kernel_type;
    const double *g_centers, *g_coefs;   // global
    const double *tail;        // smem [(d+1)][m]
    long long total;           // chunks this CTA will consume over the whole launch
    long long issued, consumed;
    int phase_bits;
};

__host__ __device__ inline size_t sdb_ring_smem_bytes(int chunk, int d, int m) {
    return 16 + sdb_align16((size_t)(d + 1) * m * 8) + 2 * (sdb_align16((size_t)chunk * d * 8) + sdb_align16((size_t)chunk * m * 8));
}

__device__ __forceinline__ void sdb_ring_issue(sdb_ring &r, int d, int m) {
    // all threads call it: thread 0 drives the TMA, a few threads copy a (< 16 B) tail of the last chunk
    const int ci = (int)(r.issued % r.nchunks), st = (int)(r.issued & 1);
    const long long c0 = (long long)ci * r.chunk;
    const int cn = (int)((r.n - c0) < r.chunk ? (r.n - c0) : r.chunk);
    unsigned char *sc = r.base + (size_t)st * (r.cbytes + r.wbytes);
    sdb_stage_table(sc, r.g_centers + c0 * d, (size_t)cn * d * 8, &r.full[st]);
    sdb_stage_table(sc + r.cbytes, r.g_coefs + c0 * m, (size_t)cn * m * 8, &r.full[st]);
    if (threadIdx.x == 0) sdb_mbar_arrive(&r.full[st]);
    ++r.issued;
}

// All threads (barriers inside).  evals = number of surrogate evaluations this launch will make.
__device__ inline sdb_ring sdb_ring_init(const sdb_model &mo, int d, int m, int chunk, long long evals, unsigned char *&base) {
    sdb_ring r;
    r.full = (uint64_t *)base; base += 16;
    double *t = (double *)base; base += sdb_align16((size_t)(d + 1) * m * 8);
    r.tail = t;
    r.n = (int)mo.n_centers; r.kernel_type = mo.kernel_type;
    r.chunk = chunk; r.nchunks = (r.n + chunk - 1) / chunk;
    r.cbytes = sdb_align16((size_t)chunk * d * 8); r.wbytes = sdb_align16((size_t)chunk * m * 8);
    r.base = base; base += 2 * (r.cbytes + r.wbytes);
    r.g_centers = mo.centers; r.g_coefs = mo.coefs;
    r.total = evals * r.nchunks; r.issued = 0; r.consumed = 0; r.phase_bits = 0;
    if (threadIdx.x == 0) {
        sdb_mbar_init(&r.full[0], 1);
        sdb_mbar_init(&r.full[1], 1);
        sdb_mbar_fence_init();
    }
    for (int i = threadIdx.x; i < (d + 1) * m; i += blockDim.x) t[i] = mo.coefs[(size_t)r.n * m + i];
    __syncthreads();
    while (r.issued < 2 && r.issued < r.total) sdb_ring_issue(r, d, m);
    return r;
}

// One evaluation of every lane's point against the whole table.  Full CTA, converged (barriers inside).
template <int D, int M, int NP>
__device__ __forceinline__ void sdb_rbf_point_team_stream(sdb_ring &r, int d, int m, const double *x, double *out, int L) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int team = (threadIdx.x & 31) & (NP * L - 1);
    double xs[NP][DD], acc[NP][MM];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) xs[p][j] = p ? sdb_shfl_xor(x[j], p * L) : x[j];
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[p][k] = 0.0;
    }
    for (int ci = 0; ci < r.nchunks; ++ci) {
        const int st = (int)(r.consumed & 1);
        sdb_mbar_wait(&r.full[st], (r.phase_bits >> st) & 1);
        r.phase_bits ^= 1 << st;
        const long long c0 = (long long)ci * r.chunk;
        sdb_smodel s;
        s.kind = SDB_MODEL_RBF; s.kernel_type = r.kernel_type;
        s.n_centers = (int)((r.n - c0) < r.chunk ? (r.n - c0) : r.chunk);
        s.centers = (const double *)(r.base + (size_t)st * (r.cbytes + r.wbytes));
        s.coefs = (const double *)(r.base + (size_t)st * (r.cbytes + r.wbytes) + r.cbytes);
        if (NP * L == 2) sdb_rbf_partialn_st<D, M, 2, NP>(s, d, m, xs, acc, team, 2);
        else if (NP * L == 4) sdb_rbf_partialn_st<D, M, 4, NP>(s, d, m, xs, acc, team, 4);
        else sdb_rbf_partialn_st<D, M, 0, NP>(s, d, m, xs, acc, team, NP * L);
        ++r.consumed;
        __syncthreads();      // every warp is done with stage st (and tail words of earlier issues are visible)
        if (r.issued < r.total) sdb_ring_issue(r, d, m);
    }
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = acc[0][k];
#pragma unroll
            for (int p = 1; p < NP; ++p) t += sdb_shfl_xor(acc[p][k], p * L);
            out[k] = t;
        }
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[k] += sdb_shfl_xor(out[k], off);
    }
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = 0.0;
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) t = fma(x[j], r.tail[j * m + k], t);
            out[k] += r.tail[d * m + k] + t;
        }
}

// ---------------------------------------------------------------------------------------
// Warp-owned chains (chain_kernel<..., WT = true>): a warp owns P <= 32 chains, lane p < P holds chain p.  For the
// surrogate of a lockstep step ALL 32 lanes split the centers (lane t takes centers t, t + 32, ...) and evaluate the
// warp's points in groups of NP = 8 (the tail of the warp's points in a group of 4 / 2 / 1): a center / weight load
// serves NP pairs, and -- the reason for this layout -- P is ANY integer, so C chains spread over exactly the warps
// the GPU holds in one wave (65 536 chains = 2368 warps x 27.67) instead of power-of-two lane groups that leave a
// sub-partition with 3 warps next to one with 4.  The 32 partial sums of a point are combined by a fixed tree
// (xor 16, 8, 4, 2, 1), whatever P and NP are: the result does not depend on how the chains are sharded.
// ---------------------------------------------------------------------------------------
// largest group: the group's points and sums live in registers (NP (d + m) doubles) beside the chain state
__host__ __device__ constexpr int sdb_wt_group_max(int D, int M) {
    return (D == 0 || M == 0) ? 2 : (D + M <= 3) ? 8 : (D + M <= 7) ? 4 : 2;
}
// v[NP]: this lane's partial sums of NP points.  Returns, on every lane, the 32-lane total of point
// (lane * NP) >> 5: halving exchanges while more than one point is held, plain butterflies after.
template <int NP>
__device__ __forceinline__ double sdb_wt_reduce(const double *v, int lane) {
    double r[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) r[i] = v[i];
    int off = 16;
#pragma unroll
    for (int h = NP / 2; h >= 1; h >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const double keep = up ? r[h + i] : r[i];
            const double send = up ? r[i] : r[h + i];
            r[i] = keep + sdb_shfl_xor(send, off);
        }
        off >>= 1;
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        if (off >= 1) r[0] += sdb_shfl_xor(r[0], off);
        off >>= 1;
    }
    return r[0];
}

// One group: points of lanes g0 .. g0 + NP - 1.  partial(xs, acc) runs this lane's share of the pair loop.
template <int D, int M, int NP, class F>
__device__ __forceinline__ void sdb_wt_group(F &partial, int d, int m, const double *x, double *out, int g0, int lane) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    double xs[NP][DD], acc[NP][MM];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) xs[p][j] = sdb_shfl(x[j], g0 + p);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[p][k] = 0.0;
    }
    partial(xs, acc, std::integral_constant<int, NP>());
    const int src = ((lane - g0) & (NP - 1)) * (32 / NP);
    const bool mine = lane >= g0 && lane < g0 + NP;
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double v[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) v[p] = acc[p][k];
            const double t = sdb_shfl(sdb_wt_reduce<NP>(v, lane), src);
            if (mine) out[k] = t;
        }
}

// All groups of a warp's P points (P warp-uniform).  Full warp, converged.  out[] of lanes >= P is left untouched.
template <int D, int M, class F>
__device__ __forceinline__ void sdb_wt_points(F &partial, int d, int m, const double *x, double *out, int P) {
    const int lane = threadIdx.x & 31;
    constexpr int GN = sdb_wt_group_max(D, M);
    int g0 = 0;
    while (g0 < P) {
        const int rem = P - g0;
        // a pass costs about its group size (the small groups a little more per point), so the binary decomposition of P
        if (GN >= 8 && rem >= 8) { sdb_wt_group<D, M, GN >= 8 ? 8 : 1>(partial, d, m, x, out, g0, lane); g0 += 8; }
        else if (GN >= 4 && rem >= 4) { sdb_wt_group<D, M, GN >= 4 ? 4 : 1>(partial, d, m, x, out, g0, lane); g0 += 4; }
        else if (rem >= 2) { sdb_wt_group<D, M, 2>(partial, d, m, x, out, g0, lane); g0 += 2; }
        else { sdb_wt_group<D, M, 1>(partial, d, m, x, out, g0, lane); g0 += 1; }
    }
}
// number of sdb_wt_group passes sdb_wt_points makes for P points
__host__ __device__ inline int sdb_wt_passes(int P, int D, int M) {
    const int GN = sdb_wt_group_max(D, M);
    int n = 0, g0 = 0;
    while (g0 < P) { const int rem = P - g0; g0 += (GN >= 8 && rem >= 8) ? 8 : (GN >= 4 && rem >= 4) ? 4 : rem >= 2 ? 2 : 1; ++n; }
    return n;
}

// resident FP64 tables
template <int D, int M>
__device__ __forceinline__ void sdb_rbf_points_wt(const sdb_smodel &s, int d, int m, const double *x, double *out, int P) {
    const int lane = threadIdx.x & 31;
    auto partial = [&](auto &xs, auto &acc, auto np) {
        sdb_rbf_partialn_st<D, M, 32, decltype(np)::value>(s, d, m, xs, acc, lane, 32);
    };
    sdb_wt_points<D, M>(partial, d, m, x, out, P);
    sdb_rbf_add_tail<D, M>(s, d, m, x, out);
}

// FP32 tables (see above): the per-lane sums are FP32, the tree and the tail FP64
template <int D, int M>
__device__ __forceinline__ void sdb_rbf32_points_wt(const sdb_smodel32 &s, int d, int m, const double *x, double *out, int P) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int lane = threadIdx.x & 31;
    auto partial = [&](auto &xs, auto &acc, auto np) {
        constexpr int NP = decltype(np)::value;
        float xf[NP][DD], af[NP][MM];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) xf[p][j] = (float)(xs[p][j] - s.shift[j]);
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) af[p][k] = 0.0f;
        }
        sdb_rbf32_partialn_st<D, M, 32, NP>(s, d, m, xf, af, lane, 32);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) acc[p][k] = (double)af[p][k];
        }
    };
    sdb_wt_points<D, M>(partial, d, m, x, out, P);
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = 0.0;
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) t = fma(x[j], s.tail[j * m + k], t);
            out[k] += s.tail[d * m + k] + t;
        }
}

// streamed tables: every group pass walks the ring once.  Full CTA, converged (barriers inside), so P is the CTA-uniform
// maximum over the launch's warps (lanes beyond a warp's own count carry a shadow point whose result is unused).
template <int D, int M>
__device__ __forceinline__ void sdb_rbf_points_wt_stream(sdb_ring &r, int d, int m, const double *x, double *out, int P) {
    constexpr int UD = D ? D : 1;
    constexpr int UM = M ? M : 1;
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int MM = M ? M : SDB_MAX_M;
    const int lane = threadIdx.x & 31;
    auto partial = [&](auto &xs, auto &acc, auto np) {
        for (int ci = 0; ci < r.nchunks; ++ci) {
            const int st = (int)(r.consumed & 1);
            sdb_mbar_wait(&r.full[st], (r.phase_bits >> st) & 1);
            r.phase_bits ^= 1 << st;
            const long long c0 = (long long)ci * r.chunk;
            sdb_smodel s;
            s.kind = SDB_MODEL_RBF; s.kernel_type = r.kernel_type;
            s.n_centers = (int)((r.n - c0) < r.chunk ? (r.n - c0) : r.chunk);
            s.centers = (const double *)(r.base + (size_t)st * (r.cbytes + r.wbytes));
            s.coefs = (const double *)(r.base + (size_t)st * (r.cbytes + r.wbytes) + r.cbytes);
            sdb_rbf_partialn_st<D, M, 32, decltype(np)::value>(s, d, m, xs, acc, lane, 32);
            ++r.consumed;
            __syncthreads();      // every warp is done with stage st
            if (r.issued < r.total) sdb_ring_issue(r, d, m);
        }
    };
    sdb_wt_points<D, M>(partial, d, m, x, out, P);
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) {
            double t = 0.0;
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) t = fma(x[j], r.tail[j * m + k], t);
            out[k] += r.tail[d * m + k] + t;
        }
}

// ---------------------------------------------------------------------------------------
// Toy model (examples/solvers/solver_examples.py:20-29), the reference's operation order with
// separately rounded operations; bit-identical to numpy given the same exp().
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double sdb_toy_G(const sdb_smodel &s, const double *x) {
    const double f = s.toy_f, L = s.toy_L, Mm = s.toy_M;
    const double k1 = exp(x[0]), k2 = exp(x[1]);
    const double D1 = __ddiv_rn(__dmul_rn(f, L), k2);
    const double C1 = __ddiv_rn(__dmul_rn(D1, k2), k1);
    const double MM2 = __dmul_rn(Mm, Mm), LL2 = __dmul_rn(L, L);
    const double a = __dmul_rn(__ddiv_rn(-f, __dmul_rn(2.0, k1)), MM2);
    const double b = __dmul_rn(C1, Mm);
    const double c = __dmul_rn(__ddiv_rn(f, __dmul_rn(2.0, k2)), MM2);
    const double e = __dmul_rn(D1, Mm);
    const double D2 = __dadd_rn(__dadd_rn(__dadd_rn(a, b), c), -e);
    const double g = __dmul_rn(__ddiv_rn(-f, __dmul_rn(2.0, k2)), LL2);
    return __dadd_rn(__dadd_rn(g, __dmul_rn(D1, L)), D2);
}

// ---------------------------------------------------------------------------------------
// Gaussian log-posterior (classes_SAMPLER.py:295-316); constants in shared memory.
// ---------------------------------------------------------------------------------------
struct sdb_sproblem {
    int d, m, prior_full, noise_full;
    const double *prior_mean, *prior_scale, *obs, *noise_scale;  // smem
    const int *prior_piv, *noise_piv;                            // smem
};

// -0.5 * v . C^{-1} v  with C given by its LU factors (row-major, unit lower) and LAPACK-style
// pivots (row i was swapped with row piv[i], i = 0..n-1).  Loops are written over the compile-time
// bound N so that for small N everything stays in registers.
template <int N>
__device__ __forceinline__ double sdb_quad_lu(const double *lu, const int *piv, int n, const double *v) {
    constexpr int UN = (N <= 4) ? N : 1;
    double b[N];
#pragma unroll UN
    for (int i = 0; i < N; ++i)
        if (i < n) b[i] = v[i];
#pragma unroll UN
    for (int i = 0; i < N; ++i)
        if (i < n) {
            const int p = piv[i];
#pragma unroll UN
            for (int q = 0; q < N; ++q)
                if (q > i && q == p) { const double t = b[i]; b[i] = b[q]; b[q] = t; }
        }
#pragma unroll UN
    for (int i = 1; i < N; ++i)            // L y = P v
        if (i < n) {
            double t = b[i];
#pragma unroll UN
            for (int j = 0; j < N; ++j)
                if (j < i) t = __dadd_rn(t, -__dmul_rn(lu[i * n + j], b[j]));
            b[i] = t;
        }
#pragma unroll UN
    for (int ii = 0; ii < N; ++ii) {       // U x = y
        const int i = N - 1 - ii;
        if (i < n) {
            double t = b[i];
#pragma unroll UN
            for (int j = 0; j < N; ++j)
                if (j > i && j < n) t = __dadd_rn(t, -__dmul_rn(lu[i * n + j], b[j]));
            b[i] = __ddiv_rn(t, lu[i * n + i]);
        }
    }
    double q = 0.0;
#pragma unroll UN
    for (int i = 0; i < N; ++i)
        if (i < n) q = __dadd_rn(q, __dmul_rn(v[i], b[i]));
    return -0.5 * q;
}

// log-likelihood and log-prior separately: the prior of a proposal is needed twice per DAMH step (surrogate and
// true posterior, classes_SAMPLER.py:189-191,79-80) and is computed once; ll + lp is rounded exactly as the
// reference's get_log_likelihood(G) + get_log_prior(u).
template <int D, int M>
__device__ __forceinline__ double sdb_log_likelihood(const sdb_sproblem &p, const double *G) {
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int m = M ? M : p.m;
    double v[MM];
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) v[k] = __dadd_rn(p.obs[k], -G[k]);
    if (p.noise_full) return sdb_quad_lu<MM>(p.noise_scale, p.noise_piv, m, v);
    double q = 0.0;
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) q = __dadd_rn(q, __dmul_rn(v[k], __ddiv_rn(v[k], p.noise_scale[k])));
    return -0.5 * q;
}

template <int D, int M>
__device__ __forceinline__ double sdb_log_prior(const sdb_sproblem &p, const double *u) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    const int d = D ? D : p.d;
    double v[DD];
#pragma unroll UD
    for (int j = 0; j < DD; ++j)
        if (j < d) v[j] = __dadd_rn(u[j], -p.prior_mean[j]);
    if (p.prior_full) return sdb_quad_lu<DD>(p.prior_scale, p.prior_piv, d, v);
    double q = 0.0;
#pragma unroll UD
    for (int j = 0; j < DD; ++j)
        if (j < d) q = __dadd_rn(q, __dmul_rn(v[j], __ddiv_rn(v[j], p.prior_scale[j])));
    return -0.5 * q;
}

template <int D, int M>
__device__ __forceinline__ double sdb_log_posterior(const sdb_sproblem &p, const double *u, const double *G) {
    return __dadd_rn(sdb_log_likelihood<D, M>(p, G), sdb_log_prior<D, M>(p, u));
}
