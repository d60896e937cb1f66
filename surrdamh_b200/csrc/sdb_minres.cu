// MINRES (Paige & Saunders 1975) for the symmetric indefinite RBF saddle-point system, on the device.
// Replaces scipy.sparse.linalg.minres at surrogate_rbf.py:131-135 -- the reference's DEFAULT solver
// (solver_type='minres', tol 1e-6).  Same recurrences, same stopping rules (istop 1/2 on
// test1/test2 <= rtol, 3 on epsx >= beta1, 4 on Acond >= 0.1/eps, 6 on the iteration limit, -1 when
// the Krylov space is exhausted after one step), no preconditioner, no shift.
//
// Per iteration two launches: a grid-wide symmetric mat-vec (one warp per row, HBM/L2-bound) and a
// single-CTA update that does the three reductions and the vector recurrences.  The host launches
// iterations in chunks and reads the stop flag between chunks (this entry point synchronises the
// stream); kernels of a chunk that start after convergence return immediately.
//
// Also here: the FP64 FMA throughput probe used by bench.py as the measured roofline denominator
// of the FP64-pipe-bound kernels.
#include <float.h>

#include "sdb_common.cuh"

struct minres_state {
    double oldb, beta, dbar, epsln, phibar, rhs1, rhs2, tnorm2, gmax, gmin, cs, sn, beta1;
    double rnorm, Anorm, Acond;
    int itn, istop;
};

// y = A v, A symmetric N x N (ld N); one warp per row.
__global__ void __launch_bounds__(256) minres_matvec_kernel(const double *__restrict__ A, int N, const double *__restrict__ v,
                                                            double *__restrict__ y, const minres_state *st) {
    if (st && st->istop != 0) return;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const double *a = A + (size_t)row * N;
    double s = 0.0;
    for (int j = lane; j < N; j += 32) s = fma(a[j], v[j], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[row] = s;
}

__device__ __forceinline__ double block_sum(double v, double *scratch) {
    // 1024 threads; result broadcast to every thread; deterministic order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = scratch[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    return t;
}

// vectors: x, v, y (also r2), r1, r2, w, w2, Av
__global__ void __launch_bounds__(1024, 1) minres_init_kernel(int N, const double *__restrict__ b, const double *__restrict__ Ax0,
                                                              const double *__restrict__ x0, double *x, double *v, double *y,
                                                              double *r1, double *r2, double *w, double *w2, minres_state *st) {
    __shared__ double scratch[32];
    double bb = 0.0, rr = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double r = Ax0 ? b[i] - Ax0[i] : b[i];
        r1[i] = r; r2[i] = r; y[i] = r; w[i] = 0.0; w2[i] = 0.0;
        x[i] = x0 ? x0[i] : 0.0;
        bb = fma(b[i], b[i], bb);
        rr = fma(r, r, rr);
    }
    bb = block_sum(bb, scratch);
    rr = block_sum(rr, scratch);
    const double beta1 = sqrt(rr);
    if (threadIdx.x == 0) {
        st->oldb = 0; st->beta = beta1; st->dbar = 0; st->epsln = 0; st->phibar = beta1; st->rhs1 = beta1; st->rhs2 = 0;
        st->tnorm2 = 0; st->gmax = 0; st->gmin = DBL_MAX; st->cs = -1; st->sn = 0; st->beta1 = beta1;
        st->rnorm = 0; st->Anorm = 0; st->Acond = 0; st->itn = 0;
        st->istop = (rr == 0.0 || bb == 0.0) ? 100 : 0;  // scipy returns x (= x0, or b when b == 0) at once
    }
    if (rr == 0.0 || bb == 0.0) {
        if (bb == 0.0) for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] = b[i];
        return;
    }
    const double s = 1.0 / beta1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) v[i] = s * y[i];
}

__global__ void __launch_bounds__(1024, 1) minres_update_kernel(int N, const double *__restrict__ Av, double *x, double *v, double *y,
                                                                double *r1, double *r2, double *w, double *w2, double rtol,
                                                                int maxiter, minres_state *st) {
    __shared__ double scratch[32];
    if (st->istop != 0) return;
    const double eps = DBL_EPSILON;
    const int itn = st->itn + 1;
    const double beta = st->beta, oldb_prev = st->oldb;
    // y = A v - (beta/oldb) r1 ; alfa = v.y
    double alfa = 0.0;
    const double c1 = itn >= 2 ? beta / oldb_prev : 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double yi = Av[i];
        if (itn >= 2) yi = yi - c1 * r1[i];
        y[i] = yi;
        alfa = fma(v[i], yi, alfa);
    }
    alfa = block_sum(alfa, scratch);
    const double c2 = alfa / beta;
    double b2 = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double yi = y[i] - c2 * r2[i];
        r1[i] = r2[i];
        r2[i] = yi;
        y[i] = yi;
        b2 = fma(yi, yi, b2);
    }
    b2 = block_sum(b2, scratch);
    const double oldb = beta;
    const double beta_new = sqrt(b2);
    const double tnorm2 = st->tnorm2 + alfa * alfa + oldb * oldb + beta_new * beta_new;
    int istop = 0;
    if (itn == 1 && beta_new / st->beta1 <= 10 * eps) istop = -1;
    const double oldeps = st->epsln;
    const double cs0 = st->cs, sn0 = st->sn, dbar0 = st->dbar;
    const double delta = cs0 * dbar0 + sn0 * alfa;
    const double gbar = sn0 * dbar0 - cs0 * alfa;
    const double epsln = sn0 * beta_new;
    const double dbar = -cs0 * beta_new;
    const double root = hypot(gbar, dbar);
    double gamma = hypot(gbar, beta_new);
    gamma = fmax(gamma, eps);
    const double cs = gbar / gamma, sn = beta_new / gamma;
    const double phi = cs * st->phibar;
    const double phibar = sn * st->phibar;
    const double denom = 1.0 / gamma;
    double xx = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double w1 = w2[i], w2o = w[i];
        const double wn = (v[i] - oldeps * w1 - delta * w2o) * denom;
        w2[i] = w2o;
        w[i] = wn;
        const double xi = x[i] + phi * wn;
        x[i] = xi;
        xx = fma(xi, xi, xx);
    }
    xx = block_sum(xx, scratch);
    const double gmax = fmax(st->gmax, gamma), gmin = fmin(st->gmin, gamma);
    const double z = st->rhs1 / gamma;
    const double rhs1 = st->rhs2 - delta * z, rhs2 = -epsln * z;
    const double Anorm = sqrt(tnorm2), ynorm = sqrt(xx);
    const double epsx = Anorm * ynorm * eps;
    const double rnorm = phibar;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double test1 = (ynorm == 0.0 || Anorm == 0.0) ? inf : rnorm / (Anorm * ynorm);
    const double test2 = (Anorm == 0.0) ? inf : root / Anorm;
    const double Acond = gmax / gmin;
    if (istop == 0) {
        const double t1 = 1 + test1, t2 = 1 + test2;
        if (t2 <= 1) istop = 2;
        if (t1 <= 1) istop = 1;
        if (itn >= maxiter) istop = 6;
        if (Acond >= 0.1 / eps) istop = 4;
        if (epsx >= st->beta1) istop = 3;
        if (test2 <= rtol) istop = 2;
        if (test1 <= rtol) istop = 1;
    }
    if (istop == 0) {
        const double s = 1.0 / beta_new;
        for (int i = threadIdx.x; i < N; i += blockDim.x) v[i] = s * y[i];
    }
    __syncthreads();  // everyone has read the old state
    if (threadIdx.x == 0) {
        st->oldb = oldb; st->beta = beta_new; st->dbar = dbar; st->epsln = epsln; st->phibar = phibar;
        st->rhs1 = rhs1; st->rhs2 = rhs2; st->tnorm2 = tnorm2; st->gmax = gmax; st->gmin = gmin; st->cs = cs; st->sn = sn;
        st->rnorm = rnorm; st->Anorm = Anorm; st->Acond = Acond; st->itn = itn; st->istop = istop;
    }
}

// Solve A x = b (A symmetric N x N).  d_x0 may be NULL.  d_work: >= 8*N + 64 doubles.
// d_info (int32[4]): [0] = 0 converged / maxiter if the limit was hit (scipy's info), [1] = iterations,
// [2] = istop.  Synchronises `stream`.
extern "C" int sdb_minres(const double *d_A, int64_t N, const double *d_b, const double *d_x0, double *d_x, double rtol,
                          int64_t maxiter, double *d_work, int32_t *d_info, void *stream) {
    SDB_CHECK_ARG(d_A && d_b && d_x && d_work && d_info, "sdb_minres: NULL pointer");
    SDB_CHECK_ARG(N >= 1 && N < (1 << 30) && maxiter >= 1, "sdb_minres: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = (int)N;
    double *v = d_work, *y = v + N, *r1 = y + N, *r2 = r1 + N, *w = r2 + N, *w2 = w + N, *Av = w2 + N;
    minres_state *state = (minres_state *)(Av + N);
    const int mv_blocks = (n + 7) / 8;
    if (d_x0) { sdb_count_launch(); minres_matvec_kernel<<<mv_blocks, 256, 0, st>>>(d_A, n, d_x0, Av, nullptr); }
    sdb_count_launch();
    minres_init_kernel<<<1, 1024, 0, st>>>(n, d_b, d_x0 ? Av : nullptr, d_x0, d_x, v, y, r1, r2, w, w2, state);
    SDB_CUDA(cudaGetLastError());
    minres_state h;
    int64_t done = 0;
    const int chunk = 32;
    for (;;) {
        SDB_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
        SDB_CUDA(cudaStreamSynchronize(st));
        if (h.istop != 0 || done >= maxiter) break;
        for (int k = 0; k < chunk && done < maxiter; ++k, ++done) {
            sdb_count_launch();
            minres_matvec_kernel<<<mv_blocks, 256, 0, st>>>(d_A, n, v, Av, state);
            sdb_count_launch();
            minres_update_kernel<<<1, 1024, 0, st>>>(n, Av, d_x, v, y, r1, r2, w, w2, rtol, (int)maxiter, state);
        }
        SDB_CUDA(cudaGetLastError());
    }
    int32_t info[4] = {h.istop == 6 ? (int32_t)maxiter : 0, h.itn, h.istop, 0};
    SDB_CUDA(cudaMemcpyAsync(d_info, info, sizeof(info), cudaMemcpyHostToDevice, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// FP64 FMA throughput probe: every thread runs 8 independent DFMA chains of `iters` steps.
// FLOPs = blocks * 256 * 8 * 2 * iters.  d_out: one double per thread (keeps the chains alive).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(int64_t iters, double *out) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Same, but every FMA reads three DIFFERENT live registers per chain (a_k = fma(a_k, b_k, c_k), b_k and c_k loaded from
// memory), the operand pattern of real FP64 code; tells register-file limits apart from pipe limits.
__global__ void __launch_bounds__(256) fp64_probe3_kernel(int64_t iters, double *out) {
    double a[8], b[8], c[8];
    const size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = 1.0 + k;
        b[k] = 1.0 + out[base] * 1e-30 + k * 1e-9;      // opaque to the compiler
        c[k] = 1e-7 + out[base] * 1e-30 + k * 1e-9;
    }
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b[k], c[k]);
    }
    out[base] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

extern "C" int sdb_fp64_probe3(int64_t iters, int32_t blocks, double *d_out, void *stream) {
    SDB_CHECK_ARG(iters >= 1 && blocks >= 1 && d_out, "sdb_fp64_probe3: bad arguments");
    sdb_count_launch();
    fp64_probe3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_out);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

extern "C" int sdb_fp64_probe(int64_t iters, int32_t blocks, double *d_out, void *stream) {
    SDB_CHECK_ARG(iters >= 1 && blocks >= 1 && d_out, "sdb_fp64_probe: bad arguments");
    sdb_count_launch();
    fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_out);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}
