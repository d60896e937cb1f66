// Stand-alone surrogate evaluation kernels: points [N x d] row-major -> out [N x m] row-major.
//   sdb_poly_eval  replaces surrogate_poly.py:18-30  (Surrogate_apply.apply)
//   sdb_rbf_eval   replaces surrogate_rbf.py:20-40   (Surrogate_apply.apply)
// The RBF kernel streams the centers / weights through a two-stage shared-memory ring filled
// by TMA bulk copies (cp.async.bulk + mbarrier), so the number of centers is not limited by
// shared memory; each point is handled by L lanes of a warp (L chosen so the grid fills the GPU).
#include "sdb_models.cuh"

// ---------------------------------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(256) poly_eval_kernel(const double *__restrict__ pts, int64_t N, int d, int m,
                                                        const sdb_model mo, double *__restrict__ out) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *p = smem + 16;
    if (threadIdx.x == 0) {
        sdb_mbar_init(bar, 1);
        sdb_mbar_fence_init();
    }
    __syncthreads();
    const sdb_smodel s = sdb_stage_model(mo, d, m, p, bar);
    if (threadIdx.x == 0) sdb_mbar_arrive(bar);
    sdb_mbar_wait(bar, 0);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double x[DD], o[MM];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) x[j] = pts[i * d + j];
        sdb_poly_point<D, M>(s, d, m, x, o);
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) out[i * m + k] = o[k];
    }
}

extern "C" int sdb_poly_eval(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model,
                             double *d_out, void *stream) {
    SDB_CHECK_ARG(model && model->kind == SDB_MODEL_POLY, "sdb_poly_eval: model must be SDB_MODEL_POLY");
    SDB_CHECK_ARG(d_points && d_out && model->coefs && model->alpha && model->herm, "sdb_poly_eval: NULL pointer");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D && m >= 1 && m <= SDB_MAX_M, "sdb_poly_eval: d=%d m=%d outside limits", d, m);
    SDB_CHECK_LIMIT(model->degree >= 0 && model->degree <= SDB_MAX_DEGREE, "sdb_poly_eval: degree %d outside [0,%d]",
                    model->degree, SDB_MAX_DEGREE);
    if (N <= 0) return SDB_OK;
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    const size_t smem = 16 + sdb_model_smem_bytes(*model, d, m);
    SDB_CHECK_LIMIT(smem <= (size_t)dp.smem_optin, "sdb_poly_eval: tables need %zu B of shared memory", smem);
    const int threads = 128;
    int64_t blocks = (N + threads - 1) / threads;
    if (blocks > (int64_t)dp.sm_count * 16) blocks = (int64_t)dp.sm_count * 16;
    auto launch = [&](auto kern) -> int {
        SDB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sdb_count_launch();
        kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(d_points, N, d, m, *model, d_out);
        SDB_CUDA(cudaGetLastError());
        return SDB_OK;
    };
    if (d == 2 && m == 1) return launch(poly_eval_kernel<2, 1>);
    if (d == 4 && m == 3) return launch(poly_eval_kernel<4, 3>);
    return launch(poly_eval_kernel<0, 0>);
}

// ---------------------------------------------------------------------------------------------------
// RBF: two-stage ring of center chunks.
// ---------------------------------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(256) rbf_eval_kernel(const double *__restrict__ pts, int64_t N, int d, int m,
                                                       const sdb_model mo, double *__restrict__ out, int L, int chunk) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = (uint64_t *)smem;  // full[0], full[1]
    double *tail = (double *)(smem + 16);  // linear tail rows [(d+1) x m]
    const size_t tail_bytes = sdb_align16((size_t)(d + 1) * m * 8);
    const size_t cbytes = sdb_align16((size_t)chunk * d * 8), wbytes = sdb_align16((size_t)chunk * m * 8);
    unsigned char *ring = smem + 16 + tail_bytes;
    const int64_t n = mo.n_centers;
    const int nchunks = (int)((n + chunk - 1) / chunk);

    if (threadIdx.x == 0) {
        sdb_mbar_init(&full[0], 1);
        sdb_mbar_init(&full[1], 1);
        sdb_mbar_fence_init();
    }
    for (int i = threadIdx.x; i < (d + 1) * m; i += blockDim.x) tail[i] = mo.coefs[n * m + i];
    __syncthreads();

    // Called by every thread: thread 0 drives the TMA, a few threads copy the <16 B tails.
    // `ci` is the chunk index inside a sweep, `st` the ring stage it lands in.
    auto issue = [&](int ci, int st) {
        const int64_t c0 = (int64_t)ci * chunk;
        const int cn = (int)((n - c0) < chunk ? (n - c0) : chunk);
        unsigned char *sc = ring + (size_t)st * (cbytes + wbytes);
        sdb_stage_table(sc, mo.centers + c0 * d, (size_t)cn * d * 8, &full[st]);
        sdb_stage_table(sc + cbytes, mo.coefs + c0 * m, (size_t)cn * m * 8, &full[st]);
        if (threadIdx.x == 0) sdb_mbar_arrive(&full[st]);
    };

    const int lanes_pts = blockDim.x / L;  // points per CTA per sweep
    const int sub = threadIdx.x % L;
    const int64_t sweeps = (N + (int64_t)gridDim.x * lanes_pts - 1) / ((int64_t)gridDim.x * lanes_pts);
    // One sweep over all chunks per tile of points; the host sizes the grid so that sweeps == 1 in
    // the common case (re-streaming the centers from L2 for every further sweep).
    const int64_t total = sweeps * nchunks;  // chunks streamed by this CTA over all sweeps
    int64_t issued = 0;                      // the q-th chunk overall lives in stage q & 1
    for (; issued < 2 && issued < total; ++issued) issue((int)(issued % nchunks), (int)(issued & 1));
    int phase_bits = 0;   // parity of full[0], full[1]
    int64_t consumed = 0; // chunks consumed so far over all sweeps (ring position)
    for (int64_t sw = 0; sw < sweeps; ++sw) {
        __syncthreads();  // tail words written by ordinary stores in issue() become visible
        const int64_t pi = (sw * gridDim.x + blockIdx.x) * lanes_pts + threadIdx.x / L;
        const bool live = pi < N;
        const int64_t pc = live ? pi : (N - 1);
        double x[DD], acc[MM];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) x[j] = pts[pc * d + j];
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) acc[k] = 0.0;
        for (int ci = 0; ci < nchunks; ++ci, ++consumed) {
            const int st = (int)(consumed & 1);
            sdb_mbar_wait(&full[st], (phase_bits >> st) & 1);
            phase_bits ^= 1 << st;
            const int64_t c0 = (int64_t)ci * chunk;
            sdb_smodel s;
            s.kind = SDB_MODEL_RBF; s.kernel_type = mo.kernel_type;
            s.n_centers = (int)((n - c0) < chunk ? (n - c0) : chunk);
            s.centers = (const double *)(ring + (size_t)st * (cbytes + wbytes));
            s.coefs = (const double *)(ring + (size_t)st * (cbytes + wbytes) + cbytes);
            sdb_rbf_partial<D, M>(s, d, m, x, acc, sub, L);
            __syncthreads();  // everyone is done with stage st
            // refill the ring with the next chunk not yet in flight (wrapping into the next sweep);
            // its stage issued & 1 is free: either st, or -- single-chunk models -- the other, idle one
            if (issued < total) {
                issue((int)(issued % nchunks), (int)(issued & 1));
                ++issued;
            }
        }
        for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) acc[k] += sdb_shfl_xor(acc[k], off);
        }
        if (live && sub == 0) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) {
                    double t = 0.0;
#pragma unroll UD
                    for (int j = 0; j < DD; ++j)
                        if (j < d) t = fma(x[j], tail[j * m + k], t);
                    out[pi * m + k] = acc[k] + (tail[d * m + k] + t);
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// RBF, many points against tables that fit in shared memory: the chain kernel's warp-owned evaluation (sdb_rbf_points_wt) as
// a stand-alone kernel.  A persistent grid of one 16-warp CTA per SM stages the tables once (TMA bulk copies); a warp takes
// 32 consecutive points at a time, one per lane, and all 32 lanes split the centers of every point (passes of 8 points: a center /
// weight load serves 8 pairs).  Round 2: 0.27 -> 0.40 of the FP64 peak at 1 M points x 4096 centers (profiles/r02_eval_kernels.md).
// ---------------------------------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(512, 1) rbf_eval_wt_kernel(const double *__restrict__ pts, int64_t N, int d_rt, int m_rt,
                                                             const sdb_model mo, double *__restrict__ out) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    const int d = D ? D : d_rt, m = M ? M : m_rt;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *p = smem + 16;
    if (threadIdx.x == 0) {
        sdb_mbar_init(bar, 1);
        sdb_mbar_fence_init();
    }
    __syncthreads();
    const sdb_smodel sur = sdb_stage_model(mo, d, m, p, bar);
    if (threadIdx.x == 0) sdb_mbar_arrive(bar);
    sdb_mbar_wait(bar, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (N + 31) / 32, n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += n_warps) {
        const int64_t p0 = t * 32;
        const int P = (int)((N - p0) < 32 ? (N - p0) : 32);
        const int64_t pi = p0 + (lane < P ? lane : P - 1);
        double x[DD], S[MM];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) x[j] = pts[pi * d + j];
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) S[k] = 0.0;
        sdb_rbf_points_wt<D, M>(sur, d, m, x, S, P);
        if (lane < P) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) out[pi * m + k] = S[k];
        }
    }
}

extern "C" int sdb_rbf_eval(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model,
                            double *d_out, void *stream) {
    SDB_CHECK_ARG(model && model->kind == SDB_MODEL_RBF, "sdb_rbf_eval: model must be SDB_MODEL_RBF");
    SDB_CHECK_ARG(d_points && d_out && model->coefs && model->centers, "sdb_rbf_eval: NULL pointer");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D && m >= 1 && m <= SDB_MAX_M, "sdb_rbf_eval: d=%d m=%d outside limits", d, m);
    SDB_CHECK_ARG(model->n_centers >= 1, "sdb_rbf_eval: n_centers must be >= 1");
    SDB_CHECK_ARG(model->kernel_type >= 0 && model->kernel_type <= 7, "sdb_rbf_eval: kernel_type %d outside [0,7]",
                  model->kernel_type);
    SDB_CHECK_ARG((((uintptr_t)model->centers) & 15) == 0 && (((uintptr_t)model->coefs) & 15) == 0,
                  "sdb_rbf_eval: centers / coefs must be 16-byte aligned");
    if (N <= 0) return SDB_OK;
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    {
        // enough points to give every warp of a persistent grid full 32-point tiles, tables resident: the warp-owned kernel
        static const bool no_wt = getenv("SDB_NO_WT") != nullptr;
        const size_t need = 16 + sdb_model_smem_bytes(*model, d, m);
        if (!no_wt && need <= (size_t)dp.smem_optin && N >= (int64_t)dp.sm_count * 16 * 32 && model->n_centers >= 256) {
            auto launch_wt = [&](auto kern) -> int {
                SDB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
                sdb_count_launch();
                kern<<<dp.sm_count, 512, need, (cudaStream_t)stream>>>(d_points, N, d, m, *model, d_out);
                SDB_CUDA(cudaGetLastError());
                return SDB_OK;
            };
            if (d == 2 && m == 1) return launch_wt(rbf_eval_wt_kernel<2, 1>);
            if (d == 4 && m == 3) return launch_wt(rbf_eval_wt_kernel<4, 3>);
            return launch_wt(rbf_eval_wt_kernel<0, 0>);
        }
    }
    const int threads = 256;
    // lanes per point: fill >= 2 CTAs per SM before going wider than one lane per point
    int L = 1;
    while (L < 32 && N * L < (int64_t)dp.sm_count * threads * 2) L <<= 1;
    const int pts_per_cta = threads / L;
    int64_t blocks = (N + pts_per_cta - 1) / pts_per_cta;
    const int64_t max_blocks = (int64_t)dp.sm_count * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    // chunk: two stages of (d+m) doubles per center in <= ~48 KiB per stage, even count
    int64_t chunk = (48 * 1024) / ((int64_t)(d + m) * 8);
    if (chunk > model->n_centers) chunk = model->n_centers;
    chunk = (chunk + 1) & ~(int64_t)1;
    const size_t smem = 16 + sdb_align16((size_t)(d + 1) * m * 8) +
                        2 * (sdb_align16((size_t)chunk * d * 8) + sdb_align16((size_t)chunk * m * 8));
    auto launch = [&](auto kern) -> int {
        SDB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sdb_count_launch();
        kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(d_points, N, d, m, *model, d_out, L, (int)chunk);
        SDB_CUDA(cudaGetLastError());
        return SDB_OK;
    };
    if (d == 2 && m == 1) return launch(rbf_eval_kernel<2, 1>);
    if (d == 4 && m == 3) return launch(rbf_eval_kernel<4, 3>);
    return launch(rbf_eval_kernel<0, 0>);
}

// ---------------------------------------------------------------------------------------------------
// FP32 evaluation (additive: Surrogate_apply(..., precision="f32")): the arithmetic the chain kernel uses with
// surrogate_f32, as a stand-alone kernel -- tables resident in shared memory as FP32, one point per lane, four
// neighbouring lanes share their center loads.
// ---------------------------------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(256) rbf_eval32_kernel(const double *__restrict__ pts, int64_t N, int d, int m,
                                                         const sdb_model mo, double *__restrict__ out) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *p = smem;
    const sdb_smodel32 s = sdb_stage_model32(mo, d, m, p);
    const int64_t per_sweep = (int64_t)gridDim.x * blockDim.x;
    const int64_t sweeps = (N + per_sweep - 1) / per_sweep;
    for (int64_t sw = 0; sw < sweeps; ++sw) {        // whole warps stay converged: idle tail lanes shadow the last point
        const int64_t i = sw * per_sweep + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t pc = i < N ? i : N - 1;
        double x[DD], o[MM];
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) x[j] = pts[pc * d + j];
        if (D && M && D <= 4 && M <= 3) sdb_rbf32_point_team<D, M, 4>(s, d, m, x, o, 1);
        else sdb_rbf32_point_team<D, M, 2>(s, d, m, x, o, 1);
        if (i < N) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) out[i * m + k] = o[k];
        }
    }
}

extern "C" int sdb_rbf_eval_f32(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model,
                                double *d_out, void *stream) {
    SDB_CHECK_ARG(model && model->kind == SDB_MODEL_RBF, "sdb_rbf_eval_f32: model must be SDB_MODEL_RBF");
    SDB_CHECK_ARG(d_points && d_out && model->coefs && model->centers, "sdb_rbf_eval_f32: NULL pointer");
    SDB_CHECK_LIMIT(d >= 1 && d <= SDB_MAX_D && m >= 1 && m <= SDB_MAX_M, "sdb_rbf_eval_f32: d=%d m=%d outside limits", d, m);
    SDB_CHECK_ARG(model->n_centers >= 1, "sdb_rbf_eval_f32: n_centers must be >= 1");
    SDB_CHECK_ARG(model->kernel_type >= 0 && model->kernel_type <= 7, "sdb_rbf_eval_f32: kernel_type %d outside [0,7]",
                  model->kernel_type);
    if (N <= 0) return SDB_OK;
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    const size_t smem = sdb_model32_smem_bytes(*model, d, m);
    SDB_CHECK_LIMIT(smem <= (size_t)dp.smem_optin, "sdb_rbf_eval_f32: FP32 tables need %zu B of shared memory (%d available)", smem,
                    dp.smem_optin);
    const int threads = 256;
    int64_t blocks = (N + threads - 1) / threads;
    if (blocks > (int64_t)dp.sm_count * 4) blocks = (int64_t)dp.sm_count * 4;
    auto launch = [&](auto kern) -> int {
        SDB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sdb_count_launch();
        kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(d_points, N, d, m, *model, d_out);
        SDB_CUDA(cudaGetLastError());
        return SDB_OK;
    };
    if (d == 2 && m == 1) return launch(rbf_eval32_kernel<2, 1>);
    if (d == 4 && m == 3) return launch(rbf_eval32_kernel<4, 3>);
    return launch(rbf_eval32_kernel<0, 0>);
}

// FP32 FMA throughput probe (roofline denominator of the FP32-surrogate chain kernel): 8 independent FFMA chains per
// thread; FLOPs = blocks * 256 * 8 * 2 * iters.  d_out: DEVICE floats [blocks * 256].
__global__ void __launch_bounds__(256) fp32_probe_kernel(long long iters, float *out) {
    float a0 = 1.0f + threadIdx.x * 1e-7f, a1 = a0 + 1e-3f, a2 = a0 + 2e-3f, a3 = a0 + 3e-3f, a4 = a0 + 4e-3f, a5 = a0 + 5e-3f,
          a6 = a0 + 6e-3f, a7 = a0 + 7e-3f;
    const float b = 0.9999999f, c = 1e-9f;
    for (long long i = 0; i < iters; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int sdb_fp32_probe(int64_t iters, int32_t blocks, float *d_out, void *stream) {
    SDB_CHECK_ARG(iters >= 1 && blocks >= 1 && d_out, "sdb_fp32_probe: bad arguments");
    sdb_count_launch();
    fp32_probe_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((long long)iters, d_out);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}
