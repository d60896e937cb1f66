// Host side of the lockstep chain kernels (validation, launch plan, C ABI) + the lean kernel of the toy shape.
// The fused kernel template lives in sdb_chain_kernel.cuh; its FP32-surrogate and streamed-surrogate instantiations
// are compiled in sdb_chain_f32.cu / sdb_chain_stream.cu.
#include "sdb_chain_kernel.cuh"

// ---------------------------------------------------------------------------------------------------------------
// Lean fused kernel for the reference's own toy shape (simple.json / BASELINE configs[0..1]): d = 2, m = 1, polynomial
// surrogate of degree <= 6 (or plain MH), analytic toy / zero truth, one thread per chain.  Few chains cannot fill
// the GPU, so a step is bound by the LATENCY of its dependent chain; this kernel shortens that chain:
//   * the random numbers of step k+1 (Philox, Box-Muller, both log-uniforms) do not depend on the chain state and
//     are computed during step k (software pipelining),
//   * the Hermite-Vandermonde evaluation is fully unrolled with the tables in registers / broadcast shared memory
//     (terms in nested-loop order; the coefficient column is permuted once per CTA to match),
//   * the cheap truth and the prior are evaluated for every proposal, next to the surrogate, not after its test.
// Same arithmetic per quantity as chain_kernel (same device functions for the likelihood, prior and toy model),
// so decisions and trace values are identical; only the order of the PHI.MAT accumulation differs (as BLAS's does).
// ---------------------------------------------------------------------------------------------------------------
struct lean_rng {
    double z0, z1, lu0, lu1;
};

__device__ __forceinline__ lean_rng lean_draw(const sdb_chain_args &a, uint64_t gchain, int64_t c, int64_t C, int k) {
    lean_rng r;
    double u0, u1;
    if (a.rng_kind == SDB_RNG_STREAMS) {
        r.z0 = a.z[((size_t)k * 2 + 0) * C + c];
        r.z1 = a.z[((size_t)k * 2 + 1) * C + c];
        u0 = a.uni[((size_t)k * 2 + 0) * C + c];
        u1 = a.uni[((size_t)k * 2 + 1) * C + c];
    } else {
        const uint64_t gstep = (uint64_t)(a.step0 + k);
        sdb_normal_pair(sdb_chain_block(a.seed, a.stream_id, gchain, gstep, 0u), r.z0, r.z1);
        const philox4 b = sdb_chain_block(a.seed, a.stream_id, gchain, gstep, SDB_SLOT_UNIFORMS);
        u0 = sdb_u01(b.x, b.y);
        u1 = sdb_u01(b.z, b.w);
    }
    r.lu0 = log(u0);
    r.lu1 = log(u1);
    return r;
}

template <int DEG, bool DAMH>
__global__ void __launch_bounds__(128) chain_lean_kernel(const sdb_chain_args a) {
    constexpr int NH = DEG + 1;
    constexpr int P = (DEG + 1) * (DEG + 2) / 2;
    const int64_t C = a.state.n_chains;
    __shared__ __align__(16) unsigned char lsm[1024];
    __shared__ double herm[NH * NH];
    __shared__ double matc[P];          // MAT in nested-loop order (a0 outer, a1 inner)
    unsigned char *p = lsm;
    sdb_sproblem pr;
    pr.d = 2; pr.m = 1; pr.prior_full = a.problem.prior_full; pr.noise_full = 0;
    {
        size_t b;
        b = 16; pr.prior_mean = (const double *)p; copy_small(p, a.problem.prior_mean, b); p += 16;
        b = (size_t)(pr.prior_full ? 4 : 2) * 8; pr.prior_scale = (const double *)p; copy_small(p, a.problem.prior_scale, b); p += 32;
        pr.prior_piv = (const int *)p; copy_small(p, a.problem.prior_piv, 8); p += 16;
        pr.obs = (const double *)p; copy_small(p, a.problem.obs, 8); p += 16;
        pr.noise_scale = (const double *)p; copy_small(p, a.problem.noise_scale, 8); p += 16;
        pr.noise_piv = nullptr;
    }
    const double *prop_shared = (const double *)p;
    if (a.prop_kind == SDB_PROP_SHARED) copy_small(p, a.prop_scale, 16);
    if (DAMH) {
        for (int e = threadIdx.x; e < NH * NH; e += blockDim.x) herm[e] = a.surrogate.herm[e];
        if (threadIdx.x < P) {           // canonical index -> (a0, a1) -> row of the reference's multi-index table
            int t = threadIdx.x, a0 = 0;
            while (t > DEG - a0) { t -= DEG - a0 + 1; ++a0; }
            const int a1 = t;
            double v = 0.0;
            for (int row = 0; row < a.surrogate.n_terms; ++row)
                if (a.surrogate.alpha[row * 2] == a0 && a.surrogate.alpha[row * 2 + 1] == a1) v = a.surrogate.coefs[row];
            matc[threadIdx.x] = v;
        }
    }
    __syncthreads();
    sdb_smodel tru;
    tru.kind = a.truth.kind; tru.toy_f = a.truth.toy_f; tru.toy_L = a.truth.toy_L; tru.toy_M = a.truth.toy_M;

    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const uint64_t gchain = (uint64_t)(a.chain0 + c);
    double u0 = a.state.u[c], u1 = a.state.u[C + c], Gu = a.state.G_u[c];
    double post_u = a.state.post_u[c], pre_u = a.state.pre_u[c];
    int n_acc = a.state.counters[0 * C + c], n_rej = a.state.counters[1 * C + c];
    int n_pre = a.state.counters[2 * C + c], rej_cur = a.state.counters[3 * C + c];
    const bool do_mom = a.state.moments != nullptr;
    double mom[5] = {0, 0, 0, 0, 0};
    if (do_mom)
#pragma unroll
        for (int i = 0; i < 5; ++i) mom[i] = a.state.moments[(size_t)i * C + c];
    const double sig0 = a.prop_kind == SDB_PROP_PER_CHAIN ? a.prop_scale[c] : prop_shared[0];
    const double sig1 = a.prop_kind == SDB_PROP_PER_CHAIN ? a.prop_scale[C + c] : prop_shared[1];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    auto surrogate_post = [&](const double *x, double lp) -> double {
        // normalised-Hermite Vandermonde x MAT, the reference's per-value operation order (surrogate_poly.py:162-175,26-29)
        double h0[NH], h1[NH];
        {
            double pw[NH];
            pw[0] = 1.0;
#pragma unroll
            for (int i = 1; i <= DEG; ++i) pw[i] = __dmul_rn(pw[i - 1], x[0]);
#pragma unroll
            for (int k = 0; k <= DEG; ++k) {
                double v = herm[k * NH];
#pragma unroll
                for (int i = 1; i <= k; ++i) v = __dadd_rn(v, __dmul_rn(herm[k * NH + i], pw[i]));
                h0[k] = v;
            }
            pw[0] = 1.0;
#pragma unroll
            for (int i = 1; i <= DEG; ++i) pw[i] = __dmul_rn(pw[i - 1], x[1]);
#pragma unroll
            for (int k = 0; k <= DEG; ++k) {
                double v = herm[k * NH];
#pragma unroll
                for (int i = 1; i <= k; ++i) v = __dadd_rn(v, __dmul_rn(herm[k * NH + i], pw[i]));
                h1[k] = v;
            }
        }
        double Sp[3] = {0.0, 0.0, 0.0};     // three interleaved partial sums: a third of the dependent FMA chain
        int t = 0;
#pragma unroll
        for (int a0 = 0; a0 <= DEG; ++a0)
#pragma unroll
            for (int a1 = 0; a1 <= DEG - a0; ++a1) {
                Sp[t % 3] = fma(__dmul_rn(h0[a0], h1[a1]), matc[t], Sp[t % 3]);
                ++t;
            }
        const double S = (Sp[0] + Sp[1]) + Sp[2];
        return __dadd_rn(sdb_log_likelihood<2, 1>(pr, &S), lp);
    };

    if (DAMH && a.refresh_pre) {
        const double x[2] = {u0, u1};
        pre_u = surrogate_post(x, sdb_log_prior<2, 1>(pr, x));
    }

    const int K = a.n_steps;
    lean_rng nxt = lean_draw(a, gchain, c, C, 0);
    for (int k = 0; k < K; ++k) {
        const lean_rng cur = nxt;
        if (k + 1 < K) nxt = lean_draw(a, gchain, c, C, k + 1);   // independent of the chain state: overlaps this step
        const size_t row = (size_t)k * C + c;
        if (a.pre_cur) a.pre_cur[row] = pre_u;
        double up[2];
        up[0] = __dadd_rn(u0, __dmul_rn(sig0, cur.z0));
        up[1] = __dadd_rn(u1, __dmul_rn(sig1, cur.z1));
        const double lp_p = sdb_log_prior<2, 1>(pr, up);
        double Gp = 0.0;
        if (tru.kind == SDB_MODEL_TOY) Gp = sdb_toy_G(tru, up);
        const double post_p = __dadd_rn(sdb_log_likelihood<2, 1>(pr, &Gp), lp_p);
        double pre_p = qnan;
        bool pass1 = true;
        if (DAMH) {
            pre_p = surrogate_post(up, lp_p);
            pass1 = cur.lu0 < __dadd_rn(pre_p, -pre_u);
        }
        const double log_ratio = __dadd_rn(post_p, -post_u);
        const bool acc = pass1 && (DAMH ? (cur.lu1 < __dadd_rn(log_ratio, -__dadd_rn(pre_p, -pre_u))) : (cur.lu0 < log_ratio));
        const int flag = !pass1 ? SDB_FLAG_PREREJECTED : (acc ? SDB_FLAG_ACCEPTED : SDB_FLAG_REJECTED);
        if (a.flags) a.flags[row] = (uint8_t)flag;
        if (a.prop) {
            a.prop[((size_t)k * 2 + 0) * C + c] = up[0];
            a.prop[((size_t)k * 2 + 1) * C + c] = up[1];
        }
        if (a.post) a.post[row] = pass1 ? post_p : qnan;
        if (a.pre) a.pre[row] = pre_p;
        if (a.snap_par && pass1) {       // displaced current sample on accept, the proposal on reject (:85,:98)
            a.snap_par[((size_t)k * 2 + 0) * C + c] = acc ? u0 : up[0];
            a.snap_par[((size_t)k * 2 + 1) * C + c] = acc ? u1 : up[1];
            a.snap_obs[row] = acc ? Gu : Gp;
        }
        n_acc += acc;
        n_rej += (pass1 && !acc);
        n_pre += !pass1;
        rej_cur = acc ? 0 : rej_cur + 1;
        u0 = acc ? up[0] : u0;
        u1 = acc ? up[1] : u1;
        Gu = acc ? Gp : Gu;
        post_u = acc ? post_p : post_u;
        if (DAMH) pre_u = acc ? pre_p : pre_u;
        if (do_mom) {
            mom[0] += u0; mom[1] += u1;
            mom[2] = fma(u0, u0, mom[2]); mom[3] = fma(u0, u1, mom[3]); mom[4] = fma(u1, u1, mom[4]);
        }
    }
    a.state.u[c] = u0; a.state.u[C + c] = u1; a.state.G_u[c] = Gu;
    a.state.post_u[c] = post_u; a.state.pre_u[c] = pre_u;
    a.state.counters[0 * C + c] = n_acc; a.state.counters[1 * C + c] = n_rej;
    a.state.counters[2 * C + c] = n_pre; a.state.counters[3 * C + c] = rej_cur;
    if (do_mom)
#pragma unroll
        for (int i = 0; i < 5; ++i) a.state.moments[(size_t)i * C + c] = mom[i];
}

static bool lean_applicable(const sdb_chain_args *a) {
    if (getenv("SDB_NO_LEAN")) return false;
    if (a->phase != SDB_PHASE_FUSED || a->problem.d != 2 || a->problem.m != 1 || a->problem.noise_full) return false;
    if (a->lanes_per_chain > 1 || a->coresident || a->chain_lo != 0 || a->chain_n != 0) return false;
    if (a->prop_kind != SDB_PROP_SHARED && a->prop_kind != SDB_PROP_PER_CHAIN) return false;
    if (a->truth.kind != SDB_MODEL_TOY && a->truth.kind != SDB_MODEL_ZERO) return false;
    if (a->algorithm == SDB_ALG_DAMH)
        return a->surrogate.kind == SDB_MODEL_POLY && a->surrogate.degree >= 1 && a->surrogate.degree <= 6;
    return true;
}

template <int DEG>
static void lean_launch_deg(const sdb_chain_args *a, int blocks, cudaStream_t st) {
    if (a->algorithm == SDB_ALG_DAMH) chain_lean_kernel<DEG, true><<<blocks, 128, 0, st>>>(*a);
    else chain_lean_kernel<1, false><<<blocks, 128, 0, st>>>(*a);
}

static int lean_launch(const sdb_chain_args *a, cudaStream_t st) {
    const int64_t C = a->state.n_chains;
    const int blocks = (int)((C + 127) / 128);
    sdb_count_launch();
    switch (a->algorithm == SDB_ALG_DAMH ? a->surrogate.degree : 1) {
        case 1: lean_launch_deg<1>(a, blocks, st); break;
        case 2: lean_launch_deg<2>(a, blocks, st); break;
        case 3: lean_launch_deg<3>(a, blocks, st); break;
        case 4: lean_launch_deg<4>(a, blocks, st); break;
        case 5: lean_launch_deg<5>(a, blocks, st); break;
        default: lean_launch_deg<6>(a, blocks, st); break;
    }
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef void (*chain_fn)(const sdb_chain_args, const int, const int);

static chain_fn pick_kernel(int d, int m) {
    if (d == 2 && m == 1) return chain_kernel<2, 1, 0>;
    if (d == 4 && m == 3) return chain_kernel<4, 3, 0>;
    return chain_kernel<0, 0, 0>;
}
// the FP32-surrogate and streamed-surrogate instantiations (sdb_chain_f32.cu, sdb_chain_stream.cu)
chain_fn sdb_chain_pick_f32(int d, int m);
chain_fn sdb_chain_pick_stream(int d, int m);
// warp-owned chains (sdb_chain_wt.cu): mode = SUR_RESIDENT / SUR_F32 / SUR_STREAM
chain_fn sdb_chain_pick_wt(int d, int m, int mode);

enum { SUR_RESIDENT = 0, SUR_F32 = 1, SUR_STREAM = 2 };

static int validate(const sdb_chain_args *a) {
    SDB_CHECK_ARG(a != nullptr, "sdb_chain_run: args is NULL");
    const sdb_problem &p = a->problem;
    SDB_CHECK_LIMIT(p.d >= 1 && p.d <= SDB_MAX_D, "no_parameters %d outside [1,%d]", p.d, SDB_MAX_D);
    SDB_CHECK_LIMIT(p.m >= 1 && p.m <= SDB_MAX_M, "no_observations %d outside [1,%d]", p.m, SDB_MAX_M);
    SDB_CHECK_ARG(a->state.n_chains >= 1, "n_chains must be >= 1");
    SDB_CHECK_ARG(a->chain_lo >= 0 && a->chain_n >= 0 && a->chain_lo < a->state.n_chains, "chain range [%lld, +%lld) outside the %lld chains",
                  (long long)a->chain_lo, (long long)a->chain_n, (long long)a->state.n_chains);
    SDB_CHECK_ARG(a->state.u && a->state.G_u && a->state.post_u && a->state.pre_u && a->state.counters,
                  "chain state pointers must not be NULL");
    SDB_CHECK_ARG(p.prior_mean && p.prior_scale && p.obs && p.noise_scale, "problem pointers must not be NULL");
    SDB_CHECK_ARG(!p.prior_full || p.prior_piv, "prior_piv required with prior_full");
    SDB_CHECK_ARG(!p.noise_full || p.noise_piv, "noise_piv required with noise_full");
    SDB_CHECK_ARG(a->algorithm == SDB_ALG_MH || a->algorithm == SDB_ALG_DAMH, "unknown algorithm %d", a->algorithm);
    SDB_CHECK_ARG(a->phase >= SDB_PHASE_FUSED && a->phase <= SDB_PHASE_PREPARE, "unknown phase %d", a->phase);
    if (a->algorithm == SDB_ALG_DAMH)
        SDB_CHECK_ARG(a->surrogate.kind == SDB_MODEL_POLY || a->surrogate.kind == SDB_MODEL_RBF,
                      "DAMH needs a poly or rbf surrogate (a DAMH stage before any surrogate was fitted)");
    for (const sdb_model *mo : {&a->surrogate, &a->truth}) {
        if (mo->kind == SDB_MODEL_POLY) {
            SDB_CHECK_ARG(mo->coefs && mo->alpha && mo->herm, "poly model pointers must not be NULL");
            SDB_CHECK_LIMIT(mo->degree >= 0 && mo->degree <= SDB_MAX_DEGREE, "poly degree %d outside [0,%d]", mo->degree,
                            SDB_MAX_DEGREE);
            SDB_CHECK_ARG(mo->n_terms >= 1, "poly n_terms must be >= 1");
        } else if (mo->kind == SDB_MODEL_RBF) {
            SDB_CHECK_ARG(mo->coefs && mo->centers, "rbf model pointers must not be NULL");
            SDB_CHECK_ARG(mo->n_centers >= 1, "rbf n_centers must be >= 1");
            SDB_CHECK_ARG(mo->kernel_type >= 0 && mo->kernel_type <= 7, "rbf kernel_type %d outside [0,7]", mo->kernel_type);
        } else if (mo->kind == SDB_MODEL_TOY) {
            SDB_CHECK_ARG(p.d == 2 && p.m == 1, "the toy model is 2 parameters -> 1 observation");
        } else {
            SDB_CHECK_ARG(mo->kind == SDB_MODEL_NONE || mo->kind == SDB_MODEL_ZERO, "unknown model kind %d", mo->kind);
        }
    }
    SDB_CHECK_ARG(!a->surrogate_f32 || a->surrogate.kind == SDB_MODEL_RBF || a->surrogate.kind == SDB_MODEL_NONE,
                  "surrogate_f32 applies to an RBF surrogate (the polynomial surrogate is evaluated in FP64 only)");
    if (a->phase == SDB_PHASE_FUSED || (a->phase == SDB_PHASE_PREPARE && !a->have_G))
        SDB_CHECK_ARG(a->truth.kind != SDB_MODEL_NONE, "fused / prepare-without-G launches need a device truth model");
    if (a->phase == SDB_PHASE_PREPARE) return SDB_OK;
    SDB_CHECK_ARG(a->n_steps >= 1, "n_steps must be >= 1");
    SDB_CHECK_ARG(a->prop_scale != nullptr, "prop_scale must not be NULL");
    SDB_CHECK_ARG(a->prop_kind >= SDB_PROP_SHARED && a->prop_kind <= SDB_PROP_FACTOR, "unknown prop_kind %d", a->prop_kind);
    if (a->rng_kind == SDB_RNG_STREAMS)
        SDB_CHECK_ARG(a->uni && (a->phase == SDB_PHASE_DECIDE || a->z), "explicit streams requested but z / uni is NULL");
    else
        SDB_CHECK_ARG(a->rng_kind == SDB_RNG_PHILOX, "unknown rng_kind %d", a->rng_kind);
    SDB_CHECK_ARG(!a->snap_par == !a->snap_obs, "snap_par and snap_obs go together");
    SDB_CHECK_ARG(!a->snap_par || a->flags, "snapshots need the flags trace (it marks the valid rows)");
    if (a->phase != SDB_PHASE_FUSED) {
        SDB_CHECK_ARG(a->n_steps == 1, "split phases run one step per launch");
        SDB_CHECK_ARG(a->flags && a->prop && a->pre, "split phases need flags, prop and pre buffers");
        SDB_CHECK_ARG(a->phase != SDB_PHASE_DECIDE || a->G_prop, "DECIDE needs G_prop");
    }
    return SDB_OK;
}

static int plan(const sdb_chain_args *a, int *L_out, int *threads_out, int *blocks_out, int *smem_out, int *mode_out = nullptr,
                int *chunk_out = nullptr, int *wt_out = nullptr) {
    sdb_device_props dp;
    int rc = sdb_get_device_props(&dp);
    if (rc) return rc;
    const int d = a->problem.d, m = a->problem.m;
    const size_t fixed = 16 + problem_smem_bytes(a->problem, a->prop_kind) + sdb_model_smem_bytes(a->truth, d, m);
    const bool rbf_sur = a->surrogate.kind == SDB_MODEL_RBF && a->algorithm == SDB_ALG_DAMH;
    int mode = SUR_RESIDENT, chunk = 0;
    size_t smem = fixed + sdb_model_smem_bytes(a->surrogate, d, m);
    if (rbf_sur && a->surrogate_f32) {
        mode = SUR_F32;
        smem = fixed + sdb_model32_smem_bytes(a->surrogate, d, m);
    } else if (rbf_sur && (smem > (size_t)dp.smem_optin || getenv("SDB_STREAM_CHUNK"))) {
        // the tables do not fit: stream them through a two-stage ring; stages as large as the rest of the shared memory
        // allows (fewer barriers per evaluation), a multiple of 32 centers (bit-identical summation order)
        mode = SUR_STREAM;
        const size_t avail = (size_t)dp.smem_optin > fixed + 4096 ? (size_t)dp.smem_optin - fixed - 4096 : 0;
        chunk = (int)((avail / 2) / ((size_t)(d + m) * 8)) & ~31;
        if (chunk > 4096) chunk = 4096;
        if (getenv("SDB_STREAM_CHUNK")) {       // development / tests: stream even tables that would fit, in pieces of this size
            const int want = atoi(getenv("SDB_STREAM_CHUNK")) & ~31;
            if (want >= 32 && want < chunk) chunk = want;
        }
        const int64_t n32 = (a->surrogate.n_centers + 31) & ~(int64_t)31;
        if (chunk > n32) chunk = (int)n32;
        SDB_CHECK_LIMIT(chunk >= 32, "no room for a ring of 32 surrogate centers next to the true model's tables (%zu B)", fixed);
        SDB_CHECK_ARG((((uintptr_t)a->surrogate.centers) & 15) == 0 && (((uintptr_t)a->surrogate.coefs) & 15) == 0,
                      "streamed surrogate: centers / coefs must be 16-byte aligned");
        smem = fixed + sdb_ring_smem_bytes(chunk, d, m);
    }
    SDB_CHECK_LIMIT(smem <= (size_t)dp.smem_optin,
                    "model tables need %zu B of shared memory, the device offers %d B per block "
                    "(only an RBF SURROGATE can be streamed; the true model's tables must fit)",
                    smem, dp.smem_optin);
    if (mode_out) *mode_out = mode;
    if (chunk_out) *chunk_out = chunk;
    int64_t C = a->state.n_chains;             // chains this launch steps
    if (a->chain_n > 0) C = a->chain_lo + a->chain_n < C ? a->chain_n : C - a->chain_lo;
    const bool splittable = a->surrogate.kind == SDB_MODEL_RBF;  // only the RBF pair loop is split across lanes
    int L = a->lanes_per_chain;
    static const bool no_wt = getenv("SDB_NO_WT") != nullptr;
    if (wt_out) *wt_out = 0;
    // Warp-owned chains pay a fixed cost per pass (shuffling the group's points in, the 32-lane tree out) and bind a chain
    // to its warp for the whole launch; with few centers per lane (cfg5: 500 centers = 16 per lane) the lane-group shape
    // below is faster (measured 8.7 against 10.5 ms).  lanes_per_chain < 0 forces the warp-owned shape.
    static const int wt_min_centers = getenv("SDB_WT_MIN_CENTERS") ? atoi(getenv("SDB_WT_MIN_CENTERS")) : 1024;
    if (rbf_sur && !no_wt && (L < 0 || (L == 0 && a->surrogate.n_centers >= wt_min_centers))) {
        // warp-owned chains: 16-warp CTAs, one per SM and wave; the chains are dealt evenly to the warps of the grid (up to
        // 32 each), so every sub-partition carries the same number of pairs whatever C is.  65 536 chains: 148 CTAs,
        // 27 or 28 chains per warp, one wave -- against 293 CTAs of 14 warps (3 or 4 per sub-partition) in two waves.
        // coresident = SMs left entirely free for the kernels of other streams (the collector's refit running beside
        // the chains): a 16-warp CTA with 128 registers per thread owns its SM, so room is made by launching fewer CTAs.
        int sms = dp.sm_count - (a->coresident > 0 ? a->coresident : 0);
        if (sms < 1) sms = 1;
        // More than 32 chains per warp: the grid stays at one CTA per SM and every warp works through several runs of
        // chains ("waves") inside the kernel, so the SMs left free are never taken by a second wave of CTAs.
        const int64_t wave_chains = (int64_t)sms * 16 * 32;
        const int64_t waves = (C + wave_chains - 1) / wave_chains;
        int64_t warps_per_sm = (C + sms - 1) / sms;                       // few chains: one warp each, spread over the SMs
        if (warps_per_sm > 16) warps_per_sm = 16;
        if (warps_per_sm < 2) warps_per_sm = 2;
        int64_t blocks = (C + warps_per_sm - 1) / warps_per_sm;
        if (blocks > sms) blocks = sms;
        *L_out = 0; *threads_out = (int)warps_per_sm * 32; *blocks_out = (int)blocks; *smem_out = (int)smem;
        if (wt_out) *wt_out = (int)waves;
        return SDB_OK;
    }
    if (L <= 0) {
        L = 1;
        // aim at >= 512 resident threads per SM before giving a chain more lanes
        const int64_t want = (int64_t)dp.sm_count * 512;
        if (splittable)
            while (L < 32 && C * L < want) L <<= 1;
    }
    SDB_CHECK_ARG(L == 1 || L == 2 || L == 4 || L == 8 || L == 16 || L == 32, "lanes_per_chain must be a power of two <= 32");
    const int64_t lanes = C * L;
    // one CTA per SM slot: the table copy is per CTA, so prefer few fat CTAs; threads a multiple of 32
    const int blocks_per_sm = (int)((size_t)dp.smem_optin / (smem + 1024)) >= 2 && smem < 100 * 1024 ? 2 : 1;
    const int64_t slots = (int64_t)dp.sm_count * blocks_per_sm;
    int64_t per = (lanes + slots - 1) / slots;
    if (per > 512) {
        // several waves of CTAs: size them so that the last wave is as full as the others (131 072 chains on 148 one-CTA
        // SMs: 293 CTAs of 448 threads in two full waves instead of 256 CTAs of 512 in a full and a 73 % wave)
        const int64_t waves = (per + 511) / 512;
        per = (lanes + slots * waves - 1) / (slots * waves);
    }
    int threads = (int)(((per + 31) / 32) * 32);
    if (threads < 64) threads = 64;
    if (threads > 512) threads = 512;
    const int64_t blocks = (lanes + threads - 1) / threads;
    size_t smem_req = smem;
    if (a->coresident) {
        // more than half of the shared memory per CTA -> at most one CTA of this kernel per SM, i.e. the CTAs run in two
        // waves and the second wave leaves a few SMs (and, on every SM, ~8 K registers and ~110 KB of shared memory) to
        // the kernels of other streams.  Measured: making more room (224-thread CTAs, so that the refit kernels really
        // co-reside) slows this kernel by 8 % and the FP64-latency-bound refit still runs 4-5 x slower beside it.
        const size_t half = (size_t)dp.smem_optin / 2 + 2048;
        if (smem_req < half) smem_req = half;
    }
    *L_out = L; *threads_out = threads; *blocks_out = (int)blocks; *smem_out = (int)smem_req;
    return SDB_OK;
}

extern "C" int sdb_chain_plan(const sdb_chain_args *args, int *lanes_per_chain, int *threads_per_block, int *blocks,
                              int *smem_bytes) {
    int rc = validate(args);
    if (rc) return rc;
    return plan(args, lanes_per_chain, threads_per_block, blocks, smem_bytes);
}

extern "C" int sdb_chain_run(const sdb_chain_args *args, void *stream) {
    int rc = validate(args);
    if (rc) return rc;
    if (lean_applicable(args)) return lean_launch(args, (cudaStream_t)stream);
    int L, threads, blocks, smem, mode, chunk, wt;
    rc = plan(args, &L, &threads, &blocks, &smem, &mode, &chunk, &wt);
    if (rc) return rc;
    const int d = args->problem.d, m = args->problem.m;
    chain_fn fn = wt ? sdb_chain_pick_wt(d, m, mode)
                     : mode == SUR_F32 ? sdb_chain_pick_f32(d, m) : mode == SUR_STREAM ? sdb_chain_pick_stream(d, m) : pick_kernel(d, m);
    if (wt) L = wt;     // warp-owned chains: the kernel's second argument is the number of waves
    SDB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    sdb_count_launch();
    fn<<<blocks, threads, smem, (cudaStream_t)stream>>>(*args, L, chunk);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}
