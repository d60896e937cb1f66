// Fused lockstep MH / delayed-acceptance MH chain kernel (sm_100a).
//
// One thread -- or a group of L = 2..32 lanes -- owns one chain and keeps its state
// (u, G(u), log posteriors, counters) in registers for the K steps of a launch.  The
// surrogate / truth model tables and the problem constants are staged once per CTA into
// shared memory with TMA bulk copies.  Per step: Philox (or explicit-stream) proposal,
// surrogate evaluation of the proposal, stage-1 test, conditional true-model evaluation
// (warp-cooperative for an RBF truth so divergence costs nothing), stage-2 test, state
// update, coalesced [K][C] trace / snapshot rows.
//
// Follows classes_SAMPLER.py:54-101,122-128,147-220 (see include/surrdamh_b200.h).
#pragma once
#include <stdarg.h>
#include <stdlib.h>

#include "sdb_models.cuh"

#define SDB_FLAG_PENDING 3  // internal: stage 1 passed, waiting for the host's G (split phases)

struct chain_launch {
    int L;             // lanes per chain
    int smem_bytes;    // dynamic shared memory
};

__host__ __device__ inline size_t problem_smem_bytes(const sdb_problem &p, int prop_kind) {
    const int d = p.d, m = p.m;
    size_t b = sdb_align16((size_t)d * 8);                                  // prior_mean
    b += sdb_align16((size_t)(p.prior_full ? d * d : d) * 8);               // prior_scale
    b += sdb_align16((size_t)d * 4);                                        // prior_piv
    b += sdb_align16((size_t)m * 8);                                        // obs
    b += sdb_align16((size_t)(p.noise_full ? m * m : m) * 8);               // noise_scale
    b += sdb_align16((size_t)m * 4);                                        // noise_piv
    b += sdb_align16((size_t)(prop_kind == SDB_PROP_FACTOR ? d * d : d) * 8);  // shared proposal scale
    return b;
}

__device__ __forceinline__ void copy_small(void *dst, const void *src, size_t bytes) {
    if (src == nullptr) return;
    for (size_t i = threadIdx.x; i < (bytes >> 2); i += blockDim.x) ((uint32_t *)dst)[i] = ((const uint32_t *)src)[i];
}

// WARP: the call is made by every lane of the warp, converged (the surrogate of a lockstep step) -- the RBF evaluation may
// then team up neighbouring chains (sdb_rbf_point_team)
template <int D, int M, bool WARP = false>
__device__ __forceinline__ void eval_model_lane(const sdb_smodel &s, int d, int m, const double *x, double *out, int sub,
                                                int L) {
    // per-lane (or L-lane split) evaluation; used for the surrogate and for cheap truths
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    switch (s.kind) {
        case SDB_MODEL_POLY: sdb_poly_point<D, M>(s, d, m, x, out); break;
        case SDB_MODEL_RBF:
            if (WARP && L <= 8 && D <= 4 && M <= 3 && D && M) sdb_rbf_point_team<D, M, 4>(s, d, m, x, out, L);
            else if (WARP && L <= 16) sdb_rbf_point_team<D, M, 2>(s, d, m, x, out, L);
            else sdb_rbf_point_split<D, M>(s, d, m, x, out, sub, L);
            break;
        case SDB_MODEL_TOY: out[0] = sdb_toy_G(s, x); break;
        default:
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) out[k] = 0.0;
    }
}

// SM: how the surrogate is held -- 0: tables resident in shared memory, FP64 (poly or RBF); 1: RBF tables resident in
// FP32, evaluated on the FP32 pipe (surrogate_f32); 2: RBF tables streamed from L2 through a TMA ring of `chunk`-center
// pieces per evaluation (they do not fit in shared memory).  Separate instantiations, so mode 0 keeps its code.
// WT: warp-owned chains (see sdb_wt_points): launched with L = 1; a warp owns a contiguous run of up to 32 chains, one per
// lane, and the warps of the grid share the launch's chains evenly (counts differ by at most one).  RBF surrogate only.
template <int D, int M, int SM = 0, bool WT = false>
__global__ void __launch_bounds__(512, 1) chain_kernel(const sdb_chain_args a, const int L_or_waves, const int chunk) {
    constexpr int DD = D ? D : SDB_MAX_D;
    constexpr int UD = D ? D : 1;  // unroll only the specialised sizes
    constexpr int MM = M ? M : SDB_MAX_M;
    constexpr int UM = M ? M : 1;
    constexpr bool MOM_REGS = (D >= 1 && D <= 4);
    constexpr int NMOM = MOM_REGS ? (D + D * (D + 1) / 2) : 1;
    const int d = D ? D : a.problem.d, m = M ? M : a.problem.m;
    const int64_t C = a.state.n_chains;

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;
    unsigned char *p = smem + 16;
    if (threadIdx.x == 0) {
        sdb_mbar_init(bar, 1);
        sdb_mbar_fence_init();
    }
    __syncthreads();

    // ---- problem constants (tiny: ordinary loads) --------------------------------------
    sdb_sproblem pr;
    pr.d = d; pr.m = m; pr.prior_full = a.problem.prior_full; pr.noise_full = a.problem.noise_full;
    {
        size_t b;
        b = (size_t)d * 8; pr.prior_mean = (const double *)p; copy_small(p, a.problem.prior_mean, b); p += sdb_align16(b);
        b = (size_t)(pr.prior_full ? d * d : d) * 8; pr.prior_scale = (const double *)p; copy_small(p, a.problem.prior_scale, b); p += sdb_align16(b);
        b = (size_t)d * 4; pr.prior_piv = (const int *)p; copy_small(p, a.problem.prior_piv, b); p += sdb_align16(b);
        b = (size_t)m * 8; pr.obs = (const double *)p; copy_small(p, a.problem.obs, b); p += sdb_align16(b);
        b = (size_t)(pr.noise_full ? m * m : m) * 8; pr.noise_scale = (const double *)p; copy_small(p, a.problem.noise_scale, b); p += sdb_align16(b);
        b = (size_t)m * 4; pr.noise_piv = (const int *)p; copy_small(p, a.problem.noise_piv, b); p += sdb_align16(b);
    }
    const double *prop_shared = (const double *)p;
    {
        const size_t b = (size_t)(a.prop_kind == SDB_PROP_FACTOR ? d * d : d) * 8;
        if (a.prop_kind != SDB_PROP_PER_CHAIN) copy_small(p, a.prop_scale, b);
        p += sdb_align16(b);
    }
    // ---- model tables: TMA bulk copies, one mbarrier phase ---------------------------------
    sdb_smodel sur;
    if (SM == 0) {
        sur = sdb_stage_model(a.surrogate, d, m, p, bar);
    } else {
        sur.kind = a.surrogate.kind; sur.kernel_type = a.surrogate.kernel_type; sur.degree = 0; sur.n_terms = 0;
        sur.n_centers = (int)a.surrogate.n_centers;
        sur.coefs = nullptr; sur.centers = nullptr; sur.alpha = nullptr; sur.herm = nullptr;
        sur.toy_f = sur.toy_L = sur.toy_M = 0.0;
    }
    const sdb_smodel tru = sdb_stage_model(a.truth, d, m, p, bar);
    if (threadIdx.x == 0) sdb_mbar_arrive(bar);  // the one arrival; the bulk pieces carry the tx counts
    sdb_mbar_wait(bar, 0);
    __syncthreads();
    sdb_smodel32 sur32;
    sdb_ring ring;
    if constexpr (SM == 1) sur32 = sdb_stage_model32(a.surrogate, d, m, p);
    // ---- chain ownership ---------------------------------------------------------------------
    // WT: the kernel's second argument is the number of WAVES -- every warp works through `waves` runs of chains one after
    // the other (a persistent grid of at most one CTA per SM, so SMs the launch leaves free stay free until it ends)
    const int waves = WT ? L_or_waves : 1;
    const int L = WT ? 1 : L_or_waves;
    const int64_t c_hi = a.chain_n > 0 ? (a.chain_lo + a.chain_n < C ? a.chain_lo + a.chain_n : C) : C;   // chains [chain_lo, c_hi)
    const int64_t wt_nW = (int64_t)gridDim.x * (blockDim.x >> 5) * waves;      // virtual warps
    const int64_t wt_q = (c_hi - a.chain_lo) / wt_nW, wt_r = (c_hi - a.chain_lo) % wt_nW;
    const int wt_Pmax = (int)(wt_q + (wt_r > 0 ? 1 : 0));
    if constexpr (SM == 2) {
        const bool is_damh = a.algorithm == SDB_ALG_DAMH;
        const long long evals = !is_damh ? 0
                                : ((a.phase == SDB_PHASE_PREPARE || a.refresh_pre) ? 1 : 0) +
                                      ((a.phase == SDB_PHASE_FUSED || a.phase == SDB_PHASE_PROPOSE) ? a.n_steps : 0);
        ring = sdb_ring_init(a.surrogate, d, m, chunk, WT ? evals * sdb_wt_passes(wt_Pmax, D, M) * waves : evals, p);
    }

    for (int wave = 0; wave < waves; ++wave) {
    int64_t grp;
    int sub, wt_P = 0;
    if constexpr (WT) {
        const int64_t w = (int64_t)wave * (wt_nW / waves) + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        wt_P = (int)(wt_q + (w < wt_r ? 1 : 0));
        grp = (threadIdx.x & 31) < wt_P ? a.chain_lo + w * wt_q + (w < wt_r ? w : wt_r) + (threadIdx.x & 31) : c_hi;
        sub = 0;
    } else {
        const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        grp = a.chain_lo + gtid / L;
        sub = (int)(gtid % L);
    }
    const bool active = grp < c_hi;
    const int64_t c = active ? grp : (c_hi - 1);  // idle tail lanes shadow the last chain, stores masked
    const bool writer = active && sub == 0;
    const uint64_t gchain = (uint64_t)(a.chain0 + c);
    const bool damh = a.algorithm == SDB_ALG_DAMH;
    // speculative evaluation of a cheap analytic truth: only in fused launches too small to fill the GPU
    const bool spec_truth = a.phase == SDB_PHASE_FUSED && (tru.kind == SDB_MODEL_TOY || tru.kind == SDB_MODEL_ZERO) &&
                            (int64_t)gridDim.x * blockDim.x <= 148 * 256;
    // the surrogate at a point, called by every lane of the CTA, converged (the proposal of a lockstep step)
    auto eval_sur = [&](const double *X, double *S) {
        if constexpr (WT) {
#pragma unroll UM
            for (int k = 0; k < MM; ++k)
                if (k < m) S[k] = 0.0;
            if constexpr (SM == 1) sdb_rbf32_points_wt<D, M>(sur32, d, m, X, S, wt_P);
            else if constexpr (SM == 2) sdb_rbf_points_wt_stream<D, M>(ring, d, m, X, S, wt_Pmax);
            else sdb_rbf_points_wt<D, M>(sur, d, m, X, S, wt_P);
        } else if constexpr (SM == 1) {
            if (L <= 8 && D <= 4 && M <= 3 && D && M) sdb_rbf32_point_team<D, M, 4>(sur32, d, m, X, S, L);
            else if (L <= 16) sdb_rbf32_point_team<D, M, 2>(sur32, d, m, X, S, L);
            else sdb_rbf32_point_team<D, M, 1>(sur32, d, m, X, S, L);
        } else if constexpr (SM == 2) {
            if (L <= 8 && D <= 4 && M <= 3 && D && M) sdb_rbf_point_team_stream<D, M, 4>(ring, d, m, X, S, L);
            else if (L <= 16) sdb_rbf_point_team_stream<D, M, 2>(ring, d, m, X, S, L);
            else sdb_rbf_point_team_stream<D, M, 1>(ring, d, m, X, S, L);
        } else {
            eval_model_lane<D, M, true>(sur, d, m, X, S, sub, L);
        }
    };

    double u[DD], Gu[MM];
#pragma unroll UD
    for (int j = 0; j < DD; ++j)
        if (j < d) u[j] = a.state.u[(size_t)j * C + c];
#pragma unroll UM
    for (int k = 0; k < MM; ++k)
        if (k < m) Gu[k] = a.state.G_u[(size_t)k * C + c];
    double post_u = a.state.post_u[c], pre_u = a.state.pre_u[c];
    int n_acc = a.state.counters[0 * C + c], n_rej = a.state.counters[1 * C + c];
    int n_pre = a.state.counters[2 * C + c], rej_cur = a.state.counters[3 * C + c];
    double mom[NMOM];
    const bool do_mom = a.state.moments != nullptr;
    if (MOM_REGS && do_mom) {
#pragma unroll
        for (int i = 0; i < NMOM; ++i) mom[i] = a.state.moments[(size_t)i * C + c];
    }
    double sig[DD];  // diagonal proposal std
    if (a.prop_kind != SDB_PROP_FACTOR && a.phase != SDB_PHASE_PREPARE) {
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) sig[j] = (a.prop_kind == SDB_PROP_PER_CHAIN) ? a.prop_scale[(size_t)j * C + c] : prop_shared[j];
    }

    // ---- stage prologue (Algorithm_PARENT.prepare, :54-75; DAMH prologue :178-180) --------------
    if (a.phase == SDB_PHASE_PREPARE) {
        if (!a.have_G) {
            if (tru.kind == SDB_MODEL_RBF) sdb_rbf_point_warp<D, M>(tru, d, m, u, Gu, true, L);
            else eval_model_lane<D, M>(tru, d, m, u, Gu, sub, L);
        }
        post_u = sdb_log_posterior<D, M>(pr, u, Gu);
        rej_cur = 0;
        pre_u = 0.0;
    }
    if (damh && (a.phase == SDB_PHASE_PREPARE || a.refresh_pre)) {
        double S[MM];
        eval_sur(u, S);
        pre_u = sdb_log_posterior<D, M>(pr, u, S);
    }

    const int K = (a.phase == SDB_PHASE_PREPARE) ? 0 : a.n_steps;
    for (int k = 0; k < K; ++k) {
        const uint64_t gstep = (uint64_t)(a.step0 + k);
        const size_t row = (size_t)k * C + c;
        double up[DD], Gp[MM];
        double pre_p = __longlong_as_double(0x7ff8000000000000LL), post_p = pre_p;  // NaN until computed
        double uni0, uni1;
        if (a.rng_kind == SDB_RNG_STREAMS) {
            uni0 = a.uni[((size_t)k * 2 + 0) * C + c];
            uni1 = a.uni[((size_t)k * 2 + 1) * C + c];
        } else {
            const philox4 b = sdb_chain_block(a.seed, a.stream_id, gchain, gstep, SDB_SLOT_UNIFORMS);
            uni0 = sdb_u01(b.x, b.y);
            uni1 = sdb_u01(b.z, b.w);
        }
        if (a.pre_cur && writer && a.phase != SDB_PHASE_DECIDE) a.pre_cur[row] = pre_u;
        bool pass1;
        double lp_p = 0.0, post_p_spec = 0.0;
        bool have_lp = false;
        if (a.phase != SDB_PHASE_DECIDE) {
            // ---- proposal: u' = u + sigma*z, separately rounded (numpy: loc + scale*gauss) -------
            double z[DD];
            if (a.rng_kind == SDB_RNG_STREAMS) {
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) z[j] = a.z[((size_t)k * d + j) * C + c];
            } else {
#pragma unroll UD
                for (int j = 0; j < DD; j += 2)
                    if (j < d) {
                        double z0, z1;
                        sdb_normal_pair(sdb_chain_block(a.seed, a.stream_id, gchain, gstep, (uint32_t)(j >> 1)), z0, z1);
                        z[j] = z0;
                        if (j + 1 < DD) z[j + 1] = z1;
                    }
            }
            if (a.prop_kind == SDB_PROP_FACTOR) {
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) {
                        double t = 0.0;
                        for (int i = 0; i <= j; ++i) t = fma(prop_shared[j * d + i], z[i], t);
                        up[j] = u[j] + t;
                    }
            } else {
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) up[j] = __dadd_rn(u[j], __dmul_rn(sig[j], z[j]));
            }
            // ---- stage 1: surrogate posterior of the proposal (:185-193) --------------------------
            pass1 = true;
            lp_p = sdb_log_prior<D, M>(pr, up);       // shared by the surrogate and the true posterior of this proposal
            have_lp = true;
            if (spec_truth) {
                // latency-bound launches (few chains): the cheap analytic truth is evaluated for every proposal, off
                // the critical path of the stage-1 test; same arithmetic, same result, the value is simply unused
                // after a pre-rejection
                eval_model_lane<D, M>(tru, d, m, up, Gp, 0, 1);
                post_p_spec = __dadd_rn(sdb_log_likelihood<D, M>(pr, Gp), lp_p);
            }
            if (damh) {
                double S[MM];
                eval_sur(up, S);
                pre_p = __dadd_rn(sdb_log_likelihood<D, M>(pr, S), lp_p);
                pass1 = log(uni0) < __dadd_rn(pre_p, -pre_u);
            }
            if (a.phase == SDB_PHASE_PROPOSE) {
                if (!pass1) { ++n_pre; ++rej_cur; }
                if (writer) {
                    a.flags[row] = pass1 ? SDB_FLAG_PENDING : SDB_FLAG_PREREJECTED;
#pragma unroll UD
                    for (int j = 0; j < DD; ++j)
                        if (j < d) a.prop[((size_t)k * d + j) * C + c] = up[j];
                    a.pre[row] = pre_p;
                    if (a.post) a.post[row] = post_p;
                }
                continue;
            }
        } else {
#pragma unroll UD
            for (int j = 0; j < DD; ++j)
                if (j < d) up[j] = a.prop[((size_t)k * d + j) * C + c];
            pre_p = a.pre[row];
            pass1 = a.flags[row] == SDB_FLAG_PENDING;
        }
        // ---- true model for the survivors (:77-81) ---------------------------------------------
        if (a.phase == SDB_PHASE_DECIDE) {
            if (pass1) {
#pragma unroll UM
                for (int kk = 0; kk < MM; ++kk)
                    if (kk < m) Gp[kk] = a.G_prop[(size_t)kk * C + c];
            }
        } else if (tru.kind == SDB_MODEL_RBF) {
            sdb_rbf_point_warp<D, M>(tru, d, m, up, Gp, pass1 && active, L);
        } else if (pass1 && !spec_truth) {
            eval_model_lane<D, M>(tru, d, m, up, Gp, 0, 1);
        }
        int flag = SDB_FLAG_PREREJECTED;
        if (pass1) {
            if (spec_truth) post_p = post_p_spec;
            else post_p = __dadd_rn(sdb_log_likelihood<D, M>(pr, Gp), have_lp ? lp_p : sdb_log_prior<D, M>(pr, up));
            const double log_ratio = __dadd_rn(post_p, -post_u);
            bool acc;
            if (damh) acc = log(uni1) < __dadd_rn(log_ratio, -__dadd_rn(pre_p, -pre_u));  // stage 2 (:195)
            else acc = log(uni0) < log_ratio;                                              // MH (:152)
            if (acc) {
                flag = SDB_FLAG_ACCEPTED;
                if (a.snap_par && writer) {  // displaced current sample goes to the collector (:85)
#pragma unroll UD
                    for (int j = 0; j < DD; ++j)
                        if (j < d) a.snap_par[((size_t)k * d + j) * C + c] = u[j];
#pragma unroll UM
                    for (int kk = 0; kk < MM; ++kk)
                        if (kk < m) a.snap_obs[((size_t)k * m + kk) * C + c] = Gu[kk];
                }
                ++n_acc;
                rej_cur = 0;
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) u[j] = up[j];
#pragma unroll UM
                for (int kk = 0; kk < MM; ++kk)
                    if (kk < m) Gu[kk] = Gp[kk];
                post_u = post_p;
                if (damh) pre_u = pre_p;  // same surrogate, same point: what :189 recomputes next step
            } else {
                flag = SDB_FLAG_REJECTED;
                ++n_rej;
                ++rej_cur;
                if (a.snap_par && writer) {  // rejected proposal goes to the collector with weight 0 (:98)
#pragma unroll UD
                    for (int j = 0; j < DD; ++j)
                        if (j < d) a.snap_par[((size_t)k * d + j) * C + c] = up[j];
#pragma unroll UM
                    for (int kk = 0; kk < MM; ++kk)
                        if (kk < m) a.snap_obs[((size_t)k * m + kk) * C + c] = Gp[kk];
                }
            }
        } else if (a.phase != SDB_PHASE_DECIDE) {  // split path: PROPOSE already counted the pre-rejection
            ++n_pre;
            ++rej_cur;
        }
        // ---- trace rows (coalesced over chains) ----------------------------------------------------
        if (writer) {
            if (a.flags) a.flags[row] = (uint8_t)flag;
            if (a.prop && a.phase != SDB_PHASE_DECIDE) {
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) a.prop[((size_t)k * d + j) * C + c] = up[j];
            }
            if (a.post) a.post[row] = post_p;
            if (a.pre && a.phase != SDB_PHASE_DECIDE) a.pre[row] = pre_p;
        }
        // ---- running moments of the chain (state after the decision) -------------------------------
        if (do_mom) {
            if (MOM_REGS) {
                int q = 0;
#pragma unroll UD
                for (int j = 0; j < DD; ++j)
                    if (j < d) mom[q++] += u[j];
#pragma unroll UD
                for (int i = 0; i < DD; ++i)
#pragma unroll UD
                    for (int j = i; j < DD; ++j)
                        if (i < d && j < d) { mom[q] = fma(u[i], u[j], mom[q]); ++q; }
            } else if (writer) {
                int q = 0;
                for (int j = 0; j < d; ++j) a.state.moments[(size_t)(q++) * C + c] += u[j];
                for (int i = 0; i < d; ++i)
                    for (int j = i; j < d; ++j) a.state.moments[(size_t)(q++) * C + c] += u[i] * u[j];
            }
        }
    }

    // ---- state back to HBM ------------------------------------------------------------------------
    if (writer) {
#pragma unroll UD
        for (int j = 0; j < DD; ++j)
            if (j < d) a.state.u[(size_t)j * C + c] = u[j];
#pragma unroll UM
        for (int k = 0; k < MM; ++k)
            if (k < m) a.state.G_u[(size_t)k * C + c] = Gu[k];
        a.state.post_u[c] = post_u;
        a.state.pre_u[c] = pre_u;
        a.state.counters[0 * C + c] = n_acc;
        a.state.counters[1 * C + c] = n_rej;
        a.state.counters[2 * C + c] = n_pre;
        a.state.counters[3 * C + c] = rej_cur;
        if (MOM_REGS && do_mom) {
#pragma unroll
            for (int i = 0; i < NMOM; ++i) a.state.moments[(size_t)i * C + c] = mom[i];
        }
    }
    }   // waves
}

