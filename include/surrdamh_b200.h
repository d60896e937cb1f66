/*
 * surrdamh_b200.h -- C ABI of the B200-native surrDAMH hot path.
 *
 * The reference (dom0015/surrDAMH) is pure Python and has no FFI; its "plugin API" for
 * this path is three duck-typed Python surfaces.  This header is what
 * a binding for those surfaces talks to: plain pointers and sizes, no torch types.
 * All pointers named d_* / marked DEVICE are device (HBM) pointers owned by the caller;
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every call
 * is asynchronous with respect to the host unless stated otherwise and returns 0 on
 * success or a non-zero code; sdb_last_error() gives the message of the last failure on
 * the calling thread.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with SDB_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   sdb_poly_eval        surrDAMH/modules/surrogate_poly.py:18-30   Surrogate_apply.apply
 *   sdb_rbf_eval         surrDAMH/modules/surrogate_rbf.py:20-40    Surrogate_apply.apply
 *   sdb_chain_run        surrDAMH/modules/classes_SAMPLER.py:54-101,122-128,147-220
 *                        Algorithm_MH.run / Algorithm_DAMH.run (+ prepare, if_accepted,
 *                        if_rejected, _acceptance_log_symmetric), Proposal_GaussRandomWalk
 *                        (:241-251), Problem_Gauss.get_log_posterior (:295-316), and the
 *                        in-process surrogate evaluation of
 *                        surrDAMH/modules/classes_communication.py:112-119
 *   sdb_snapshot_compact surrDAMH/modules/classes_communication.py:121-135 (snapshot batching)
 *   sdb_poly_basis/_fit  surrDAMH/modules/surrogate_poly.py:68-97    Surrogate_update.update
 *   sdb_rbf_system/_fit  surrDAMH/modules/surrogate_rbf.py:76-137    Surrogate_update.update
 *   sdb_lu_solve         np.linalg.solve at surrogate_rbf.py:130     (solver_type 'direct')
 *   sdb_minres           scipy minres at surrogate_rbf.py:131-135    (solver_type 'minres', the default)
 *   sdb_rbf_thin         surrDAMH/modules/surrogate_rbf.py:97-120    greedy thinning to no_keep
 */
#ifndef SURRDAMH_B200_H
#define SURRDAMH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 2

/* exported symbols (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define SDB_API __attribute__((visibility("default")))
#else
#define SDB_API
#endif

/* error codes */
#define SDB_OK 0
#define SDB_ERR_ARG 1      /* invalid argument (message says which) */
#define SDB_ERR_CUDA 2     /* CUDA runtime error (message carries cudaGetErrorString) */
#define SDB_ERR_LIMIT 3    /* shape exceeds a compiled limit (d, m, degree, shared memory) */
#define SDB_ERR_SINGULAR 4 /* factorisation met an exactly zero pivot */

/* limits of the compiled kernels */
#define SDB_MAX_D 32      /* no_parameters */
#define SDB_MAX_M 64      /* no_observations */
#define SDB_MAX_DEGREE 10 /* polynomial total degree */

/* per-step outcome codes written to the trace (classes_SAMPLER.py raw_data rows) */
#define SDB_FLAG_PREREJECTED 0
#define SDB_FLAG_ACCEPTED 1
#define SDB_FLAG_REJECTED 2

/* ---- forward models usable on the device (surrogate and "true" G) ------------------- */
#define SDB_MODEL_NONE 0 /* no model (MH has no surrogate; G evaluated by the host) */
#define SDB_MODEL_POLY 1 /* SOL = [MAT, degree] of surrogate_poly.py */
#define SDB_MODEL_RBF 2  /* SOL = [COEFS, centers] of surrogate_rbf.py */
#define SDB_MODEL_TOY 3  /* examples/solvers/solver_examples.py:12-29 Solver_linela2exp_local */
#define SDB_MODEL_ZERO 4 /* examples/solvers/solver_examples.py:66-75 Solver_uninformative */

typedef struct sdb_model {
    int32_t kind;        /* SDB_MODEL_* */
    int32_t kernel_type; /* RBF: 0 r, 1 r^3, 2 r^5, 3 r^7, 7 r^9, 4 exp(-r^2), 5 (r^2+1)^-1/2, 6 == 0 */
    int32_t degree;      /* POLY: total degree of the basis */
    int32_t n_terms;     /* POLY: number of basis terms P = C(d+degree, d) */
    int64_t n_centers;   /* RBF: number of centers n */
    const double *coefs;   /* DEVICE. RBF: COEFS [(n+d+1) x m] row-major; POLY: MAT [P x m] row-major */
    const double *centers; /* DEVICE. RBF: centers [n x d] row-major */
    const int32_t *alpha;  /* DEVICE. POLY: multi-indices [P x d] row-major, reference order */
    const double *herm;    /* DEVICE. POLY: normalised Hermite coefficient table [(degree+1)^2] */
    double toy_f, toy_L, toy_M; /* TOY: constructor arguments f, L, M */
} sdb_model;

/* ---- Gaussian problem (Problem_Gauss) ------------------------------------------------ */
typedef struct sdb_problem {
    int32_t d, m;
    const double *prior_mean; /* DEVICE [d] */
    int32_t prior_full;       /* 0: prior_scale = variances std^2 [d]; 1: LU factors of the covariance [d x d] */
    const double *prior_scale;
    const int32_t *prior_piv; /* DEVICE [d] row interchanges of the LU (prior_full == 1) */
    const double *obs;        /* DEVICE [m] observations y */
    int32_t noise_full;       /* as prior_full, for the noise */
    const double *noise_scale;
    const int32_t *noise_piv;
} sdb_problem;

/* ---- chain state, struct of arrays over the C chains of this GPU ---------------------- */
typedef struct sdb_chain_state {
    int64_t n_chains; /* C */
    double *u;        /* DEVICE [d][C] current sample */
    double *G_u;      /* DEVICE [m][C] G(current) */
    double *post_u;   /* DEVICE [C] log posterior of current (true model) */
    double *pre_u;    /* DEVICE [C] log posterior of current under the surrogate (pre_posterior_current_sample) */
    int32_t *counters; /* DEVICE [4][C]: accepted, rejected, prerejected, rejected-since-last-accept */
    double *moments;  /* DEVICE [(d + d(d+1)/2)][C] running sums of u and of u_i*u_j (i<=j) over steps; may be NULL */
} sdb_chain_state;

#define SDB_ALG_MH 0
#define SDB_ALG_DAMH 1

#define SDB_RNG_PHILOX 0  /* counter-based Philox-4x32-10 keyed by (seed, stream), counter (chain, step, slot) */
#define SDB_RNG_STREAMS 1 /* explicit normals / uniforms read from memory (parity runs) */

#define SDB_PROP_SHARED 0    /* proposal std [d], same for every chain */
#define SDB_PROP_PER_CHAIN 1 /* proposal std [d][C] (proposal_differs) */
#define SDB_PROP_FACTOR 2    /* lower-triangular factor L [d x d] row-major, u' = u + L z */

/* which part of a step the launch executes */
#define SDB_PHASE_FUSED 0   /* whole steps, G evaluated on the device */
#define SDB_PHASE_PROPOSE 1 /* one step, first half: propose (+ surrogate + stage-1 test); host evaluates G */
#define SDB_PHASE_DECIDE 2  /* one step, second half: consumes G of the proposals, decides, updates */
#define SDB_PHASE_PREPARE 3 /* stage prologue (Algorithm_PARENT.prepare + DAMH prologue), no step */

typedef struct sdb_chain_args {
    int32_t algorithm; /* SDB_ALG_* */
    int32_t phase;     /* SDB_PHASE_* */
    int32_t n_steps;   /* K lockstep steps in this launch (1 for the split phases) */
    int32_t lanes_per_chain; /* 1,2,4,8,16 or 32 threads cooperating on one chain; 0 = library picks: with an RBF
                              * surrogate of >= 1024 centers warp-owned chains (a warp owns up to 32 chains, one per lane,
                              * and all 32 lanes split the centers of every evaluation; sdb_chain_plan then reports
                              * lanes_per_chain 0), else lane groups.  < 0 forces the warp-owned shape. */
    int32_t refresh_pre;     /* recompute pre_u = log posterior(surrogate(u)) first (surrogate was swapped) */
    int32_t have_G;          /* PREPARE: G_u already holds G(u) (host-evaluated); else the device model fills it */
    sdb_problem problem;
    sdb_model surrogate; /* kind NONE for MH */
    sdb_model truth;     /* device-evaluable G; kind NONE for the split (host G) phases */
    sdb_chain_state state;
    /* proposal */
    int32_t prop_kind;
    const double *prop_scale; /* DEVICE, see SDB_PROP_* */
    /* randomness */
    int32_t rng_kind;
    uint64_t seed;       /* PHILOX key */
    uint32_t stream_id;  /* PHILOX: sub-stream (counter word 3) */
    int64_t chain0;      /* global index of this GPU's first chain (keys Philox; results independent of sharding) */
    int64_t step0;       /* global index of the first step of this launch */
    const double *z;     /* STREAMS: DEVICE [K][d][C] standard normals */
    const double *uni;   /* STREAMS: DEVICE [K][2][C] uniforms in [0,1): slot 0 stage-1/MH test, slot 1 stage-2 test */
    /* per-step trace, each may be NULL (no trace) */
    uint8_t *flags;   /* DEVICE [K][C] SDB_FLAG_* */
    double *prop;     /* DEVICE [K][d][C] proposed samples */
    double *post;     /* DEVICE [K][C] log posterior of the proposal (NaN where G was not evaluated) */
    double *pre;      /* DEVICE [K][C] surrogate log posterior of the proposal (NaN for MH) */
    /* snapshots for the surrogate update: rows valid where flags != PREREJECTED; need flags != NULL */
    double *snap_par; /* DEVICE [K][d][C]: displaced current (accept) or the proposal (reject) */
    double *snap_obs; /* DEVICE [K][m][C]: its G */
    /* split phases: G of the proposals supplied by the host for DECIDE */
    const double *G_prop; /* DEVICE [m][C] */
    /* optional trace: surrogate log posterior of the CURRENT sample as used by step k (the
     * pre_posterior_current_sample the reference writes into its data rows); 0 for MH */
    double *pre_cur;      /* DEVICE [K][C], may be NULL */
    /* > 0: leave room for kernels of another stream (the collector's refit running concurrently).  Warp-owned launch
     * shape (RBF surrogate, lanes_per_chain 0): the number of SMs left entirely free -- the launch uses sm_count -
     * coresident CTAs per wave.  Lane-group shape: at most one CTA of this kernel per SM (its shared-memory request is
     * padded past half of the SM's). */
    int32_t coresident;
    /* != 0 (additive, RBF surrogate only): the surrogate is evaluated in FP32 -- distances, kernel and the sum over
     * centers on the FP32 pipe, linear tail and the log-posterior in FP64.  The reference evaluates in float64
     * (surrogate_rbf.py:20-40); in a DAMH chain the surrogate only screens proposals, the accepted chain is corrected by
     * the true posterior in stage 2 (classes_SAMPLER.py:193-195), so the sampled distribution does not depend on it. */
    int32_t surrogate_f32;
    /* Restrict the launch to the chains [chain_lo, chain_lo + chain_n) of the state (chain_n == 0: all).  The split
     * phases use it to pipeline two halves of the chains against the host's solver pool: while the pool evaluates G for
     * the survivors of one half, the other half is decided / proposed (process_SOLVER.py:60-106 keeps its solvers busy
     * the same way by queueing requests).  Results do not depend on the split: the RNG is keyed by the global chain id. */
    int64_t chain_lo;
    int64_t chain_n;
} sdb_chain_args;

/* ---- library ------------------------------------------------------------------------- */
SDB_API int sdb_abi_version(void);
SDB_API const char *sdb_last_error(void);
/* number of SMs, compute capability and opt-in shared memory per block of the current device */
SDB_API int sdb_device_info(int *sm_count, int *cc_major, int *cc_minor, int *smem_optin_bytes);

/* ---- surrogate evaluation: points [N x d] row-major -> out [N x m] row-major ------------ */
SDB_API int sdb_poly_eval(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model, double *d_out,
                  void *stream);
SDB_API int sdb_rbf_eval(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model, double *d_out,
                 void *stream);

/* The same contraction on the FP32 pipe (additive; Surrogate_apply(..., precision="f32")): centers relative to the first
 * center and weights held in FP32 in shared memory (all of them: SDB_ERR_LIMIT when they do not fit), per-lane partial
 * sums combined and the linear tail added in FP64.  |error| is about 1e-7 * sum_j |w_j phi_j|. */
SDB_API int sdb_rbf_eval_f32(const double *d_points, int64_t N, int32_t d, int32_t m, const sdb_model *model, double *d_out,
                     void *stream);

/* ---- lockstep chains ------------------------------------------------------------------ */
SDB_API int sdb_chain_run(const sdb_chain_args *args, void *stream);
/* Threads per block / blocks the library would use for these args (for reporting). */
SDB_API int sdb_chain_plan(const sdb_chain_args *args, int *lanes_per_chain, int *threads_per_block, int *blocks,
                   int *smem_bytes);
/* Fill z [K][d][C] and uni [K][2][C] with exactly the values SDB_RNG_PHILOX would use. */
SDB_API int sdb_philox_streams(uint64_t seed, uint32_t stream_id, int64_t chain0, int64_t step0, int32_t K, int32_t d,
                       int64_t C, double *d_z, double *d_uni, void *stream);
/* Raw Philox-4x32-10 blocks: out[i] = philox(key=(seed lo,hi), ctr=(c0+i, c1, c2, c3)) as 4 uint32. */
SDB_API int sdb_philox_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int64_t n, uint32_t *d_out,
                   void *stream);

/* Gather the valid snapshot rows of one launch in (step, chain) order.
 * flags [K][C]; snap_par [K][d][C]; snap_obs [K][m][C] -> out_par [cap x d], out_obs [cap x m] row-major.
 * d_count (DEVICE int64) receives the number of valid rows; rows beyond `cap` are dropped.
 * d_scratch: DEVICE int64 [ceil(K*C/1024) + 1]. */
SDB_API int sdb_snapshot_compact(const uint8_t *d_flags, const double *d_snap_par, const double *d_snap_obs, int32_t K,
                         int64_t C, int32_t d, int32_t m, double *d_out_par, double *d_out_obs, int64_t cap,
                         int64_t *d_count, int64_t *d_scratch, void *stream);

/* ---- collector round with fixed shapes (no host synchronisation) --------------------------
 * The sampler side of classes_communication.py:121-135 and the collector's add_data (process_COLLECTOR.py:87-104,
 * surrogate_rbf.py:76-90) when every buffer has a static size: each rank contributes at most `quota` snapshots per
 * round (additive configuration key snapshots_per_refit; the reference forwards every snapshot). */
/* At most `quota` of the valid snapshots of one launch, evenly spaced over their (step, chain) order, into
 * d_block [(quota + 1) x (d + m)] row-major: row 0 = {rows kept, valid snapshots in the launch, 0...}, rows 1.. =
 * [parameters | observations], zero beyond the rows kept.  d_scratch: DEVICE int64 [ceil(K*C/1024) + 2]. */
SDB_API int sdb_snapshot_select(const uint8_t *d_flags, const double *d_snap_par, const double *d_snap_obs, int32_t K,
                                int64_t C, int32_t d, int32_t m, double *d_block, int64_t quota, int64_t *d_scratch,
                                void *stream);
/* Candidate set of a refit: the n0 rows the updater holds (X0 [n0 x d], Y0 [n0 x m]) followed by the gathered blocks of
 * `world` ranks (each laid out as above) -> d_X [(n0 + world*quota) x d], d_Y [... x m], d_alive (uint8) = 1 for the held
 * rows and for each rank's kept rows.  d_total (DEVICE int64, may be NULL) = number of new snapshots. */
SDB_API int sdb_round_candidates(const double *d_X0, const double *d_Y0, int64_t n0, const double *d_gathered, int32_t world,
                                 int64_t quota, int32_t d, int32_t m, double *d_X, double *d_Y, uint8_t *d_alive,
                                 int64_t *d_total, void *stream);
/* Rows with d_keep != 0, in order, into [n_out x d] / [n_out x m] (the deletion at surrogate_rbf.py:117-119 on fixed-size
 * buffers).  d_count: DEVICE int32 = rows kept (may exceed or fall short of n_out; the caller checks). */
SDB_API int sdb_rows_compact(const double *d_X, const double *d_Y, const uint8_t *d_keep, int64_t n, int32_t d, int32_t m,
                             int64_t n_out, double *d_outX, double *d_outY, int32_t *d_count, void *stream);

/* ---- surrogate update ------------------------------------------------------------------ */
/* Basis matrix PHI [N x P] row-major of the normalised-Hermite tensor basis (surrogate_poly.py:76-79). */
SDB_API int sdb_poly_basis(const double *d_points, int64_t N, int32_t d, const sdb_model *model, double *d_PHI, void *stream);
/* MAT = pinv(PHI^T PHI) PHI^T Y with numpy's default cutoff (1e-15 * largest singular value),
 * surrogate_poly.py:93-95.  X [n x d], Y [n x m] row-major; model supplies alpha / herm / degree /
 * n_terms (coefs unused); d_MAT [P x m] row-major.  d_work: DEVICE doubles, at least
 * sdb_poly_fit_work(n, P, m).  d_info: DEVICE int32[4] = {directions kept, Jacobi sweeps, converged, 0}. */
SDB_API int64_t sdb_poly_fit_work(int64_t n, int32_t P, int32_t m);
SDB_API int sdb_poly_fit(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, const sdb_model *model,
                         double *d_MAT, double *d_work, int32_t *d_info, void *stream);
/* Saddle-point interpolation system S = [[phi(D), P],[P^T, 0]], P = [X, 1] (surrogate_rbf.py:91-96,
 * 121-128): d_S [(n+d+1)^2] (symmetric), d_RHS [m][n+d+1] = m contiguous columns [Y_k; 0] (may be NULL). */
SDB_API int sdb_rbf_system(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                           double *d_S, double *d_RHS, void *stream);
/* LU with partial pivoting (LAPACK dgesv conventions) of the column-major N x N matrix d_A, in
 * place, then A X = B for the nrhs columns of d_B (column-major, leading dimension N, overwritten).
 * d_piv: DEVICE int32[N].  d_info: DEVICE int32, 0 or k+1 if U(k,k) is exactly zero. */
SDB_API int sdb_lu_solve(double *d_A, int64_t N, double *d_B, int32_t nrhs, int32_t *d_piv, int32_t *d_info, void *stream);
/* sdb_rbf_system + sdb_lu_solve (the reference's solver_type='direct', surrogate_rbf.py:129-130).
 * X [n x d], Y [n x m] -> d_COEFS [(n+d+1) x m] row-major.  d_work: at least sdb_rbf_fit_work(n, d, m). */
SDB_API int64_t sdb_rbf_fit_work(int64_t n, int32_t d, int32_t m);
SDB_API int sdb_rbf_fit(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                        double *d_COEFS, double *d_work, int32_t *d_info, void *stream);
/* The same system solved by the null-space method + Cholesky (no pivoting): Householder QR of P = [X, 1],
 * B = Q^T phi(D) Q, Cholesky of the conditionally definite block.  Kernel types 0, 1, 4, 5, 6 only
 * (SDB_ERR_ARG otherwise).  d_info: 0, or k+1 if the k-th Cholesky pivot was not positive (fall back to
 * sdb_rbf_fit).  d_work: at least sdb_rbf_fit_nullspace_work(n, d, m).  An additive solver
 * (Surrogate_update(solver_type='nullspace')); the reference has 'direct' and 'minres' only. */
SDB_API int64_t sdb_rbf_fit_nullspace_work(int64_t n, int32_t d, int32_t m);
SDB_API int sdb_rbf_fit_nullspace(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type,
                                  double *d_COEFS, double *d_work, int32_t *d_info, void *stream);
/* MINRES on a symmetric N x N matrix (scipy.sparse.linalg.minres as called at surrogate_rbf.py:131-135:
 * no shift, no preconditioner, stop on rtol).  d_x0 may be NULL.  d_work: at least 8*N + 64 doubles.
 * d_info: DEVICE int32[4] = {0 or maxiter (scipy's info), iterations, istop, 0}.  SYNCHRONISES the stream. */
SDB_API int sdb_minres(const double *d_A, int64_t N, const double *d_b, const double *d_x0, double *d_x, double rtol,
                       int64_t maxiter, double *d_work, int32_t *d_info, void *stream);
/* Greedy thinning to no_keep points (surrogate_rbf.py:97-120): d_keep (DEVICE uint8 [n]) 1 = kept.
 * d_work: DEVICE doubles, at least sdb_rbf_thin_work(n). */
SDB_API int64_t sdb_rbf_thin_work(int64_t n);
SDB_API int sdb_rbf_thin(const double *d_X, int64_t n, int32_t d, int64_t no_keep, uint8_t *d_keep, double *d_work,
                         void *stream);
/* The same on a fixed-size candidate block: rows with d_alive[i] == 0 (d_alive may be NULL = all alive) are absent from
 * the start (d_keep = 0) and do not count towards no_keep.  drop_coincident != 0 (additive): once no_keep is reached the
 * removal goes on while the closest remaining pair has distance exactly 0 -- identical snapshots (every lockstep chain
 * leaving the same start point) would make the saddle matrix singular, where np.linalg.solve raises. */
SDB_API int sdb_rbf_thin_masked(const double *d_X, int64_t n, int32_t d, int64_t no_keep, const uint8_t *d_alive,
                                int32_t drop_coincident, uint8_t *d_keep, double *d_work, void *stream);

/* FP64 tensor-pipe GEMM used by the refit: C (column-major, ldc) = beta C + alpha A B,
 * A(i,k) = A[i*sa_m + k*sa_k], B(k,j) = B[k*sb_k + j*sb_n].  nslab > 1: deterministic split-K through
 * d_ws (>= nslab*M*N doubles). */
SDB_API int sdb_dgemm(int32_t M, int32_t N, int32_t K, double alpha, const double *d_A, int64_t sa_m, int64_t sa_k,
                      const double *d_B, int64_t sb_k, int64_t sb_n, double beta, double *d_C, int64_t ldc, int32_t nslab,
                      double *d_ws, void *stream);

/* ---- measurement ----------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process (reset != 0 zeroes the counter). */
SDB_API int64_t sdb_launch_count(int32_t reset);
/* FP64 FMA throughput probe (roofline denominator of the FP64-pipe-bound kernels):
 * FLOPs = blocks * 256 * 8 * 2 * iters.  d_out: DEVICE doubles [blocks * 256]. */
SDB_API int sdb_fp64_probe(int64_t iters, int32_t blocks, double *d_out, void *stream);
/* FP32 FMA throughput (denominator for the FP32-surrogate chain kernel), same FLOP count.  d_out: DEVICE floats [blocks * 256]. */
SDB_API int sdb_fp32_probe(int64_t iters, int32_t blocks, float *d_out, void *stream);
/* Same FLOP count, but each FMA reads three distinct live registers (the operand pattern of real code). */
SDB_API int sdb_fp64_probe3(int64_t iters, int32_t blocks, double *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SURRDAMH_B200_H */
