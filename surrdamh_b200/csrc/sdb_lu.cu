// Dense LU with partial pivoting + triangular solves, FP64, on the device (sm_100a).
// Replaces np.linalg.solve (LAPACK dgesv) at surrogate_rbf.py:129-130 (solver_type='direct').
//
// Column-major N x N matrix A (leading dimension lda), factored in place as P A = L U with the
// LAPACK conventions (unit lower L, pivot row indices piv[k] >= k applied in order).
//
//   * blocked right-looking outer loop, panel width LU_NB
//   * inside a panel: the (N - j) x w sub-panel is factored by ONE kernel that keeps it in the shared
//     memory of a thread-block cluster of 16 CTAs (16 SMs); per column one block-level arg-max, one
//     exchange of the candidate rows through distributed shared memory and ONE cluster barrier.  w is 64
//     while the rows fit (N <= ~5800), else the panel is halved recursively (w = 32, 16, 8)
//   * everything else (row interchanges, triangular solves with the 8..64-wide diagonal blocks,
//     rank-k updates) runs on the whole GPU; the updates use the DMMA GEMM of sdb_gemm.cuh
//   * solve phase: warp-per-right-hand-side triangular solves on 64 x 64 diagonal blocks + GEMM
#include <cooperative_groups.h>
#include <limits.h>
#include <stdlib.h>

#include "sdb_gemm.cuh"
#include "sdb_graph.cuh"

namespace cg = cooperative_groups;

#define LU_T 512   // threads per CTA of the base case
#define LU_NB 64   // outer panel width

// ---------------------------------------------------------------------------------------------------
// Base case: factor the (N - j0) x w sub-panel (w <= W) starting at A(j0, j0); row interchanges are
// applied to these w columns only.  The sub-panel lives in the SHARED MEMORY of one thread-block cluster
// (CL = 8 or 16 CTAs): CTA `rank` holds rows [j0 + rank*rpc, j0 + (rank+1)*rpc) as tile[rpc][W+1].
// Per column: block arg-max -> warp `r` of the winning CTA writes the candidate row (w values + |pivot| +
// row index) into the mailbox of CTA r over distributed shared memory, the CTA holding the diagonal row
// does the same with that row -> ONE cluster barrier -> every thread picks the global winner from its
// own CTA's mailbox, the two owners swap their rows, every thread eliminates its rows.  Mailboxes are
// double-buffered on the column parity.  Loops over columns are runtime loops: the code stays small
// (the fully unrolled register variant of this kernel fell out of the instruction cache beyond 8 columns).
// ---------------------------------------------------------------------------------------------------
#define LU_CLMAX 16
#ifdef SDB_LU_PROFILE
__device__ long long g_lu_prof[8];
#define LU_STAMP(k) do { const long long t_ = clock64(); t_acc[k] += t_ - t_last; t_last = t_; } while (0)
#else
#define LU_STAMP(k) do { } while (0)
#endif
template <int W>
struct lu_mail {
    double val[LU_CLMAX];
    int row[LU_CLMAX];
    double vals[LU_CLMAX][W];
};

template <int W>
__global__ void __launch_bounds__(LU_T, 1)
    lu_base_kernel(double *__restrict__ A, int64_t lda, int N, int j0, int w, int rpc, int *__restrict__ piv, int *__restrict__ info) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), CL = (int)cluster.num_blocks();
    extern __shared__ __align__(16) double tile[];   // [rpc][W + 2], then int lrow[rpc], int done[rpc]
    __shared__ lu_mail<W> mail[2];
    __shared__ double rcp_s[W];
    __shared__ int piv_s[W];
    __shared__ int sing_s;
    __shared__ double red_val[LU_T / 32];
    __shared__ int red_row[LU_T / 32], red_slot[LU_T / 32];
    __shared__ double win_val;
    __shared__ int win_row, win_slot;
    constexpr int TS = W + 2;                           // even: (row, even column) is 16-byte aligned; 33 i mod 8 banks
    // elimination work item = (row, 16-column chunk); a chunk must not reach past the row: with W = 8 (sub-panels of more than
    // ~20 000 rows) a 16-wide chunk read AND WROTE BACK the first columns of the next slot's row, racing with that row's own
    // thread (backward error 5e-7 at n = 32768, found by the round-2 refit sweep against np.linalg.solve)
    constexpr int CH = W < 16 ? W : 16, NCH = (W + CH - 1) / CH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = j0 + rank * rpc;                       // first row of this CTA
    const int nloc = max(0, min(rpc, N - row0));            // rows it really holds
    int *lrow = (int *)(tile + (size_t)rpc * TS);           // logical row of every slot (rows are never moved)
    int *done = lrow + rpc;                                 // column at which the slot became a pivot row (w: not yet)

    for (int c = 0; c < w; ++c)
        for (int i = tid; i < nloc; i += LU_T) tile[i * TS + c] = A[(int64_t)(j0 + c) * lda + row0 + i];
    for (int i = tid; i < nloc; i += LU_T) { lrow[i] = row0 + i; done[i] = w; }
    if (tid == 0) sing_s = 0;
    __syncthreads();

#ifdef SDB_LU_PROFILE
    long long t_acc[5] = {0, 0, 0, 0, 0};
    long long t_last = clock64();
#endif
    int prow = -1, drow = -1, pcol = -1;   // the interchange of the previous column, applied by the slot owners
    for (int c = 0; c < w; ++c) {
        lu_mail<W> &mb = mail[c & 1];
        LU_STAMP(0);
        // ---- owners apply the previous interchange to their slots' logical rows, then scan column c ---------
        double best = -1.0;
        int brow = INT_MAX, bslot = -1;
        for (int i = tid; i < nloc; i += LU_T) {
            if (done[i] == w && pcol >= 0) {
                const int lr = lrow[i];
                if (lr == prow) { lrow[i] = drow; done[i] = pcol; }
                else if (lr == drow) lrow[i] = prow;
            }
            if (done[i] == w) {
                const double v = fabs(tile[i * TS + c]);
                const int lr = lrow[i];
                if (v > best || (v == best && lr < brow)) { best = v; brow = lr; bslot = i; }
            }
        }
        drow = j0 + c;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
            const int oslot = __shfl_xor_sync(0xffffffffu, bslot, off);
            if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; bslot = oslot; }
        }
        if (lane == 0) { red_val[warp] = best; red_row[warp] = brow; red_slot[warp] = bslot; }
        __syncthreads();
        if (warp == 0) {
            double v = lane < LU_T / 32 ? red_val[lane] : -1.0;
            int r = lane < LU_T / 32 ? red_row[lane] : INT_MAX;
            int sl = lane < LU_T / 32 ? red_slot[lane] : -1;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, off);
                const int orow = __shfl_xor_sync(0xffffffffu, r, off);
                const int oslot = __shfl_xor_sync(0xffffffffu, sl, off);
                if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; sl = oslot; }
            }
            if (lane == 0) { win_val = v; win_row = r; win_slot = sl; }
        }
        __syncthreads();
        LU_STAMP(1);
        // ---- publish: warp r writes this CTA's candidate row into the mailbox of CTA r ------------------------
        if (warp < CL) {
            lu_mail<W> *dst = cluster.map_shared_rank(&mb, warp);
            const int wslot = win_slot;
            if (wslot < 0) {
                if (lane == 0) { dst->val[rank] = -1.0; dst->row[rank] = INT_MAX; }
            } else {
                const double *src = tile + (size_t)wslot * TS;
                for (int k = lane; k < w; k += 32) dst->vals[rank][k] = src[k];
                if (lane == 0) { dst->val[rank] = win_val; dst->row[rank] = win_row; }
            }
        }
        LU_STAMP(2);
        // cluster barrier with release / acquire at CLUSTER scope: orders the distributed-shared-memory writes
        // above.  (cooperative_groups' cluster.sync() adds a GPU-scope MEMBAR and an L1 invalidate, ~1 us each.)
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        LU_STAMP(3);
        // ---- every thread picks the global winner -----------------------------------------------------------
        double pv = -1.0;
        int pr = 0;
        prow = INT_MAX;
        for (int r = 0; r < CL; ++r) {
            const double v = mb.val[r];
            const int row = mb.row[r];
            if (v > pv || (v == pv && row < prow)) { pv = v; prow = row; pr = r; }
        }
        const bool singular = !(pv > 0.0);  // exactly zero (or NaN) column: LAPACK records info and moves on
        if (singular) prow = drow;
        pcol = c;
        const double *__restrict__ pvals = mb.vals[pr];
        const double rcp = singular ? 1.0 : 1.0 / pvals[c];  // dgetf2 scales by the reciprocal
        if (tid == 0) {   // kept in shared memory; global memory is not touched inside the column loop
            rcp_s[c] = rcp;
            piv_s[c] = prow;
            if (singular && sing_s == 0) sing_s = drow + 1;
        }
        if (!singular) {
            // rows still in play, except the new pivot row: a(i, c2) -= l_i * pivot(c2), l_i = a(i, c) / pivot(c).
            // The scaling of column c itself is deferred to the write-back (nobody else reads it any more).
            for (int it = tid; it < nloc * NCH; it += LU_T) {
                // consecutive threads take consecutive ROWS of one chunk (conflict-free 16-byte accesses)
                const int ch = it / nloc, i = it - ch * nloc;
                const int base = ch * CH;
                if (base + CH <= c + 1 || base >= w) continue;          // chunk entirely left of the pivot column
                if (done[i] != w || lrow[i] == prow) continue;
                double *__restrict__ t = tile + (size_t)i * TS;
                const double nl = -(t[c] * rcp);
                double2 tv[CH / 2], pvv[CH / 2];
#pragma unroll
                for (int k = 0; k < CH / 2; ++k) {      // all loads first: independent, pipelined
                    tv[k] = *reinterpret_cast<const double2 *>(t + base + 2 * k);
                    pvv[k] = *reinterpret_cast<const double2 *>(pvals + base + 2 * k);
                }
#pragma unroll
                for (int k = 0; k < CH / 2; ++k) {
                    const int c2 = base + 2 * k;
                    double2 o = tv[k];
                    if (c2 > c && c2 < w) o.x = fma(nl, pvv[k].x, o.x);
                    if (c2 + 1 > c && c2 + 1 < w) o.y = fma(nl, pvv[k].y, o.y);
                    *reinterpret_cast<double2 *>(t + base + 2 * k) = o;
                }
            }
        }
        __syncthreads();   // eliminated values and the slot states are read by other threads in the next column
        LU_STAMP(4);
    }
#ifdef SDB_LU_PROFILE
    if (rank == 0 && tid == 0)
        for (int k = 0; k < 5; ++k) g_lu_prof[k] += t_acc[k];
#endif
    // last interchange, then write back to the LOGICAL rows, scaling the sub-diagonal (L) entries
    for (int i = tid; i < nloc; i += LU_T) {
        if (done[i] == w && pcol >= 0) {
            const int lr = lrow[i];
            if (lr == prow) { lrow[i] = drow; done[i] = pcol; }
            else if (lr == drow) lrow[i] = prow;
        }
    }
    __syncthreads();
    if (rank == 0) {
        for (int c = tid; c < w; c += LU_T) piv[j0 + c] = piv_s[c];
        if (tid == 0 && sing_s != 0 && *info == 0) *info = sing_s;
    }
    for (int c = 0; c < w; ++c)
        for (int i = tid; i < nloc; i += LU_T) {
            double v = tile[i * TS + c];
            if (c < done[i]) v *= rcp_s[c];
            A[(int64_t)(j0 + c) * lda + lrow[i]] = v;
        }
}

// ---------------------------------------------------------------------------------------------------
// Row interchanges k = k0..k1-1 (row k <-> piv[k], in order; k1 - k0 <= 64) applied to columns [c0, c1).
// The sequence of swaps is first folded into its net permutation on the <= 2 (k1-k0) rows it touches
// (one warp, shared memory); every thread then moves one column: all loads are independent (no chain of
// dependent swaps), staged through shared memory, then stored.
// ---------------------------------------------------------------------------------------------------
#define LASWP_T 64
__global__ void __launch_bounds__(LASWP_T) lu_laswp_kernel(double *__restrict__ A, int64_t lda, int c0, int c1,
                                                          const int *__restrict__ piv, int k0, int k1) {
    extern __shared__ double stage[];      // [nS][LASWP_T]
    __shared__ int rows[128], cont[128];   // destination row, source row
    __shared__ int nS_s;
    const int w = k1 - k0, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        for (int i = lane; i < w; i += 32) { rows[i] = k0 + i; cont[i] = k0 + i; }
        const int piv_lo = lane < w ? piv[k0 + lane] : 0, piv_hi = lane + 32 < w ? piv[k0 + 32 + lane] : 0;  // one coalesced load
        __syncwarp();
        int nS = w;
        for (int k = 0; k < w; ++k) {
            const int p = __shfl_sync(0xffffffffu, k < 32 ? piv_lo : piv_hi, k & 31);
            if (p == k0 + k) continue;
            int idx;
            if (p < k1) idx = p - k0;
            else {
                unsigned hit = 0;
                int found = -1;
                for (int base = w; base < nS; base += 32) {
                    const int e = base + lane;
                    hit = __ballot_sync(0xffffffffu, e < nS && rows[e] == p);
                    if (hit) { found = base + __ffs(hit) - 1; break; }
                }
                if (found < 0) {
                    found = nS++;
                    if (lane == 0) { rows[found] = p; cont[found] = p; }
                }
                idx = found;
            }
            __syncwarp();
            if (lane == 0) { const int t = cont[k]; cont[k] = cont[idx]; cont[idx] = t; }
            __syncwarp();
        }
        if (lane == 0) nS_s = nS;
    }
    __syncthreads();
    const int nS = nS_s;
    const int col = c0 + blockIdx.x * LASWP_T + threadIdx.x;
    if (col >= c1) return;
    double *a = A + (int64_t)col * lda;
    for (int i0 = 0; i0 < nS; i0 += 16) {     // 16 independent loads in flight per thread
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = i0 + k;
            v[k] = (i < nS && cont[i] != rows[i]) ? a[cont[i]] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (i0 + k < nS) stage[(i0 + k) * LASWP_T + threadIdx.x] = v[k];
    }
    for (int i = 0; i < nS; ++i)
        if (cont[i] != rows[i]) a[rows[i]] = stage[i * LASWP_T + threadIdx.x];
}

static int lu_laswp(cudaStream_t st, double *A, int64_t lda, int c0, int c1, const int *piv, int k0, int k1) {
    if (c1 <= c0 || k1 <= k0) return SDB_OK;
    static bool attr_set = false;
    if (!attr_set) {
        SDB_CUDA(cudaFuncSetAttribute((const void *)lu_laswp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      128 * LASWP_T * 8));
        attr_set = true;
    }
    sdb_count_launch();
    lu_laswp_kernel<<<(c1 - c0 + LASWP_T - 1) / LASWP_T, LASWP_T, 2 * (k1 - k0) * LASWP_T * 8, st>>>(A, lda, c0, c1, piv, k0, k1);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
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

// ---------------------------------------------------------------------------------------------------
// Panel-internal rank-h update with h, ncols <= 32:  C[M x ncols] -= L[M x h] * U[h x ncols]
// (column-major, common leading dimension lda).  64 rows x 4 column groups per CTA; U in shared memory;
// each thread keeps its row of L in registers.  Replaces a GEMM launch whose 128 x 64 tiles would be
// mostly empty for these shapes.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lu_skinny_update_kernel(const double *__restrict__ Lp, const double *__restrict__ Up,
                                                               double *__restrict__ Cp, int64_t lda, int M, int h, int ncols) {
    __shared__ double Us[32][33];
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
        const int k = e & 31, c = e >> 5;
        Us[k][c] = (k < h && c < ncols) ? Up[(int64_t)c * lda + k] : 0.0;
    }
    __syncthreads();
    const int r = blockIdx.x * 64 + (threadIdx.x & 63), cg = threadIdx.x >> 6;  // columns cg*8 .. cg*8+7
    if (r >= M || cg * 8 >= ncols) return;
    double l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) l[k] = k < h ? Lp[(int64_t)k * lda + r] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        if (c < ncols) {
            double acc = Cp[(int64_t)c * lda + r];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = fma(-l[k], Us[k][c], acc);
            Cp[(int64_t)c * lda + r] = acc;
        }
    }
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

// Cluster size and widest base case the shared memory allows for `rows` rows.
struct lu_base_plan { int W, CL, rpc; size_t smem; };

static int lu_cluster_size() {
    // 16-CTA clusters are the non-portable maximum on sm_100; SDB_LU_CLUSTER overrides (8 is always legal)
    static int v = 0;
    if (!v) {
        const char *e = getenv("SDB_LU_CLUSTER");
        v = e ? atoi(e) : 16;
        if (v != 16 && v != 8 && v != 4 && v != 2 && v != 1) v = 8;
    }
    return v;
}

static lu_base_plan lu_plan_base(int rows) {
    const size_t budget = 180 * 1024;
    lu_base_plan p;
    p.CL = lu_cluster_size();
    p.rpc = (rows + p.CL - 1) / p.CL;
    const int widths[4] = {64, 32, 16, 8};
    static const int wmax = getenv("SDB_LU_BASE_MAX") ? atoi(getenv("SDB_LU_BASE_MAX")) : 64;   // tests: force the narrow bases
    for (int k = 0; k < 4; ++k) {
        p.W = widths[k];
        if (p.W > wmax && k < 3) continue;
        p.smem = (size_t)p.rpc * (p.W + 2) * sizeof(double) + (size_t)p.rpc * 2 * sizeof(int) + 16;
        if (p.smem <= budget) return p;
    }
    p.W = 0;
    return p;
}

template <int W>
static int lu_launch_base(cudaStream_t st, const lu_base_plan &p, double *A, int64_t lda, int N, int j0, int w, int *piv,
                          int *info) {
    static bool attr_set = false;
    if (!attr_set) {
        SDB_CUDA(cudaFuncSetAttribute((const void *)lu_base_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SDB_CUDA(cudaFuncSetAttribute((const void *)lu_base_kernel<W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.CL, 1, 1);
    cfg.blockDim = dim3(LU_T, 1, 1);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    sdb_count_launch();
    SDB_CUDA(cudaLaunchKernelEx(&cfg, lu_base_kernel<W>, A, lda, N, j0, w, p.rpc, piv, info));
    return SDB_OK;
}

static int lu_base_width(int rows) { return lu_plan_base(rows).W; }

static int lu_base(cudaStream_t st, double *A, int64_t lda, int N, int j0, int w, int *piv, int *info) {
    const lu_base_plan p = lu_plan_base(N - j0);
    if (p.W == 0) {
        sdb_set_error("sdb_lu: %d rows exceed the shared memory of one cluster", N - j0);
        return SDB_ERR_LIMIT;
    }
    switch (p.W) {
        case 64: return lu_launch_base<64>(st, p, A, lda, N, j0, w, piv, info);
        case 32: return lu_launch_base<32>(st, p, A, lda, N, j0, w, piv, info);
        case 16: return lu_launch_base<16>(st, p, A, lda, N, j0, w, piv, info);
        default: return lu_launch_base<8>(st, p, A, lda, N, j0, w, piv, info);
    }
}

// Recursive panel: columns [j0, j0 + w), rows [j0, N); interchanges applied inside the panel only.
static int lu_panel(cudaStream_t st, double *A, int64_t lda, int N, int j0, int w, int *piv, int *info) {
    const int bw = lu_base_width(N - j0);
    if (w <= bw) return lu_base(st, A, lda, N, j0, w, piv, info);
    const int h = ((w / 2 + bw - 1) / bw) * bw;  // left half, a multiple of the base width
    int rc = lu_panel(st, A, lda, N, j0, h, piv, info);
    if (rc) return rc;
    const int rcols = w - h;
    rc = lu_laswp(st, A, lda, j0 + h, j0 + w, piv, j0, j0 + h);
    if (rc) return rc;
    double *A11 = A + (int64_t)j0 * lda + j0;
    double *A12 = A + (int64_t)(j0 + h) * lda + j0;
    double *A21 = A11 + h;
    double *A22 = A12 + h;
    rc = lu_trsm_lunit(st, A11, lda, A12, lda, h, rcols);
    if (rc) return rc;
    if (h <= 32 && rcols <= 32) {
        const int M = N - j0 - h;
        if (M > 0) {
            sdb_count_launch();
            lu_skinny_update_kernel<<<(M + 63) / 64, 256, 0, st>>>(A21, A12, A22, lda, M, h, rcols);
        }
    } else {
        rc = sdb_gemm(st, N - j0 - h, rcols, h, -1.0, A21, 1, lda, A12, 1, lda, 1.0, A22, lda);
        if (rc) return rc;
    }
    rc = lu_panel(st, A, lda, N, j0 + h, rcols, piv, info);
    if (rc) return rc;
    return lu_laswp(st, A, lda, j0, j0 + h, piv, j0 + h, j0 + w);
}

static int lu_factor(cudaStream_t st, double *A, int64_t lda, int N, int *piv, int *info) {
    for (int j0 = 0; j0 < N; j0 += LU_NB) {
        const int w = (N - j0) < LU_NB ? (N - j0) : LU_NB;
        int rc = lu_panel(st, A, lda, N, j0, w, piv, info);
        if (rc) return rc;
        rc = lu_laswp(st, A, lda, 0, j0, piv, j0, j0 + w);
        if (rc) return rc;
        const int rest = N - j0 - w;
        if (rest > 0) {
            rc = lu_laswp(st, A, lda, j0 + w, N, piv, j0, j0 + w);
            if (rc) return rc;
            double *A11 = A + (int64_t)j0 * lda + j0;
            double *A12 = A + (int64_t)(j0 + w) * lda + j0;
            rc = lu_trsm_lunit(st, A11, lda, A12, lda, w, rest);
            if (rc) return rc;
            rc = sdb_gemm_k64_update(st, rest, rest, w, A11 + w, A12, A12 + w, lda);
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

static int lu_enqueue(cudaStream_t st, double *d_A, int64_t N, double *d_B, int nrhs, int32_t *d_piv, int32_t *d_info) {
    SDB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    int rc = lu_factor(st, d_A, N, (int)N, d_piv, d_info);
    if (rc) return rc;
    if (nrhs > 0) rc = lu_solve_factored(st, d_A, N, (int)N, d_piv, d_B, N, nrhs);
    return rc;
}

// Factor A (column-major N x N, overwritten by L and U) and solve A X = B for nrhs columns
// (B column-major N x nrhs, leading dimension N, overwritten by X).  d_piv: N ints.  d_info: 0, or
// k+1 when U(k,k) is exactly zero (LAPACK's info).  The kernel chain is replayed as a CUDA graph (sdb_graph.cuh).
extern "C" int sdb_lu_solve(double *d_A, int64_t N, double *d_B, int32_t nrhs, int32_t *d_piv, int32_t *d_info,
                            void *stream) {
    SDB_CHECK_ARG(d_A && d_piv && d_info, "sdb_lu_solve: NULL pointer");
    SDB_CHECK_ARG(N >= 1 && N < (1 << 30), "sdb_lu_solve: N = %lld out of range", (long long)N);
    SDB_CHECK_ARG(nrhs >= 0 && (nrhs == 0 || d_B), "sdb_lu_solve: bad right-hand sides");
    cudaStream_t st = (cudaStream_t)stream;
    if (N < 256) return lu_enqueue(st, d_A, N, d_B, nrhs, d_piv, d_info);   // not launch-bound enough to pay for an instantiation
    sdb_graph_key key = {{d_A, d_B, d_piv, d_info, nullptr}, {N, nrhs, 0, 0}, 1, 0};
    return sdb_graph_run(st, key, [&]() { return lu_enqueue(st, d_A, N, d_B, nrhs, d_piv, d_info); });
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

#ifdef SDB_LU_PROFILE
extern "C" SDB_API int sdb_lu_profile(long long *out8, int reset) {
    long long z[8] = {0};
    if (out8) cudaMemcpyFromSymbol(out8, g_lu_prof, sizeof(z));
    if (reset) cudaMemcpyToSymbol(g_lu_prof, z, sizeof(z));
    return 0;
}
#endif
