// Library plumbing: error text, device properties, the Philox stream dump used by the parity
// tests, and the deterministic snapshot compaction.
#include <stdarg.h>
#include <string.h>

#include "sdb_common.cuh"

static thread_local char g_err[512] = "";

void sdb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void sdb_count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
extern "C" int64_t sdb_launch_count(int32_t reset) {
    const long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, 0, __ATOMIC_RELAXED);
    return v;
}

extern "C" const char *sdb_last_error(void) { return g_err; }
extern "C" int sdb_abi_version(void) { return SDB_ABI_VERSION; }

int sdb_get_device_props(sdb_device_props *out) {
    static sdb_device_props cache[64];
    static bool have[64] = {false};
    int dev = 0;
    SDB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !have[dev]) {
        sdb_device_props p;
        SDB_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SDB_CUDA(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        SDB_CUDA(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        SDB_CUDA(cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (dev < 0 || dev >= 64) { *out = p; return SDB_OK; }
        cache[dev] = p;
        have[dev] = true;
    }
    *out = cache[dev];
    return SDB_OK;
}

extern "C" int sdb_device_info(int *sm_count, int *cc_major, int *cc_minor, int *smem_optin_bytes) {
    sdb_device_props p;
    int rc = sdb_get_device_props(&p);
    if (rc) return rc;
    if (sm_count) *sm_count = p.sm_count;
    if (cc_major) *cc_major = p.cc_major;
    if (cc_minor) *cc_minor = p.cc_minor;
    if (smem_optin_bytes) *smem_optin_bytes = p.smem_optin;
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void philox_streams_kernel(uint64_t seed, uint32_t stream_id, int64_t chain0, int64_t step0, int K, int d,
                                      int64_t C, double *z, double *uni) {
    const int64_t total = (int64_t)K * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / C);
        const int64_t c = i % C;
        const uint64_t gchain = (uint64_t)(chain0 + c), gstep = (uint64_t)(step0 + k);
        if (uni) {
            const philox4 b = sdb_chain_block(seed, stream_id, gchain, gstep, SDB_SLOT_UNIFORMS);
            uni[((size_t)k * 2 + 0) * C + c] = sdb_u01(b.x, b.y);
            uni[((size_t)k * 2 + 1) * C + c] = sdb_u01(b.z, b.w);
        }
        if (z) {
            for (int j = 0; j < d; j += 2) {
                double z0, z1;
                sdb_normal_pair(sdb_chain_block(seed, stream_id, gchain, gstep, (uint32_t)(j >> 1)), z0, z1);
                z[((size_t)k * d + j) * C + c] = z0;
                if (j + 1 < d) z[((size_t)k * d + j + 1) * C + c] = z1;
            }
        }
    }
}

extern "C" int sdb_philox_streams(uint64_t seed, uint32_t stream_id, int64_t chain0, int64_t step0, int32_t K, int32_t d,
                                  int64_t C, double *d_z, double *d_uni, void *stream) {
    SDB_CHECK_ARG(K >= 1 && d >= 1 && C >= 1, "sdb_philox_streams: K, d, C must be >= 1");
    const int64_t total = (int64_t)K * C;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    sdb_count_launch();
    philox_streams_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, chain0, step0, K, d, C, d_z, d_uni);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

__global__ void philox_raw_kernel(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int64_t n, uint32_t *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const philox4 b = philox4x32_10(c0 + (uint32_t)i, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
        out[i * 4 + 0] = b.x; out[i * 4 + 1] = b.y; out[i * 4 + 2] = b.z; out[i * 4 + 3] = b.w;
    }
}

extern "C" int sdb_philox_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int64_t n, uint32_t *d_out,
                              void *stream) {
    SDB_CHECK_ARG(n >= 1 && d_out, "sdb_philox_raw: n >= 1 and a destination are required");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    sdb_count_launch();
    philox_raw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, c0, c1, c2, c3, n, d_out);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Snapshot compaction in (step, chain) order: per-1024-entry block counts -> exclusive scan by one
// CTA -> ordered scatter.  Deterministic (no atomics).
// ---------------------------------------------------------------------------------------------------
#define CB 1024

__global__ void __launch_bounds__(256) compact_count(const uint8_t *__restrict__ flags, int64_t total, int64_t *counts) {
    const int64_t base = (int64_t)blockIdx.x * CB;
    int cnt = 0;
    for (int i = threadIdx.x; i < CB; i += 256) {
        const int64_t e = base + i;
        if (e < total) cnt += (flags[e] == SDB_FLAG_ACCEPTED || flags[e] == SDB_FLAG_REJECTED);
    }
    __shared__ int ws[8];
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) compact_scan(int64_t *counts, int64_t nblocks, int64_t *total_out) {
    // single CTA exclusive scan, in place; counts[nblocks] receives the total
    __shared__ int64_t carry;
    __shared__ int64_t ws[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int64_t v = i < nblocks ? counts[i] : 0, x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t w = ws[threadIdx.x], xs = w;
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, xs, off);
                if (threadIdx.x >= off) xs += y;
            }
            ws[threadIdx.x] = xs - w;  // exclusive warp offsets
        }
        __syncthreads();
        const int64_t incl = x + ws[threadIdx.x >> 5] + carry;
        if (i < nblocks) counts[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[nblocks] = carry; *total_out = carry; }
}

__global__ void __launch_bounds__(256) compact_scatter(const uint8_t *__restrict__ flags, const double *__restrict__ sp,
                                                       const double *__restrict__ so, int64_t total, int64_t C, int d, int m,
                                                       const int64_t *__restrict__ offsets, double *op, double *oo, int64_t cap) {
    // one warp walks 32 entries at a time in order; ballot gives the in-order rank
    __shared__ int wcount[8];
    const int64_t base = (int64_t)blockIdx.x * CB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // each warp owns a contiguous 128-entry slice of the block
    const int64_t wbase = base + warp * 128;
    unsigned masks[4];
    int cnt = 0;
    for (int r = 0; r < 4; ++r) {
        const int64_t e = wbase + r * 32 + lane;
        const bool v = e < total && (flags[e] == SDB_FLAG_ACCEPTED || flags[e] == SDB_FLAG_REJECTED);
        masks[r] = __ballot_sync(0xffffffffu, v);
        cnt += __popc(masks[r]);
    }
    if (lane == 0) wcount[warp] = cnt;
    __syncthreads();
    int64_t pos = offsets[blockIdx.x];
    for (int w = 0; w < warp; ++w) pos += wcount[w];
    for (int r = 0; r < 4; ++r) {
        const int64_t e = wbase + r * 32 + lane;
        if (masks[r] & (1u << lane)) {
            const int64_t o = pos + __popc(masks[r] & ((1u << lane) - 1));
            if (o < cap) {
                const int64_t k = e / C, c = e % C;
                for (int j = 0; j < d; ++j) op[o * d + j] = sp[(k * d + j) * C + c];
                for (int j = 0; j < m; ++j) oo[o * m + j] = so[(k * m + j) * C + c];
            }
        }
        pos += __popc(masks[r]);
    }
}

extern "C" int sdb_snapshot_compact(const uint8_t *d_flags, const double *d_snap_par, const double *d_snap_obs, int32_t K,
                                    int64_t C, int32_t d, int32_t m, double *d_out_par, double *d_out_obs, int64_t cap,
                                    int64_t *d_count, int64_t *d_scratch, void *stream) {
    SDB_CHECK_ARG(d_flags && d_snap_par && d_snap_obs && d_out_par && d_out_obs && d_count && d_scratch,
                  "sdb_snapshot_compact: NULL pointer");
    SDB_CHECK_ARG(K >= 1 && C >= 1 && d >= 1 && m >= 1 && cap >= 0, "sdb_snapshot_compact: bad sizes");
    const int64_t total = (int64_t)K * C;
    const int64_t nblocks = (total + CB - 1) / CB;
    cudaStream_t st = (cudaStream_t)stream;
    sdb_count_launch();
    compact_count<<<(unsigned)nblocks, 256, 0, st>>>(d_flags, total, d_scratch);
    sdb_count_launch();
    compact_scan<<<1, 1024, 0, st>>>(d_scratch, nblocks, d_count);
    sdb_count_launch();
    compact_scatter<<<(unsigned)nblocks, 256, 0, st>>>(d_flags, d_snap_par, d_snap_obs, total, C, d, m, d_scratch, d_out_par,
                                                       d_out_obs, cap);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Fixed-shape collector round (no host synchronisation): fair snapshot selection, candidate block, row compaction.
// ---------------------------------------------------------------------------------------------------
// Of the T valid snapshots of a launch, in (step, chain) order, at most `quota` are kept, evenly spaced over that order:
// ordinal o is kept iff floor((o+1) quota / T) > floor(o quota / T) and lands in slot floor((o+1) quota / T) - 1
// (all of them when T <= quota).  Block layout [(quota + 1) x (d + m)] row-major: row 0 = header {rows kept, T, 0...},
// rows 1.. = [parameters | observations].
__global__ void __launch_bounds__(256) select_scatter(const uint8_t *__restrict__ flags, const double *__restrict__ sp,
                                                      const double *__restrict__ so, int64_t total, int64_t C, int d, int m,
                                                      const int64_t *__restrict__ offsets, int64_t nblocks, double *block,
                                                      int64_t quota) {
    __shared__ int wcount[8];
    const int64_t base = (int64_t)blockIdx.x * CB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = base + warp * 128;
    const int64_t T = offsets[nblocks];
    const int w = d + m;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        block[0] = (double)(T < quota ? T : quota);
        block[1] = (double)T;
        for (int j = 2; j < w; ++j) block[j] = 0.0;
    }
    unsigned masks[4];
    int cnt = 0;
    for (int r = 0; r < 4; ++r) {
        const int64_t e = wbase + r * 32 + lane;
        const bool v = e < total && (flags[e] == SDB_FLAG_ACCEPTED || flags[e] == SDB_FLAG_REJECTED);
        masks[r] = __ballot_sync(0xffffffffu, v);
        cnt += __popc(masks[r]);
    }
    if (lane == 0) wcount[warp] = cnt;
    __syncthreads();
    int64_t pos = offsets[blockIdx.x];
    for (int k = 0; k < warp; ++k) pos += wcount[k];
    for (int r = 0; r < 4; ++r) {
        const int64_t e = wbase + r * 32 + lane;
        if (masks[r] & (1u << lane)) {
            const int64_t o = pos + __popc(masks[r] & ((1u << lane) - 1));
            int64_t slot = o;
            bool take = true;
            if (T > quota) {
                const int64_t a = (o * quota) / T, b = ((o + 1) * quota) / T;
                take = b > a;
                slot = b - 1;
            }
            if (take) {
                const int64_t k = e / C, c = e % C;
                double *row = block + (slot + 1) * w;
                for (int j = 0; j < d; ++j) row[j] = sp[(k * d + j) * C + c];
                for (int j = 0; j < m; ++j) row[d + j] = so[(k * m + j) * C + c];
            }
        }
        pos += __popc(masks[r]);
    }
}

extern "C" int sdb_snapshot_select(const uint8_t *d_flags, const double *d_snap_par, const double *d_snap_obs, int32_t K,
                                   int64_t C, int32_t d, int32_t m, double *d_block, int64_t quota, int64_t *d_scratch,
                                   void *stream) {
    SDB_CHECK_ARG(d_flags && d_snap_par && d_snap_obs && d_block && d_scratch, "sdb_snapshot_select: NULL pointer");
    SDB_CHECK_ARG(K >= 1 && C >= 1 && d >= 1 && m >= 1 && quota >= 1, "sdb_snapshot_select: bad sizes");
    const int64_t total = (int64_t)K * C;
    const int64_t nblocks = (total + CB - 1) / CB;
    cudaStream_t st = (cudaStream_t)stream;
    SDB_CUDA(cudaMemsetAsync(d_block, 0, (size_t)(quota + 1) * (d + m) * sizeof(double), st));
    sdb_count_launch();
    compact_count<<<(unsigned)nblocks, 256, 0, st>>>(d_flags, total, d_scratch);
    sdb_count_launch();
    compact_scan<<<1, 1024, 0, st>>>(d_scratch, nblocks, d_scratch + nblocks + 1);
    sdb_count_launch();
    select_scatter<<<(unsigned)nblocks, 256, 0, st>>>(d_flags, d_snap_par, d_snap_obs, total, C, d, m, d_scratch, nblocks, d_block,
                                                      quota);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// Candidate set of one refit round: the n0 centers the updater holds, then the gathered per-rank blocks
// (world x [(quota + 1) x (d + m)], the layout above).  alive[i] = 1 for the centers and for the first `rows kept` rows
// of every rank.  d_total (DEVICE int64, may be NULL) receives the number of new snapshots.
__global__ void round_candidates_kernel(const double *__restrict__ X0, const double *__restrict__ Y0, int n0,
                                        const double *__restrict__ gathered, int world, int quota, int d, int m,
                                        double *__restrict__ X, double *__restrict__ Y, uint8_t *__restrict__ alive,
                                        int64_t *total) {
    const int n = n0 + world * quota;
    const int w = d + m;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < n0) {
            for (int j = 0; j < d; ++j) X[(size_t)i * d + j] = X0[(size_t)i * d + j];
            for (int j = 0; j < m; ++j) Y[(size_t)i * m + j] = Y0[(size_t)i * m + j];
            alive[i] = 1;
        } else {
            const int r = (i - n0) / quota, k = (i - n0) % quota;
            const double *blk = gathered + (size_t)r * (quota + 1) * w;
            const double *row = blk + (size_t)(k + 1) * w;
            for (int j = 0; j < d; ++j) X[(size_t)i * d + j] = row[j];
            for (int j = 0; j < m; ++j) Y[(size_t)i * m + j] = row[d + j];
            alive[i] = (double)k < blk[0];
        }
    }
    if (total && blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t t = 0;
        for (int r = 0; r < world; ++r) t += (int64_t)gathered[(size_t)r * (quota + 1) * w];
        *total = t;
    }
}

extern "C" int sdb_round_candidates(const double *d_X0, const double *d_Y0, int64_t n0, const double *d_gathered, int32_t world,
                                    int64_t quota, int32_t d, int32_t m, double *d_X, double *d_Y, uint8_t *d_alive,
                                    int64_t *d_total, void *stream) {
    SDB_CHECK_ARG(d_gathered && d_X && d_Y && d_alive && (n0 == 0 || (d_X0 && d_Y0)), "sdb_round_candidates: NULL pointer");
    SDB_CHECK_ARG(n0 >= 0 && world >= 1 && quota >= 1 && d >= 1 && m >= 1 && n0 + world * quota < (1 << 30),
                  "sdb_round_candidates: bad sizes");
    const int64_t n = n0 + (int64_t)world * quota;
    sdb_count_launch();
    round_candidates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_X0, d_Y0, (int)n0, d_gathered, world, (int)quota, d, m, d_X, d_Y, d_alive, d_total);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}

// Rows with keep != 0, in order, into fixed-size outputs [n_out x d], [n_out x m] (rows beyond the number kept are left
// untouched, rows beyond n_out dropped).  d_count: DEVICE int32 = number of kept rows.  One CTA (n is a few thousand).
__global__ void __launch_bounds__(1024) rows_compact_kernel(const double *__restrict__ X, const double *__restrict__ Y,
                                                            const uint8_t *__restrict__ keep, int n, int d, int m, int n_out,
                                                            double *__restrict__ oX, double *__restrict__ oY, int *count) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const bool k = i < n && keep[i];
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int pos = carry;
        for (int w = 0; w < warp; ++w) pos += wsum[w];
        pos += __popc(bal & ((1u << lane) - 1));
        if (k && pos < n_out) {
            for (int j = 0; j < d; ++j) oX[(size_t)pos * d + j] = X[(size_t)i * d + j];
            for (int j = 0; j < m; ++j) oY[(size_t)pos * m + j] = Y[(size_t)i * m + j];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += wsum[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && count) *count = carry;
}

extern "C" int sdb_rows_compact(const double *d_X, const double *d_Y, const uint8_t *d_keep, int64_t n, int32_t d, int32_t m,
                                int64_t n_out, double *d_outX, double *d_outY, int32_t *d_count, void *stream) {
    SDB_CHECK_ARG(d_X && d_Y && d_keep && d_outX && d_outY, "sdb_rows_compact: NULL pointer");
    SDB_CHECK_ARG(n >= 1 && n < (1 << 30) && d >= 1 && m >= 1 && n_out >= 0, "sdb_rows_compact: bad sizes");
    sdb_count_launch();
    rows_compact_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_X, d_Y, d_keep, (int)n, d, m, (int)n_out, d_outX, d_outY, d_count);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
}
