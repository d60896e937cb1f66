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
