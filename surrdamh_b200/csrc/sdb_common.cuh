// Common device/host helpers for the surrdamh_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/surrdamh_b200.h"

// ---------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------
void sdb_set_error(const char *fmt, ...);

#define SDB_CHECK_ARG(cond, ...)         \
    do {                                 \
        if (!(cond)) {                   \
            sdb_set_error(__VA_ARGS__);  \
            return SDB_ERR_ARG;          \
        }                                \
    } while (0)

#define SDB_CHECK_LIMIT(cond, ...)       \
    do {                                 \
        if (!(cond)) {                   \
            sdb_set_error(__VA_ARGS__);  \
            return SDB_ERR_LIMIT;        \
        }                                \
    } while (0)

#define SDB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            sdb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return SDB_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

void sdb_count_launch();  // every kernel launch of the library is counted (sdb_launch_count)
extern "C" SDB_API int64_t sdb_launch_count(int32_t reset);

struct sdb_device_props {
    int sm_count, cc_major, cc_minor, smem_optin;
};
int sdb_get_device_props(sdb_device_props *out);  // cached per device
// sdb_rbf_system with an explicit leading dimension (sdb_fit.cu; used by the null-space solver)
int sdb_rbf_system_ld(const double *d_X, const double *d_Y, int64_t n, int32_t d, int32_t m, int32_t kernel_type, double *d_S,
                      double *d_RHS, int64_t ld, void *stream);

// ---------------------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al. 2011), counter-based: no state is carried between steps.
// ---------------------------------------------------------------------------------------
struct philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 53-bit uniform in [0,1) from two 32-bit words (hi, lo).
__host__ __device__ __forceinline__ double sdb_u01(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * 1.1102230246251565e-16;  // 2^-53
}
// same bits mapped to (0,1] (argument of the Box-Muller logarithm)
__host__ __device__ __forceinline__ double sdb_u01_open0(uint32_t hi, uint32_t lo) {
    const uint64_t v = ((((uint64_t)hi << 32) | lo) >> 11) + 1;
    return (double)v * 1.1102230246251565e-16;
}

#define SDB_SLOT_UNIFORMS 0xFFFFu

// Counter layout of the chain RNG: (chain, step lo, slot | step hi << 16, stream id).
// slot b < 0xFFFF : normals 2b, 2b+1 of that (chain, step);  slot 0xFFFF : the two uniforms.
__device__ __forceinline__ philox4 sdb_chain_block(uint64_t seed, uint32_t stream_id, uint64_t chain, uint64_t step,
                                                   uint32_t slot) {
    return philox4x32_10((uint32_t)chain, (uint32_t)step, slot | ((uint32_t)(step >> 32) << 16), stream_id,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Box-Muller pair from one Philox block.
__device__ __forceinline__ void sdb_normal_pair(const philox4 &b, double &z0, double &z1) {
    const double u1 = sdb_u01_open0(b.x, b.y);
    const double u2 = sdb_u01(b.z, b.w);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// ---------------------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared with an mbarrier (cp.async.bulk; SASS: UBLKCP).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sdb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sdb_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sdb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sdb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// raise the pending transaction count (no arrival)
__device__ __forceinline__ void sdb_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(sdb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sdb_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sdb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sdb_mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(sdb_smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void sdb_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sdb_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(sdb_smem_u32(bar))
                 : "memory");
}

// 2-D tensor-map TMA: one box of the mapped tensor, starting at element coordinates (c0 = inner, c1 = outer), into shared memory
// (128-byte aligned), completion by transaction count on `bar`.  `tmap` is a __grid_constant__ kernel parameter.
__device__ __forceinline__ void sdb_tensor_g2s_2d(void *smem_dst, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     sdb_smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(sdb_smem_u32(bar))
                 : "memory");
}

// One elected thread stages `bytes` (a multiple of 16, both pointers 16-byte aligned) in
// pieces of at most 32 KiB, raising the barrier's transaction count piece by piece.  The
// caller arrives on `bar` once after its last piece; the phase completes when that arrival
// and every byte have landed.
__device__ __forceinline__ void sdb_bulk_stage(void *smem_dst, const void *gmem_src, size_t bytes, uint64_t *bar) {
    const size_t PIECE = 32768;
    size_t off = 0;
    while (off < bytes) {
        const uint32_t n = (uint32_t)((bytes - off) < PIECE ? (bytes - off) : PIECE);
        sdb_mbar_expect_tx(bar, n);
        sdb_bulk_g2s((char *)smem_dst + off, (const char *)gmem_src + off, n, bar);
        off += n;
    }
}

// ---------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------
// Square roots for the pair loops.  y0 = MUFU.RSQ64H(s) (relative error e ~ 2^-22) and ONE third-order correction:
//   1/sqrt(s) = y0 (1 + q),  sqrt(s) = t (1 + q),  t = s y0,  q = e/2 + 3 e^2/8,  e = 1 - s y0^2
// (error ~ e^3: rounded-to-nearest quality).  The callers fold the (1 + q) factor into their last multiply, so
// r^3 = s^(3/2) costs 6 FP64-pipe operations from s:  t, e, p, q, a = s t, fma(a, q, a).
// The s == 0 / subnormal guard is an integer test on the high word (ALU pipe, not the FP64 pipe that bounds
// these loops); it yields y0 = 0, hence t = 0 and every kernel value 0 without a NaN.
struct sdb_rs {
    double y0, t, q;
};
__device__ __forceinline__ sdb_rs sdb_rsqrt_parts(double s) {
    sdb_rs o;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const int hi = __double2hiint(s);
    o.y0 = __hiloint2double(hi < 0x00100000 ? 0 : __double2hiint(y), 0);   // RSQ64H fills the high word only
    o.t = s * o.y0;
    const double e = fma(-o.t, o.y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    o.q = e * p;
    return o;
}
// The same without the guard, for arguments known to be normal numbers (the pair loops start their sum of squares
// from SDB_R2_FLOOR instead of 0): two ALU instructions fewer per pair in loops whose issue slots are the limit.
#define SDB_R2_FLOOR 1e-300
__device__ __forceinline__ sdb_rs sdb_rsqrt_parts_pos(double s) {
    sdb_rs o;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(o.y0) : "d"(s));
    o.t = s * o.y0;
    const double e = fma(-o.t, o.y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    o.q = e * p;
    return o;
}
__device__ __forceinline__ double sdb_fast_rsqrt(double s) {
    const sdb_rs r = sdb_rsqrt_parts(s);
    return fma(r.y0, r.q, r.y0);
}

__device__ __forceinline__ double sdb_shfl_xor(double v, int lanemask, unsigned mask = 0xffffffffu) {
    return __shfl_xor_sync(mask, v, lanemask);
}
__device__ __forceinline__ double sdb_shfl(double v, int src, unsigned mask = 0xffffffffu) {
    return __shfl_sync(mask, v, src);
}

// phi(r) from s = r^2 for the reference's kernel types (surrogate_rbf.py:146-173), for s >= SDB_R2_FLOOR (never zero or
// subnormal): the pair loops of the RBF surrogate.  A coincident point gives s = 1e-300 instead of 0, i.e. phi = 1e-150 for
// the linear kernel and 0 (underflow) for the odd powers.
__device__ __forceinline__ double sdb_rbf_phi(double s, int kernel_type) {
    switch (kernel_type) {
        case 1: { const sdb_rs r = sdb_rsqrt_parts_pos(s); const double a = s * r.t; return fma(a, r.q, a); }
        case 2: { const sdb_rs r = sdb_rsqrt_parts_pos(s); const double a = s * s * r.t; return fma(a, r.q, a); }
        case 3: { const sdb_rs r = sdb_rsqrt_parts_pos(s); const double a = s * s * s * r.t; return fma(a, r.q, a); }
        case 7: { const sdb_rs r = sdb_rsqrt_parts_pos(s); const double s2 = s * s; const double a = s2 * s2 * r.t; return fma(a, r.q, a); }
        case 4: return exp(-s);
        case 5: return sdb_fast_rsqrt(s + 1.0);
        default: { const sdb_rs r = sdb_rsqrt_parts_pos(s); return fma(r.t, r.q, r.t); }
    }
}
