// Shared device/host helpers for the apertis_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/apertis_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing (the C ABI never throws / aborts: every entry point returns a code, the text is
// fetched with ab_last_error)
// ---------------------------------------------------------------------------------------------
void ab_set_error(const char* fmt, ...);

#define AB_CHECK_CUDA(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ab_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),  \
                         cudaGetErrorString(_e));                                              \
            return AB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define AB_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ab_set_error(__VA_ARGS__);                                                         \
            return AB_ERR_INVALID;                                                             \
        }                                                                                      \
    } while (0)

#define AB_LAUNCH_CHECK() AB_CHECK_CUDA(cudaGetLastError())

static inline int64_t ab_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t ab_round_up(int64_t a, int64_t b) { return ab_ceil_div(a, b) * b; }

int ab_num_sms();   // cached cudaDevAttrMultiProcessorCount of the current device

// TMA tensor-map encode (driver entry point resolved at run time; no link-time libcuda dependency)
int ab_encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* base,
                   const uint64_t* dims, const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box,
                   CUtensorMapSwizzle swz);

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
template <typename T> struct ab_dtype;
template <> struct ab_dtype<float> { static constexpr int id = AB_F32; static constexpr int vec = 4; };
template <> struct ab_dtype<__nv_bfloat16> { static constexpr int id = AB_BF16; static constexpr int vec = 8; };

__device__ __forceinline__ float ab_to_float(float v) { return v; }
__device__ __forceinline__ float ab_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T ab_from_float(float v);
template <> __device__ __forceinline__ float ab_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 ab_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// A 16-byte vector of T (4 x f32 or 8 x bf16) unpacked to / packed from floats.
template <typename T> struct ab_vec16;
template <> struct ab_vec16<float> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void unpack(const uint4& r, float* f) {
        f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
    }
    __device__ __forceinline__ static uint4 pack(const float* f) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
};
template <> struct ab_vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void unpack(const uint4& r, float* f) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static uint4 pack(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&p);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
#define AB_LOG2E 1.4426950408889634f
__device__ __forceinline__ float ab_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ab_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ab_sigmoid(float x) { return ab_rcp(1.0f + ab_ex2(-x * AB_LOG2E)); }
// torch.nn.functional.softplus (beta 1, threshold 20)
__device__ __forceinline__ float ab_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
// same function with two MUFU ops: log1p(e) via its series for small e (where 1 + e would lose the digits), lg2 otherwise;
// relative error < 2e-6 over the whole range
__device__ __forceinline__ float ab_softplus_fast(float x) {
    if (x > 20.0f) return x;
    const float e = ab_ex2(x * AB_LOG2E);
    if (e < 0.03125f) return e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.0f);
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return l * 0.6931471805599453f;
}

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two adjacent channels / columns per instruction
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f2;     // two packed fp32

__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_bcast(float v) { return f2_pack(v, v); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_ex2(f2 v) { float a, b; f2_unpack(v, a, b); return f2_pack(ab_ex2(a), ab_ex2(b)); }
__device__ __forceinline__ f2 f2_rcp(f2 v) { float a, b; f2_unpack(v, a, b); return f2_pack(ab_rcp(a), ab_rcp(b)); }
__device__ __forceinline__ float f2_hsum(f2 v) { float a, b; f2_unpack(v, a, b); return a + b; }

// keep decision of the wrappers' output dropout (core.py:836, 918): counter-based hash of the element index and a 64-bit seed
__device__ __forceinline__ bool ab_out_keep(uint32_t s0, uint32_t s1, uint64_t i, uint32_t thresh) {
    uint32_t x = ((uint32_t)i * 0x9E3779B1u) ^ ((uint32_t)(i >> 32) * 0x85EBCA77u) ^ s0;
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16; x += s1;
    x ^= x >> 15; x *= 0x2C1B3C6Du;
    x ^= x >> 12;
    return x >= thresh;
}

__device__ __forceinline__ float ab_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float ab_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ab_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool ab_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void ab_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ab_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ab_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ab_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ab_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ab_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ab_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ab_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool ab_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(ab_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ab_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!ab_mbar_try_wait(bar, parity)) {}
}
// same, for single-thread roles that run ahead of the rest of the CTA (TMA producer, MMA issuer): sleep between polls so
// the spin does not take issue slots from the warps that share the scheduler
__device__ __forceinline__ void ab_mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!ab_mbar_try_wait(bar, parity)) __nanosleep(40);
}

__device__ __forceinline__ void ab_prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D / 3-D tiled TMA loads, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void ab_tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ab_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ab_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void ab_tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ab_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ab_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void ab_tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ab_smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void ab_tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ab_tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void ab_tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// tcgen05 / TMEM
__device__ __forceinline__ void ab_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ab_smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ab_tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ab_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ab_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ab_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with f32 accumulation
__device__ __forceinline__ void ab_umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void ab_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ab_smem_u32(bar)) : "memory");
}
// ---- thread-block clusters and CTA pairs (tcgen05 cta_group::2): the two CTAs of a pair run one 256-row MMA; only the
//      even CTA (cluster rank 0, the "leader") issues it, both CTAs load operands and drain their half of the accumulator
__device__ __forceinline__ uint32_t ab_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void ab_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void ab_mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(ab_smem_u32(bar)), "r"(cta) : "memory");
}
// TMA load issued by either CTA of a pair into its OWN shared memory; the bytes are counted on the LEADER's mbarrier
// (the pair-rank bit of the barrier address is cleared)
__device__ __forceinline__ void ab_tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ab_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ab_smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void ab_tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ab_smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ab_tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ab_tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[128 rows from each CTA's smem] * B[N/2 columns from each CTA's smem]
__device__ __forceinline__ void ab_umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `cta_mask` once all previously issued pair MMAs have completed
__device__ __forceinline__ void ab_umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ab_smem_u32(bar)), "h"(cta_mask) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void ab_tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ab_tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ab_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// global-memory release/acquire words for inter-CTA protocols
__device__ __forceinline__ unsigned long long ab_ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ab_st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// same store without the compiler barrier: for self-validating words whose order against other accesses is irrelevant
__device__ __forceinline__ void ab_st_relaxed_u64_unordered(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}
