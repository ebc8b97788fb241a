// Selective scan of the Apertis SSM layer (core.py:324-353 scans, :383 softplus, :394-396 skip + gate), forward and
// backward, as ONE persistent kernel each: "rounds" schedule.
//
// The recurrence  h_t = abar_t * h_{t-1} + B_t  (abar = exp(-exp(A_log) * softplus(dt)), one chain per channel) has only
// B * Di independent chains, so the sequence is cut into chunks of Tc tokens.  A warp owns one chunk of one 64-channel
// slab of one sequence at a time: a lane owns two adjacent channels (packed f32x2 arithmetic) and walks the chunk's
// tokens serially, straight from global memory (coalesced 128-byte rows per warp and token, operands of the next half
// group of tokens in flight while the current one is computed), no shared memory, no CTA-level synchronisation.
//
// Every chunk is visited twice:
//   P1  reads only dt and B (forward) / C, z, dout (backward), computes delta = softplus(dt) (32 lanes = 8 tokens x 4 heads,
//       distributed by shuffles, saved for the second visit and for the backward) and the chunk aggregate (P = prod abar,
//       S = state the chunk produces from 0), publishes it;
//   P2  streams the chunk once with the state entering it known: h, y = (C*h + D*x) * silu(z)  (backward: forward
//       recompute of 8 states from the saved checkpoint, reverse sweep, all gradients).
// The persistent grid works in rounds: in round r warp w handles chunk r * cpr + w / nchains of chain w % nchains, and the
// order per warp is P1(0), P1(1), P2(0), P1(2), P2(1), ...: the aggregates of round r are published a whole P1 pass before
// anybody needs their prefix, and the second read of the P1 operands (one round later) comes from the 126 MB L2, so DRAM
// sees every operand byte once.  The prefix over a round's chunk aggregates is a two-level scan done by whoever arrives
// last (segments of 32 chunks, then the segment totals chained to the carry of the previous round): fixed association,
// bitwise reproducible.  All counters / flags are indexed by round and zeroed by a memset node ahead of the launch (graph
// capture safe); waits are bounded (4 s) and trap, so a protocol error is a CUDA error, never a silent wrong answer.
//
// The forward saves the state entering every group of 8 tokens (hck, fp32, +0.5 B per element) so that the backward
// needs no forward prefix of its own and recomputes states 8 at a time in registers.
#include "common.cuh"

namespace {

typedef unsigned long long f2;     // two packed fp32 (a channel pair)

__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_bcast(float v) { return f2_pack(v, v); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_ex2(f2 v) { float a, b; f2_unpack(v, a, b); return f2_pack(ab_ex2(a), ab_ex2(b)); }
__device__ __forceinline__ float f2_hsum(f2 v) { float a, b; f2_unpack(v, a, b); return a + b; }

// sigmoid of a channel pair.  bf16 activations: single-MUFU tanh form (relative error ~5e-4, below bf16 resolution);
// fp32 activations: ex2 + rcp
template <typename T>
__device__ __forceinline__ f2 f2_sigmoid(f2 z) {
    if constexpr (sizeof(T) == 2) {
        float a, b, ta, tb;
        f2_unpack(f2_mul(z, f2_bcast(0.5f)), a, b);
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
        return f2_fma(f2_pack(ta, tb), f2_bcast(0.5f), f2_bcast(0.5f));
    } else {
        float a, b;
        f2_unpack(z, a, b);
        return f2_pack(ab_sigmoid(a), ab_sigmoid(b));
    }
}

// ---- global memory access: channel pairs with an L2 eviction policy ------------------------------------------------
template <typename T> struct Raw;
template <> struct Raw<__nv_bfloat16> { typedef uint32_t type; };
template <> struct Raw<float> { typedef f2 type; };

__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

__device__ __forceinline__ void ldg_raw(uint32_t& r, const void* p, uint64_t pol) {
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ldg_raw(f2& r, const void* p, uint64_t pol) {
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void stg_raw(void* p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_raw(void* p, f2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ f2 up(uint32_t r) { return f2_pack(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u)); }
__device__ __forceinline__ f2 up(f2 r) { return r; }
template <typename T> __device__ __forceinline__ typename Raw<T>::type down(f2 v);
template <> __device__ __forceinline__ uint32_t down<__nv_bfloat16>(f2 v) {
    float a, b;
    f2_unpack(v, a, b);
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
template <> __device__ __forceinline__ f2 down<float>(f2 v) { return v; }

// pair access of row `t` of a [rows, stride] array (base already points at this lane's channel pair).  Loads are
// unconditional: the callers clamp rows past the sequence to its last row and lanes past the width alias channel 0, so
// every address is valid and no load sits behind a divergent branch; stores carry their predicate inside the instruction.
template <typename T>
__device__ __forceinline__ typename Raw<T>::type ld_pair(const T* base, int stride, int t, uint64_t pol) {
    typename Raw<T>::type r;
    ldg_raw(r, base + (int64_t)t * stride, pol);
    return r;
}
__device__ __forceinline__ void stg_pred(void* p, uint32_t v, uint64_t pol, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.L2::cache_hint.b32 [%0], %1, %2;\n\t}" ::"l"(p), "r"(v), "l"(pol), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void stg_pred(void* p, f2 v, uint64_t pol, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.L2::cache_hint.b64 [%0], %1, %2;\n\t}" ::"l"(p), "l"(v), "l"(pol), "r"((int)ok) : "memory");
}
template <typename T>
__device__ __forceinline__ void st_pair(T* base, int stride, int t, bool ok, f2 v, uint64_t pol) {
    stg_pred(base + (int64_t)t * stride, down<T>(v), pol, ok);
}
template <typename R> __device__ __forceinline__ R keep_if(R v, bool ok) { return ok ? v : (R)0; }

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

constexpr int RG = 8;                 // tokens per group: spacing of the saved states, unit of the delta shuffles
constexpr int RSEG = 32;              // chunks per level-1 segment of a round's prefix
constexpr int RWARPS = 8;             // warps per CTA (warps are independent: the CTA is only an occupancy unit)
constexpr unsigned long long WAIT_LIMIT_NS = 4000000000ull;

struct RoundsParams {
    int B, L, Di, H, nslab, nchains;
    int Tc, nck, cpr, nrounds, nseg, nw_used, nck8, L8;
    const void *xa, *dlog, *Bm, *Cm, *z, *dout, *dyssm;
    void *y, *yssm, *dxa, *dBm, *dCm, *dz, *ddlog;
    int xa_stride, dlog_stride, bc_stride, z_stride, y_stride, dbc_stride, dz_stride, dxa_stride, ddlog_stride;   // elements (< 2^30, checked on the host)
    const float *dt_bias, *A_log, *D, *h0;
    float* h_last;
    float* hck;        // [B, nck8, Di]     state entering every group of 8 tokens
    float* delta;      // [B, nslab, L8, 4] softplus(dt) of the slab's four heads
    float4* agg;       // [2][nchains][cpr][32]   chunk aggregate (P0, P1, S0, S1) per lane; level 1 rewrites it as the exclusive prefix inside the segment
    float4* segagg;    // [2][nchains][nseg][32]
    float2* segcarry;  // [2][nchains][nseg][32]  state entering each segment
    float2* carry;     // [nchains][32]           state at the end of the last prefixed round
    unsigned* cnt1;    // [nrounds][nchains][nseg]
    unsigned* cnt2;    // [nrounds][nchains]
    unsigned* flag;    // [nrounds][nchains]
    float4* part;      // backward: [nw_used][32]  (dA0, dA1, dD0, dD1) accumulated by each warp over its chunks
    float* part_b;     // backward: [nw_used][32]  d dt_bias share of lane (head lane >> 3, token lane & 7)
    int ddlog_cols;    // backward: columns [H, ddlog_cols) of the d dlog rows are written as zero (padding of a fused buffer)
};

__device__ __forceinline__ void wait_flag(const unsigned* f, int lane) {
    if (lane == 0) {
        unsigned v;
        unsigned long long t0 = 0;
        for (int it = 0;; ++it) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v != 0u) break;
            if (it >= 32) {
                __nanosleep(100);
                if (t0 == 0) t0 = globaltimer_ns();
                else if (globaltimer_ns() - t0 > WAIT_LIMIT_NS) {
                    __trap();          // protocol error: surfaces as a CUDA error at the next synchronisation, never as a wrong answer
                }
            }
        }
    }
    __syncwarp();
}

// Publishes this warp's chunk aggregate and, when it is the last of its segment / of the round, runs the level-1 /
// level-2 prefix.  `init` = state entering the chain (forward: h0; backward: 0), used by round 0.
__device__ __forceinline__ void publish_and_prefix(const RoundsParams& p, int r, int chain, int slot, int n_r, int lane, f2 P, f2 S, f2 init,
                                                   float* final_out /* this lane's pair of the chain's final state, or null */) {
    const int par = r & 1;
    const int seg = slot / RSEG;
    const int nseg_r = (n_r + RSEG - 1) / RSEG;
    float4* agg = p.agg + ((size_t)(par * p.nchains + chain) * p.cpr) * 32;
    {
        float p0, p1, s0, s1;
        f2_unpack(P, p0, p1);
        f2_unpack(S, s0, s1);
        __stcg(&agg[(size_t)slot * 32 + lane], make_float4(p0, p1, s0, s1));
    }
    __syncwarp();
    unsigned last1 = 0;
    if (lane == 0) {
        __threadfence();
        const int n_in_seg = min(RSEG, n_r - seg * RSEG);
        const unsigned old = atomicAdd(&p.cnt1[((size_t)r * p.nchains + chain) * p.nseg + seg], 1u);
        last1 = (old == (unsigned)(n_in_seg - 1));
        if (last1) __threadfence();
    }
    last1 = __shfl_sync(0xffffffffu, last1, 0);
    if (!last1) return;
    // ---- level 1: exclusive prefixes inside the segment, segment total
    f2 Pex = f2_bcast(1.f), Sex = f2_bcast(0.f);
    {
        const int n_in_seg = min(RSEG, n_r - seg * RSEG);
        float4* a = agg + (size_t)(seg * RSEG) * 32 + lane;
        float4 v = __ldcg(a);
        for (int j = 0; j < n_in_seg; ++j) {
            float4 nv = v;
            if (j + 1 < n_in_seg) nv = __ldcg(a + (size_t)(j + 1) * 32);
            float e0, e1, g0, g1;
            f2_unpack(Pex, e0, e1);
            f2_unpack(Sex, g0, g1);
            __stcg(a + (size_t)j * 32, make_float4(e0, e1, g0, g1));
            const f2 vp = f2_pack(v.x, v.y), vs = f2_pack(v.z, v.w);
            Sex = f2_fma(vp, Sex, vs);
            Pex = f2_mul(Pex, vp);
            v = nv;
        }
    }
    float4* segagg = p.segagg + ((size_t)(par * p.nchains + chain) * p.nseg) * 32;
    {
        float e0, e1, g0, g1;
        f2_unpack(Pex, e0, e1);
        f2_unpack(Sex, g0, g1);
        __stcg(&segagg[(size_t)seg * 32 + lane], make_float4(e0, e1, g0, g1));
    }
    __syncwarp();
    unsigned last2 = 0;
    if (lane == 0) {
        __threadfence();
        const unsigned old = atomicAdd(&p.cnt2[(size_t)r * p.nchains + chain], 1u);
        last2 = (old == (unsigned)(nseg_r - 1));
        if (last2) __threadfence();
    }
    last2 = __shfl_sync(0xffffffffu, last2, 0);
    if (!last2) return;
    // ---- level 2: chain the segment totals to the carry of the previous round
    f2 h = init;
    if (r > 0) {
        wait_flag(&p.flag[(size_t)(r - 1) * p.nchains + chain], lane);
        const float2 c = __ldcg(&p.carry[(size_t)chain * 32 + lane]);
        h = f2_pack(c.x, c.y);
    }
    float2* segcarry = p.segcarry + ((size_t)(par * p.nchains + chain) * p.nseg) * 32;
    for (int s = 0; s < nseg_r; ++s) {
        const float4 v = __ldcg(&segagg[(size_t)s * 32 + lane]);
        float h0, h1;
        f2_unpack(h, h0, h1);
        __stcg(&segcarry[(size_t)s * 32 + lane], make_float2(h0, h1));
        h = f2_fma(f2_pack(v.x, v.y), h, f2_pack(v.z, v.w));
    }
    {
        float h0, h1;
        f2_unpack(h, h0, h1);
        __stcg(&p.carry[(size_t)chain * 32 + lane], make_float2(h0, h1));
        if (final_out != nullptr && r == p.nrounds - 1) { final_out[0] = h0; final_out[1] = h1; }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&p.flag[(size_t)r * p.nchains + chain]), "r"(1u) : "memory");
    }
}

// state entering chunk `slot` of round r (after wait_flag)
__device__ __forceinline__ f2 incoming_state(const RoundsParams& p, int r, int chain, int slot, int lane) {
    const int par = r & 1;
    const float4 ex = __ldcg(&p.agg[((size_t)(par * p.nchains + chain) * p.cpr + slot) * 32 + lane]);
    const float2 sc = __ldcg(&p.segcarry[((size_t)(par * p.nchains + chain) * p.nseg + slot / RSEG) * 32 + lane]);
    return f2_fma(f2_pack(ex.x, ex.y), f2_pack(sc.x, sc.y), f2_pack(ex.z, ex.w));
}

// ====================================================================================================================
// forward
// ====================================================================================================================
template <typename T>
struct FwdCtx {
    const T *xa, *Bm, *Cm, *z, *dlog;     // this lane's channel pair of row 0 of the sequence
    T *y, *yssm;
    float* delta;      // this (sequence, slab)'s [L8, 4] block + head (lane & 3)
    float* hck;        // this sequence's [nck8, Di] block + channel pair
    f2 A2, Dv;
    float bias;        // dt bias of the head this lane computes delta for
    bool cv, hv;       // channel pair valid; head (lane & 3) valid
    int lane;
    int L, sx, sbc, sz, sy;     // sequence length and row strides (elements)
};

// delta of 8 tokens x 4 heads, one (token, head) per lane; saved for P2 and the backward
template <typename T>
__device__ __forceinline__ float delta_compute(const FwdCtx<T>& c, int dlog_stride, int tb) {
    const int tok = tb + (c.lane >> 2);
    float d = 0.f;
    if (c.hv && tok < c.L) d = ab_softplus_fast(ab_to_float(c.dlog[(int64_t)tok * dlog_stride]) + c.bias);
    c.delta[(int64_t)tok * 4] = d;        // tok < L8 always (tb < L, tb a multiple of 8)
    return d;
}

// B rows of one group; rows past the sequence read as zero (they must not change the state)
template <typename T, bool FULL>
__device__ __forceinline__ void p1_load(const FwdCtx<T>& c, typename Raw<T>::type (&bv)[RG], int tb, uint64_t pol) {
#pragma unroll
    for (int j = 0; j < RG; ++j) {
        if (FULL) bv[j] = ld_pair<T>(c.Bm, c.sbc, tb + j, pol);
        else bv[j] = keep_if(ld_pair<T>(c.Bm, c.sbc, min(tb + j, c.L - 1), pol), tb + j < c.L);
    }
}

template <typename T>
__device__ __forceinline__ void fwd_p1(const RoundsParams& p, const FwdCtx<T>& c, int r, int chain, int slot, int b, uint64_t pol_keep, f2 init) {
    typedef typename Raw<T>::type raw_t;
    const int n_r = min(p.cpr, p.nck - r * p.cpr);
    if (slot >= n_r) return;
    const int t0 = (r * p.cpr + slot) * p.Tc;
    const int tend = min(t0 + p.Tc, c.L);
    const int dls = p.dlog_stride;
    f2 S = f2_bcast(0.f);
    float sumd = 0.f;
    raw_t bv[RG];
    const int hsel = c.lane >> 3;
    if (t0 + RG <= c.L) p1_load<T, true>(c, bv, t0, pol_keep);
    else p1_load<T, false>(c, bv, t0, pol_keep);
    for (int tb = t0; tb < tend; tb += RG) {
        const float dmine = delta_compute<T>(c, dls, tb);
        raw_t nb[RG];
        const int tn = tb + RG;
        if (tn < tend) {
            if (tn + RG <= c.L) p1_load<T, true>(c, nb, tn, pol_keep);
            else p1_load<T, false>(c, nb, tn, pol_keep);
        }
#pragma unroll
        for (int j = 0; j < RG; ++j) {
            const float d = __shfl_sync(0xffffffffu, dmine, j * 4 + hsel);
            const f2 a = f2_ex2(f2_mul(f2_bcast(d), c.A2));
            S = f2_fma(a, S, up(bv[j]));
            sumd += d;
        }
#pragma unroll
        for (int j = 0; j < RG; ++j) bv[j] = nb[j];
    }
    const f2 P = f2_ex2(f2_mul(f2_bcast(sumd), c.A2));
    float* fin = (p.h_last != nullptr && c.cv) ? p.h_last + (size_t)b * p.Di + (chain % p.nslab) * 64 + 2 * c.lane : nullptr;
    publish_and_prefix(p, r, chain, slot, n_r, c.lane, P, S, init, fin);
}

template <typename T>
struct FwdBuf {
    typename Raw<T>::type x[4], b[4], c[4], z[4];
};

// operands of 4 tokens.  Rows past the sequence are clamped to its last row: what they produce is never stored and the
// state after the last token of the sequence is not used by P2 (h_last comes from the prefix of the P1 aggregates).
template <typename T, bool FULL>
__device__ __forceinline__ void fwd_load4(const FwdCtx<T>& c, FwdBuf<T>& f, int t, uint64_t pol) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tt = FULL ? t + j : min(t + j, c.L - 1);
        f.b[j] = ld_pair<T>(c.Bm, c.sbc, tt, pol);
        f.x[j] = ld_pair<T>(c.xa, c.sx, tt, pol);
        f.c[j] = ld_pair<T>(c.Cm, c.sbc, tt, pol);
        f.z[j] = ld_pair<T>(c.z, c.sz, tt, pol);
    }
}
template <typename T>
__device__ __forceinline__ void fwd_load4_any(const FwdCtx<T>& c, FwdBuf<T>& f, int t, uint64_t pol) {
    if (t + 4 <= c.L) fwd_load4<T, true>(c, f, t, pol);
    else fwd_load4<T, false>(c, f, t, pol);
}

template <typename T, bool YSSM>
__device__ __forceinline__ void fwd_compute4(const FwdCtx<T>& c, const FwdBuf<T>& f, int t, float dmine, int jbase, f2& h, uint64_t pol) {
    const int hsel = c.lane >> 3;
    const int nvalid = c.cv ? c.L - t : 0;          // tokens t + j with j < nvalid are stored
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float d = __shfl_sync(0xffffffffu, dmine, (jbase + j) * 4 + hsel);
        const f2 a = f2_ex2(f2_mul(f2_bcast(d), c.A2));
        h = f2_fma(a, h, up(f.b[j]));
        const f2 ys = f2_mul(up(f.c[j]), h);
        const bool ok = j < nvalid;
        if (YSSM) st_pair<T>(c.yssm, c.sy, t + j, ok, ys, pol);
        const f2 yv = f2_fma(c.Dv, up(f.x[j]), ys);
        const f2 zv = up(f.z[j]);
        const f2 out = f2_mul(yv, f2_mul(zv, f2_sigmoid<T>(zv)));
        st_pair<T>(c.y, c.sy, t + j, ok, out, pol);
    }
}

template <typename T, bool YSSM>
__device__ __forceinline__ void fwd_p2(const RoundsParams& p, const FwdCtx<T>& c, int r, int chain, int slot, uint64_t pol_stream) {
    const int n_r = min(p.cpr, p.nck - r * p.cpr);
    if (slot >= n_r) return;
    const int t0 = (r * p.cpr + slot) * p.Tc;
    const int tend = min(t0 + p.Tc, c.L);
    FwdBuf<T> fa, fb;
    fwd_load4_any<T>(c, fa, t0, pol_stream);         // in flight during the wait
    fwd_load4_any<T>(c, fb, t0 + 4, pol_stream);
    wait_flag(&p.flag[(size_t)r * p.nchains + chain], c.lane);
    f2 h = incoming_state(p, r, chain, slot, c.lane);
    const int Di = p.Di;
    for (int tb = t0; tb < tend; tb += RG) {
        if (c.cv) {
            float h0, h1;
            f2_unpack(h, h0, h1);
            *reinterpret_cast<float2*>(c.hck + (size_t)(tb >> 3) * Di) = make_float2(h0, h1);
        }
        const float dmine = __ldcg(c.delta + (int64_t)(tb + (c.lane >> 2)) * 4);
        const bool more = tb + RG < tend;
        fwd_compute4<T, YSSM>(c, fa, tb, dmine, 0, h, pol_stream);
        if (more) fwd_load4_any<T>(c, fa, tb + RG, pol_stream);
        fwd_compute4<T, YSSM>(c, fb, tb + 4, dmine, 4, h, pol_stream);
        if (more) fwd_load4_any<T>(c, fb, tb + RG + 4, pol_stream);
    }
}

template <typename T, bool YSSM>
__global__ void __launch_bounds__(RWARPS * 32, sizeof(T) == 2 ? 3 : 2) scan_rounds_fwd_kernel(const RoundsParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (gw >= p.nw_used) return;
    const int chain = gw % p.nchains, slot = gw / p.nchains;
    const int b = chain / p.nslab, slab = chain % p.nslab;
    const int c0 = slab * 64 + 2 * lane;
    FwdCtx<T> c;
    c.lane = lane;
    c.L = p.L; c.sx = p.xa_stride; c.sbc = p.bc_stride; c.sz = p.z_stride; c.sy = p.y_stride;
    c.cv = c0 < p.Di;
    const int head = slab * 4 + (lane & 3);
    c.hv = head < p.H;
    c.bias = (c.hv && p.dt_bias != nullptr) ? p.dt_bias[head] : 0.f;
    const int cs = c.cv ? c0 : 0;          // lanes past the width alias channel 0 (loads stay valid, nothing is stored)
    c.A2 = c.cv ? f2_pack(-__expf(p.A_log[c0]) * AB_LOG2E, -__expf(p.A_log[c0 + 1]) * AB_LOG2E) : f2_bcast(0.f);
    c.Dv = c.cv ? f2_pack(p.D[c0], p.D[c0 + 1]) : f2_bcast(0.f);
    const int64_t row0 = (int64_t)b * p.L;
    c.xa = reinterpret_cast<const T*>(p.xa) + row0 * p.xa_stride + cs;
    c.Bm = reinterpret_cast<const T*>(p.Bm) + row0 * p.bc_stride + cs;
    c.Cm = reinterpret_cast<const T*>(p.Cm) + row0 * p.bc_stride + cs;
    c.z = reinterpret_cast<const T*>(p.z) + row0 * p.z_stride + cs;
    c.y = reinterpret_cast<T*>(p.y) + row0 * p.y_stride + cs;
    c.yssm = YSSM ? reinterpret_cast<T*>(p.yssm) + row0 * p.y_stride + cs : nullptr;
    c.dlog = reinterpret_cast<const T*>(p.dlog) + row0 * p.dlog_stride + (c.hv ? head : 0);
    c.delta = p.delta + (size_t)chain * p.L8 * 4 + (lane & 3);
    c.hck = p.hck + (size_t)b * p.nck8 * p.Di + cs;
    f2 init = f2_bcast(0.f);
    if (p.h0 != nullptr && c.cv) init = f2_pack(p.h0[(size_t)b * p.Di + c0], p.h0[(size_t)b * p.Di + c0 + 1]);
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    for (int r = -1; r < p.nrounds; ++r) {
        if (r + 1 < p.nrounds) fwd_p1<T>(p, c, r + 1, chain, slot, b, pol_keep, init);
        if (r >= 0) fwd_p2<T, YSSM>(p, c, r, chain, slot, pol_stream);
    }
}

// ====================================================================================================================
// backward
//   dyv = dout * silu(z);  dz = dout * (C h + D x) * silu'(z);  dys = dyv (+ dyssm);  dxa = dyv * D;  dD += dyv * x
//   dC = dys * h;  E_t = C_t dys_t + F_{t+1};  dB = E;  F_t = abar_t E_t   (F: what a token hands to its predecessor)
//   d abar = E * h_{t-1};  w = d abar * abar;  d delta[head] += sum_c w * A_c;  dA_log_c += w * delta * A_c
//   d dlog = d delta * sigmoid(dlog) = d delta * (1 - exp(-delta))
// Chunks are visited from the end of the sequence; the reverse chunk aggregate is F_first = Pr * F_in + Sr.
// Rows past the sequence: loads are clamped to the last row and dout (dyssm) read as zero, so they hand F on unchanged.
// ====================================================================================================================
template <typename T>
struct BwdCtx {
    const T *xa, *Bm, *Cm, *z, *dout, *dyssm;
    T *dxa, *dBm, *dCm, *dz, *ddlog;
    const float* delta;
    const float* hck;
    f2 A2, Dv, An;      // An = natural-log A = -exp(A_log)
    bool cv;
    int lane;
    int L, sx, sbc, sz, sy, sdx, sdbc, sdz;
};

template <typename T>
struct BwdP1Buf {
    typename Raw<T>::type c[RG], z[RG], g[RG], s[RG];
};

template <typename T, bool YSSM, bool FULL>
__device__ __forceinline__ void bwd_p1_load(const BwdCtx<T>& c, BwdP1Buf<T>& f, int tb, uint64_t pol) {
#pragma unroll
    for (int j = RG - 1; j >= 0; --j) {
        const int tt = FULL ? tb + j : min(tb + j, c.L - 1);
        f.c[j] = ld_pair<T>(c.Cm, c.sbc, tt, pol);
        f.z[j] = ld_pair<T>(c.z, c.sz, tt, pol);
        f.g[j] = ld_pair<T>(c.dout, c.sy, tt, pol);
        if (YSSM) f.s[j] = ld_pair<T>(c.dyssm, c.sy, tt, pol);
        if (!FULL) {
            f.g[j] = keep_if(f.g[j], tb + j < c.L);
            if (YSSM) f.s[j] = keep_if(f.s[j], tb + j < c.L);
        }
    }
}

template <typename T, bool YSSM>
__device__ __forceinline__ void bwd_p1(const RoundsParams& p, const BwdCtx<T>& c, int r, int chain, int slot, uint64_t pol_keep) {
    const int n_r = min(p.cpr, p.nck - r * p.cpr);
    if (slot >= n_r) return;
    const int pos = p.nck - 1 - (r * p.cpr + slot);
    const int t0 = pos * p.Tc;
    const int hsel = c.lane >> 3;
    f2 F = f2_bcast(0.f);
    float sumd = 0.f;
    // groups from the last one that starts inside the sequence down to the first
    int tb = t0 + p.Tc - RG;
    while (tb >= c.L) tb -= RG;
    BwdP1Buf<T> f;
    if (tb + RG <= c.L) bwd_p1_load<T, YSSM, true>(c, f, tb, pol_keep);
    else bwd_p1_load<T, YSSM, false>(c, f, tb, pol_keep);
    for (; tb >= t0; tb -= RG) {
        const float dmine = __ldg(c.delta + (int64_t)(tb + (c.lane >> 2)) * 4);
        BwdP1Buf<T> nf;
        if (tb - RG >= t0) bwd_p1_load<T, YSSM, true>(c, nf, tb - RG, pol_keep);       // every group below the last one is full
#pragma unroll
        for (int j = RG - 1; j >= 0; --j) {
            const float d = __shfl_sync(0xffffffffu, dmine, j * 4 + hsel);
            const f2 a = f2_ex2(f2_mul(f2_bcast(d), c.A2));
            const f2 zv = up(f.z[j]);
            f2 dys = f2_mul(up(f.g[j]), f2_mul(zv, f2_sigmoid<T>(zv)));
            if (YSSM) dys = f2_add(dys, up(f.s[j]));
            F = f2_mul(a, f2_fma(up(f.c[j]), dys, F));
            sumd += d;
        }
        f = nf;
    }
    const f2 P = f2_ex2(f2_mul(f2_bcast(sumd), c.A2));
    publish_and_prefix(p, r, chain, slot, n_r, c.lane, P, F, f2_bcast(0.f), nullptr);
}

template <typename T>
struct BwdP2Buf {
    typename Raw<T>::type b[RG], x[RG], c[RG], z[RG], g[RG], s[RG];
};

template <typename T, bool YSSM, bool FULL>
__device__ __forceinline__ void bwd_p2_load(const BwdCtx<T>& c, BwdP2Buf<T>& f, int tb, uint64_t pol) {
#pragma unroll
    for (int j = 0; j < RG; ++j) f.b[j] = ld_pair<T>(c.Bm, c.sbc, FULL ? tb + j : min(tb + j, c.L - 1), pol);
#pragma unroll
    for (int j = RG - 1; j >= 0; --j) {
        const int tt = FULL ? tb + j : min(tb + j, c.L - 1);
        f.g[j] = ld_pair<T>(c.dout, c.sy, tt, pol);
        f.z[j] = ld_pair<T>(c.z, c.sz, tt, pol);
        f.c[j] = ld_pair<T>(c.Cm, c.sbc, tt, pol);
        f.x[j] = ld_pair<T>(c.xa, c.sx, tt, pol);
        if (YSSM) f.s[j] = ld_pair<T>(c.dyssm, c.sy, tt, pol);
        if (!FULL) {
            f.g[j] = keep_if(f.g[j], tb + j < c.L);
            if (YSSM) f.s[j] = keep_if(f.s[j], tb + j < c.L);
        }
    }
}

template <typename T, bool YSSM>
__device__ __forceinline__ void bwd_p2(const RoundsParams& p, const BwdCtx<T>& c, int r, int chain, int slot, f2& accA, f2& accD, float& accB, uint64_t pol_stream) {
    const int n_r = min(p.cpr, p.nck - r * p.cpr);
    if (slot >= n_r) return;
    const int pos = p.nck - 1 - (r * p.cpr + slot);
    const int t0 = pos * p.Tc;
    const int hsel = c.lane >> 3;
    const int Di = p.Di;
    int tb = t0 + p.Tc - RG;
    while (tb >= c.L) tb -= RG;
    BwdP2Buf<T> f;
    if (tb + RG <= c.L) bwd_p2_load<T, YSSM, true>(c, f, tb, pol_stream);       // in flight during the wait
    else bwd_p2_load<T, YSSM, false>(c, f, tb, pol_stream);
    wait_flag(&p.flag[(size_t)r * p.nchains + chain], c.lane);
    f2 F = incoming_state(p, r, chain, slot, c.lane);
    for (; tb >= t0; tb -= RG) {
        const float2 hp = __ldg(reinterpret_cast<const float2*>(c.hck + (size_t)(tb >> 3) * Di));
        const float dmine = __ldg(c.delta + (int64_t)(tb + (c.lane >> 2)) * 4);
        // ---- forward recompute of the 8 states
        f2 a[RG], h[RG + 1];
        float dl[RG];
        h[0] = f2_pack(hp.x, hp.y);
#pragma unroll
        for (int j = 0; j < RG; ++j) {
            dl[j] = __shfl_sync(0xffffffffu, dmine, j * 4 + hsel);
            a[j] = f2_ex2(f2_mul(f2_bcast(dl[j]), c.A2));
            h[j + 1] = f2_fma(a[j], h[j], up(f.b[j]));
        }
        // ---- reverse sweep
        const int nvalid = c.cv ? c.L - tb : 0;
        float dd[RG];          // this lane's share (2 channels) of d delta of each token
#pragma unroll
        for (int j = RG - 1; j >= 0; --j) {
            const bool ok = j < nvalid;
            const f2 zv = up(f.z[j]), go = up(f.g[j]), xx = up(f.x[j]), cc = up(f.c[j]);
            const f2 sg = f2_sigmoid<T>(zv);
            const f2 dyv = f2_mul(go, f2_mul(zv, sg));
            const f2 yv = f2_fma(c.Dv, xx, f2_mul(cc, h[j + 1]));
            // silu'(z) = sg * (1 + z * (1 - sg)) = sg * (1 + z - z * sg)
            const f2 dsilu = f2_mul(sg, f2_fma(f2_mul(zv, sg), f2_bcast(-1.f), f2_add(zv, f2_bcast(1.f))));
            st_pair<T>(c.dz, c.sdz, tb + j, ok, f2_mul(f2_mul(go, yv), dsilu), pol_stream);
            st_pair<T>(c.dxa, c.sdx, tb + j, ok, f2_mul(dyv, c.Dv), pol_stream);
            accD = f2_fma(dyv, xx, accD);
            f2 dys = dyv;
            if (YSSM) dys = f2_add(dys, up(f.s[j]));
            st_pair<T>(c.dCm, c.sdbc, tb + j, ok, f2_mul(dys, h[j + 1]), pol_stream);
            const f2 E = f2_fma(cc, dys, F);
            st_pair<T>(c.dBm, c.sdbc, tb + j, ok, E, pol_stream);
            const f2 w = f2_mul(f2_mul(E, h[j]), a[j]);
            dd[j] = f2_hsum(f2_mul(w, c.An));
            accA = f2_fma(w, f2_bcast(dl[j]), accA);
            F = f2_mul(a[j], E);
        }
        // next group's operands: in flight during the reduction below and the next iteration's recompute
        if (tb - RG >= t0) bwd_p2_load<T, YSSM, true>(c, f, tb - RG, pol_stream);
        // ---- d delta: sum over the 8 lanes of a head (16 channels), transposing butterfly over the 8 tokens so that lane
        //      (head hsel, k = lane & 7) ends with the total of token k
        {
            const int k = c.lane & 7;
            float s4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float keep = (k & 4) ? dd[i + 4] : dd[i];
                const float send = (k & 4) ? dd[i] : dd[i + 4];
                s4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            float s2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float keep = (k & 2) ? s4[i + 2] : s4[i];
                const float send = (k & 2) ? s4[i] : s4[i + 2];
                s2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            const float keep = (k & 1) ? s2[1] : s2[0];
            const float send = (k & 1) ? s2[0] : s2[1];
            const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);      // token k of head hsel
            const float dk = __shfl_sync(0xffffffffu, dmine, k * 4 + hsel);       // delta of (token k, head hsel)
            const int head = (chain % p.nslab) * 4 + hsel;
            const int tok = tb + k;
            if (tok < c.L) {
                if (head < p.H) {
                    // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x))
                    const float g = tot * (1.0f - ab_ex2(-dk * AB_LOG2E));
                    accB += g;
                    c.ddlog[(int64_t)tok * p.ddlog_stride + head] = ab_from_float<T>(g);
                }
                for (int hh = head; hh < p.ddlog_cols; hh += 4)       // padding columns (covered by the last slab's lanes)
                    if (hh >= p.H) c.ddlog[(int64_t)tok * p.ddlog_stride + hh] = ab_from_float<T>(0.f);
            }
        }
    }
}

template <typename T, bool YSSM>
__global__ void __launch_bounds__(RWARPS * 32, sizeof(T) == 2 ? 2 : 1) scan_rounds_bwd_kernel(const RoundsParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (gw >= p.nw_used) return;
    const int chain = gw % p.nchains, slot = gw / p.nchains;
    const int b = chain / p.nslab, slab = chain % p.nslab;
    const int c0 = slab * 64 + 2 * lane;
    BwdCtx<T> c;
    c.lane = lane;
    c.L = p.L; c.sx = p.xa_stride; c.sbc = p.bc_stride; c.sz = p.z_stride; c.sy = p.y_stride;
    c.sdx = p.dxa_stride; c.sdbc = p.dbc_stride; c.sdz = p.dz_stride;
    c.cv = c0 < p.Di;
    const int cs = c.cv ? c0 : 0;
    const float a0 = c.cv ? -__expf(p.A_log[c0]) : 0.f, a1 = c.cv ? -__expf(p.A_log[c0 + 1]) : 0.f;
    c.An = f2_pack(a0, a1);
    c.A2 = f2_pack(a0 * AB_LOG2E, a1 * AB_LOG2E);
    c.Dv = c.cv ? f2_pack(p.D[c0], p.D[c0 + 1]) : f2_bcast(0.f);
    const int64_t row0 = (int64_t)b * p.L;
    c.xa = reinterpret_cast<const T*>(p.xa) + row0 * p.xa_stride + cs;
    c.Bm = reinterpret_cast<const T*>(p.Bm) + row0 * p.bc_stride + cs;
    c.Cm = reinterpret_cast<const T*>(p.Cm) + row0 * p.bc_stride + cs;
    c.z = reinterpret_cast<const T*>(p.z) + row0 * p.z_stride + cs;
    c.dout = reinterpret_cast<const T*>(p.dout) + row0 * p.y_stride + cs;
    c.dyssm = YSSM ? reinterpret_cast<const T*>(p.dyssm) + row0 * p.y_stride + cs : nullptr;
    c.dxa = reinterpret_cast<T*>(p.dxa) + row0 * p.dxa_stride + cs;
    c.dBm = reinterpret_cast<T*>(p.dBm) + row0 * p.dbc_stride + cs;
    c.dCm = reinterpret_cast<T*>(p.dCm) + row0 * p.dbc_stride + cs;
    c.dz = reinterpret_cast<T*>(p.dz) + row0 * p.dz_stride + cs;
    c.ddlog = reinterpret_cast<T*>(p.ddlog) + row0 * p.ddlog_stride;
    c.delta = p.delta + (size_t)chain * p.L8 * 4 + (lane & 3);
    c.hck = p.hck + (size_t)b * p.nck8 * p.Di + cs;
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    f2 accA = f2_bcast(0.f), accD = f2_bcast(0.f);
    float accB = 0.f;
    for (int r = -1; r < p.nrounds; ++r) {
        if (r + 1 < p.nrounds) bwd_p1<T, YSSM>(p, c, r + 1, chain, slot, pol_keep);
        if (r >= 0) bwd_p2<T, YSSM>(p, c, r, chain, slot, accA, accD, accB, pol_stream);
    }
    {
        // dA_log = A * sum(w * delta): accA holds sum(w * delta)
        const f2 da = f2_mul(accA, c.An);
        float x0, x1, y0, y1;
        f2_unpack(da, x0, x1);
        f2_unpack(accD, y0, y1);
        p.part[(size_t)gw * 32 + lane] = make_float4(x0, x1, y0, y1);
        p.part_b[(size_t)gw * 32 + lane] = accB;
    }
}

// dA_log[c], dD[c] (and d dt_bias[head]) = sum over the warps that worked on the channel's slab (all sequences, all chunk
// slots), fixed order: bitwise reproducible
__global__ void scan_rounds_param_reduce_kernel(const float4* __restrict__ part, const float* __restrict__ part_b, float* __restrict__ dA,
                                                float* __restrict__ dD, float* __restrict__ dbias, int Di, int H, int nslab, int nw_used) {
    __shared__ float4 red[8][32];
    __shared__ float redb[8][32];
    const int slab = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float accb = 0.f;
    // warp gw works on slab (gw % nchains) % nslab = gw % nslab (nchains is a multiple of nslab)
    int idx = 0;
    for (int gw = slab; gw < nw_used; gw += nslab, ++idx) {
        if ((idx & 7) != w) continue;
        const float4 v = part[(size_t)gw * 32 + lane];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        accb += part_b[(size_t)gw * 32 + lane];
    }
    red[w][lane] = acc;
    redb[w][lane] = accb;
    __syncthreads();
    if (w == 0) {
        float4 s = red[0][lane];
        float sb = redb[0][lane];
        for (int i = 1; i < 8; ++i) { s.x += red[i][lane].x; s.y += red[i][lane].y; s.z += red[i][lane].z; s.w += red[i][lane].w; sb += redb[i][lane]; }
        const int c0 = slab * 64 + 2 * lane;
        if (c0 < Di) { dA[c0] = s.x; dA[c0 + 1] = s.y; dD[c0] = s.z; dD[c0 + 1] = s.w; }
        // lanes (head lane >> 3, token lane & 7): sum the 8 token lanes of a head
        sb += __shfl_xor_sync(0xffffffffu, sb, 4);
        sb += __shfl_xor_sync(0xffffffffu, sb, 2);
        sb += __shfl_xor_sync(0xffffffffu, sb, 1);
        const int head = slab * 4 + (lane >> 3);
        if (dbias != nullptr && (lane & 7) == 0 && head < H) dbias[head] = sb;
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------
struct RoundsCfg {
    int nslab, nchains, nw, cpr, nw_used, Tc, nck, nrounds, nseg, nck8, L8;
    size_t off_agg, off_segagg, off_segcarry, off_carry, off_cnt, cnt_bytes, off_part, off_part_b, total;
};

template <typename K>
int occupancy_of(K kernel, int* out) {
    int n = 0;
    AB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, RWARPS * 32, 0));
    *out = n < 1 ? 1 : n;
    return AB_OK;
}

int ctas_per_sm(int dtype, bool bwd, bool yssm, int* out) {
    static int cache[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
    int& slot = cache[dtype == AB_F32 ? 0 : 1][bwd ? 1 : 0][yssm ? 1 : 0];
    if (slot == 0) {
        int n = 0, e;
        if (dtype == AB_F32) {
            if (!bwd) e = yssm ? occupancy_of(scan_rounds_fwd_kernel<float, true>, &n) : occupancy_of(scan_rounds_fwd_kernel<float, false>, &n);
            else e = yssm ? occupancy_of(scan_rounds_bwd_kernel<float, true>, &n) : occupancy_of(scan_rounds_bwd_kernel<float, false>, &n);
        } else {
            if (!bwd) e = yssm ? occupancy_of(scan_rounds_fwd_kernel<__nv_bfloat16, true>, &n) : occupancy_of(scan_rounds_fwd_kernel<__nv_bfloat16, false>, &n);
            else e = yssm ? occupancy_of(scan_rounds_bwd_kernel<__nv_bfloat16, true>, &n) : occupancy_of(scan_rounds_bwd_kernel<__nv_bfloat16, false>, &n);
        }
        if (e) return e;
        slot = n;
    }
    *out = slot;
    return AB_OK;
}

// Chunk length: few, well-filled rounds (the grid is persistent: a partly filled last round idles warps), at least two
// rounds where the sequence allows it (the prefix of round r hides behind P1 of round r + 1), and a round's P1 operands
// small against L2.
int choose_tc(int L, int cpr, int tc_hint) {
    if (tc_hint >= RG) return (tc_hint / RG) * RG;
    int best = RG;
    double best_score = -1.0;
    for (int tc = RG; tc <= 64; tc += RG) {
        const int nck = (int)ab_ceil_div(L, tc);
        const int rounds = (int)ab_ceil_div(nck, cpr);
        double score = (double)L / ((double)rounds * cpr * tc);            // fill
        score *= (double)tc / (tc + 3.0);                                  // per-chunk overhead ~ 3 tokens of work
        if (rounds == 1 && nck > 1) score *= 0.9;                          // exposed prefix
        if (score > best_score) { best_score = score; best = tc; }
    }
    return best;
}

int make_cfg(int B, int L, int Di, int warps_per_sm_cap, int tc_hint, RoundsCfg& c) {
    c.nslab = (int)ab_ceil_div(Di, 64);
    c.nchains = B * c.nslab;
    c.nw = ab_num_sms() * warps_per_sm_cap;
    AB_REQUIRE(c.nchains <= c.nw, "selective scan: %d chains exceed the %d resident warps; split the batch", c.nchains, c.nw);
    c.cpr = c.nw / c.nchains;
    c.nck8 = (int)ab_ceil_div(L, RG);
    c.L8 = c.nck8 * RG;
    if (c.cpr > c.nck8) c.cpr = c.nck8;
    c.Tc = choose_tc(L, c.cpr, tc_hint);
    c.nck = (int)ab_ceil_div(L, c.Tc);
    if (c.cpr > c.nck) c.cpr = c.nck;
    c.nw_used = c.cpr * c.nchains;
    c.nrounds = (int)ab_ceil_div(c.nck, c.cpr);
    c.nseg = (int)ab_ceil_div(c.cpr, RSEG);
    size_t o = 0;
    c.off_agg = o;      o += (size_t)2 * c.nchains * c.cpr * 32 * sizeof(float4);
    c.off_segagg = o;   o += (size_t)2 * c.nchains * c.nseg * 32 * sizeof(float4);
    c.off_segcarry = o; o += (size_t)2 * c.nchains * c.nseg * 32 * sizeof(float2);
    c.off_carry = o;    o += (size_t)c.nchains * 32 * sizeof(float2);
    c.off_part = o;     o += (size_t)c.nw_used * 32 * sizeof(float4);
    c.off_part_b = o;   o += (size_t)c.nw_used * 32 * sizeof(float);
    c.off_cnt = o;
    c.cnt_bytes = (size_t)c.nrounds * c.nchains * (c.nseg + 2) * sizeof(unsigned);
    o += ab_round_up((int64_t)c.cnt_bytes, 256);
    c.total = o;
    return AB_OK;
}

void fill_sync(RoundsParams& p, const RoundsCfg& c, void* ws) {
    unsigned char* w = reinterpret_cast<unsigned char*>(ws);
    p.agg = reinterpret_cast<float4*>(w + c.off_agg);
    p.segagg = reinterpret_cast<float4*>(w + c.off_segagg);
    p.segcarry = reinterpret_cast<float2*>(w + c.off_segcarry);
    p.carry = reinterpret_cast<float2*>(w + c.off_carry);
    p.part = reinterpret_cast<float4*>(w + c.off_part);
    p.part_b = reinterpret_cast<float*>(w + c.off_part_b);
    p.cnt1 = reinterpret_cast<unsigned*>(w + c.off_cnt);
    p.cnt2 = p.cnt1 + (size_t)c.nrounds * c.nchains * c.nseg;
    p.flag = p.cnt2 + (size_t)c.nrounds * c.nchains;
    p.nslab = c.nslab; p.nchains = c.nchains; p.Tc = c.Tc; p.nck = c.nck; p.cpr = c.cpr; p.nrounds = c.nrounds; p.nseg = c.nseg;
    p.nw_used = c.nw_used; p.nck8 = c.nck8; p.L8 = c.L8;
}

int g_tc_hint_fwd = 0, g_tc_hint_bwd = 0, g_warp_cap = 0;

int common_checks(const char* who, int B, int L, int Di, int H, int dtype) {
    AB_REQUIRE(B > 0 && L > 0 && Di > 0 && H > 0, "%s: B, L, Di, H must be positive", who);
    AB_REQUIRE(Di == 16 * H, "%s: Di (%d) must equal 16 * H (%d): the kernel is specialised for ssm_d_state 16", who, Di, H);
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "%s: bad dtype %d", who, dtype);
    return AB_OK;
}

int resident_warps(int dtype, bool bwd, bool yssm, int* wps) {
    int ctas = 0;
    if (int e = ctas_per_sm(dtype, bwd, yssm, &ctas)) return e;
    int w = ctas * RWARPS;
    if (g_warp_cap > 0 && w > g_warp_cap) w = (g_warp_cap / RWARPS) * RWARPS;
    if (w < RWARPS) w = RWARPS;
    *wps = w;
    return AB_OK;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------------------------
extern "C" int ab_ssm_scan_tune(int tc_fwd, int tc_bwd, int warps_per_sm) {
    g_tc_hint_fwd = tc_fwd; g_tc_hint_bwd = tc_bwd; g_warp_cap = warps_per_sm;
    return AB_OK;
}

extern "C" int ab_ssm_scan_plan(int B, int L, int Di, int dtype, int64_t* state_floats, size_t* ws_bytes) {
    if (int e = common_checks("ssm_scan_plan", B, L, Di, Di / 16, dtype)) return e;
    size_t total = 0;
    for (int bwd = 0; bwd < 2; ++bwd)
        for (int ys = 0; ys < 2; ++ys) {
            int wps = 0;
            if (int e = resident_warps(dtype, bwd != 0, ys != 0, &wps)) return e;
            RoundsCfg c;
            if (int e = make_cfg(B, L, Di, wps, bwd ? g_tc_hint_bwd : g_tc_hint_fwd, c)) return e;
            if (c.total > total) total = c.total;
        }
    const int64_t nck8 = ab_ceil_div(L, RG), nslab = ab_ceil_div(Di, 64);
    if (state_floats) *state_floats = (int64_t)B * nck8 * Di + (int64_t)B * nslab * nck8 * RG * 4;
    if (ws_bytes) *ws_bytes = total;
    return AB_OK;
}

extern "C" int ab_ssm_scan_fwd(const void* xa, int64_t xa_stride, const void* dlog, int64_t dlog_stride, const float* dt_bias,
                               const void* Bm, const void* Cm, int64_t bc_stride, const void* z, int64_t z_stride,
                               const float* A_log, const float* D, const float* h0, void* y, void* y_ssm, float* h_last,
                               float* state, void* ws, size_t ws_bytes, int B, int L, int Di, int H, int dtype,
                               cudaStream_t stream) {
    if (int e = common_checks("ssm_scan_fwd", B, L, Di, H, dtype)) return e;
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE(xa && dlog && Bm && Cm && z && A_log && D && y && state && ws, "ssm_scan_fwd: null argument");
    AB_REQUIRE(xa_stride >= Di && bc_stride >= Di && z_stride >= Di && dlog_stride >= H, "ssm_scan_fwd: a row stride is smaller than its row");
    AB_REQUIRE(((xa_stride | bc_stride | z_stride) % 2) == 0, "ssm_scan_fwd: row strides must be even (channel pairs are loaded as one word)");
    AB_REQUIRE((xa_stride | bc_stride | z_stride | dlog_stride) < (1 << 30), "ssm_scan_fwd: row strides must be below 2^30 elements");
    AB_REQUIRE(((uintptr_t)xa | (uintptr_t)Bm | (uintptr_t)Cm | (uintptr_t)z | (uintptr_t)y | (uintptr_t)y_ssm) % (2 * es) == 0,
               "ssm_scan_fwd: activations must be aligned to a channel pair");
    int wps = 0;
    if (int e = resident_warps(dtype, false, y_ssm != nullptr, &wps)) return e;
    RoundsCfg c;
    if (int e = make_cfg(B, L, Di, wps, g_tc_hint_fwd, c)) return e;
    AB_REQUIRE(ws_bytes >= c.total, "ssm_scan_fwd: workspace too small (%zu < %zu)", ws_bytes, c.total);
    RoundsParams p;
    memset(&p, 0, sizeof(p));
    fill_sync(p, c, ws);
    p.B = B; p.L = L; p.Di = Di; p.H = H;
    p.xa = xa; p.dlog = dlog; p.Bm = Bm; p.Cm = Cm; p.z = z; p.y = y; p.yssm = y_ssm;
    p.xa_stride = (int)xa_stride; p.dlog_stride = (int)dlog_stride; p.bc_stride = (int)bc_stride; p.z_stride = (int)z_stride; p.y_stride = Di;
    p.dt_bias = dt_bias; p.A_log = A_log; p.D = D; p.h0 = h0; p.h_last = h_last;
    p.hck = state;
    p.delta = state + (size_t)B * c.nck8 * Di;
    AB_CHECK_CUDA(cudaMemsetAsync(p.cnt1, 0, c.cnt_bytes, stream));
    const unsigned grid = (unsigned)ab_ceil_div(c.nw_used, RWARPS);
    if (dtype == AB_F32) {
        if (y_ssm) scan_rounds_fwd_kernel<float, true><<<grid, RWARPS * 32, 0, stream>>>(p);
        else scan_rounds_fwd_kernel<float, false><<<grid, RWARPS * 32, 0, stream>>>(p);
    } else {
        if (y_ssm) scan_rounds_fwd_kernel<__nv_bfloat16, true><<<grid, RWARPS * 32, 0, stream>>>(p);
        else scan_rounds_fwd_kernel<__nv_bfloat16, false><<<grid, RWARPS * 32, 0, stream>>>(p);
    }
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_ssm_scan_bwd(const void* xa, int64_t xa_stride, const void* Bm, const void* Cm, int64_t bc_stride,
                               const void* z, int64_t z_stride, const void* dout, const void* dyssm, const float* A_log,
                               const float* D, const float* state, void* dxa, int64_t dxa_stride, void* dBm, void* dCm,
                               int64_t dbc_stride, void* dz, int64_t dz_stride, void* ddlog, int64_t ddlog_stride, int ddlog_cols,
                               float* ddt_bias, float* dA_log, float* dD, void* ws, size_t ws_bytes, int B, int L, int Di, int H, int dtype,
                               cudaStream_t stream) {
    if (int e = common_checks("ssm_scan_bwd", B, L, Di, H, dtype)) return e;
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE(xa && Bm && Cm && z && dout && A_log && D && state && dxa && dBm && dCm && dz && ddlog && dA_log && dD && ws,
               "ssm_scan_bwd: null argument");
    AB_REQUIRE(xa_stride >= Di && bc_stride >= Di && z_stride >= Di && dxa_stride >= Di && dbc_stride >= Di && dz_stride >= Di && ddlog_stride >= H && ddlog_cols <= ddlog_stride,
               "ssm_scan_bwd: a row stride is smaller than its row");
    AB_REQUIRE(((xa_stride | bc_stride | z_stride | dxa_stride | dbc_stride | dz_stride) % 2) == 0, "ssm_scan_bwd: row strides must be even");
    AB_REQUIRE((xa_stride | bc_stride | z_stride | dxa_stride | dbc_stride | dz_stride | ddlog_stride) < (1 << 30), "ssm_scan_bwd: row strides must be below 2^30 elements");
    AB_REQUIRE(((uintptr_t)xa | (uintptr_t)Bm | (uintptr_t)Cm | (uintptr_t)z | (uintptr_t)dout | (uintptr_t)dyssm | (uintptr_t)dxa | (uintptr_t)dBm |
                (uintptr_t)dCm | (uintptr_t)dz) % (2 * es) == 0, "ssm_scan_bwd: activations must be aligned to a channel pair");
    int wps = 0;
    if (int e = resident_warps(dtype, true, dyssm != nullptr, &wps)) return e;
    RoundsCfg c;
    if (int e = make_cfg(B, L, Di, wps, g_tc_hint_bwd, c)) return e;
    AB_REQUIRE(ws_bytes >= c.total, "ssm_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, c.total);
    RoundsParams p;
    memset(&p, 0, sizeof(p));
    fill_sync(p, c, ws);
    p.B = B; p.L = L; p.Di = Di; p.H = H;
    p.xa = xa; p.Bm = Bm; p.Cm = Cm; p.z = z; p.dout = dout; p.dyssm = dyssm;
    p.dxa = dxa; p.dBm = dBm; p.dCm = dCm; p.dz = dz; p.ddlog = ddlog;
    p.xa_stride = (int)xa_stride; p.bc_stride = (int)bc_stride; p.z_stride = (int)z_stride; p.y_stride = Di;
    p.dxa_stride = (int)dxa_stride; p.dbc_stride = (int)dbc_stride; p.dz_stride = (int)dz_stride; p.ddlog_stride = (int)ddlog_stride;
    p.A_log = A_log; p.D = D;
    p.ddlog_cols = ddlog_cols;
    p.hck = const_cast<float*>(state);
    p.delta = const_cast<float*>(state) + (size_t)B * c.nck8 * Di;
    AB_CHECK_CUDA(cudaMemsetAsync(p.cnt1, 0, c.cnt_bytes, stream));
    const unsigned grid = (unsigned)ab_ceil_div(c.nw_used, RWARPS);
    if (dtype == AB_F32) {
        if (dyssm) scan_rounds_bwd_kernel<float, true><<<grid, RWARPS * 32, 0, stream>>>(p);
        else scan_rounds_bwd_kernel<float, false><<<grid, RWARPS * 32, 0, stream>>>(p);
    } else {
        if (dyssm) scan_rounds_bwd_kernel<__nv_bfloat16, true><<<grid, RWARPS * 32, 0, stream>>>(p);
        else scan_rounds_bwd_kernel<__nv_bfloat16, false><<<grid, RWARPS * 32, 0, stream>>>(p);
    }
    AB_LAUNCH_CHECK();
    scan_rounds_param_reduce_kernel<<<c.nslab, 256, 0, stream>>>(p.part, p.part_b, dA_log, dD, ddt_bias, Di, H, c.nslab, c.nw_used);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
