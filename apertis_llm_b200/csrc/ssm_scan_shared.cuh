// Pieces shared by the selective-scan translation units (ssm_scan.cu, ssm_scan_pipe.cu): launch parameters, the
// self-validating hand-shake words, the scanner role and the small vector load/store helpers.
#pragma once
#include "common.cuh"

namespace ab_scan {

constexpr int TS = 4;            // tokens per thread run
constexpr int SCAN_K = 32;       // tile aggregates the scanner polls per round trip (shared-memory ring, cp.async)
constexpr int MODE_FUSED = 0, MODE_AGG = 1, MODE_APPLY = 2;
constexpr uint32_t ST_AGG = 1, ST_INCL = 2;
constexpr int SPIN_LIMIT = 1 << 20;   // polls (about a microsecond each) before a wait gives up: error flag + trap, i.e. a
                                      // protocol error is a CUDA error at the next synchronisation, never a wrong answer

struct ScanParams {
    int B, L, Di, H;
    int Cs, T, n_s, nslab, nchunks, nchains;
    const void* dlog;        // [B, L, H] activation dtype
    const float* A_log;      // [Di]
    const float* Dp;         // [Di]
    const float* h0;         // [B, Di] or null
    void* y; void* y_ssm;    // [B, L, Di]
    float* h_last;           // [B, Di] or null
    float* hstart;           // [B, nchunks, Di] or null
    unsigned long long* words;   // [nchains*nchunks][Cs][2]  tile aggregates (P, S)
    unsigned long long* inclw;   // [nchains*nchunks][Cs]     state entering each tile, published by the scanners
    int n_scan;                  // scanner CTAs (take the first tickets)
    unsigned int* ticket;
    unsigned int* err_flag;
    float* aggP; float* aggS;    // two-pass: [nchains*nchunks][Cs]
    uint32_t epoch;
    // backward only
    const void* dyssm;       // optional grad of y_ssm, [B, L, Di]
    void* dxa; void* dBm; void* dCm; void* dz; int64_t dbc_stride;
    float* ddlog_parts;      // [B, L, Di / V_b]
    float* part;             // [ntiles][2][Cs] partial dA_log / dD sums
};

__device__ __forceinline__ unsigned long long pack_word(uint32_t epoch, uint32_t st, float v) {
    return ((unsigned long long)((epoch << 2) | st) << 32) | (unsigned long long)__float_as_uint(v);
}

// deepest power-of-two ring (<= SCAN_K tiles x Cs channels x 16 B) that fits in the tile area of a scanner CTA
__device__ __forceinline__ int ring_depth(size_t avail_bytes, int Cs, int kmax) {
    int k = kmax;
    while (k >= 4 && (size_t)k * Cs * sizeof(uint4) > avail_bytes) k >>= 1;
    return k >= 4 ? k : 0;
}

__device__ __forceinline__ bool word_valid(unsigned long long w, uint32_t epoch) { return (uint32_t)(w >> 34) == epoch; }

#ifdef AB_SCAN_TRACE
// debug build only (make TRACE=1): per-tile phase timestamps of the forward kernel, read back by tools/scan_trace.py
constexpr int TRACE_SLOTS = 8, TRACE_TILES = 16384;
__device__ unsigned long long g_scan_trace[TRACE_TILES * TRACE_SLOTS];
__device__ __forceinline__ void trace_mark(size_t tile_lin, int slot) {
    if (threadIdx.x == 0 && tile_lin < TRACE_TILES) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_scan_trace[tile_lin * TRACE_SLOTS + slot] = t;
    }
}
#define TRACE_MARK(tl, s) trace_mark(tl, s)
constexpr int STRACE_ROUNDS = 2048;
__device__ unsigned long long g_scanner_trace[64 * STRACE_ROUNDS * 4];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define TRACE_MARK(tl, s)
#endif

// Tile side: spin until the scanner has published the state entering this tile.
__device__ __forceinline__ float wait_incoming(const ScanParams& p, uint32_t epoch, size_t tile_lin, int c) {
    const unsigned long long* w = p.inclw + tile_lin * p.Cs + c;
    unsigned long long v = ab_ld_relaxed_u64(w);
    int spins = 0;
    while (!word_valid(v, epoch)) {
        if (++spins > SPIN_LIMIT) { atomicExch(p.err_flag, 1u); __trap(); }
        v = ab_ld_relaxed_u64(w);
    }
    return __uint_as_float((uint32_t)v);
}

// Scanner role.  One CTA per chain (b, slab); DIR = +1 walks chunk 0 -> last (forward scan, starts from h0),
// DIR = -1 walks last -> 0 (reverse scan of the backward, starts from 0).  It publishes, for every tile, the state
// entering it (which does not depend on the tile's own aggregate).
// Fast path: the CTA's threads form R replicas of the slab's channels.  Every round polls the aggregate words of the
// next K tiles with cp.async into a shared-memory ring (the tile buffers are free in a scanner CTA), K / R consecutive
// tiles per replica.  Each replica composes the valid prefix of its segment, the segment aggregates are chained
// through shared memory (every thread derives the same new head and state), and each replica then walks its own
// segment again to publish the per-tile states: loads, FMA chains and stores of a round are spread over R warps sets.
// The composition order is fixed by the tile order, so results are bitwise reproducible.
template <int DIR>
__device__ __forceinline__ int scanner_role(const ScanParams& p, uint32_t epoch, int kmax, int rmax, int chain, float* hs /*smem [Cs]*/, uint4* ring /*smem below hs*/, size_t ring_bytes) {
    const int n = p.nchunks, Cs = p.Cs;
    const int slab = chain % p.nslab, b = chain / p.nslab;
    int spins = 0;
    int R = (int)blockDim.x / Cs;
    R = R >= rmax ? rmax : (R >= 4 ? 4 : (R >= 2 ? 2 : R));
    // shared memory: ring [K][Cs] uint4, then segP / segS / segN [8][Cs]
    const size_t seg_bytes = (size_t)3 * 8 * Cs * sizeof(float);
    const int K = ring_bytes > seg_bytes ? ring_depth(ring_bytes - seg_bytes, Cs, kmax) : 0;
    if (R >= 1 && K >= 8) {
        const int S = K / R;                           // tiles per replica and round (>= 2)
        float* segP = reinterpret_cast<float*>(ring + (size_t)K * Cs);
        float* segS = segP + 8 * Cs;
        int* segN = reinterpret_cast<int*>(segS + 8 * Cs);
        const int c = threadIdx.x % Cs, r = threadIdx.x / Cs;
        if (r >= R) return 0;                          // surplus warps leave: they would only take issue slots and barrier time
        const bool active = true;
        const int cg = slab * Cs + c;
        float h = (DIR > 0 && p.h0) ? p.h0[(size_t)b * p.Di + cg] : 0.f;
        uint4* myring = ring + c;
        unsigned long long* incl_base = p.inclw + (size_t)chain * n * Cs + c;
        const unsigned long long* word_base = p.words + ((size_t)chain * n * Cs + c) * 2;
        const unsigned long long tag_incl = (unsigned long long)((epoch << 2) | ST_INCL) << 32;
        if (r == 0) ab_st_relaxed_u64_unordered(incl_base + (size_t)(DIR > 0 ? 0 : n - 1) * Cs, tag_incl | __float_as_uint(h));
        // Fixed two-level association (independent of timing, hence bitwise reproducible): the chain is cut into
        // aligned segments of S tiles; H(g+1) = Pseg(g) * H(g) + Sseg(g) with the segment aggregate composed in tile
        // order, and the states inside a segment follow sequentially from H(g).
        int gh = 0;                                    // first segment that is not folded into h yet
        const int nseg = (n + S - 1) / S;
#ifdef AB_SCAN_TRACE
        int round = 0;
#endif
        while (gh < nseg) {
#ifdef AB_SCAN_TRACE
            const unsigned long long tr0 = gtime();
#endif
            const int lo = min((gh + r) * S, n), hi = active ? min(lo + S, n) : lo;     // my segment of this round
            for (int s2 = lo; s2 < hi; ++s2) {
                const int j = DIR > 0 ? s2 : n - 1 - s2;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ab_smem_u32(myring + (size_t)(s2 & (K - 1)) * Cs)),
                             "l"(word_base + (size_t)j * Cs * 2) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef AB_SCAN_TRACE
            const unsigned long long tr1 = gtime();
#endif
            // valid prefix of my segment and its aggregate
            float Pa = 1.f, Sa = 0.f;
            int cnt = 0;
            {
                bool ok = true;
                for (int u0 = lo; u0 < hi && ok; u0 += 8) {
                    uint4 w4[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) w4[u] = myring[(size_t)((u0 + u) & (K - 1)) * Cs];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        ok = ok && u0 + u < hi && (w4[u].y >> 2) == epoch && (w4[u].w >> 2) == epoch;
                        if (ok) {
                            Sa = fmaf(__uint_as_float(w4[u].x), Sa, __uint_as_float(w4[u].z));
                            Pa *= __uint_as_float(w4[u].x);
                            ++cnt;
                        }
                    }
                }
            }
            const bool full = cnt == hi - lo;          // every tile of the segment has published
            if (active) { segP[r * Cs + c] = Pa; segS[r * Cs + c] = Sa; segN[r * Cs + c] = full ? 1 : 0; }
            __syncthreads();
            // chain the complete segments: every replica of a channel derives the same new gh / state
            float hin = h, hnew = h;
            bool reach = true, mine = false;
            int adv = 0;
            for (int r2 = 0; r2 < R; ++r2) {
                if (r2 == r) { hin = hnew; mine = reach; }
                if (reach && gh + r2 < nseg && segN[r2 * Cs + c]) {
                    hnew = fmaf(segP[r2 * Cs + c], hnew, segS[r2 * Cs + c]);
                    ++adv;
                } else {
                    reach = false;
                }
            }
            // publish the states entering the tiles behind my valid prefix (idempotent when a segment is polled again)
            if (mine) {
                float hh = hin;
                for (int u0 = 0; u0 < cnt; u0 += 8) {
                    uint4 w4[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) w4[u] = myring[(size_t)((lo + u0 + u) & (K - 1)) * Cs];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int jn = lo + u0 + u + 1;             // state entering tile jn
                        if (u0 + u < cnt) {
                            hh = fmaf(__uint_as_float(w4[u].x), hh, __uint_as_float(w4[u].z));
                            // the first tile of the next segment gets H(g+1) exactly as the chain carries it on
                            const float hv = (full && jn == hi) ? fmaf(Pa, hin, Sa) : hh;
                            if (jn < n) ab_st_relaxed_u64_unordered(incl_base + (size_t)(DIR > 0 ? jn : n - 1 - jn) * Cs, tag_incl | __float_as_uint(hv));
                        }
                    }
                }
            }
            gh += adv;
            h = hnew;
#ifdef AB_SCAN_TRACE
            if (threadIdx.x == 0 && chain < 64 && round < STRACE_ROUNDS) {
                unsigned long long* o = g_scanner_trace + ((size_t)chain * STRACE_ROUNDS + round) * 4;
                o[0] = tr0; o[1] = tr1; o[2] = gtime(); o[3] = (unsigned long long)(adv * S);
            }
            ++round;
#endif
            if (adv == 0) {
                if (++spins > SPIN_LIMIT) { atomicExch(p.err_flag, 1u); __trap(); }
                __nanosleep(100);   // nothing new yet: leave the issue slots to the CTA that shares this SM
            }
            __syncthreads();        // ring slots and segment words are rewritten next round
        }
        if (DIR > 0 && p.h_last && r == 0) p.h_last[(size_t)b * p.Di + cg] = h;
        return R * Cs;
    }
    const int c = threadIdx.x;
    // generic path (blocks narrower than the slab: tiny sequences): state in shared memory, no prefetch
    for (int cc = c; cc < Cs; cc += blockDim.x) hs[cc] = (DIR > 0 && p.h0) ? p.h0[(size_t)b * p.Di + slab * Cs + cc] : 0.f;
    for (int step = 0; step < n; ++step) {
        const int j = DIR > 0 ? step : n - 1 - step;
        const size_t tl = (size_t)chain * n + j;
        for (int cc = c; cc < Cs; cc += blockDim.x) {
            const float h = hs[cc];
            ab_st_relaxed_u64(p.inclw + tl * Cs + cc, pack_word(epoch, ST_INCL, h));
        }
        for (int cc = c; cc < Cs; cc += blockDim.x) {
            const unsigned long long* w = p.words + (tl * Cs + cc) * 2;
            unsigned long long wp = ab_ld_relaxed_u64(w), wsv = ab_ld_relaxed_u64(w + 1);
            while (!word_valid(wp, epoch) || !word_valid(wsv, epoch)) {
                if (++spins > SPIN_LIMIT) { atomicExch(p.err_flag, 1u); __trap(); }
                wp = ab_ld_relaxed_u64(w); wsv = ab_ld_relaxed_u64(w + 1);
            }
            hs[cc] = fmaf(__uint_as_float((uint32_t)wp), hs[cc], __uint_as_float((uint32_t)wsv));
        }
    }
    if (DIR > 0 && p.h_last)
        for (int cc = c; cc < Cs; cc += blockDim.x) p.h_last[(size_t)b * p.Di + slab * Cs + cc] = hs[cc];
    return (int)blockDim.x;
}

// softplus'd delta rows of this tile (plus one extra row for the reverse scan) into shared memory
template <typename T>
__device__ __forceinline__ void stage_delta(const ScanParams& p, float* sdel, int b, int row0, int nrows, int h_lo, int nh) {
    const T* dl = reinterpret_cast<const T*>(p.dlog);
    // thread -> (row, head) without integer division: nh <= blockDim in every tiling
    const int hh = threadIdx.x % nh, r0 = threadIdx.x / nh, rstep = blockDim.x / nh;
    if (r0 >= rstep) return;                      // the last partial group of threads sits out
    const bool head_ok = h_lo + hh < p.H;
    for (int r = r0; r < nrows; r += rstep) {
        const int row = row0 + r;
        float d = 0.f;
        if (row < p.L && head_ok) d = ab_softplus_fast(ab_to_float(dl[((size_t)b * p.L + row) * p.H + h_lo + hh]));
        sdel[r * nh + hh] = d;
    }
}

// sigmoid for the SiLU gate: f32 activations use ex2 + rcp; bf16 activations use the single-MUFU tanh.approx form
// (relative error ~5e-4, an order below bf16 resolution)
template <typename T>
__device__ __forceinline__ float gate_sigmoid(float x) {
    if constexpr (sizeof(T) == 2) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    } else {
        return ab_sigmoid(x);
    }
}

template <typename T, int V>
__device__ __forceinline__ void lds_vec(const T* p, float* f) {
    if constexpr (sizeof(T) * V == 16) {
        ab_vec16<T>::unpack(*reinterpret_cast<const uint4*>(p), f);
    } else {   // 4 x bf16 = 8 bytes
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    }
}
template <typename T, int V>
__device__ __forceinline__ void st_vec(T* p, const float* f) {
    if constexpr (sizeof(T) * V == 16) {
        *reinterpret_cast<uint4*>(p) = ab_vec16<T>::pack(f);
    } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b2 = __floats2bfloat162_rn(f[2], f[3]);
        uint2 r; r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b2);
        *reinterpret_cast<uint2*>(p) = r;
    }
}
template <typename T, int V>
__device__ __forceinline__ void ldg_vec(const T* p, float* f) {
    if constexpr (sizeof(T) * V == 16) {
        ab_vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(p)), f);
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    }
}

}  // namespace ab_scan
