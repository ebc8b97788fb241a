// Chunked selective scan, forward and recompute-based backward (sm_100a).
//
// Replaces core.py:324-353 (both scan formulations), :383 (softplus), :394-396 (transpose copy, D skip,
// SiLU(z) gate) and their autograd.  Math per channel c = head*16 + n (fp32 throughout):
//     delta = softplus(dlog[b,t,head]);  abar = exp2(A2[c] * delta),  A2 = -exp(A_log[c]) * log2(e)
//     s_t = abar_t * s_{t-1} + Bm_t ;  y_ssm = Cm_t * s_t ;  y = (y_ssm + D * xa_t) * silu(z_t)
//
// Decomposition.  Tensors are [B, L, Di] channels-last.  A tile is `T` consecutive tokens x `Cs`
// channels of one sequence; its operands are staged in shared memory by TMA (3-D tiled maps, OOB rows
// zero-filled, completion on an mbarrier).  Inside the tile a thread owns one channel vector and a run
// of TS=4 tokens; run aggregates (P = prod abar, S = state contribution) are combined across runs in
// shared memory.  Across tiles of a sequence the incoming state is resolved
//   * single pass: each tile publishes (P, S) per channel as two self-validating 64-bit words
//     {epoch|status, fp32}.  One SCANNER CTA per chain (the first tickets) polls the aggregate words of the
//     next tiles into a shared-memory ring, composes them in token order (its threads split every round
//     into consecutive segments that are chained through shared memory) and publishes every tile's
//     incoming state as one more tagged word the tile spins on.  The wait per tile is O(1) regardless of
//     how many tiles of a chain are in flight, the composition order is fixed (bitwise reproducible), and
//     there is no deadlock by construction: CTAs take their role/tile from an atomic ticket, so the
//     scanners and every predecessor tile are resident or finished before a tile can wait on them.
//   * two pass: aggregate kernel -> segmented combine kernel -> apply kernel (no inter-CTA waits; used when
//     there are more chains than SMs, and selectable for debugging).
// The backward recomputes the in-tile states from the saved per-tile incoming state (hstart) and runs the
// same machinery in reverse time for G_t = g_t + abar_{t+1} G_{t+1}; the reverse aggregates are published
// first, so the hand-shake overlaps the forward recompute.
// tools/scan_trace.py (with `make TRACE=1`) prints the per-phase timeline of the tiles and the scanner.
#include "common.cuh"

#include "ssm_scan_shared.cuh"

namespace {
using namespace ab_scan;

struct ScanTiling {
    int V_f, V_b;     // channel-vector width of the forward / backward kernels
    int Cs, T, n_s;   // slab channels, tile rows, runs per tile
    int nslab, nchunks;
    int esize;
};

int make_tiling(int L, int Di, int dtype, ScanTiling& t) {
    t.esize = dtype == AB_F32 ? 4 : 2;
    t.V_f = 4;
    t.V_b = 4;
    const int unit = 16 / t.esize > t.V_b ? 16 / t.esize : t.V_b;    // slab rows are whole 16-byte units (TMA box)
    int Cs = 0;
    for (int k = Di / unit; k >= 1; --k) {           // smallest slab first
        if (Di % k) continue;
        const int c = Di / k;
        if (c % unit) continue;
        if (c * t.esize >= 128 && c <= 256) { Cs = c; break; }
    }
    if (!Cs) {
        if (Di % unit == 0 && Di <= 256) Cs = Di;    // narrow model: one slab
        else return 0;
    }
    t.Cs = Cs;
    t.nslab = Di / Cs;
    int T = (56 * 1024) / (5 * Cs * t.esize);
    const int Tthr = (256 * TS * t.V_b) / Cs;
    if (T > Tthr) T = Tthr;
    if (T > 256) T = 256;
    const int Lr = (int)ab_round_up(L, TS);
    if (T > Lr) T = Lr;
    T = (T / TS) * TS;
    if (T < TS) T = TS;
    t.T = T;
    t.n_s = T / TS;
    t.nchunks = (int)ab_ceil_div(L, T);
    return 1;
}


// ---------------------------------------------------------------------------------------------
// tile coordinates from the ticket / block index: chains interleaved, chunks in scan order
// ---------------------------------------------------------------------------------------------
// floats reserved for the staged dt rows, rounded so that the aggregate arrays behind them stay 16-byte aligned
__host__ __device__ __forceinline__ int sdel_floats(int T, int nh_max) { return ((T + 1) * nh_max + 3) & ~3; }

struct TileCoord { int chain, j, b, c0, row0, h_lo, nh; size_t tile_lin; };
__device__ __forceinline__ TileCoord decode_tile(const ScanParams& p, int tile, bool reverse) {
    TileCoord t;
    t.chain = tile % p.nchains;
    const int jj = tile / p.nchains;
    t.j = reverse ? p.nchunks - 1 - jj : jj;
    const int slab = t.chain % p.nslab;
    t.b = t.chain / p.nslab;
    t.c0 = slab * p.Cs;
    t.row0 = t.j * p.T;
    t.h_lo = t.c0 / 16;
    t.nh = (t.c0 + p.Cs - 1) / 16 - t.h_lo + 1;
    t.tile_lin = (size_t)t.chain * p.nchunks + t.j;
    return t;
}

// ---------------------------------------------------------------------------------------------
// forward, one tile per CTA (many small CTAs per SM hide the TMA latency and the scanner hand-shake)
// smem: [NT tiles of T x Cs] [sdel: (T+1) x nh] [sP, sS: n_s x Cs] [hT: Cs] [mbarrier]
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int V, int CS>
__global__ void __launch_bounds__(256, 4) scan_fwd_kernel(const __grid_constant__ CUtensorMap tm_xa,
                                                          const __grid_constant__ CUtensorMap tm_b,
                                                          const __grid_constant__ CUtensorMap tm_c,
                                                          const __grid_constant__ CUtensorMap tm_z, const ScanParams p) {
    constexpr int NT = MODE == MODE_AGG ? 1 : 4;      // the aggregate pass only needs Bm
    extern __shared__ __align__(128) unsigned char smem[];
    const int Cs = CS ? CS : p.Cs;                          // CS != 0: slab width known at compile time
    const int Tt = p.T, n_s = p.n_s;
    const size_t tile_bytes = (size_t)Tt * Cs * sizeof(T);
    const size_t pitch = (tile_bytes + 127) / 128 * 128;     // TMA destinations are 128-byte aligned
    const int nh_max = Cs / 16 + 2;
    float* sdel = reinterpret_cast<float*>(smem + (size_t)NT * pitch);
    float* sP = sdel + sdel_floats(Tt, nh_max);
    float* sS = sP + (size_t)n_s * Cs;
    float* hT = sS + (size_t)n_s * Cs;
    uint64_t* bars = reinterpret_cast<uint64_t*>(hT + Cs + (((uintptr_t)(hT + Cs)) % 8 ? 1 : 0));
    __shared__ unsigned int s_ticket;

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_ticket = MODE == MODE_FUSED ? atomicAdd(p.ticket, 1u) : blockIdx.x;
        ab_mbar_init(&bars[0], 1);
        ab_fence_mbar_init();
    }
    __syncthreads();
    int tile = (int)s_ticket;
    if (MODE == MODE_FUSED) {
        if (tile < p.n_scan) {
            scanner_role<+1>(p, p.epoch, SCAN_K, 4, tile, hT, reinterpret_cast<uint4*>(smem), (size_t)(reinterpret_cast<unsigned char*>(hT) - smem));
            return;
        }
        tile -= p.n_scan;
    }
    if (tile >= p.nchains * p.nchunks) return;

    const TileCoord tc = decode_tile(p, tile, false);
    const size_t tile_lin = tc.tile_lin;
    const int c0 = tc.c0, row0 = tc.row0, b = tc.b, nh = tc.nh, j = tc.j;
    TRACE_MARK(tile_lin, 0);
    if (tid == 0) {
        ab_mbar_expect_tx(&bars[0], (uint32_t)(NT * tile_bytes));
        ab_tma_load_3d(smem, &tm_b, &bars[0], c0, row0, b);
        if (MODE != MODE_AGG) {
            ab_tma_load_3d(smem + pitch, &tm_xa, &bars[0], c0, row0, b);
            ab_tma_load_3d(smem + 2 * pitch, &tm_c, &bars[0], c0, row0, b);
            ab_tma_load_3d(smem + 3 * pitch, &tm_z, &bars[0], c0, row0, b);
        }
    }
    stage_delta<T>(p, sdel, b, row0, Tt, tc.h_lo, nh);

    const int ncv = Cs / V;
    const int i_run = tid / ncv, cv = tid % ncv;       // blockDim.x == n_s * ncv
    const int cl = cv * V;                             // channel offset inside the slab
    const T* s_b = reinterpret_cast<const T*>(smem);
    const T* s_xa = reinterpret_cast<const T*>(smem + pitch);
    const T* s_c = reinterpret_cast<const T*>(smem + 2 * pitch);
    const T* s_z = reinterpret_cast<const T*>(smem + 3 * pitch);
    const int hh = (c0 + cl) / 16 - tc.h_lo;           // head slot of this thread's channels (V divides 16)
    float A2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) A2[v] = -__expf(__ldg(p.A_log + c0 + cl + v)) * AB_LOG2E;

    __syncthreads();            // sdel visible
    TRACE_MARK(tile_lin, 1);
    ab_mbar_wait(&bars[0], 0);  // TMA tiles landed
    TRACE_MARK(tile_lin, 2);

    // ---- sweep 1: per-run aggregates (the decay factors are recomputed in sweep 2: one MUFU is cheaper than
    //      holding TS x V registers across the hand-shake with the scanner)
    {
        float P[V], S[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { P[v] = 1.f; S[v] = 0.f; }
#pragma unroll
        for (int t = 0; t < TS; ++t) {
            const int r = i_run * TS + t;
            const float d = sdel[r * nh + hh];
            float bv[V];
            lds_vec<T, V>(s_b + (size_t)r * Cs + cl, bv);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float at = ab_ex2(A2[v] * d);
                P[v] *= at;
                S[v] = fmaf(at, S[v], bv[v]);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { sP[i_run * Cs + cl + v] = P[v]; sS[i_run * Cs + cl + v] = S[v]; }
    }
    __syncthreads();
    TRACE_MARK(tile_lin, 3);

    // ---- tile aggregate, publish, resolve incoming state (one thread per channel)
    for (int c = tid; c < Cs; c += blockDim.x) {
        // run aggregates are read 8 at a time (independent loads first, then the dependent FMA chain)
        float Pt = 1.f, St = 0.f;
        for (int i0 = 0; i0 < n_s; i0 += 8) {
            float P8[8], S8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = i0 + u < n_s;
                P8[u] = ok ? sP[(i0 + u) * Cs + c] : 1.f;
                S8[u] = ok ? sS[(i0 + u) * Cs + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { St = fmaf(St, P8[u], S8[u]); Pt *= P8[u]; }
        }
        float hin;
        if (MODE == MODE_AGG) {
            p.aggP[tile_lin * Cs + c] = Pt;
            p.aggS[tile_lin * Cs + c] = St;
            continue;
        } else if (MODE == MODE_APPLY) {
            hin = p.hstart[((size_t)b * p.nchunks + j) * p.Di + c0 + c];
        } else {
            unsigned long long* w = p.words + (tile_lin * Cs + c) * 2;
            ab_st_relaxed_u64(w, pack_word(p.epoch, ST_AGG, Pt));
            ab_st_relaxed_u64(w + 1, pack_word(p.epoch, ST_AGG, St));
            TRACE_MARK(tile_lin, 4);
            hin = wait_incoming(p, p.epoch, tile_lin, c);        // h_last is written by the scanner
            TRACE_MARK(tile_lin, 5);
            if (p.hstart) p.hstart[((size_t)b * p.nchunks + j) * p.Di + c0 + c] = hin;     // kept for the backward
        }
        // state entering every run of this channel, in place of the run aggregates S (one serial pass by the
        // channel's thread instead of an O(n_s) prefix loop in every thread)
        float hr = hin;
        for (int i0 = 0; i0 < n_s; i0 += 8) {
            float P8[8], S8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = i0 + u < n_s;
                P8[u] = ok ? sP[(i0 + u) * Cs + c] : 1.f;
                S8[u] = ok ? sS[(i0 + u) * Cs + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (i0 + u < n_s) sS[(i0 + u) * Cs + c] = hr;
                hr = fmaf(P8[u], hr, S8[u]);
            }
        }
    }
    if (MODE == MODE_AGG) return;
    __syncthreads();
    TRACE_MARK(tile_lin, 6);

    // ---- state entering this thread's run, then sweep 2
    float h[V], Dv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { h[v] = sS[i_run * Cs + cl + v]; Dv[v] = __ldg(p.Dp + c0 + cl + v); }
    const size_t off0 = ((size_t)b * p.L + row0 + i_run * TS) * p.Di + c0 + cl;
    T* yo = reinterpret_cast<T*>(p.y) + off0;
    T* yso = p.y_ssm ? reinterpret_cast<T*>(p.y_ssm) + off0 : nullptr;
#pragma unroll
    for (int t = 0; t < TS; ++t) {
        const int r = i_run * TS + t;
        const int row = row0 + r;
        float bv[V], cvv[V], xv[V], zv[V], o[V], os[V];
        lds_vec<T, V>(s_b + (size_t)r * Cs + cl, bv);
        lds_vec<T, V>(s_c + (size_t)r * Cs + cl, cvv);
        lds_vec<T, V>(s_xa + (size_t)r * Cs + cl, xv);
        lds_vec<T, V>(s_z + (size_t)r * Cs + cl, zv);
        const float d = sdel[r * nh + hh];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            h[v] = fmaf(ab_ex2(A2[v] * d), h[v], bv[v]);
            os[v] = cvv[v] * h[v];
            o[v] = fmaf(Dv[v], xv[v], os[v]) * (zv[v] * gate_sigmoid<T>(zv[v]));
        }
        if (row < p.L) {
            st_vec<T, V>(yo + (size_t)t * p.Di, o);
            if (yso) st_vec<T, V>(yso + (size_t)t * p.Di, os);
        }
    }
    TRACE_MARK(tile_lin, 7);
}

// ---------------------------------------------------------------------------------------------
// two-pass forward, aggregate pass: reads dt and Bm only.  A CTA owns AGG_G consecutive tiles of one chain; all of their
// TMA loads and dt rows are in flight from the start (one exposed load latency per AGG_G tiles), then the tiles are
// reduced one after the other.  smem: [AGG_G tiles of T x Cs] [sdel: AGG_G*T x nh] [sP, sS: n_s x Cs] [mbarrier]
// ---------------------------------------------------------------------------------------------
constexpr int AGG_G = 4;
template <typename T, int CS>
__global__ void __launch_bounds__(256, 4) scan_fwd_agg_kernel(const __grid_constant__ CUtensorMap tm_b, const ScanParams p) {
    constexpr int V = 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Cs = CS ? CS : p.Cs;
    const int Tt = p.T, n_s = p.n_s;
    const size_t tile_bytes = (size_t)Tt * Cs * sizeof(T);
    const size_t pitch = (tile_bytes + 127) / 128 * 128;
    const int nh_max = Cs / 16 + 2;
    float* sdel = reinterpret_cast<float*>(smem + (size_t)AGG_G * pitch);
    float* sP = sdel + sdel_floats(AGG_G * Tt, nh_max);
    float* sS = sP + (size_t)n_s * Cs;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + (size_t)n_s * Cs + (((uintptr_t)(sS + (size_t)n_s * Cs)) % 8 ? 1 : 0));
    const int tid = threadIdx.x;
    const int ngrp = (p.nchunks + AGG_G - 1) / AGG_G;
    const int chain = blockIdx.x % p.nchains, jg = blockIdx.x / p.nchains;
    if (jg >= ngrp) return;
    const int j0 = jg * AGG_G, ng = min(AGG_G, p.nchunks - j0);
    const int slab = chain % p.nslab, b = chain / p.nslab;
    const int c0 = slab * Cs, row0 = j0 * Tt;
    const int h_lo = c0 / 16, nh = (c0 + Cs - 1) / 16 - h_lo + 1;
    if (tid == 0) {
        ab_mbar_init(&bars[0], 1);
        ab_fence_mbar_init();
        ab_mbar_expect_tx(&bars[0], (uint32_t)(ng * tile_bytes));
        for (int g = 0; g < ng; ++g) ab_tma_load_3d(smem + (size_t)g * pitch, &tm_b, &bars[0], c0, row0 + g * Tt, b);
    }
    stage_delta<T>(p, sdel, b, row0, ng * Tt, h_lo, nh);
    const int ncv = Cs / V;
    const int i_run = tid / ncv, cv = tid % ncv;
    const int cl = cv * V;
    const int hh = (c0 + cl) / 16 - h_lo;
    float A2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) A2[v] = -__expf(__ldg(p.A_log + c0 + cl + v)) * AB_LOG2E;
    __syncthreads();            // sdel and the barrier initialisation visible
    ab_mbar_wait(&bars[0], 0);
    for (int g = 0; g < ng; ++g) {
        const T* s_b = reinterpret_cast<const T*>(smem + (size_t)g * pitch);
        float P[V], S[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { P[v] = 1.f; S[v] = 0.f; }
#pragma unroll
        for (int t = 0; t < TS; ++t) {
            const int r = i_run * TS + t;
            const float d = sdel[(g * Tt + r) * nh + hh];
            float bv[V];
            lds_vec<T, V>(s_b + (size_t)r * Cs + cl, bv);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float at = ab_ex2(A2[v] * d);
                P[v] *= at;
                S[v] = fmaf(at, S[v], bv[v]);
            }
        }
        if (g) __syncthreads();                     // the previous tile's aggregates have been consumed
        *reinterpret_cast<float4*>(sP + i_run * Cs + cl) = make_float4(P[0], P[1], P[2], P[3]);
        *reinterpret_cast<float4*>(sS + i_run * Cs + cl) = make_float4(S[0], S[1], S[2], S[3]);
        __syncthreads();
        const size_t tile_lin = (size_t)chain * p.nchunks + j0 + g;
        for (int c = tid; c < Cs; c += blockDim.x) {
            float Pt = 1.f, St = 0.f;
            for (int i0 = 0; i0 < n_s; i0 += 8) {
                float P8[8], S8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = i0 + u < n_s;
                    P8[u] = ok ? sP[(i0 + u) * Cs + c] : 1.f;
                    S8[u] = ok ? sS[(i0 + u) * Cs + c] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) { St = fmaf(St, P8[u], S8[u]); Pt *= P8[u]; }
            }
            p.aggP[tile_lin * Cs + c] = Pt;
            p.aggS[tile_lin * Cs + c] = St;
        }
    }
}

// two-pass: state entering every tile.  A block owns COMB_CH channels; the chunk axis is cut into COMB_SEG segments that
// are composed in parallel (pass 1), chained through shared memory, and re-walked to emit the per-tile states (pass 2;
// the aggregates are L2 hits by then).  Loads of 8 tiles are issued ahead of the 8 dependent FMAs.
constexpr int COMB_CH = 16, COMB_SEG = 32;
__global__ void __launch_bounds__(COMB_CH * COMB_SEG) scan_combine_kernel(
    const float* __restrict__ aggP, const float* __restrict__ aggS, const float* __restrict__ h0,
    float* __restrict__ hstart, float* __restrict__ h_last, int B, int Di, int Cs, int nslab, int nchunks, int reverse) {
    __shared__ float segP[COMB_SEG][COMB_CH], segS[COMB_SEG][COMB_CH];
    const int cx = threadIdx.x % COMB_CH, sg = threadIdx.x / COMB_CH;
    const int g = blockIdx.x * COMB_CH + cx;
    const bool live = g < B * Di;
    const int b = live ? g / Di : 0, cg = live ? g % Di : 0;
    const int slab = cg / Cs, c = cg % Cs;
    const size_t chain_base = (size_t)(b * nslab + slab) * nchunks;
    const int seglen = (nchunks + COMB_SEG - 1) / COMB_SEG;
    const int lo = min(sg * seglen, nchunks), hi = min(lo + seglen, nchunks);
    float Pa = 1.f, Sa = 0.f;
    if (live) {
        for (int j0 = lo; j0 < hi; j0 += 8) {
            float P[8], S[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int jj = j0 + u;
                P[u] = 1.f; S[u] = 0.f;
                if (jj < hi) {
                    const size_t o = (chain_base + (reverse ? nchunks - 1 - jj : jj)) * Cs + c;
                    P[u] = aggP[o]; S[u] = aggS[o];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { Sa = fmaf(P[u], Sa, S[u]); Pa *= P[u]; }
        }
    }
    segP[sg][cx] = Pa; segS[sg][cx] = Sa;
    __syncthreads();
    if (!live) return;
    float h = (!reverse && h0) ? h0[(size_t)b * Di + cg] : 0.f;
    for (int i = 0; i < sg; ++i) h = fmaf(segP[i][cx], h, segS[i][cx]);
    for (int j0 = lo; j0 < hi; j0 += 8) {
        float P[8], S[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int jj = j0 + u;
            P[u] = 1.f; S[u] = 0.f;
            if (jj < hi) {
                const size_t o = (chain_base + (reverse ? nchunks - 1 - jj : jj)) * Cs + c;
                P[u] = aggP[o]; S[u] = aggS[o];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int jj = j0 + u;
            if (jj < hi) {
                const int j = reverse ? nchunks - 1 - jj : jj;
                hstart[((size_t)b * nchunks + j) * Di + cg] = h;
                h = fmaf(P[u], h, S[u]);
            }
        }
    }
    if (h_last && sg == COMB_SEG - 1) h_last[(size_t)b * Di + cg] = h;
}

// ---------------------------------------------------------------------------------------------
// backward (V = 4 channels per thread for both dtypes), one tile per CTA, tiles taken in REVERSE time order.
// Phase order (the hand-shake of the reverse scan is hidden behind the forward recompute):
//   A  one sweep: decay factors, g_t = d(y_ssm)_t * C_t, reverse run aggregates (sP/sS) and forward run aggregates (fP/fS)
//   B  channel threads: reverse tile aggregate -> publish (single pass) / store (aggregate pass)
//   C  forward recompute from hstart: states, dxa / dCm / dz, keeps h_{t-1}
//   D  channel threads: wait for G entering the tile (scanner / combine kernel)
//   E  reverse sweep: G_t, dBm, d(dt), per-tile partials of dA_log and dD
// smem: [5 tiles of T x Cs] [sdel: (T+1) x nh] [sP, sS, fP, fS: n_s x Cs] [hT, gT: Cs] [mbarrier]
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CS>
__global__ void __launch_bounds__(256, 3) scan_bwd_kernel(const __grid_constant__ CUtensorMap tm_xa,
                                                       const __grid_constant__ CUtensorMap tm_b,
                                                       const __grid_constant__ CUtensorMap tm_c,
                                                       const __grid_constant__ CUtensorMap tm_z,
                                                       const __grid_constant__ CUtensorMap tm_do, const ScanParams p,
                                                       float* __restrict__ gin_ws) {
    constexpr int V = 4;
    constexpr int NT = 5;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Cs = CS ? CS : p.Cs;
    const int Tt = p.T, n_s = p.n_s;
    const size_t tile_bytes = (size_t)Tt * Cs * sizeof(T);
    const size_t pitch = (tile_bytes + 127) / 128 * 128;
    constexpr int n_staged = MODE == MODE_AGG ? 3 : 5;      // the aggregate pass needs Cm, z, dout only
    const int nh_max = Cs / 16 + 2;
    float* sdel = reinterpret_cast<float*>(smem + (size_t)NT * pitch);
    float* sP = sdel + sdel_floats(Tt, nh_max);
    float* sS = sP + (size_t)n_s * Cs;
    float* fP = sS + (size_t)n_s * Cs;
    float* fS = fP + (size_t)n_s * Cs;
    float* hT = fS + (size_t)n_s * Cs;
    float* gT = hT + Cs;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gT + Cs + (((uintptr_t)(gT + Cs)) % 8 ? 1 : 0));
    __shared__ unsigned int s_ticket;

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_ticket = MODE == MODE_FUSED ? atomicAdd(p.ticket, 1u) : blockIdx.x;
        ab_mbar_init(&bars[0], 1);
        ab_fence_mbar_init();
    }
    __syncthreads();
    int tile = (int)s_ticket;
    if (MODE == MODE_FUSED) {
        if (tile < p.n_scan) {
            scanner_role<-1>(p, p.epoch, SCAN_K, 4, tile, hT, reinterpret_cast<uint4*>(smem), (size_t)(reinterpret_cast<unsigned char*>(hT) - smem));
            return;
        }
        tile -= p.n_scan;
    }
    if (tile >= p.nchains * p.nchunks) return;

    const TileCoord tc = decode_tile(p, tile, true);
    const size_t tile_lin = tc.tile_lin;
    const int c0 = tc.c0, row0 = tc.row0, b = tc.b, nh = tc.nh, j = tc.j;
    TRACE_MARK(tile_lin, 0);
    if (tid == 0) {
        ab_mbar_expect_tx(&bars[0], (uint32_t)(n_staged * tile_bytes));
        if (MODE != MODE_AGG) ab_tma_load_3d(smem, &tm_b, &bars[0], c0, row0, b);
        ab_tma_load_3d(smem + pitch, &tm_c, &bars[0], c0, row0, b);
        ab_tma_load_3d(smem + 2 * pitch, &tm_z, &bars[0], c0, row0, b);
        ab_tma_load_3d(smem + 3 * pitch, &tm_do, &bars[0], c0, row0, b);
        if (MODE != MODE_AGG) ab_tma_load_3d(smem + 4 * pitch, &tm_xa, &bars[0], c0, row0, b);
    }
    stage_delta<T>(p, sdel, b, row0, Tt + 1, tc.h_lo, nh);

    const int ncv = Cs / V;
    const int i_run = tid / ncv, cv = tid % ncv;
    const int cl = cv * V;
    const T* s_b = reinterpret_cast<const T*>(smem);
    const T* s_c = reinterpret_cast<const T*>(smem + pitch);
    const T* s_z = reinterpret_cast<const T*>(smem + 2 * pitch);
    const T* s_do = reinterpret_cast<const T*>(smem + 3 * pitch);
    const T* s_xa = reinterpret_cast<const T*>(smem + 4 * pitch);
    const T* dys_i = reinterpret_cast<const T*>(p.dyssm);
    const int hh = (c0 + cl) / 16 - tc.h_lo;
    float A2[V], Dv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        A2[v] = -__expf(__ldg(p.A_log + c0 + cl + v)) * AB_LOG2E;
        Dv[v] = __ldg(p.Dp + c0 + cl + v);
    }
    if (MODE != MODE_AGG)
        for (int c = tid; c < Cs; c += blockDim.x) hT[c] = p.hstart[((size_t)b * p.nchunks + j) * p.Di + c0 + c];

    __syncthreads();            // sdel + hT visible
    ab_mbar_wait(&bars[0], 0);
    TRACE_MARK(tile_lin, 1);

    // ---- A: decay factors, g, reverse and forward run aggregates
    float a[TS][V], g[TS][V];
    float dl[TS];
    float anext[V];
    {
        float P[V], S[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { P[v] = 1.f; S[v] = 0.f; }
#pragma unroll
        for (int t = 0; t < TS; ++t) {
            const int r = i_run * TS + t;
            const int row = row0 + r;
            dl[t] = sdel[r * nh + hh];
            float bv[V], cvv[V], zv[V], dov[V], dys[V];
            lds_vec<T, V>(s_c + (size_t)r * Cs + cl, cvv);
            lds_vec<T, V>(s_z + (size_t)r * Cs + cl, zv);
            lds_vec<T, V>(s_do + (size_t)r * Cs + cl, dov);
            if (MODE != MODE_AGG) lds_vec<T, V>(s_b + (size_t)r * Cs + cl, bv);
#pragma unroll
            for (int v = 0; v < V; ++v) dys[v] = 0.f;
            if (dys_i && row < p.L) ldg_vec<T, V>(dys_i + ((size_t)b * p.L + row) * p.Di + c0 + cl, dys);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                a[t][v] = ab_ex2(A2[v] * dl[t]);
                g[t][v] = fmaf(dov[v], zv[v] * gate_sigmoid<T>(zv[v]), dys[v]) * cvv[v];
                if (MODE != MODE_AGG) {
                    P[v] *= a[t][v];
                    S[v] = fmaf(a[t][v], S[v], bv[v]);
                }
            }
        }
        if (MODE != MODE_AGG) {
            *reinterpret_cast<float4*>(fP + i_run * Cs + cl) = make_float4(P[0], P[1], P[2], P[3]);
            *reinterpret_cast<float4*>(fS + i_run * Cs + cl) = make_float4(S[0], S[1], S[2], S[3]);
        }
        // reverse:  G(run start) = Gs + Pr * G(next run start)
        const float dn = sdel[(i_run * TS + TS) * nh + hh];
        float Gs[V], Pr[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            anext[v] = ab_ex2(A2[v] * dn);
            Gs[v] = g[TS - 1][v];
            Pr[v] = anext[v];
        }
#pragma unroll
        for (int t = TS - 2; t >= 0; --t) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                Gs[v] = fmaf(a[t + 1][v], Gs[v], g[t][v]);
                Pr[v] *= a[t + 1][v];
            }
        }
        *reinterpret_cast<float4*>(sP + i_run * Cs + cl) = make_float4(Pr[0], Pr[1], Pr[2], Pr[3]);
        *reinterpret_cast<float4*>(sS + i_run * Cs + cl) = make_float4(Gs[0], Gs[1], Gs[2], Gs[3]);
    }
    __syncthreads();
    TRACE_MARK(tile_lin, 2);

    // ---- B: reverse tile aggregate (one thread per channel); the single pass publishes it and collects the answer in D
    for (int c = tid; c < Cs; c += blockDim.x) {
        float Pt = 1.f, St = 0.f;
        for (int i0 = n_s - 1; i0 >= 0; i0 -= 8) {
            float P8[8], S8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = i0 - u >= 0;
                P8[u] = ok ? sP[(i0 - u) * Cs + c] : 1.f;
                S8[u] = ok ? sS[(i0 - u) * Cs + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { St = fmaf(St, P8[u], S8[u]); Pt *= P8[u]; }
        }
        if (MODE == MODE_AGG) {
            p.aggP[tile_lin * Cs + c] = Pt;
            p.aggS[tile_lin * Cs + c] = St;
        } else if (MODE == MODE_FUSED) {
            unsigned long long* w = p.words + (tile_lin * Cs + c) * 2;
            ab_st_relaxed_u64(w, pack_word(p.epoch, ST_AGG, Pt));
            ab_st_relaxed_u64(w + 1, pack_word(p.epoch, ST_AGG, St));
        }
    }
    if (MODE == MODE_AGG) return;
    TRACE_MARK(tile_lin, 3);

    // ---- C: recompute states, emit dxa / dCm / dz, keep h_{t-1}
    float hprev[TS][V];
    float accD[V];
#pragma unroll
    for (int v = 0; v < V; ++v) accD[v] = 0.f;
    {
        float h[V];
#pragma unroll
        for (int v = 0; v < V; ++v) h[v] = hT[cl + v];
#pragma unroll 4
        for (int i = 0; i < i_run; ++i) {
            const float4 pp = *reinterpret_cast<const float4*>(fP + i * Cs + cl), ss = *reinterpret_cast<const float4*>(fS + i * Cs + cl);
            h[0] = fmaf(h[0], pp.x, ss.x); h[1] = fmaf(h[1], pp.y, ss.y); h[2] = fmaf(h[2], pp.z, ss.z); h[3] = fmaf(h[3], pp.w, ss.w);
        }
        const size_t tok0 = (size_t)b * p.L + row0 + i_run * TS;          // first token of this run
        T* dxa_o = reinterpret_cast<T*>(p.dxa) + tok0 * p.Di + c0 + cl;
        T* dz_o = reinterpret_cast<T*>(p.dz) + tok0 * p.Di + c0 + cl;
        T* dc_o = reinterpret_cast<T*>(p.dCm) + tok0 * p.dbc_stride + c0 + cl;
#pragma unroll
        for (int t = 0; t < TS; ++t) {
            const int r = i_run * TS + t;
            const int row = row0 + r;
            float bv[V], cvv[V], zv[V], dov[V], xv[V], dys[V];
            lds_vec<T, V>(s_c + (size_t)r * Cs + cl, cvv);
            lds_vec<T, V>(s_z + (size_t)r * Cs + cl, zv);
            lds_vec<T, V>(s_do + (size_t)r * Cs + cl, dov);
            lds_vec<T, V>(s_b + (size_t)r * Cs + cl, bv);
            lds_vec<T, V>(s_xa + (size_t)r * Cs + cl, xv);
#pragma unroll
            for (int v = 0; v < V; ++v) dys[v] = 0.f;
            if (dys_i && row < p.L) ldg_vec<T, V>(dys_i + ((size_t)b * p.L + row) * p.Di + c0 + cl, dys);
            float o_dxa[V], o_dc[V], o_dz[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float sg = gate_sigmoid<T>(zv[v]);
                const float dyv = dov[v] * (zv[v] * sg);       // grad of (y_ssm + D*xa)
                const float dtot = dyv + dys[v];               // grad of y_ssm
                hprev[t][v] = h[v];
                h[v] = fmaf(a[t][v], h[v], bv[v]);
                const float yv = fmaf(Dv[v], xv[v], cvv[v] * h[v]);
                o_dxa[v] = dyv * Dv[v];
                o_dc[v] = dtot * h[v];
                o_dz[v] = dov[v] * yv * (sg * fmaf(zv[v], 1.f - sg, 1.f));
                accD[v] = fmaf(dyv, xv[v], accD[v]);
            }
            if (row < p.L) {
                st_vec<T, V>(dxa_o + (size_t)t * p.Di, o_dxa);
                st_vec<T, V>(dz_o + (size_t)t * p.Di, o_dz);
                st_vec<T, V>(dc_o + (size_t)t * p.dbc_stride, o_dc);
            }
        }
    }
    TRACE_MARK(tile_lin, 4);

    // ---- D: G entering the tile from the later tiles
    for (int c = tid; c < Cs; c += blockDim.x)
        gT[c] = MODE == MODE_APPLY ? gin_ws[((size_t)b * p.nchunks + j) * p.Di + c0 + c] : wait_incoming(p, p.epoch, tile_lin, c);
    TRACE_MARK(tile_lin, 5);
    __syncthreads();
    TRACE_MARK(tile_lin, 6);

    // ---- E: G entering this run from later runs, then the reverse sweep
    float accA[V];
    {
        float G[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { G[v] = gT[cl + v]; accA[v] = 0.f; }
#pragma unroll 4
        for (int i = n_s - 1; i > i_run; --i) {
            const float4 pp = *reinterpret_cast<const float4*>(sP + i * Cs + cl), ss = *reinterpret_cast<const float4*>(sS + i * Cs + cl);
            G[0] = fmaf(G[0], pp.x, ss.x); G[1] = fmaf(G[1], pp.y, ss.y); G[2] = fmaf(G[2], pp.z, ss.z); G[3] = fmaf(G[3], pp.w, ss.w);
        }
        float q[V];
#pragma unroll
        for (int v = 0; v < V; ++v) q[v] = anext[v] * G[v];
        const size_t tok0 = (size_t)b * p.L + row0 + i_run * TS;
        T* db_o = reinterpret_cast<T*>(p.dBm) + tok0 * p.dbc_stride + c0 + cl;
        const int nparts = p.Di / V;
        float* dd_o = p.ddlog_parts + tok0 * nparts + (c0 + cl) / V;
#pragma unroll
        for (int t = TS - 1; t >= 0; --t) {
            const int row = row0 + i_run * TS + t;
            float o_db[V];
            float dd = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float Gt = g[t][v] + q[v];
                o_db[v] = Gt;
                const float e = Gt * hprev[t][v] * a[t][v];      // d abar * abar
                dd = fmaf(e, A2[v], dd);
                accA[v] = fmaf(e, dl[t], accA[v]);
                q[v] = a[t][v] * Gt;
            }
            if (row < p.L) {
                st_vec<T, V>(db_o + (size_t)t * p.dbc_stride, o_db);
                // d delta = sum_n e * A  (A = A2 / log2e);  d dlog = d delta * sigmoid(dlog) = d delta * (1 - exp(-delta))
                const float sp = 1.f - __expf(-dl[t]);
                dd_o[(size_t)t * nparts] = dd * (1.f / AB_LOG2E) * sp;
            }
        }
    }
    // ---- per-tile partial sums of dA_log (= A * sum e*delta) and dD; the forward aggregates are dead (every thread is
    //      past phase C once it left the barrier above)
    *reinterpret_cast<float4*>(fP + i_run * Cs + cl) = make_float4(accA[0] * A2[0] * (1.f / AB_LOG2E), accA[1] * A2[1] * (1.f / AB_LOG2E),
                                                                  accA[2] * A2[2] * (1.f / AB_LOG2E), accA[3] * A2[3] * (1.f / AB_LOG2E));
    *reinterpret_cast<float4*>(fS + i_run * Cs + cl) = make_float4(accD[0], accD[1], accD[2], accD[3]);
    __syncthreads();
    for (int c = tid; c < Cs; c += blockDim.x) {
        float sa = 0.f, sd = 0.f;
#pragma unroll 4
        for (int i = 0; i < n_s; ++i) { sa += fP[i * Cs + c]; sd += fS[i * Cs + c]; }
        p.part[(tile_lin * 2 + 0) * Cs + c] = sa;
        p.part[(tile_lin * 2 + 1) * Cs + c] = sd;
    }
    TRACE_MARK(tile_lin, 7);
}

// dA_log[c], dD[c] = sum over (b, j) of the tile partials.  A block owns PR_CH channels (adjacent threads = adjacent
// channels: coalesced rows); PR_Q threads per channel each add every PR_Q-th tile, then one thread per channel adds
// the PR_Q partial sums in index order: fixed order -> deterministic.
constexpr int PR_CH = 16, PR_Q = 64;
__global__ void __launch_bounds__(PR_CH * PR_Q) scan_param_reduce_kernel(const float* __restrict__ part, float* __restrict__ dA,
                                                                         float* __restrict__ dD, int B, int Di, int Cs, int nslab,
                                                                         int nchunks) {
    __shared__ float sA[PR_Q][PR_CH], sD[PR_Q][PR_CH];
    const int cx = threadIdx.x % PR_CH, q = threadIdx.x / PR_CH;
    const int cg = blockIdx.x * PR_CH + cx;
    float sa = 0.f, sd = 0.f;
    if (cg < Di) {
        const int slab = cg / Cs, c = cg % Cs;
        const int n = B * nchunks;
        for (int i0 = q; i0 < n; i0 += PR_Q * 4) {
            float va[4], vd[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * PR_Q;
                va[u] = 0.f; vd[u] = 0.f;
                if (i < n) {
                    const int b = i / nchunks, j = i % nchunks;
                    const size_t tl = (size_t)(b * nslab + slab) * nchunks + j;
                    va[u] = part[(tl * 2 + 0) * Cs + c];
                    vd[u] = part[(tl * 2 + 1) * Cs + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { sa += va[u]; sd += vd[u]; }
        }
    }
    sA[q][cx] = sa; sD[q][cx] = sd;
    __syncthreads();
    if (q == 0 && cg < Di) {
        float ta = 0.f, td = 0.f;
#pragma unroll 8
        for (int i = 0; i < PR_Q; ++i) { ta += sA[i][cx]; td += sD[i][cx]; }
        dA[cg] = ta; dD[cg] = td;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct WsLayout {
    size_t off_ticket, off_err, off_words, off_incl, off_aggP, off_aggS, off_gin, off_part, total;
};
WsLayout ws_layout(const ScanTiling& t, int B, int Di) {
    WsLayout w;
    const size_t ntiles = (size_t)B * t.nslab * t.nchunks;
    size_t o = 0;
    w.off_ticket = o; o += 64;
    w.off_err = o; o += 64;
    w.off_words = o; o += ntiles * t.Cs * 2 * sizeof(unsigned long long);
    w.off_incl = o; o += ntiles * t.Cs * sizeof(unsigned long long);
    w.off_aggP = o; o += ntiles * t.Cs * sizeof(float);
    w.off_aggS = o; o += ntiles * t.Cs * sizeof(float);
    w.off_gin = o; o += (size_t)B * t.nchunks * Di * sizeof(float);
    w.off_part = o; o += ntiles * 2 * t.Cs * sizeof(float);
    w.total = ab_round_up((int64_t)o, 256);
    return w;
}

int make_map3(CUtensorMap* m, const void* base, int dtype, int B, int L, int Di, int64_t row_stride, int Cs, int T) {
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE(((uintptr_t)base % 16) == 0 && (row_stride * es) % 16 == 0,
               "selective_scan: tensor base and row stride must be 16-byte aligned for TMA (stride %lld elems)", (long long)row_stride);
    uint64_t dims[3] = {(uint64_t)Di, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)row_stride * es, (uint64_t)row_stride * es * L};
    uint32_t box[3] = {(uint32_t)Cs, (uint32_t)T, 1};
    return ab_encode_tmap(m, dtype == AB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base,
                          dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

size_t smem_bytes(const ScanTiling& t, int ntiles_staged, bool bwd) {
    const size_t tile_bytes = (size_t)t.T * t.Cs * t.esize;
    const size_t pitch = (tile_bytes + 127) / 128 * 128;
    const int nh_max = t.Cs / 16 + 2;
    size_t o = (size_t)ntiles_staged * pitch;
    o += (size_t)sdel_floats(t.T, nh_max) * 4;
    o += (size_t)(bwd ? 4 : 2) * t.n_s * t.Cs * 4;      // run aggregates (the backward keeps forward and reverse ones)
    o += (size_t)(bwd ? 2 : 1) * t.Cs * 4 + 8 + 32;
    return o;
}

// scanner CTAs: one per chain, all resident for the whole launch
int scanner_ctas(const ScanParams& p, int threads) { (void)threads; return p.nchains; }
// single pass only while the scanners occupy a small part of the machine; with that many independent chains the
// two-pass schedule has all the parallelism it needs
bool single_pass_ok(int B, const ScanTiling& t, int threads) {
    ScanParams q;
    q.nchains = B * t.nslab; q.Cs = t.Cs;
    return scanner_ctas(q, threads) <= ab_num_sms();
}

template <typename K>
int prepare_kernel(K kfn, size_t smem) {
    AB_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // without this the driver may pick a carve-out that admits fewer CTAs per SM than the shared memory allows
    AB_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return AB_OK;
}

template <typename T, int MODE, int V, int CS>
int launch_fwd_cs(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, int n_staged, cudaStream_t st) {
    ScanParams q = p;
    const size_t smem = smem_bytes(t, n_staged, false);
    auto kfn = scan_fwd_kernel<T, MODE, V, CS>;
    if (int e = prepare_kernel(kfn, smem)) return e;
    const int threads = t.n_s * (t.Cs / V);
    q.n_scan = MODE == MODE_FUSED ? scanner_ctas(p, threads) : 0;
    kfn<<<(unsigned)(p.nchains * p.nchunks + q.n_scan), threads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], q);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
// slab widths with a specialised kernel: 64 (d_inner a multiple of 64) and 88 (the 1.5B text block, d_inner 176)
template <typename T, int MODE, int V>
int launch_fwd_v(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, int n_staged, cudaStream_t st) {
    if (t.Cs == 64) return launch_fwd_cs<T, MODE, V, 64>(maps, p, t, n_staged, st);
    if (t.Cs == 88) return launch_fwd_cs<T, MODE, V, 88>(maps, p, t, n_staged, st);
    return launch_fwd_cs<T, MODE, V, 0>(maps, p, t, n_staged, st);
}
template <typename T, int MODE>
int launch_fwd_mode(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, int n_staged, cudaStream_t st) {
    return launch_fwd_v<T, MODE, 4>(maps, p, t, n_staged, st);
}

template <typename T, int MODE, int CS>
int launch_bwd_cs(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, float* gin, cudaStream_t st) {
    ScanParams q = p;
    const size_t smem = smem_bytes(t, 5, true);
    auto kfn = scan_bwd_kernel<T, MODE, CS>;
    if (int e = prepare_kernel(kfn, smem)) return e;
    const int threads = t.n_s * (t.Cs / 4);
    q.n_scan = MODE == MODE_FUSED ? scanner_ctas(p, threads) : 0;
    kfn<<<(unsigned)(p.nchains * p.nchunks + q.n_scan), threads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], q, gin);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
template <typename T, int CS>
int launch_fwd_agg_cs(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, cudaStream_t st) {
    const size_t tile_bytes = (size_t)t.T * t.Cs * t.esize;
    const size_t pitch = (tile_bytes + 127) / 128 * 128;
    const size_t smem = (size_t)AGG_G * pitch + (size_t)sdel_floats(AGG_G * t.T, t.Cs / 16 + 2) * 4 + (size_t)2 * t.n_s * t.Cs * 4 + 8 + 32;
    auto kfn = scan_fwd_agg_kernel<T, CS>;
    if (int e = prepare_kernel(kfn, smem)) return e;
    const int ngrp = (p.nchunks + AGG_G - 1) / AGG_G;
    kfn<<<(unsigned)(p.nchains * ngrp), t.n_s * (t.Cs / 4), smem, st>>>(maps[1], p);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
template <typename T>
int launch_fwd_agg(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, cudaStream_t st) {
    if (t.Cs == 64) return launch_fwd_agg_cs<T, 64>(maps, p, t, st);
    if (t.Cs == 88) return launch_fwd_agg_cs<T, 88>(maps, p, t, st);
    return launch_fwd_agg_cs<T, 0>(maps, p, t, st);
}

template <typename T, int MODE>
int launch_bwd_mode(const CUtensorMap* maps, const ScanParams& p, const ScanTiling& t, float* gin, cudaStream_t st) {
    if (t.Cs == 64) return launch_bwd_cs<T, MODE, 64>(maps, p, t, gin, st);
    if (t.Cs == 88) return launch_bwd_cs<T, MODE, 88>(maps, p, t, gin, st);
    return launch_bwd_cs<T, MODE, 0>(maps, p, t, gin, st);
}

int check_common(int B, int L, int Di, int H, int dtype, ScanTiling& t) {
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "selective_scan: bad dtype %d", dtype);
    AB_REQUIRE(B > 0 && L > 0 && Di > 0 && H > 0 && Di == H * 16, "selective_scan: need Di == 16*H (ssm_d_state 16); got Di=%d H=%d", Di, H);
    AB_REQUIRE(make_tiling(L, Di, dtype, t), "selective_scan: no tiling for Di=%d", Di);
    return AB_OK;
}

}  // namespace

// pipelined persistent schedule (ssm_scan_pipe.cu)
bool ab_scan_pipe_plan(int B, int L, int Di, int dtype, int* tile_rows, int* slab, int* n_states, size_t* ws_bytes,
                       int* bwd_tile_rows, int* bwd_tiles_per_chain);
int ab_scan_pipe_fwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride, const void* z,
                     int64_t z_stride, const float* A_log, const float* D, const float* h0, void* y, float* h_last,
                     float* hrun, void* ws, size_t ws_bytes, int B, int L, int Di, int H, int dtype, cudaStream_t stream);
int ab_scan_pipe_bwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride, const void* z,
                     int64_t z_stride, const void* dout, const float* A_log, const float* D, const float* hrun, void* dxa,
                     void* dBm, void* dCm, int64_t dbc_stride, void* dz, float* ddlog, float** part_out, void* ws,
                     size_t ws_bytes, int B, int L, int Di, int H, int dtype, cudaStream_t stream);

extern "C" int ab_selective_scan_plan(int B, int L, int Di, int dtype, int* mode, int* tile_rows, int* slab, int* n_chunks,
                                      size_t* ws_bytes) {
    ScanTiling t;
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "selective_scan_plan: bad dtype");
    AB_REQUIRE(mode && (*mode == AB_SCAN_SINGLE_PASS || *mode == AB_SCAN_TWO_PASS || *mode == AB_SCAN_PIPELINED), "selective_scan_plan: bad mode");
    if (*mode == AB_SCAN_PIPELINED) {
        if (B > 0 && L > 0 && Di > 0 && Di % 16 == 0 && ab_scan_pipe_plan(B, L, Di, dtype, tile_rows, slab, n_chunks, ws_bytes, nullptr, nullptr)) return AB_OK;
        *mode = AB_SCAN_SINGLE_PASS;          // shapes the pipelined schedule does not cover
    }
    AB_REQUIRE(B > 0 && L > 0 && Di > 0 && make_tiling(L, Di, dtype, t), "selective_scan_plan: no tiling for Di=%d", Di);
    if (tile_rows) *tile_rows = t.T;
    if (slab) *slab = t.Cs;
    if (n_chunks) *n_chunks = t.nchunks;
    if (ws_bytes) *ws_bytes = ws_layout(t, B, Di).total;
    return AB_OK;
}

extern "C" int ab_selective_scan_fwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride,
                                     const void* z, int64_t z_stride, const float* A_log, const float* D, const float* h0,
                                     void* y, void* y_ssm, float* h_last, float* hstart, void* ws, size_t ws_bytes,
                                     uint32_t epoch, int mode, int B, int L, int Di, int H, int dtype, cudaStream_t stream) {
    ScanTiling t;
    if (int e = check_common(B, L, Di, H, dtype, t)) return e;
    if (mode == AB_SCAN_PIPELINED) {
        AB_REQUIRE(y_ssm == nullptr, "selective_scan_fwd: the pipelined mode does not emit y_ssm (plan another mode)");
        return ab_scan_pipe_fwd(xa, dlog, Bm, Cm, bc_stride, z, z_stride, A_log, D, h0, y, h_last, hstart, ws, ws_bytes, B, L, Di, H,
                                dtype, stream);
    }
    const WsLayout wl = ws_layout(t, B, Di);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "selective_scan_fwd: workspace too small (%zu < %zu)", ws_bytes, wl.total);
    AB_REQUIRE(mode == AB_SCAN_SINGLE_PASS || mode == AB_SCAN_TWO_PASS, "selective_scan_fwd: bad mode %d", mode);
    AB_REQUIRE(mode == AB_SCAN_SINGLE_PASS ? (epoch > 0 && epoch < (1u << 30)) : hstart != nullptr,
               "selective_scan_fwd: single-pass needs 0 < epoch < 2^30; two-pass needs hstart");
    CUtensorMap maps[4];
    if (int e = make_map3(&maps[0], xa, dtype, B, L, Di, Di, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[1], Bm, dtype, B, L, Di, bc_stride, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[2], Cm, dtype, B, L, Di, bc_stride, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[3], z, dtype, B, L, Di, z_stride, t.Cs, t.T)) return e;
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.L = L; p.Di = Di; p.H = H;
    p.Cs = t.Cs; p.T = t.T; p.n_s = t.n_s; p.nslab = t.nslab; p.nchunks = t.nchunks; p.nchains = B * t.nslab;
    p.dlog = dlog; p.A_log = A_log; p.Dp = D; p.h0 = h0; p.y = y; p.y_ssm = y_ssm; p.h_last = h_last; p.hstart = hstart;
    unsigned char* w8 = (unsigned char*)ws;
    p.ticket = (unsigned int*)(w8 + wl.off_ticket);
    p.err_flag = (unsigned int*)(w8 + wl.off_err);
    p.words = (unsigned long long*)(w8 + wl.off_words);
    p.inclw = (unsigned long long*)(w8 + wl.off_incl);
    p.aggP = (float*)(w8 + wl.off_aggP);
    p.aggS = (float*)(w8 + wl.off_aggS);
    p.epoch = epoch;
    const bool f32 = dtype == AB_F32;
    if (mode == AB_SCAN_SINGLE_PASS && !single_pass_ok(B, t, t.n_s * (t.Cs / t.V_f))) {
        AB_REQUIRE(hstart != nullptr, "selective_scan_fwd: this many chains run two-pass, which needs hstart");
        mode = AB_SCAN_TWO_PASS;
    }
    if (mode == AB_SCAN_SINGLE_PASS) {
        AB_CHECK_CUDA(cudaMemsetAsync(p.ticket, 0, sizeof(unsigned int), stream));
        return f32 ? launch_fwd_mode<float, MODE_FUSED>(maps, p, t, 4, stream)
                   : launch_fwd_mode<__nv_bfloat16, MODE_FUSED>(maps, p, t, 4, stream);
    }
    if (int e = f32 ? launch_fwd_agg<float>(maps, p, t, stream) : launch_fwd_agg<__nv_bfloat16>(maps, p, t, stream)) return e;
    scan_combine_kernel<<<(unsigned)ab_ceil_div((int64_t)B * Di, COMB_CH), COMB_CH * COMB_SEG, 0, stream>>>(p.aggP, p.aggS, h0, hstart, h_last, B, Di,
                                                                                   t.Cs, t.nslab, t.nchunks, 0);
    AB_LAUNCH_CHECK();
    p.h_last = nullptr;
    return f32 ? launch_fwd_mode<float, MODE_APPLY>(maps, p, t, 4, stream)
               : launch_fwd_mode<__nv_bfloat16, MODE_APPLY>(maps, p, t, 4, stream);
}

extern "C" int ab_selective_scan_bwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride,
                                     const void* z, int64_t z_stride, const void* dout, const void* dyssm,
                                     const float* A_log, const float* D, const float* hstart, void* dxa, void* dBm,
                                     void* dCm, int64_t dbc_stride, void* dz, float* ddlog_parts, float* dA_log, float* dD,
                                     void* ws, size_t ws_bytes, uint32_t epoch, int mode, int B, int L, int Di, int H,
                                     int dtype, cudaStream_t stream) {
    ScanTiling t;
    if (int e = check_common(B, L, Di, H, dtype, t)) return e;
    if (mode == AB_SCAN_PIPELINED) {
        AB_REQUIRE(dyssm == nullptr && hstart != nullptr, "selective_scan_bwd: the pipelined mode takes no dyssm and needs the run states");
        float* part = nullptr;
        if (int e = ab_scan_pipe_bwd(xa, dlog, Bm, Cm, bc_stride, z, z_stride, dout, A_log, D, hstart, dxa, dBm, dCm, dbc_stride, dz,
                                     ddlog_parts, &part, ws, ws_bytes, B, L, Di, H, dtype, stream)) return e;
        int T2 = 0, Cs2 = 0, ntile2 = 0;
        ab_scan_pipe_plan(B, L, Di, dtype, nullptr, &Cs2, nullptr, nullptr, &T2, &ntile2);
        scan_param_reduce_kernel<<<(unsigned)ab_ceil_div(Di, PR_CH), PR_CH * PR_Q, 0, stream>>>(part, dA_log, dD, B, Di, Cs2, Di / Cs2, ntile2);
        AB_LAUNCH_CHECK();
        return AB_OK;
    }
    const WsLayout wl = ws_layout(t, B, Di);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "selective_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, wl.total);
    AB_REQUIRE(hstart != nullptr, "selective_scan_bwd: hstart (saved by the forward) is required");
    AB_REQUIRE(mode == AB_SCAN_SINGLE_PASS || mode == AB_SCAN_TWO_PASS, "selective_scan_bwd: bad mode %d", mode);
    AB_REQUIRE(mode != AB_SCAN_SINGLE_PASS || (epoch > 0 && epoch < (1u << 30)), "selective_scan_bwd: need 0 < epoch < 2^30");
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE((dbc_stride * es) % 8 == 0, "selective_scan_bwd: dB/dC row stride must be 8-byte aligned");
    CUtensorMap maps[5];
    if (int e = make_map3(&maps[0], xa, dtype, B, L, Di, Di, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[1], Bm, dtype, B, L, Di, bc_stride, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[2], Cm, dtype, B, L, Di, bc_stride, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[3], z, dtype, B, L, Di, z_stride, t.Cs, t.T)) return e;
    if (int e = make_map3(&maps[4], dout, dtype, B, L, Di, Di, t.Cs, t.T)) return e;
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.L = L; p.Di = Di; p.H = H;
    p.Cs = t.Cs; p.T = t.T; p.n_s = t.n_s; p.nslab = t.nslab; p.nchunks = t.nchunks; p.nchains = B * t.nslab;
    p.dlog = dlog; p.A_log = A_log; p.Dp = D; p.hstart = const_cast<float*>(hstart);
    p.dyssm = dyssm; p.dxa = dxa; p.dBm = dBm; p.dCm = dCm; p.dz = dz; p.dbc_stride = dbc_stride; p.ddlog_parts = ddlog_parts;
    unsigned char* w8 = (unsigned char*)ws;
    p.ticket = (unsigned int*)(w8 + wl.off_ticket);
    p.err_flag = (unsigned int*)(w8 + wl.off_err);
    p.words = (unsigned long long*)(w8 + wl.off_words);
    p.inclw = (unsigned long long*)(w8 + wl.off_incl);
    p.aggP = (float*)(w8 + wl.off_aggP);
    p.aggS = (float*)(w8 + wl.off_aggS);
    p.part = (float*)(w8 + wl.off_part);
    float* gin = (float*)(w8 + wl.off_gin);
    p.epoch = epoch;
    const bool f32 = dtype == AB_F32;
    if (mode == AB_SCAN_SINGLE_PASS && !single_pass_ok(B, t, t.n_s * (t.Cs / 4))) mode = AB_SCAN_TWO_PASS;
    if (mode == AB_SCAN_SINGLE_PASS) {
        AB_CHECK_CUDA(cudaMemsetAsync(p.ticket, 0, sizeof(unsigned int), stream));
        if (int e = f32 ? launch_bwd_mode<float, MODE_FUSED>(maps, p, t, gin, stream)
                        : launch_bwd_mode<__nv_bfloat16, MODE_FUSED>(maps, p, t, gin, stream)) return e;
    } else {
        if (int e = f32 ? launch_bwd_mode<float, MODE_AGG>(maps, p, t, gin, stream)
                        : launch_bwd_mode<__nv_bfloat16, MODE_AGG>(maps, p, t, gin, stream)) return e;
        scan_combine_kernel<<<(unsigned)ab_ceil_div((int64_t)B * Di, COMB_CH), COMB_CH * COMB_SEG, 0, stream>>>(p.aggP, p.aggS, nullptr, gin, nullptr, B,
                                                                                       Di, t.Cs, t.nslab, t.nchunks, 1);
        AB_LAUNCH_CHECK();
        if (int e = f32 ? launch_bwd_mode<float, MODE_APPLY>(maps, p, t, gin, stream)
                        : launch_bwd_mode<__nv_bfloat16, MODE_APPLY>(maps, p, t, gin, stream)) return e;
    }
    scan_param_reduce_kernel<<<(unsigned)ab_ceil_div(Di, PR_CH), PR_CH * PR_Q, 0, stream>>>(p.part, dA_log, dD, B, Di, t.Cs, t.nslab, t.nchunks);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

#ifdef AB_SCAN_TRACE
extern "C" int ab_scan_trace_dump(unsigned long long* host, int ntiles) {
    if (ntiles > TRACE_TILES) ntiles = TRACE_TILES;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host, g_scan_trace, (size_t)ntiles * TRACE_SLOTS * sizeof(unsigned long long));
}
extern "C" int ab_scanner_trace_dump(unsigned long long* host, int clear) {
    cudaDeviceSynchronize();
    int rc = (int)cudaMemcpyFromSymbol(host, g_scanner_trace, sizeof(unsigned long long) * 64 * STRACE_ROUNDS * 4);
    if (clear) { void* d; cudaGetSymbolAddress(&d, g_scanner_trace); cudaMemset(d, 0, sizeof(unsigned long long) * 64 * STRACE_ROUNDS * 4); }
    return rc;
}
#endif
