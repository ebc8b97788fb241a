// Pipelined persistent selective scan, forward and backward (sm_100a): AB_SCAN_PIPELINED.
//
// Same math and the same scanner hand-shake as ssm_scan.cu (core.py:324-353, :383, :394-396), different schedule.
// The one-tile-per-CTA kernels spend half of a tile's life waiting for the state that enters it.  Here a CTA is
// persistent and takes SUPER-TILES (SG = 4 consecutive tiles of one chain) from an atomic ticket, in scan order with
// the chains interleaved.  Every iteration it runs two passes over two different tiles, one __syncthreads per iteration:
//
//   main pass, tile i:       the state entering the tile is already known - for the first tile of a super-tile it is
//            the scanner's word (requested at the top of the iteration), for the others the state the previous tile's
//            last run left in shared memory - so the tile is streamed exactly once.  Forward: h = abar*h + B,
//            y = (C*h + D*x) * silu(z).  Backward: forward recompute from the run state saved by the forward (dxa, dC,
//            dz; no intra-tile dependency), then the reverse sweep (dB, d dt, dA_log / dD partials) with (abar, g,
//            h_{t-1}) of the thread's 4 tokens in registers.
//   prepass, tile i + PD:    run aggregates only, from Bm and delta (forward) or C, z, dout and delta (backward).
//   barrier.                 The stage the main pass used and one prepass slot are free: thread 0 issues the TMA loads
//            of tile i + 2 (main ring, 2 stages) and tile i + PD + 2 (prepass ring, 3 slots) into them.
//   duty warp (one per 64-channel column, rotating): composes the prepass's run aggregates in run order, leaves the
//            per-run coefficients (state entering the run = P * incoming + S) in shared memory, folds the tile into
//            the aggregate of its super-tile; the super-tile's last tile publishes it (self-validating 64-bit words).
//   every warp picks up the coefficients of tile i + PD - 1 into a small register queue: they are used PD - 1
//            iterations later by the main pass.
//
// The prepass reads its operands a second time (L2 eviction hints keep about half of those reads out of DRAM); in
// exchange nothing state-dependent has to be carried between the passes, the scanner's latency is covered by
// PD - (SG - 1) tiles of work and a super-tile costs one hand-shake.
//
// A thread owns 2 adjacent channels x TSP = 4 consecutive tokens (a "run"); a warp is one run of a 64-channel column,
// so every shared-memory access is a conflict-free 128-byte row and every global store a full 128-byte line.  Channel
// pairs are processed with packed f32x2 arithmetic (FFMA2 / FMUL2).  delta = softplus(dt) is computed once by a small
// kernel into the saved-state buffer and arrives by TMA with the operands.  d dt is reduced over the 16 channels of a
// head with a transposing shuffle butterfly inside the warp and written once, final, as [B, L, H].
//
// Deadlock freedom: a CTA's tiles increase with its pipeline position and an aggregate is published before the CTA
// waits on any smaller tile, so by induction over the tile index every wait is on words whose producers are running;
// the scanners hold the first tickets.  Every spin is bounded (error flag, no hang).  The launch epoch lives in device
// memory and is advanced by the last CTA to finish, which also resets the ticket: replayed CUDA graphs stay valid.
// tools/scan_trace_pipe.py (with `make TRACE=1`) prints the per-phase timeline of the forward kernel.
#include "ssm_scan_shared.cuh"

namespace {
using namespace ab_scan;

constexpr int TSP = 4;             // tokens per run
constexpr int PIPE_SCAN_K = 32;    // tile aggregates a scanner polls per round
constexpr int INFO_RING = 16;
constexpr int PIPE_SPIN_LIMIT = 1 << 19;   // ~0.5 s of polling: a protocol error raises the flag and traps (a CUDA error, not a hang,
                                           // and never a silently wrong result)

// sigmoid of a channel pair.  bf16 activations: single-MUFU tanh form (rel. error ~5e-4, below bf16 resolution);
// f32 activations: ex2 + rcp
template <typename T>
__device__ __forceinline__ f2 f2_sigmoid(f2 z) {
    if constexpr (sizeof(T) == 2) {
        float a, b;
        f2_unpack(f2_mul(z, f2_bcast(0.5f)), a, b);
        float ta, tb;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
        return f2_fma(f2_pack(ta, tb), f2_bcast(0.5f), f2_bcast(0.5f));
    } else {
        float a, b;
        f2_unpack(z, a, b);
        return f2_pack(ab_sigmoid(a), ab_sigmoid(b));
    }
}

// two adjacent channels from shared memory / to global memory
template <typename T>
__device__ __forceinline__ f2 lds_pair(const T* p) {
    if constexpr (sizeof(T) == 2) {
        const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
        return f2_pack(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
    } else {
        return *reinterpret_cast<const f2*>(p);
    }
}
template <typename T>
__device__ __forceinline__ void stg_pair(T* p, f2 v) {
    if constexpr (sizeof(T) == 2) {
        float a, b;
        f2_unpack(v, a, b);
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h);
    } else {
        *reinterpret_cast<f2*>(p) = v;
    }
}

__device__ __forceinline__ void ld_relaxed_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// L2 residency control.  The prepass reads a tile's operands PD tiles before the main pass reads them again: those loads
// ask L2 to keep the lines (evict_last) and everything that is touched once (the main pass's loads, the outputs) asks to
// be evicted first, so that the second read is served by L2 instead of DRAM.
__device__ __forceinline__ uint64_t l2_policy_keep() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_stream() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(ab_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ab_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

struct __align__(16) TileInfo { int b, c0, row0, tile_lin; };      // b < 0: no tile; tile_lin = super-tile * SG + position inside it

#ifndef PIPE_PD
#define PIPE_PD 8
#endif
#ifndef PIPE_PD_BWD
#define PIPE_PD_BWD 8
#endif
#ifndef PIPE_LOG_SG
#define PIPE_LOG_SG 2
#endif
#ifndef PIPE_NPRE
#define PIPE_NPRE 3
#endif
constexpr int PD_FWD = PIPE_PD;     // the prepass runs PD tiles ahead of the main pass.  A super-tile's aggregate is published with the
constexpr int PD_BWD = PIPE_PD_BWD; // prepass of its LAST tile, so PD - (SG - 1) tiles of work hide the scanner's latency: keep PD > SG
constexpr int NPRE = PIPE_NPRE;    // prepass ring slots: its loads are issued NPRE - 1 tiles ahead of the prepass
constexpr int LOG_SG = PIPE_LOG_SG, SG = 1 << LOG_SG;   // consecutive tiles of a chain one CTA takes per ticket (a super-tile):
                                                      // one hand-shake with the scanner per SG tiles
constexpr int NMAIN = 2;           // main ring stages

struct PipeParams {
    ScanParams s;
    unsigned int* sync;          // [0] ticket, [1] finished CTAs, [2] launch epoch - 1
    float* hrun;                 // per batch: [ceil(L/TSP), Di] state entering every run (written by the forward), then delta
    size_t batch_stride;         // floats between the batches of hrun
    float* ddlog;                // backward: [B, L, H] fp32, final
    int ntiles;
    int scan_k, scan_r;          // tile aggregates a scanner polls per round, replicas (segments) it splits them into
    unsigned int poll_ns;        // back-off between polls of the incoming-state word
};

__device__ __forceinline__ TileInfo decode_tile_info(const PipeParams& p, int ticket, int TT, bool reverse) {
    TileInfo t;
    if (ticket < 0 || ticket >= p.ntiles) { t.b = -1; t.c0 = 0; t.row0 = 0; t.tile_lin = 0; return t; }
    const int chain = ticket % p.s.nchains, jj = ticket / p.s.nchains;
    const int js = reverse ? p.s.nchunks - 1 - jj : jj;                 // p.s.nchunks counts super-tiles (the scanner's units)
    t.b = chain / p.s.nslab;
    t.c0 = (chain % p.s.nslab) * p.s.Cs;
    t.row0 = (js * SG + (reverse ? SG - 1 : 0)) * TT;                   // rows beyond L are zero-filled by TMA: identity tiles
    t.tile_lin = (chain * p.s.nchunks + js) * SG;
    return t;
}
// thread 0: tile of pipeline position q (a new ticket every SG positions, else the successor of position q - 1)
__device__ __forceinline__ TileInfo next_tile_info(const PipeParams& p, const TileInfo* s_info, int q, int& pending, int TT, bool reverse) {
    TileInfo t;
    if ((q & (SG - 1)) == 0) {
        t = decode_tile_info(p, pending, TT, reverse);
        if (pending < p.ntiles) pending = (int)atomicAdd(p.sync, 1u) - p.s.n_scan;
    } else {
        t = s_info[(q - 1) & (INFO_RING - 1)];
        if (t.b >= 0) { t.row0 += reverse ? -TT : TT; t.tile_lin += 1; }
    }
    return t;
}

// shared-memory carve-up, identical on host and device.  A main stage holds `nmain` operand tiles and the delta tile,
// a prepass slot `npre` operand tiles and the delta tile.
struct PipeSmem {
    uint32_t pitch, dpitch, nhp, main_stride, pre_stride, off_pre, off_runP, off_runS, off_partA, off_partD, off_carry, off_acc, off_info, off_bars, total;
};
__host__ __device__ inline PipeSmem pipe_smem(int Cs, int TT, int NR, int nmain, int npre, int esize, bool bwd) {
    PipeSmem m;
    const uint32_t tile_bytes = (uint32_t)TT * Cs * esize;
    m.pitch = (tile_bytes + 127u) & ~127u;
    m.nhp = (((uint32_t)Cs >> 4) + 3u) & ~3u;                            // heads per delta row, whole 16-byte units
    m.dpitch = ((uint32_t)(TT + 1) * m.nhp * 4u + 127u) & ~127u;
    m.main_stride = nmain * m.pitch + m.dpitch;
    m.pre_stride = npre * m.pitch + m.dpitch;
    uint32_t o = NMAIN * m.main_stride;
    m.off_pre = o; o += NPRE * m.pre_stride;
    m.off_runP = o; o += 2u * NR * Cs * 4;
    m.off_runS = o; o += 2u * NR * Cs * 4;
    m.off_partA = o; if (bwd) o += 2u * NR * Cs * 4;
    m.off_partD = o; if (bwd) o += 2u * NR * Cs * 4;
    m.off_carry = o; o += 2u * Cs * 4;        // state leaving a tile -> next tile of the super-tile (two buffers)
    m.off_acc = o; o += 2u * Cs * 4;          // aggregate of the super-tile so far (P, S)
    o = (o + 15u) & ~15u;
    m.off_info = o; o += INFO_RING * 16;
    m.off_bars = o; o += (NMAIN + NPRE) * 8;
    m.total = (o + 127u) & ~127u;
    return m;
}

// last CTA out advances the device epoch and resets the ticket / finished counters for the next launch
__device__ __forceinline__ void grid_finish(unsigned int* sync) {
    __threadfence();
    const unsigned int done = atomicAdd(sync + 1, 1u);
    if (done == gridDim.x - 1) {
        const unsigned int e = *reinterpret_cast<volatile unsigned int*>(sync + 2);
        sync[0] = 0; sync[1] = 0;
        sync[2] = (e + 1u) % ((1u << 30) - 2u);
        __threadfence();
    }
}
__device__ __forceinline__ void cta_finish(unsigned int* sync) {
    __syncthreads();
    if (threadIdx.x == 0) grid_finish(sync);
}
// The scanner's threads leave their loop at different times (a channel is done when its last tile has published), so
// lanes of one warp may arrive here while their siblings still execute the loop's barriers: no __syncthreads on this
// path (two different aligned barriers inside one diverged warp are undefined behaviour); the last thread to arrive,
// counted in shared memory, signs the CTA off.
__device__ __forceinline__ void scanner_finish(unsigned int* sync, unsigned int* s_left, int participants) {
    __threadfence();
    if (atomicAdd(s_left, 1u) == (unsigned)participants - 1u) grid_finish(sync);
}

// delta = softplus(dt logits) as fp32 rows of Hp (whole 16-byte units, TMA-loadable) heads; kept for the backward
template <typename T>
__global__ void __launch_bounds__(256) scan_delta_kernel(const T* __restrict__ dlog, float* __restrict__ delta, int L, int H, int Hp,
                                                         size_t batch_stride, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % Hp);
        const size_t tok = i / Hp;
        const size_t b = tok / L, t = tok % L;
        delta[b * batch_stride + t * Hp + h] = h < H ? ab_softplus_fast(ab_to_float(dlog[tok * H + h])) : 0.f;
    }
}

#ifdef AB_SCAN_TRACE
// debug build only (make TRACE=1): per-warp phase timestamps of the first worker CTAs, read back by tools/scan_trace_pipe.py
constexpr int PT_CTAS = 8, PT_ITERS = 48, PT_WARPS = 16, PT_SLOTS = 8;
__device__ unsigned long long g_pipe_trace[PT_CTAS * PT_ITERS * PT_WARPS * PT_SLOTS];
__device__ unsigned int g_pipe_trace_cta;
__device__ __forceinline__ void pt_mark(int cta, int it, int slot) {
    if ((threadIdx.x & 31) == 0 && cta >= 0 && cta < PT_CTAS && it >= 0 && it < PT_ITERS && (threadIdx.x >> 5) < PT_WARPS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_pipe_trace[(((size_t)cta * PT_ITERS + it) * PT_WARPS + (threadIdx.x >> 5)) * PT_SLOTS + slot] = t;
    }
}
#define PT_MARK(slot) pt_mark(pt_cta, i + PD, slot)
#else
#define PT_MARK(slot)
#endif

// duty warp: compose the run aggregates of one tile in run order (forward) or reverse run order (backward), leave the
// per-run coefficients (state entering the run = P * incoming + S) in their place and publish the tile aggregate
template <int NR, bool REVERSE>
__device__ __forceinline__ void compose_and_publish(float* rp, float* rs, int Cs, float* acc, int sub, unsigned long long* w, uint32_t epoch) {
    f2 Pa = f2_bcast(1.f), Sa = f2_bcast(0.f);
#pragma unroll
    for (int q0 = 0; q0 < NR; q0 += 4) {
        f2 P4[4], S4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (q0 + u < NR) {
                const int r = REVERSE ? NR - 1 - (q0 + u) : q0 + u;
                P4[u] = *reinterpret_cast<const f2*>(rp + r * Cs); S4[u] = *reinterpret_cast<const f2*>(rs + r * Cs);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (q0 + u < NR) {
                const int r = REVERSE ? NR - 1 - (q0 + u) : q0 + u;
                *reinterpret_cast<f2*>(rp + r * Cs) = Pa;
                *reinterpret_cast<f2*>(rs + r * Cs) = Sa;
                Sa = f2_fma(P4[u], Sa, S4[u]);
                Pa = f2_mul(Pa, P4[u]);
            }
    }
    // fold the tile into the aggregate of its super-tile; the last tile publishes it
    if (sub > 0) {
        const f2 aP = *reinterpret_cast<const f2*>(acc), aS = *reinterpret_cast<const f2*>(acc + Cs);
        Sa = f2_fma(Pa, aS, Sa);
        Pa = f2_mul(aP, Pa);
    }
    if (sub < SG - 1) {
        *reinterpret_cast<f2*>(acc) = Pa;
        *reinterpret_cast<f2*>(acc + Cs) = Sa;
        return;
    }
    float p0, p1, s0, s1;
    f2_unpack(Pa, p0, p1); f2_unpack(Sa, s0, s1);
    st_relaxed_v2(w, pack_word(epoch, ST_AGG, p0), pack_word(epoch, ST_AGG, s0));
    st_relaxed_v2(w + 2, pack_word(epoch, ST_AGG, p1), pack_word(epoch, ST_AGG, s1));
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int NWC, int NR, int CS>
__global__ void __launch_bounds__(32 * NWC * NR, (1024 / (32 * NWC * NR) > 0 ? 1024 / (32 * NWC * NR) : 1))
scan_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tm_xa, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_z,
                     const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ PipeParams p) {
    constexpr int TT = NR * TSP, PD = PD_FWD, LA = PD + NPRE - 1;       // LA: ticket look-ahead
    extern __shared__ __align__(128) unsigned char smem[];
    const ScanParams& sp = p.s;
    const int Cs = CS ? CS : sp.Cs;          // CS != 0: slab width known at compile time
    const PipeSmem lay = pipe_smem(Cs, TT, NR, 4, 1, (int)sizeof(T), false);
    const int nhp = (int)lay.nhp;
    const uint32_t tile_bytes = (uint32_t)TT * Cs * sizeof(T), dbytes = (uint32_t)TT * nhp * 4u;
    float* runP = reinterpret_cast<float*>(smem + lay.off_runP);
    float* runS = reinterpret_cast<float*>(smem + lay.off_runS);
    float* carry = reinterpret_cast<float*>(smem + lay.off_carry);
    float* accb = reinterpret_cast<float*>(smem + lay.off_acc);
    TileInfo* s_info = reinterpret_cast<TileInfo*>(smem + lay.off_info);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* pbar = mbar + NMAIN;
    __shared__ unsigned int s_first, s_epoch, s_left;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wc = warp % NWC, run = warp / NWC;
    int cl = wc * 64 + 2 * lane;
    const bool act = cl < Cs;
    if (!act) cl = Cs - 2;
    const int hh = cl >> 4;

    if (tid == 0) {
        s_epoch = *reinterpret_cast<volatile unsigned int*>(p.sync + 2) + 1u;
        s_first = atomicAdd(p.sync, 1u);
        s_left = 0;
        for (int s = 0; s < NMAIN + NPRE; ++s) ab_mbar_init(&mbar[s], 1);
        ab_fence_mbar_init();
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if ((int)s_first < sp.n_scan) {
        float* hs = runS + 2 * NR * Cs - Cs;          // generic scanner path only; the ring may use everything below
        const int np = scanner_role<+1>(sp, epoch, p.scan_k, p.scan_r, (int)s_first, hs, reinterpret_cast<uint4*>(smem), (size_t)(reinterpret_cast<unsigned char*>(hs) - smem));
        if (np) scanner_finish(p.sync, &s_left, np);
        return;
    }

    auto issue_main = [&](int s, const TileInfo& ti) {         // thread 0 only
        ab_mbar_expect_tx(&mbar[s], 4u * tile_bytes + dbytes);
        unsigned char* dst = smem + (size_t)s * lay.main_stride;
        const uint64_t pol = l2_policy_stream();                // last use of every line
        tma_load_3d_hint(dst, &tm_b, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + lay.pitch, &tm_xa, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 2 * lay.pitch, &tm_c, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 3 * lay.pitch, &tm_z, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 4 * lay.pitch, &tm_d, &mbar[s], ti.c0 >> 4, ti.row0, ti.b, pol);
    };
    auto issue_pre = [&](int s, const TileInfo& ti) {
        ab_mbar_expect_tx(&pbar[s], tile_bytes + dbytes);
        unsigned char* dst = smem + lay.off_pre + (size_t)s * lay.pre_stride;
        const uint64_t pol = l2_policy_keep();                  // read again by the main pass
        tma_load_3d_hint(dst, &tm_b, &pbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + lay.pitch, &tm_d, &pbar[s], ti.c0 >> 4, ti.row0, ti.b, pol);
    };

    // ---- prologue: the first two pipeline positions
    int pending = -1;
    if (tid == 0) {
        pending = (int)s_first - sp.n_scan;
        for (int q = 0; q < NPRE - 1; ++q) {
            s_info[q] = next_tile_info(p, s_info, q, pending, TT, false);
            if (s_info[q].b >= 0) issue_pre(q, s_info[q]);
        }
    }
    __syncthreads();

    const int nrt = (sp.L + TSP - 1) / TSP;   // saved run states per sequence
    uint32_t mphase = 0, pphase = 0;          // mbarrier parities per main stage / prepass slot
    f2 cqP[PD - 1], cqS[PD - 1];          // run coefficients of the next PD - 1 main tiles
#pragma unroll
    for (int q = 0; q < PD - 1; ++q) { cqP[q] = f2_bcast(1.f); cqS[q] = f2_bcast(0.f); }
    int ps = 0;                               // prepass slot of position i + PD

#ifdef AB_SCAN_TRACE
    __shared__ int s_pt_cta;
    if (tid == 0) s_pt_cta = (int)atomicAdd(&g_pipe_trace_cta, 1u);
    __syncthreads();
    const int pt_cta = s_pt_cta;
#endif
    for (int i = -PD;; ++i) {
        const TileInfo mi = s_info[(i < 0 ? 0 : i) & (INFO_RING - 1)];
        if (mi.b < 0) break;                  // positions are handed out in increasing order: nothing left for this CTA
        PT_MARK(0);
        const TileInfo pi = s_info[(i + PD) & (INFO_RING - 1)];
        if (tid == 0) {
            s_info[(i + LA) & (INFO_RING - 1)] = next_tile_info(p, s_info, i + LA, pending, TT, false);
        }
        // per-channel parameters of both tiles: requested here, consumed behind the waits (their latency stays hidden)
        float2 m_al = make_float2(0.f, 0.f), m_dv = m_al, p_al = m_al;
        if (i >= 0) {
            m_al = __ldg(reinterpret_cast<const float2*>(sp.A_log + mi.c0 + cl));
            m_dv = __ldg(reinterpret_cast<const float2*>(sp.Dp + mi.c0 + cl));
        }
        if (pi.b >= 0) p_al = __ldg(reinterpret_cast<const float2*>(sp.A_log + pi.c0 + cl));
        // ---- main pass of tile i: the state entering it was requested PD iterations ago
        if (i >= 0) {
            const int s = i & 1;
            const bool first = (mi.tile_lin & (SG - 1)) == 0;          // first tile of a super-tile: state from the scanner
            unsigned long long w0 = 0, w1 = 0;
            const unsigned long long* wp = sp.inclw + (size_t)(mi.tile_lin >> LOG_SG) * Cs + cl;
            if (first) ld_relaxed_v2(wp, w0, w1);
            const unsigned char* st = smem + (size_t)s * lay.main_stride;
            const T* sB = reinterpret_cast<const T*>(st) + cl;
            const T* sX = reinterpret_cast<const T*>(st + lay.pitch) + cl;
            const T* sC = reinterpret_cast<const T*>(st + 2 * lay.pitch) + cl;
            const T* sZ = reinterpret_cast<const T*>(st + 3 * lay.pitch) + cl;
            const float* sd = reinterpret_cast<const float*>(st + 4 * lay.pitch) + hh;
            f2 hin;
            if (first) {
                int spins = 0;
                while (!(word_valid(w0, epoch) && word_valid(w1, epoch))) {
                    if (++spins > PIPE_SPIN_LIMIT) { atomicExch(sp.err_flag, 1u); __trap(); }
                    __nanosleep(p.poll_ns);
                    ld_relaxed_v2(wp, w0, w1);
                }
                hin = f2_pack(__uint_as_float((uint32_t)w0), __uint_as_float((uint32_t)w1));
            } else {
                hin = *reinterpret_cast<const f2*>(carry + ((i - 1) & 1) * Cs + cl);       // left by the previous tile's last run
            }
            PT_MARK(1);
            f2 h = f2_fma(cqP[0], hin, cqS[0]);
            const int row = mi.row0 + run * TSP;
            const int rg = mi.row0 / TSP + run;
            if (act && rg < nrt) *reinterpret_cast<f2*>(p.hrun + (size_t)mi.b * p.batch_stride + (size_t)rg * sp.Di + mi.c0 + cl) = h;
            T* yo = reinterpret_cast<T*>(sp.y) + ((size_t)mi.b * sp.L + row) * sp.Di + mi.c0 + cl;
            ab_mbar_wait(&mbar[s], (mphase >> s) & 1u);
            mphase ^= 1u << s;
            PT_MARK(2);
            asm volatile("" : "+f"(m_al.x), "+f"(m_al.y), "+f"(m_dv.x), "+f"(m_dv.y));
            const f2 A2 = f2_pack(-__expf(m_al.x) * AB_LOG2E, -__expf(m_al.y) * AB_LOG2E), Dv = f2_pack(m_dv.x, m_dv.y);
#pragma unroll
            for (int t = 0; t < TSP; ++t) {
                const int r = run * TSP + t;
                const f2 a = f2_ex2(f2_mul(A2, f2_bcast(sd[r * nhp])));
                const f2 bv = lds_pair<T>(sB + (size_t)r * Cs), cv = lds_pair<T>(sC + (size_t)r * Cs);
                const f2 xv = lds_pair<T>(sX + (size_t)r * Cs), zv = lds_pair<T>(sZ + (size_t)r * Cs);
                h = f2_fma(a, h, bv);
                const f2 o = f2_mul(f2_fma(Dv, xv, f2_mul(cv, h)), f2_mul(zv, f2_sigmoid<T>(zv)));
                if (act && row + t < sp.L) stg_pair<T>(yo + (size_t)t * sp.Di, o);
            }
            if (run == NR - 1 && act) *reinterpret_cast<f2*>(carry + (i & 1) * Cs + cl) = h;
        }
        PT_MARK(3);
        // ---- prepass of tile i + PD: run aggregates from Bm and delta only
        if (pi.b >= 0) {
            const unsigned char* st = smem + lay.off_pre + (size_t)ps * lay.pre_stride;
            const T* sB = reinterpret_cast<const T*>(st) + cl;
            const float* sd = reinterpret_cast<const float*>(st + lay.pitch) + hh;
            ab_mbar_wait(&pbar[ps], (pphase >> ps) & 1u);
            pphase ^= 1u << ps;
            PT_MARK(4);
            asm volatile("" : "+f"(p_al.x), "+f"(p_al.y));
            const f2 A2 = f2_pack(-__expf(p_al.x) * AB_LOG2E, -__expf(p_al.y) * AB_LOG2E);
            f2 S = f2_bcast(0.f), P = f2_bcast(1.f);
#pragma unroll
            for (int t = 0; t < TSP; ++t) {
                const int r = run * TSP + t;
                const f2 a = f2_ex2(f2_mul(A2, f2_bcast(sd[r * nhp])));
                P = f2_mul(P, a);
                S = f2_fma(a, S, lds_pair<T>(sB + (size_t)r * Cs));
            }
            if (act) {
                const int bi = (((i + PD) & 1) * NR + run) * Cs + cl;
                *reinterpret_cast<f2*>(runP + bi) = P;
                *reinterpret_cast<f2*>(runS + bi) = S;
            }
        }
        PT_MARK(5);
        __syncthreads();
        PT_MARK(6);

        // ---- refill: the main stage tile i leaves goes to tile i + 2, the free prepass slot to tile i + PD + 2
        // (the serial chores after the barrier go to different warps, so that no warp starts its next main pass late by
        //  the sum of them: warp 0 refills the main ring, warp 1 the prepass ring, the duty rotates over the others)
        if (tid == 0 && i + 2 >= 0) {
            const TileInfo nm = s_info[(i + 2) & (INFO_RING - 1)];
            if (nm.b >= 0) issue_main(i & 1, nm);
        }
        if (tid == 32) {
            const TileInfo np = s_info[(i + LA) & (INFO_RING - 1)];
            if (np.b >= 0) issue_pre((ps + NPRE - 1) % NPRE, np);
        }
        // ---- duty warp of this column: run prefixes in place, tile aggregate -> scanner
        const int duty_run = NR > 2 ? 2 + (i + PD) % (NR - 2) : (i + PD) % NR;
        if (pi.b >= 0 && run == duty_run && act)
            compose_and_publish<NR, false>(runP + ((i + PD) & 1) * NR * Cs + cl, runS + ((i + PD) & 1) * NR * Cs + cl, Cs, accb + cl,
                                           pi.tile_lin & (SG - 1), sp.words + ((size_t)(pi.tile_lin >> LOG_SG) * Cs + cl) * 2, epoch);
        // ---- run coefficients of tile i + PD - 1 (composed during the previous iteration)
#pragma unroll
        for (int q = 0; q + 1 < PD - 1; ++q) { cqP[q] = cqP[q + 1]; cqS[q] = cqS[q + 1]; }
        if (i + PD - 1 >= 0) {
            const int bi = (((i + PD - 1) & 1) * NR + run) * Cs + cl;
            cqP[PD - 2] = *reinterpret_cast<const f2*>(runP + bi);
            cqS[PD - 2] = *reinterpret_cast<const f2*>(runS + bi);
        }
        ps = ps + 1 == NPRE ? 0 : ps + 1;
        PT_MARK(7);
    }
    cta_finish(p.sync);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <typename T, int NWC, int NR, int CS>
__global__ void __launch_bounds__(32 * NWC * NR, (512 / (32 * NWC * NR) > 0 ? 512 / (32 * NWC * NR) : 1))
scan_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tm_xa, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_z,
                     const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_d,
                     const __grid_constant__ PipeParams p) {
    constexpr int TT = NR * TSP, PD = PD_BWD, LA = PD + NPRE - 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const ScanParams& sp = p.s;
    const int Cs = CS ? CS : sp.Cs;
    const PipeSmem lay = pipe_smem(Cs, TT, NR, 5, 3, (int)sizeof(T), true);
    const int nhp = (int)lay.nhp;
    const uint32_t tile_bytes = (uint32_t)TT * Cs * sizeof(T), dbytes = (uint32_t)(TT + 1) * nhp * 4u;
    float* runP = reinterpret_cast<float*>(smem + lay.off_runP);
    float* runS = reinterpret_cast<float*>(smem + lay.off_runS);
    float* carry = reinterpret_cast<float*>(smem + lay.off_carry);
    float* accb = reinterpret_cast<float*>(smem + lay.off_acc);
    float* partA = reinterpret_cast<float*>(smem + lay.off_partA);
    float* partD = reinterpret_cast<float*>(smem + lay.off_partD);
    TileInfo* s_info = reinterpret_cast<TileInfo*>(smem + lay.off_info);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* pbar = mbar + NMAIN;
    __shared__ unsigned int s_first, s_epoch, s_left;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wc = warp % NWC, run = warp / NWC;
    int cl = wc * 64 + 2 * lane;
    const bool act = cl < Cs;
    if (!act) cl = Cs - 2;
    const int hh = cl >> 4;

    if (tid == 0) {
        s_epoch = *reinterpret_cast<volatile unsigned int*>(p.sync + 2) + 1u;
        s_first = atomicAdd(p.sync, 1u);
        s_left = 0;
        for (int s = 0; s < NMAIN + NPRE; ++s) ab_mbar_init(&mbar[s], 1);
        ab_fence_mbar_init();
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if ((int)s_first < sp.n_scan) {
        float* hs = partD + 2 * NR * Cs - Cs;
        const int np = scanner_role<-1>(sp, epoch, p.scan_k, p.scan_r, (int)s_first, hs, reinterpret_cast<uint4*>(smem), (size_t)(reinterpret_cast<unsigned char*>(hs) - smem));
        if (np) scanner_finish(p.sync, &s_left, np);
        return;
    }

    auto issue_main = [&](int s, const TileInfo& ti) {
        ab_mbar_expect_tx(&mbar[s], 5u * tile_bytes + dbytes);
        unsigned char* dst = smem + (size_t)s * lay.main_stride;
        const uint64_t pol = l2_policy_stream();
        tma_load_3d_hint(dst, &tm_b, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + lay.pitch, &tm_xa, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 2 * lay.pitch, &tm_c, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 3 * lay.pitch, &tm_z, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 4 * lay.pitch, &tm_do, &mbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 5 * lay.pitch, &tm_d, &mbar[s], ti.c0 >> 4, ti.row0, ti.b, pol);
    };
    auto issue_pre = [&](int s, const TileInfo& ti) {
        ab_mbar_expect_tx(&pbar[s], 3u * tile_bytes + dbytes);
        unsigned char* dst = smem + lay.off_pre + (size_t)s * lay.pre_stride;
        const uint64_t pol = l2_policy_keep();
        tma_load_3d_hint(dst, &tm_c, &pbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + lay.pitch, &tm_z, &pbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 2 * lay.pitch, &tm_do, &pbar[s], ti.c0, ti.row0, ti.b, pol);
        tma_load_3d_hint(dst + 3 * lay.pitch, &tm_d, &pbar[s], ti.c0 >> 4, ti.row0, ti.b, pol);
    };

    int pending = -1;
    if (tid == 0) {
        pending = (int)s_first - sp.n_scan;
        for (int q = 0; q < NPRE - 1; ++q) {
            s_info[q] = next_tile_info(p, s_info, q, pending, TT, true);
            if (s_info[q].b >= 0) issue_pre(q, s_info[q]);
        }
    }
    __syncthreads();

    const int nrt = (sp.L + TSP - 1) / TSP;
    uint32_t mphase = 0, pphase = 0;
    f2 cqP[PD - 1], cqS[PD - 1];
#pragma unroll
    for (int q = 0; q < PD - 1; ++q) { cqP[q] = f2_bcast(1.f); cqS[q] = f2_bcast(0.f); }
    int ps = 0;
    const float inv_log2e = 1.f / AB_LOG2E;

    // sum the per-run partials of a finished tile in run order (deterministic) -> per-tile partials in global memory
    auto reduce_parts = [&](int tile_lin, int buf) {
        const float* pa = partA + buf * NR * Cs + cl;
        const float* pd = partD + buf * NR * Cs + cl;
        f2 sa = f2_bcast(0.f), sd2 = f2_bcast(0.f);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            sa = f2_add(sa, *reinterpret_cast<const f2*>(pa + r * Cs));
            sd2 = f2_add(sd2, *reinterpret_cast<const f2*>(pd + r * Cs));
        }
        *reinterpret_cast<f2*>(sp.part + ((size_t)tile_lin * 2 + 0) * Cs + cl) = sa;
        *reinterpret_cast<f2*>(sp.part + ((size_t)tile_lin * 2 + 1) * Cs + cl) = sd2;
    };

    for (int i = -PD;; ++i) {
        const TileInfo mi = s_info[(i < 0 ? 0 : i) & (INFO_RING - 1)];
        if (mi.b < 0) break;
        const TileInfo pi = s_info[(i + PD) & (INFO_RING - 1)];
        if (tid == 0) {
            s_info[(i + LA) & (INFO_RING - 1)] = next_tile_info(p, s_info, i + LA, pending, TT, true);
        }
        float2 m_al = make_float2(0.f, 0.f), m_dv = m_al, p_al = m_al, m_h = m_al;
        if (i >= 0) {
            const int cg0 = mi.c0 + cl;
            m_al = __ldg(reinterpret_cast<const float2*>(sp.A_log + cg0));
            m_dv = __ldg(reinterpret_cast<const float2*>(sp.Dp + cg0));
            const int rg = mi.row0 / TSP + run;
            if (rg < nrt) m_h = __ldg(reinterpret_cast<const float2*>(p.hrun + (size_t)mi.b * p.batch_stride + (size_t)rg * sp.Di + cg0));
        }
        if (pi.b >= 0) p_al = __ldg(reinterpret_cast<const float2*>(sp.A_log + pi.c0 + cl));
        // ---- main pass of tile i: forward recompute from the saved run state, then the reverse sweep
        if (i >= 0) {
            const int s = i & 1;
            const bool first = (mi.tile_lin & (SG - 1)) == 0;
            unsigned long long w0 = 0, w1 = 0;
            const unsigned long long* wp = sp.inclw + (size_t)(mi.tile_lin >> LOG_SG) * Cs + cl;
            if (first) ld_relaxed_v2(wp, w0, w1);
            const int cg0 = mi.c0 + cl;
            const unsigned char* st = smem + (size_t)s * lay.main_stride;
            const T* sB = reinterpret_cast<const T*>(st) + cl;
            const T* sX = reinterpret_cast<const T*>(st + lay.pitch) + cl;
            const T* sC = reinterpret_cast<const T*>(st + 2 * lay.pitch) + cl;
            const T* sZ = reinterpret_cast<const T*>(st + 3 * lay.pitch) + cl;
            const T* sO = reinterpret_cast<const T*>(st + 4 * lay.pitch) + cl;
            const float* sd = reinterpret_cast<const float*>(st + 5 * lay.pitch) + hh;
            const int row = mi.row0 + run * TSP;
            const size_t tok0 = (size_t)mi.b * sp.L + row;
            T* dxa_o = reinterpret_cast<T*>(sp.dxa) + tok0 * sp.Di + cg0;
            T* dz_o = reinterpret_cast<T*>(sp.dz) + tok0 * sp.Di + cg0;
            T* dc_o = reinterpret_cast<T*>(sp.dCm) + tok0 * sp.dbc_stride + cg0;
            T* db_o = reinterpret_cast<T*>(sp.dBm) + tok0 * sp.dbc_stride + cg0;
            ab_mbar_wait(&mbar[s], (mphase >> s) & 1u);
            mphase ^= 1u << s;
            asm volatile("" : "+f"(m_al.x), "+f"(m_al.y), "+f"(m_dv.x), "+f"(m_dv.y), "+f"(m_h.x), "+f"(m_h.y));
            const f2 A2 = f2_pack(-__expf(m_al.x) * AB_LOG2E, -__expf(m_al.y) * AB_LOG2E), Dv = f2_pack(m_dv.x, m_dv.y);
            f2 h = f2_pack(m_h.x, m_h.y);
            f2 a[TSP], g[TSP], hp[TSP];
            float dl[TSP];
            f2 accD = f2_bcast(0.f);
            const f2 one = f2_bcast(1.f), neg1 = f2_bcast(-1.f);
#pragma unroll
            for (int t = 0; t < TSP; ++t) {
                const int r = run * TSP + t;
                dl[t] = sd[r * nhp];
                a[t] = f2_ex2(f2_mul(A2, f2_bcast(dl[t])));
                const f2 bv = lds_pair<T>(sB + (size_t)r * Cs), cv = lds_pair<T>(sC + (size_t)r * Cs);
                const f2 xv = lds_pair<T>(sX + (size_t)r * Cs), zv = lds_pair<T>(sZ + (size_t)r * Cs);
                const f2 dov = lds_pair<T>(sO + (size_t)r * Cs);
                const f2 sg = f2_sigmoid<T>(zv);
                const f2 dyv = f2_mul(dov, f2_mul(zv, sg));            // grad of (y_ssm + D*xa)
                hp[t] = h;
                h = f2_fma(a[t], h, bv);
                const f2 yv = f2_fma(Dv, xv, f2_mul(cv, h));
                const f2 dsilu = f2_mul(sg, f2_fma(zv, f2_fma(sg, neg1, one), one));     // silu'(z) = sg * (1 + z * (1 - sg))
                accD = f2_fma(dyv, xv, accD);
                g[t] = f2_mul(dyv, cv);
                if (act && row + t < sp.L) {
                    stg_pair<T>(dxa_o + (size_t)t * sp.Di, f2_mul(dyv, Dv));
                    stg_pair<T>(dz_o + (size_t)t * sp.Di, f2_mul(f2_mul(dov, yv), dsilu));
                    stg_pair<T>(dc_o + (size_t)t * sp.dbc_stride, f2_mul(dyv, h));
                }
            }
            const f2 anext = f2_ex2(f2_mul(A2, f2_bcast(sd[(run * TSP + TSP) * nhp])));
            // G entering this run from the later ones
            f2 Gin;
            if (first) {
                int spins = 0;
                while (!(word_valid(w0, epoch) && word_valid(w1, epoch))) {
                    if (++spins > PIPE_SPIN_LIMIT) { atomicExch(sp.err_flag, 1u); __trap(); }
                    __nanosleep(p.poll_ns);
                    ld_relaxed_v2(wp, w0, w1);
                }
                Gin = f2_pack(__uint_as_float((uint32_t)w0), __uint_as_float((uint32_t)w1));
            } else {
                Gin = *reinterpret_cast<const f2*>(carry + ((i - 1) & 1) * Cs + cl);       // G at the first token of the previous (later) tile
            }
            const f2 Gn = f2_fma(cqP[0], Gin, cqS[0]);
            f2 q = f2_mul(anext, Gn);
            f2 accA = f2_bcast(0.f);
            float dd[TSP];
#pragma unroll
            for (int t = TSP - 1; t >= 0; --t) {
                const f2 Gt = f2_add(g[t], q);
                if (act && row + t < sp.L) stg_pair<T>(db_o + (size_t)t * sp.dbc_stride, Gt);
                const f2 e = f2_mul(f2_mul(Gt, hp[t]), a[t]);          // d abar * abar
                float e0, e1;
                f2_unpack(f2_mul(e, A2), e0, e1);
                dd[t] = e0 + e1;
                accA = f2_fma(e, f2_bcast(dl[t]), accA);
                q = f2_mul(a[t], Gt);
                if (t == 0 && run == 0 && act) *reinterpret_cast<f2*>(carry + (i & 1) * Cs + cl) = Gt;
            }
            // d delta of a head = sum over its 16 channels = 8 lanes: transposing butterfly, one token per lane pair
            {
                const bool b2 = lane & 4, b1 = lane & 2;
                float v0 = b2 ? dd[2] : dd[0], v1 = b2 ? dd[3] : dd[1];
                const float s0 = b2 ? dd[0] : dd[2], s1 = b2 ? dd[1] : dd[3];
                v0 += __shfl_xor_sync(0xffffffffu, s0, 4);
                v1 += __shfl_xor_sync(0xffffffffu, s1, 4);
                float wv = b1 ? v1 : v0;
                const float sv = b1 ? v0 : v1;
                wv += __shfl_xor_sync(0xffffffffu, sv, 2);
                wv += __shfl_xor_sync(0xffffffffu, wv, 1);
                const int tt = (b2 ? 2 : 0) + (b1 ? 1 : 0);
                const float dlt = b2 ? (b1 ? dl[3] : dl[2]) : (b1 ? dl[1] : dl[0]);
                // d dlog = d delta * sigmoid(dlog) = d delta * (1 - exp(-delta));  A = A2 / log2e
                if (act && !(lane & 1) && row + tt < sp.L)
                    p.ddlog[(tok0 + tt) * sp.H + (cg0 >> 4)] = wv * inv_log2e * (1.f - __expf(-dlt));
            }
            if (act) {
                const int qb = ((i & 1) * NR + run) * Cs + cl;
                *reinterpret_cast<f2*>(partA + qb) = f2_mul(f2_mul(accA, A2), f2_bcast(inv_log2e));
                *reinterpret_cast<f2*>(partD + qb) = accD;
            }
        }
        // ---- prepass of tile i + PD: g = d(y_ssm) * C and the reverse run aggregates
        if (pi.b >= 0) {
            const unsigned char* st = smem + lay.off_pre + (size_t)ps * lay.pre_stride;
            const T* sC = reinterpret_cast<const T*>(st) + cl;
            const T* sZ = reinterpret_cast<const T*>(st + lay.pitch) + cl;
            const T* sO = reinterpret_cast<const T*>(st + 2 * lay.pitch) + cl;
            const float* sd = reinterpret_cast<const float*>(st + 3 * lay.pitch) + hh;
            ab_mbar_wait(&pbar[ps], (pphase >> ps) & 1u);
            pphase ^= 1u << ps;
            asm volatile("" : "+f"(p_al.x), "+f"(p_al.y));
            const f2 A2 = f2_pack(-__expf(p_al.x) * AB_LOG2E, -__expf(p_al.y) * AB_LOG2E);
            // reverse:  G(first token of the run) = Gs + Pr * G(first token of the next run)
            f2 Gs = f2_bcast(0.f), Pr = f2_bcast(1.f);
            f2 an = f2_ex2(f2_mul(A2, f2_bcast(sd[(run * TSP + TSP) * nhp])));
#pragma unroll
            for (int t = TSP - 1; t >= 0; --t) {
                const int r = run * TSP + t;
                const f2 cv = lds_pair<T>(sC + (size_t)r * Cs), zv = lds_pair<T>(sZ + (size_t)r * Cs), dov = lds_pair<T>(sO + (size_t)r * Cs);
                const f2 gt = f2_mul(f2_mul(dov, f2_mul(zv, f2_sigmoid<T>(zv))), cv);
                // G_t = g_t + a_{t+1} G_{t+1}:  with G_{t+1} = Gs + Pr * Gin
                Gs = f2_fma(an, Gs, gt);
                Pr = f2_mul(Pr, an);
                an = f2_ex2(f2_mul(A2, f2_bcast(sd[r * nhp])));
            }
            if (act) {
                const int bi = (((i + PD) & 1) * NR + run) * Cs + cl;
                *reinterpret_cast<f2*>(runP + bi) = Pr;
                *reinterpret_cast<f2*>(runS + bi) = Gs;
            }
        }
        __syncthreads();

        if (tid == 0 && i + 2 >= 0) {
            const TileInfo nm = s_info[(i + 2) & (INFO_RING - 1)];
            if (nm.b >= 0) issue_main(i & 1, nm);
        }
        if (tid == 32) {
            const TileInfo np = s_info[(i + LA) & (INFO_RING - 1)];
            if (np.b >= 0) issue_pre((ps + NPRE - 1) % NPRE, np);
        }
        // two rotating duties on different warps: compose + publish the prepass tile, reduce the main tile's partials
        const int duty_run = NR > 2 ? 2 + (i + PD) % (NR - 2) : (i + PD) % NR;
        const int parts_run = NR > 3 ? 2 + (i + PD + (NR - 2) / 2) % (NR - 2) : (i + PD + 1) % NR;
        if (run == duty_run && act && pi.b >= 0)
            compose_and_publish<NR, true>(runP + ((i + PD) & 1) * NR * Cs + cl, runS + ((i + PD) & 1) * NR * Cs + cl, Cs, accb + cl,
                                          pi.tile_lin & (SG - 1), sp.words + ((size_t)(pi.tile_lin >> LOG_SG) * Cs + cl) * 2, epoch);
        if (run == parts_run && act && i >= 0) reduce_parts(mi.tile_lin, i & 1);
#pragma unroll
        for (int q = 0; q + 1 < PD - 1; ++q) { cqP[q] = cqP[q + 1]; cqS[q] = cqS[q + 1]; }
        if (i + PD - 1 >= 0) {
            const int bi = (((i + PD - 1) & 1) * NR + run) * Cs + cl;
            cqP[PD - 2] = *reinterpret_cast<const f2*>(runP + bi);
            cqS[PD - 2] = *reinterpret_cast<const f2*>(runS + bi);
        }
        ps = ps + 1 == NPRE ? 0 : ps + 1;
    }
    cta_finish(p.sync);
}

}  // namespace

namespace {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PipeTiling { int Cs, NWC, NR, TT, nslab, nchunks, nsuper, esize; };

// slab = 64 channels when the width allows it, else the widest head-aligned divisor of Di that one CTA row covers.
// Forward and backward tile the sequence independently (the saved states are per run of TSP tokens).
bool pipe_tiling(int L, int Di, int dtype, bool bwd, PipeTiling& t) {
    t.esize = dtype == AB_F32 ? 4 : 2;
    int Cs = 0;
    if (Di % 64 == 0) Cs = 64;
    else
        for (int c = 256; c >= 16; c -= 16)
            if (Di % c == 0) { Cs = c; break; }
    if (!Cs) return false;
    t.Cs = Cs;
    t.NWC = (Cs + 63) / 64;
    if (Cs * 2 < t.NWC * 64) return false;                      // more than half of the lanes would idle
    static const int nr_bf16[5] = {0, 16, 8, 5, 4}, nr_f32[5] = {0, 8, 4, 3, 2};
    t.NR = dtype == AB_F32 ? nr_f32[t.NWC] : nr_bf16[t.NWC];
    (void)bwd;      // both directions use the same tiling (256-thread CTAs with twice as many per SM measured slower)
    t.TT = t.NR * TSP;
    t.nslab = Di / Cs;
    t.nchunks = (int)ab_ceil_div(L, t.TT);
    t.nsuper = (int)ab_ceil_div(t.nchunks, SG);          // the scanner's units; tiles are padded to nsuper * SG per chain
    return true;
}

constexpr int PIPE_SMS = 148;          // B200; the plan must not need a device
// CTAs per SM the shared memory and the register budget of the launch bounds admit
int pipe_occ_static(const PipeTiling& t, bool bwd) {
    const PipeSmem m = pipe_smem(t.Cs, t.TT, t.NR, bwd ? 5 : 4, bwd ? 3 : 1, t.esize, bwd);
    int occ = (int)((227u * 1024u) / (m.total + 1024u));
    const int thr = 32 * t.NWC * t.NR;
    const int by_reg = bwd ? (512 / thr > 0 ? 512 / thr : 1) : (1024 / thr > 0 ? 1024 / thr : 1);
    if (occ > by_reg) occ = by_reg;
    return occ;
}
// one scanner CTA per chain: worth it while the scanners are a small part of the persistent grid
bool pipe_supported(int B, int L, int Di, int dtype, PipeTiling& tf, PipeTiling& tb) {
    if (!pipe_tiling(L, Di, dtype, false, tf) || !pipe_tiling(L, Di, dtype, true, tb)) return false;
    const int occ_b = pipe_occ_static(tb, true), occ_f = pipe_occ_static(tf, false);
    if (occ_b < 1 || occ_f < 1) return false;
    const int nchains = B * tf.nslab;
    return nchains * 4 <= PIPE_SMS * occ_b && nchains * 4 <= PIPE_SMS * occ_f;
}

// saved-state buffer per batch (fp32, rows of Di): ceil(L/TSP) run states, then delta [L, Hp]
struct PipeSaved { int Hp, n_run, n_rows; };
PipeSaved pipe_saved(int L, int Di) {
    PipeSaved v;
    const int H = Di / 16;
    v.Hp = (H + 3) & ~3;
    v.n_run = (int)ab_ceil_div(L, TSP);
    v.n_rows = v.n_run + (int)ab_ceil_div((int64_t)L * v.Hp, Di);
    return v;
}

struct PipeWs { size_t off_sync, off_err, off_words, off_incl, off_part, total; };
PipeWs pipe_ws(const PipeTiling& tf, const PipeTiling& tb, int B) {
    PipeWs w;
    const PipeTiling& t = tf;
    const size_t ntiles = (size_t)B * t.nslab * (tf.nsuper > tb.nsuper ? tf.nsuper : tb.nsuper) * SG;
    size_t o = 0;
    w.off_sync = o; o += 64;
    w.off_err = o; o += 64;
    w.off_words = o; o += ntiles * t.Cs * 2 * sizeof(unsigned long long);
    w.off_incl = o; o += ntiles * t.Cs * sizeof(unsigned long long);
    w.off_part = o; o += ntiles * 2 * t.Cs * sizeof(float);
    w.total = (size_t)ab_round_up((int64_t)o, 256);
    return w;
}

int pipe_map3(CUtensorMap* m, const void* base, int dtype, int B, int L, int Di, int64_t row_stride, int Cs, int T) {
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE(((uintptr_t)base % 16) == 0 && (row_stride * es) % 16 == 0,
               "selective_scan: tensor base and row stride must be 16-byte aligned for TMA (stride %lld elems)", (long long)row_stride);
    uint64_t dims[3] = {(uint64_t)Di, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)row_stride * es, (uint64_t)row_stride * es * L};
    uint32_t box[3] = {(uint32_t)Cs, (uint32_t)T, 1};
    return ab_encode_tmap(m, dtype == AB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base,
                          dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}
// delta [B][L, Hp] fp32 inside the saved-state buffer: box = (heads of the slab in whole 16-byte units) x rows
int pipe_map_delta(CUtensorMap* m, const float* delta, int B, int L, int Hp, size_t batch_stride, int Cs, int rows) {
    AB_REQUIRE(((uintptr_t)delta % 16) == 0 && (batch_stride * 4) % 16 == 0, "selective_scan: saved-state buffer must be 16-byte aligned");
    uint64_t dims[3] = {(uint64_t)Hp, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)Hp * 4, (uint64_t)batch_stride * 4};
    uint32_t box[3] = {(uint32_t)((((uint32_t)Cs >> 4) + 3u) & ~3u), (uint32_t)rows, 1};
    return ab_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, delta, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

template <typename K>
int pipe_grid(K kfn, int threads, size_t smem, int want, int* grid) {
    // attributes and occupancy are queried once per kernel and shared-memory size (the calls cost tens of microseconds)
    static thread_local size_t cached_smem = 0;
    static thread_local int cached_occ = 0, cached_dev = -1;
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (cached_smem != smem || cached_dev != dev) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AB_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int o = 0;
        AB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kfn, threads, smem));
        cached_occ = o; cached_smem = smem; cached_dev = dev;
    }
    const int occ = cached_occ;
    AB_REQUIRE(occ >= 1, "selective_scan (pipelined): kernel does not fit on an SM (smem %zu)", smem);
    int g = occ * ab_num_sms();
    if (g > want) g = want;
    *grid = g;
    return AB_OK;
}

template <typename T, int NWC, int NR, int CS>
int launch_fwd_pipe_cs(const CUtensorMap* maps, const PipeParams& p, const PipeTiling& t, cudaStream_t st) {
    const PipeSmem m = pipe_smem(t.Cs, t.TT, NR, 4, 1, (int)sizeof(T), false);
    auto kfn = scan_fwd_pipe_kernel<T, NWC, NR, CS>;
    int grid = 0;
    if (int e = pipe_grid(kfn, 32 * NWC * NR, m.total, p.s.n_scan + p.ntiles, &grid)) return e;
    AB_REQUIRE(grid > p.s.n_scan, "selective_scan (pipelined): %d chains leave no worker CTA", p.s.n_scan);
    kfn<<<grid, 32 * NWC * NR, m.total, st>>>(maps[0], maps[1], maps[2], maps[3], maps[5], p);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
template <typename T, int NWC, int NR, int CS>
int launch_bwd_pipe_cs(const CUtensorMap* maps, const PipeParams& p, const PipeTiling& t, cudaStream_t st) {
    const PipeSmem m = pipe_smem(t.Cs, t.TT, NR, 5, 3, (int)sizeof(T), true);
    auto kfn = scan_bwd_pipe_kernel<T, NWC, NR, CS>;
    int grid = 0;
    if (int e = pipe_grid(kfn, 32 * NWC * NR, m.total, p.s.n_scan + p.ntiles, &grid)) return e;
    AB_REQUIRE(grid > p.s.n_scan, "selective_scan (pipelined): %d chains leave no worker CTA", p.s.n_scan);
    kfn<<<grid, 32 * NWC * NR, m.total, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
// slab widths with a specialised kernel: 64 (d_inner a multiple of 64) and 176 (the 1.5B text block)
template <typename T, int NWC, int NR>
int launch_pipe(bool bwd, const CUtensorMap* maps, const PipeParams& p, const PipeTiling& t, cudaStream_t st) {
    constexpr int CSS = NWC == 1 ? 64 : (NWC == 3 ? 176 : 0);
    if (CSS && t.Cs == CSS) return bwd ? launch_bwd_pipe_cs<T, NWC, NR, CSS>(maps, p, t, st) : launch_fwd_pipe_cs<T, NWC, NR, CSS>(maps, p, t, st);
    return bwd ? launch_bwd_pipe_cs<T, NWC, NR, 0>(maps, p, t, st) : launch_fwd_pipe_cs<T, NWC, NR, 0>(maps, p, t, st);
}
int dispatch_pipe(bool bwd, int dtype, const CUtensorMap* maps, const PipeParams& p, const PipeTiling& t, cudaStream_t st) {
    if (dtype == AB_BF16) switch (t.NWC) {
        case 1: return launch_pipe<__nv_bfloat16, 1, 16>(bwd, maps, p, t, st);
        case 2: return launch_pipe<__nv_bfloat16, 2, 8>(bwd, maps, p, t, st);
        case 3: return launch_pipe<__nv_bfloat16, 3, 5>(bwd, maps, p, t, st);
        default: return launch_pipe<__nv_bfloat16, 4, 4>(bwd, maps, p, t, st);
    }
    switch (t.NWC) {
        case 1: return launch_pipe<float, 1, 8>(bwd, maps, p, t, st);
        case 2: return launch_pipe<float, 2, 4>(bwd, maps, p, t, st);
        case 3: return launch_pipe<float, 3, 3>(bwd, maps, p, t, st);
        default: return launch_pipe<float, 4, 2>(bwd, maps, p, t, st);
    }
}

void fill_common(PipeParams& p, const PipeTiling& t, const PipeWs& wl, void* ws, int B, int L, int Di, int H) {
    memset(&p, 0, sizeof(p));
    ScanParams& s = p.s;
    s.B = B; s.L = L; s.Di = Di; s.H = H;
    s.Cs = t.Cs; s.T = t.TT; s.n_s = t.NR; s.nslab = t.nslab; s.nchunks = t.nsuper; s.nchains = B * t.nslab;
    unsigned char* w8 = (unsigned char*)ws;
    p.sync = (unsigned int*)(w8 + wl.off_sync);
    s.ticket = p.sync;
    s.err_flag = (unsigned int*)(w8 + wl.off_err);
    s.words = (unsigned long long*)(w8 + wl.off_words);
    s.inclw = (unsigned long long*)(w8 + wl.off_incl);
    s.part = (float*)(w8 + wl.off_part);
    s.n_scan = s.nchains;
    p.ntiles = s.nchains * t.nsuper;          // tickets = super-tiles
    static const char* ke = getenv("APERTIS_B200_SCAN_K");
    static const char* re = getenv("APERTIS_B200_SCAN_R");
    p.scan_k = ke ? atoi(ke) : PIPE_SCAN_K;
    if (p.scan_k != 16 && p.scan_k != 32 && p.scan_k != 64 && p.scan_k != 128) p.scan_k = PIPE_SCAN_K;
    p.scan_r = re && atoi(re) == 8 ? 8 : 4;
    static const char* pe = getenv("APERTIS_B200_SCAN_POLL_NS");
    p.poll_ns = pe ? (unsigned)atoi(pe) : 100u;
}

}  // namespace

// entry points used by ssm_scan.cu
bool ab_scan_pipe_plan(int B, int L, int Di, int dtype, int* tile_rows, int* slab, int* n_states, size_t* ws_bytes,
                       int* bwd_tile_rows, int* bwd_tiles_per_chain) {
    PipeTiling tf, tb;
    if (!pipe_supported(B, L, Di, dtype, tf, tb)) return false;
    if (tile_rows) *tile_rows = tf.TT;
    if (bwd_tile_rows) *bwd_tile_rows = tb.TT;
    if (bwd_tiles_per_chain) *bwd_tiles_per_chain = tb.nsuper * SG;
    if (slab) *slab = tf.Cs;
    if (n_states) *n_states = pipe_saved(L, Di).n_rows;          // rows of Di floats per batch: run states, then delta
    if (ws_bytes) *ws_bytes = pipe_ws(tf, tb, B).total;
    return true;
}

int ab_scan_pipe_fwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride, const void* z,
                     int64_t z_stride, const float* A_log, const float* D, const float* h0, void* y, float* h_last,
                     float* hrun, void* ws, size_t ws_bytes, int B, int L, int Di, int H, int dtype, cudaStream_t stream) {
    PipeTiling t, tb;
    AB_REQUIRE(pipe_supported(B, L, Di, dtype, t, tb), "selective_scan_fwd: the pipelined mode does not cover B=%d L=%d Di=%d (ask ab_selective_scan_plan)", B, L, Di);
    const PipeWs wl = pipe_ws(t, tb, B);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "selective_scan_fwd: workspace too small (%zu < %zu)", ws_bytes, wl.total);
    AB_REQUIRE(hrun != nullptr, "selective_scan_fwd: the pipelined mode needs the saved-state buffer (hstart)");
    const PipeSaved sv = pipe_saved(L, Di);
    const size_t batch_stride = (size_t)sv.n_rows * Di;
    float* delta = hrun + (size_t)sv.n_run * Di;
    {
        const size_t total = (size_t)B * L * sv.Hp;
        const unsigned blocks = (unsigned)(ab_ceil_div((int64_t)total, 256) < 4096 ? ab_ceil_div((int64_t)total, 256) : 4096);
        if (dtype == AB_F32) scan_delta_kernel<float><<<blocks, 256, 0, stream>>>((const float*)dlog, delta, L, H, sv.Hp, batch_stride, total);
        else scan_delta_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)dlog, delta, L, H, sv.Hp, batch_stride, total);
        AB_LAUNCH_CHECK();
    }
    CUtensorMap maps[6];
    if (int e = pipe_map3(&maps[0], xa, dtype, B, L, Di, Di, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[1], Bm, dtype, B, L, Di, bc_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[2], Cm, dtype, B, L, Di, bc_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[3], z, dtype, B, L, Di, z_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map_delta(&maps[5], delta, B, L, sv.Hp, batch_stride, t.Cs, t.TT)) return e;
    PipeParams p;
    fill_common(p, t, wl, ws, B, L, Di, H);
    p.s.dlog = dlog; p.s.A_log = A_log; p.s.Dp = D; p.s.h0 = h0; p.s.y = y; p.s.h_last = h_last;
    p.hrun = hrun; p.batch_stride = batch_stride;
    return dispatch_pipe(false, dtype, maps, p, t, stream);
}

int ab_scan_pipe_bwd(const void* xa, const void* dlog, const void* Bm, const void* Cm, int64_t bc_stride, const void* z,
                     int64_t z_stride, const void* dout, const float* A_log, const float* D, const float* hrun, void* dxa,
                     void* dBm, void* dCm, int64_t dbc_stride, void* dz, float* ddlog, float** part_out, void* ws,
                     size_t ws_bytes, int B, int L, int Di, int H, int dtype, cudaStream_t stream) {
    PipeTiling tf, t;
    AB_REQUIRE(pipe_supported(B, L, Di, dtype, tf, t), "selective_scan_bwd: the pipelined mode does not cover B=%d L=%d Di=%d", B, L, Di);
    const PipeWs wl = pipe_ws(tf, t, B);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "selective_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, wl.total);
    const int es = dtype == AB_F32 ? 4 : 2;
    AB_REQUIRE((dbc_stride * es) % 8 == 0, "selective_scan_bwd: dB/dC row stride must be 8-byte aligned");
    const PipeSaved sv = pipe_saved(L, Di);
    const size_t batch_stride = (size_t)sv.n_rows * Di;
    const float* delta = hrun + (size_t)sv.n_run * Di;
    CUtensorMap maps[6];
    if (int e = pipe_map3(&maps[0], xa, dtype, B, L, Di, Di, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[1], Bm, dtype, B, L, Di, bc_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[2], Cm, dtype, B, L, Di, bc_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[3], z, dtype, B, L, Di, z_stride, t.Cs, t.TT)) return e;
    if (int e = pipe_map3(&maps[4], dout, dtype, B, L, Di, Di, t.Cs, t.TT)) return e;
    if (int e = pipe_map_delta(&maps[5], delta, B, L, sv.Hp, batch_stride, t.Cs, t.TT + 1)) return e;
    PipeParams p;
    fill_common(p, t, wl, ws, B, L, Di, H);
    (void)dlog;
    p.s.A_log = A_log; p.s.Dp = D;
    p.s.dxa = dxa; p.s.dBm = dBm; p.s.dCm = dCm; p.s.dz = dz; p.s.dbc_stride = dbc_stride;
    p.hrun = const_cast<float*>(hrun); p.batch_stride = batch_stride;
    p.ddlog = ddlog;
    *part_out = p.s.part;
    return dispatch_pipe(true, dtype, maps, p, t, stream);
}

#ifdef AB_SCAN_TRACE
extern "C" int ab_pipe_trace_dump(unsigned long long* host, int clear) {
    cudaDeviceSynchronize();
    int rc = (int)cudaMemcpyFromSymbol(host, g_pipe_trace, sizeof(unsigned long long) * PT_CTAS * PT_ITERS * PT_WARPS * PT_SLOTS);
    if (clear) {
        void* d; cudaGetSymbolAddress(&d, g_pipe_trace); cudaMemset(d, 0, sizeof(unsigned long long) * PT_CTAS * PT_ITERS * PT_WARPS * PT_SLOTS);
        cudaGetSymbolAddress(&d, g_pipe_trace_cta); cudaMemset(d, 0, 4);
    }
    return rc;
}
#endif
