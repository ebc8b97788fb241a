// Grouped expert GEMM on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA into 128B-swizzled shared memory, warp-specialised and persistent, on CTA PAIRS.
// Replaces the per-(slot, expert) nn.Linear calls of core.py:596 (experts[e] = LN, Linear, act, Dropout,
// Linear; core.py:434-442) and their autograd (dgrad / wgrad).
//
//   MODE_NT  C[r,n] = epi(sum_k A[r,k] * W[e,n,k])      A [rows,K] K-major,   W [E*N, K] K-major   (forward)
//   MODE_NN  C[r,n] = epi(sum_k A[r,k] * W[e,k,n])      A [rows,K] K-major,   W [E*K, N] N-major   (dgrad)
//   MODE_TN  Cw[e,m,n] = sum_{r in seg e} A[r,m]*B[r,n] A [rows,M] M-major,   B [rows,N] N-major   (wgrad)
//
// The kernel runs as clusters of two CTAs (the two SMs of a TPC) and issues `tcgen05.mma.cta_group::2`: one MMA covers
// 256 rows x BN columns x 16, rows 0..127 accumulate in the TMEM of the even CTA (the leader, which issues the MMAs),
// rows 128..255 in the TMEM of the odd CTA.  Each CTA loads its own 128 rows of A and only HALF of the B tile (BN/2
// columns); the tensor core reads both halves.  Against one-CTA 128 x BN tiles this cuts the operand bytes that travel
// L2 -> shared memory per flop by a third at BN = 256 (the one-CTA kernel moved 14 TB/s of operands, r2a ncu capture).
// Per CTA: BK = 64, 5-stage TMA -> MMA ring of (16 KB A + up to 16 KB B-half), 2 accumulator stages in TMEM (2 x 256
// columns) so the epilogue of tile i overlaps the MMAs of tile i+1.  Warp roles in each CTA: 0 = TMA producer, 1 = MMA
// issuer (leader only; + TMEM alloc), 2..17 = epilogue: four warps per TMEM lane quarter, each owning one 64-column group
// of the CTA's 128 x BN accumulator (tcgen05.ld -> bias / activation math in registers -> swizzled per-warp staging ->
// coalesced 16-byte global stores; no CTA-wide synchronisation in the epilogue).  Barriers: `full` lives in the leader and
// counts the TMA bytes of both CTAs; `empty` and `tfull` exist in both CTAs and are signalled by multicast
// tcgen05.commit; `tempty` lives in the leader and is arrived on by the epilogue warps of both CTAs.
// Row tiles (256 permuted rows) belong to one expert; the number of valid rows is read from device memory (n_rows[0])
// so no host synchronisation is needed after the routing plan.
#include "common.cuh"

namespace {

constexpr int BM = 128;                       // accumulator rows per CTA (TMEM lanes)
constexpr int PM = 2 * BM;                    // rows of a pair tile (AB_GEMM_ROW_TILE)
constexpr int BK = 64, NSTAGE = 5, NACC = 2;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int BH_BYTES_MAX = 128 * BK * 2;    // 16 KB: this CTA's half of a 256-column B tile
constexpr int STAGE_BYTES = A_BYTES + BH_BYTES_MAX;
constexpr int ATOM_BYTES = 64 * BK * 2;       // one 64(MN) x 64(K) swizzle-128B MN-major atom = 8 KB
constexpr int EPI_WARP0 = 2, EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + EPI_THREADS; // warp 0 TMA, warp 1 MMA, 16 epilogue warps
constexpr int GW = 64;                        // epilogue column group: one per warp of a TMEM lane quarter
constexpr int MODE_NT = 0, MODE_NN = 1, MODE_TN = 2;
constexpr int WARP_STG_BYTES = 32 * 64;       // per-warp epilogue staging: 32 rows x 64 B
constexpr int WARP_BIAS_BYTES = GW * 4;       // per-warp bias slice of the current tile
constexpr size_t SMEM_BYTES = 1024 + (size_t)NSTAGE * STAGE_BYTES + (size_t)EPI_WARPS * (WARP_STG_BYTES + WARP_BIAS_BYTES) + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(PM == AB_GEMM_ROW_TILE, "row tile of the ABI");

struct GemmParams {
    int N, K, E, M;          // M only for MODE_TN (rows of each expert's output)
    int bn;                  // N tile (both CTAs together)
    int num_n_tiles, num_m_tiles;   // num_m_tiles: MODE_TN, 256-row tiles of the output
    int epi, act, c_f32;
    int tn_nsrc;             // MODE_TN: an expert's rows are nsrc blocks [src*stride + seg_off[e], src*stride + seg_off[e+1])
    int tn_e_real;           // MODE_TN split over the source blocks: E counts (slice, expert) pairs, expert = index % tn_e_real,
                             // slice = index / tn_e_real covers tn_nsrc consecutive blocks; 0 = no split
    int64_t tn_src_stride;
    // dense use (one "expert", plain row-major matrices; tile_expert / n_rows / seg_off are NULL):
    int64_t rows_valid;      // MODE_NT / MODE_NN: rows >= rows_valid are never stored (the last row tile may be partial)
    int dense_m_tiles;       // MODE_NT / MODE_NN: number of 256-row tiles when n_rows is NULL
    int64_t tn_rows;         // MODE_TN with seg_off NULL: contraction over rows [0, tn_rows), TMA zero-fills past the end
    int64_t tn_split;        // ... cut into E slices of tn_split rows (a multiple of BK): Cw[e] holds slice e's partial product
    const int32_t* tile_expert;
    const int32_t* n_rows;
    const int32_t* seg_off;
    const float* bias;
    const void* aux;
    const uint32_t* drop_seed;   // device [2]: seed of the expert-internal Dropout mask (core.py:439); NULL = no dropout
    uint32_t drop_thresh;        // drop when the 16-bit uniform < thresh (= p * 65536)
    float drop_scale;            // 1 / (1 - p)
    void* c;
    void* c2;
    float* cw;
    // expert parallelism: c is written into OTHER ranks' memory (NVLink): row r of the local [W sources][rows_per_peer] layout
    // goes to rank r / rows_per_peer, block `peer_rank` of that rank's [W owners][rows_per_peer] buffer.  peer_n == 0: local.
    unsigned char* peer_base[16];
    int peer_n, peer_rows, peer_rank;
};

// dense weight gradient: 64-row blocks of slice e of the contraction (the last block of the last slice may be partial: TMA zero-fills)
__device__ __forceinline__ int dense_tn_blocks(const GemmParams& p, int e) {
    const int64_t lo = (int64_t)e * p.tn_split;
    const int64_t hi = lo + p.tn_split < p.tn_rows ? lo + p.tn_split : p.tn_rows;
    return hi > lo ? (int)((hi - lo + BK - 1) / BK) : 0;
}

// ---- math for the epilogues -------------------------------------------------------------------
// fp32 data path (outputs in fp32: the fp32-parity mode): erf with |error| < 1.5e-7 (Abramowitz & Stegun 7.1.26):
// erf(x/sqrt2) from t = 1/(1+p|x|/sqrt2) and g = exp(-x^2/2)
__device__ __forceinline__ float erf_core(float ax_s, float g) {     // ax_s = |x|/sqrt(2), g = exp(-ax_s^2)
    const float t = ab_rcp(fmaf(0.3275911f, ax_s, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    return 1.0f - p * t * g;
}
// exact-erf GELU as  max(x, 0) - h,  h = 0.5 |x| (1 - erf(|x|/sqrt2)) = (|x|/sqrt2) * t * q(t) * exp(-x^2/2)  with the
// 1/sqrt2 folded into the polynomial q: no sign handling, 11 fp32 ops + 2 MUFU per element
__device__ __forceinline__ float gelu_fwd(float x) {
    const float a = fabsf(x) * 0.70710678118654752f;
    const float g = ab_ex2(x * (x * (-0.5f * AB_LOG2E)));
    const float t = ab_rcp(fmaf(0.3275911f, a, 1.0f));
    float q = fmaf(0.750526976f, t, -1.027533652f);        // A&S 7.1.26 coefficients x 1/sqrt2
    q = fmaf(q, t, 1.005091295f);
    q = fmaf(q, t, -0.201169571f);
    q = fmaf(q, t, 0.180191733f);
    return fmaxf(x, 0.f) - (a * t) * (q * g);
}
__device__ __forceinline__ float act_fwd(float x, int act) {
    if (act == AB_ACT_GELU) return gelu_fwd(x);
    if (act == AB_ACT_RELU) return fmaxf(x, 0.f);
    return x * ab_sigmoid(x);
}
__device__ __forceinline__ float act_bwd(float x, int act) {
    if (act == AB_ACT_GELU) {        // cdf(x) + x*pdf(x); the erf and the pdf share exp(-x^2/2)
        const float a = fabsf(x) * 0.70710678118654752f;
        const float g = ab_ex2(-a * a * AB_LOG2E);
        const float cdf = 0.5f * (1.0f + copysignf(erf_core(a, g), x));
        return fmaf(x, 0.3989422804014327f * g, cdf);
    }
    if (act == AB_ACT_RELU) return x > 0.f ? 1.f : 0.f;
    const float s = ab_sigmoid(x);
    return s * fmaf(x, 1.f - s, 1.f);
}

// ---- the same activations on column pairs (FFMA2 / FMUL2): packed arithmetic halves the floating-point instruction count
//      of the epilogue, which at K = 704 has as many issue slots to fill as the tensor core has cycles ---------------------
__device__ __forceinline__ f2 gelu_fwd2(f2 x) {
    float x0, x1;
    f2_unpack(x, x0, x1);
    const f2 a = f2_pack(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
    const f2 g = f2_ex2(f2_mul(x, f2_mul(x, f2_bcast(-0.5f * AB_LOG2E))));
    const f2 t = f2_rcp(f2_fma(f2_bcast(0.3275911f), a, f2_bcast(1.0f)));
    f2 q = f2_fma(f2_bcast(0.750526976f), t, f2_bcast(-1.027533652f));        // A&S 7.1.26 coefficients x 1/sqrt2
    q = f2_fma(q, t, f2_bcast(1.005091295f));
    q = f2_fma(q, t, f2_bcast(-0.201169571f));
    q = f2_fma(q, t, f2_bcast(0.180191733f));
    const f2 h = f2_mul(f2_mul(a, t), f2_mul(q, g));
    return f2_fma(h, f2_bcast(-1.0f), f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}
__device__ __forceinline__ f2 gelu_bwd2(f2 x) {        // cdf(x) + x * pdf(x)
    float x0, x1;
    f2_unpack(x, x0, x1);
    const f2 a = f2_pack(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
    const f2 g = f2_ex2(f2_mul(f2_mul(a, a), f2_bcast(-AB_LOG2E)));
    const f2 t = f2_rcp(f2_fma(f2_bcast(0.3275911f), a, f2_bcast(1.0f)));
    f2 pl = f2_fma(f2_bcast(1.061405429f), t, f2_bcast(-1.453152027f));
    pl = f2_fma(pl, t, f2_bcast(1.421413741f));
    pl = f2_fma(pl, t, f2_bcast(-0.284496736f));
    pl = f2_fma(pl, t, f2_bcast(0.254829592f));
    const f2 erfv = f2_fma(f2_mul(pl, t), f2_mul(g, f2_bcast(-1.0f)), f2_bcast(1.0f));      // erf(|x| / sqrt2)
    float e0, e1;
    f2_unpack(erfv, e0, e1);
    const f2 cdf = f2_fma(f2_pack(copysignf(e0, x0), copysignf(e1, x1)), f2_bcast(0.5f), f2_bcast(0.5f));
    return f2_fma(x, f2_mul(g, f2_bcast(0.3989422804014327f)), cdf);
}

// ---- bf16 data path: the normal cdf as 0.5 (1 + tanh(u)), u = x (c0 + c1 x^2 + c2 x^4) (least-squares fit on |x| <= 6 with
//      x^2 clamped at 81; against the erf form |d gelu| < 3e-5 and |d gelu'| < 1.1e-4 before the hardware tanh's own 2^-11,
//      i.e. together < |x| 2.6e-4: a sixteenth of a bf16 ulp of the value being rounded).  One MUFU and 6 packed
//      operations per column pair instead of four MUFU and 12; the derivative is the exact derivative of the same
//      approximation, so forward and backward stay consistent.  `hs` = 0.5 x (the dropout scale 1 / (1 - p), or 1).
#define AB_GF_C0 0.79746994f
#define AB_GF_C1 0.037015004f
#define AB_GF_C2 -3.5035629e-4f
__device__ __forceinline__ float ab_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ f2 f2_tanh(f2 v) { float a, b; f2_unpack(v, a, b); return f2_pack(ab_tanh(a), ab_tanh(b)); }
__device__ __forceinline__ f2 gelu_fast_x2(f2 x) {
    float a, b;
    f2_unpack(f2_mul(x, x), a, b);
    return f2_pack(fminf(a, 81.f), fminf(b, 81.f));
}
__device__ __forceinline__ f2 gelu_fwd_fast2(f2 x, f2 hs) {
    const f2 x2 = gelu_fast_x2(x);
    f2 p = f2_fma(f2_bcast(AB_GF_C2), x2, f2_bcast(AB_GF_C1));
    p = f2_fma(p, x2, f2_bcast(AB_GF_C0));
    const f2 t = f2_tanh(f2_mul(p, x));
    const f2 hx = f2_mul(x, hs);
    return f2_fma(hx, t, hx);
}
// d/dx [x cdf(x)] = 0.5 (1 + t) [1 + (1 - t) x u'(x)],  u' = c0 + 3 c1 x^2 + 5 c2 x^4
__device__ __forceinline__ f2 gelu_bwd_fast2(f2 x, f2 hs) {
    const f2 x2 = gelu_fast_x2(x);
    f2 p = f2_fma(f2_bcast(AB_GF_C2), x2, f2_bcast(AB_GF_C1));
    p = f2_fma(p, x2, f2_bcast(AB_GF_C0));
    const f2 t = f2_tanh(f2_mul(p, x));
    f2 up = f2_fma(f2_bcast(5.f * AB_GF_C2), x2, f2_bcast(3.f * AB_GF_C1));
    up = f2_fma(up, x2, f2_bcast(AB_GF_C0));
    const f2 w = f2_mul(x, up);
    const f2 e = f2_fma(f2_mul(t, f2_bcast(-1.0f)), w, w);       // (1 - t) w
    const f2 A = f2_fma(t, hs, hs);                              // 0.5 s (1 + t)
    return f2_fma(A, e, A);
}

// activation of a row in place, times `scale` (the expert-internal Dropout's 1 / (1 - p), or 1)
template <int U, bool F32>
__device__ __forceinline__ void act_fwd_row(float (&f)[U], int act, float scale) {
    if (act == AB_ACT_GELU && !F32) {
        const f2 hs = f2_bcast(0.5f * scale);
#pragma unroll
        for (int i = 0; i < U; i += 2) f2_unpack(gelu_fwd_fast2(f2_pack(f[i], f[i + 1]), hs), f[i], f[i + 1]);
    } else if (act == AB_ACT_GELU) {
        const f2 sc = f2_bcast(scale);
#pragma unroll
        for (int i = 0; i < U; i += 2) f2_unpack(f2_mul(gelu_fwd2(f2_pack(f[i], f[i + 1])), sc), f[i], f[i + 1]);
    } else if (act == AB_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < U; ++i) f[i] = fmaxf(f[i], 0.f) * scale;
    } else {
#pragma unroll
        for (int i = 0; i < U; ++i) f[i] = act_fwd(f[i], AB_ACT_SILU) * scale;
    }
}
// bf16 round trip of a row (the value the next kernel will read), two columns per conversion
template <int U>
__device__ __forceinline__ void round_row_bf16(float (&f)[U]) {
#pragma unroll
    for (int i = 0; i < U; i += 2) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f[i + 1]), "f"(f[i]));
        f[i] = __uint_as_float(r << 16);
        f[i + 1] = __uint_as_float(r & 0xffff0000u);
    }
}

// Dropout keep-mask of the expert hidden activation: counter-based hash of the element index and a 64-bit seed, so the
// backward regenerates exactly the forward's mask without storing it.  One hash serves 4 consecutive columns (four
// 16-bit uniforms, drop when < p * 65536).  Both seed words enter between multiplication rounds, so masks of different
// seeds are unrelated sequences, not shifted copies.  The survivors' scale 1 / (1 - p) is applied by the activation.
__device__ __forceinline__ void drop_keep4(uint32_t xb, uint32_t grp, uint32_t s1, uint32_t thi, bool (&keep)[4]) {
    uint32_t x = xb + grp * 0x9E3779B1u;             // = (index of the 4-column group) * golden ratio + seed word 0
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x = x ^ (x >> 13) ^ s1; x *= 0xC2B2AE35u;
    uint32_t y = x * 0x27D4EB2Fu + s1;
    y ^= y >> 15;
    keep[0] = (x << 16) >= thi;      // 16-bit uniforms compared in place: low halves shifted up, high halves as they are
    keep[1] = x >= thi;
    keep[2] = (y << 16) >= thi;
    keep[3] = y >= thi;
}
template <int U>
__device__ __forceinline__ void dropout_zero_row(float (&f)[U], uint32_t drop_thresh, uint32_t s0, uint32_t s1, uint32_t grow, int ncol, int n_cols) {
    const uint32_t xb = ((grow * (uint32_t)n_cols + (uint32_t)ncol) >> 2) * 0x9E3779B1u + s0;
    const uint32_t thi = drop_thresh << 16;
#pragma unroll
    for (int i = 0; i < U; i += 4) {
        bool keep[4];
        drop_keep4(xb, (uint32_t)(i >> 2), s1, thi, keep);
#pragma unroll
        for (int j = 0; j < 4; ++j) f[i + j] = keep[j] ? f[i + j] : 0.f;
    }
}

// ---- descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}
// instruction descriptor: bf16 x bf16 -> f32, M = 256 (cta_group::2: 128 rows per CTA)
__host__ __device__ inline uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format f32
    d |= 1u << 7;                    // A format bf16
    d |= 1u << 10;                   // B format bf16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(PM >> 4) << 24;
    return d;
}

// ---- tile schedule (identical in every role of both CTAs of a pair) -------------------------------
struct Tile { int m_pair, n_tile, e, k_begin, nk; };

template <int MODE>
__device__ __forceinline__ int total_tiles_of(const GemmParams& p) {
    if (MODE == MODE_TN) return p.E * p.num_m_tiles * p.num_n_tiles;
    return (p.n_rows != nullptr ? p.n_rows[0] / PM : p.dense_m_tiles) * p.num_n_tiles;
}
template <int MODE>
__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int tile) {
    Tile t;
    if (MODE == MODE_TN) {
        const int per_e = p.num_m_tiles * p.num_n_tiles;
        t.e = tile / per_e;
        const int rem = tile % per_e;
        t.m_pair = rem / p.num_n_tiles; t.n_tile = rem % p.num_n_tiles;
        if (p.seg_off != nullptr) {
            const int e = p.tn_e_real > 0 ? t.e % p.tn_e_real : t.e, slice = p.tn_e_real > 0 ? t.e / p.tn_e_real : 0;
            const int s0 = p.seg_off[e];
            t.k_begin = s0 + (int)((int64_t)slice * p.tn_nsrc * p.tn_src_stride);
            t.nk = p.tn_nsrc * ((p.seg_off[e + 1] - s0) / BK);
        } else {
            t.k_begin = (int)(t.e * p.tn_split);
            t.nk = dense_tn_blocks(p, t.e);
        }
    } else {
        t.m_pair = tile / p.num_n_tiles; t.n_tile = tile % p.num_n_tiles;
        t.e = p.tile_expert != nullptr ? p.tile_expert[t.m_pair] : 0;
        t.k_begin = 0;
        t.nk = (p.K + BK - 1) / BK;
    }
    return t;
}

// ---- epilogue -----------------------------------------------------------------------------------
// One warp owns 32 accumulator rows (its TMEM lane quarter) x one 64-column group and walks it in units of 64 output
// bytes per row (32 bf16 or 16 f32 columns): tcgen05.ld -> math -> swizzled per-warp staging -> coalesced 16-byte
// global stores (8 rows x 64 B per instruction).  No CTA-wide synchronisation; everything is compile-time indexed so the
// fragments stay in registers.  Operands the math needs from global memory never sit on the critical path: the tile's bias
// slice is parked in shared memory before the accumulator is waited for, and the saved pre-activation / addend tile of the
// DACT / ADD epilogues is fetched one unit ahead into registers.
__device__ __forceinline__ uint32_t stg_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

template <int MODE, bool F32>
__device__ __forceinline__ void stage_and_store(const GemmParams& p, unsigned char* stg, const float (&src)[F32 ? 16 : 32],
                                                unsigned char* base, int lane, size_t row0, int ncol, int ncol_end) {
    constexpr int ES = F32 ? 4 : 2;
    constexpr int CPV = 16 / ES;
    const int N = p.N;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 q;
        if (F32) q = make_uint4(__float_as_uint(src[j * 4]), __float_as_uint(src[j * 4 + 1]), __float_as_uint(src[j * 4 + 2]), __float_as_uint(src[j * 4 + 3]));
        else q = ab_vec16<__nv_bfloat16>::pack(&src[j * 8]);
        *reinterpret_cast<uint4*>(stg + stg_off(lane, j)) = q;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2), j = lane & 3;
        const int col = ncol + j * CPV;
        const size_t grow = row0 + r;
        const bool ok = col < ncol_end && (MODE == MODE_TN ? (int)grow < p.M : (int64_t)grow < p.rows_valid);
        const uint4 q = *reinterpret_cast<const uint4*>(stg + stg_off(r, j));
        if (ok) *reinterpret_cast<uint4*>(base + (grow * N + col) * ES) = q;
    }
    __syncwarp();
}

// 32 rows x 64 B of a row-major [rows, N] tensor (aux), coalesced: 8 rows x 64 B per instruction
template <bool F32>
__device__ __forceinline__ void load_aux_unit(const GemmParams& p, uint4 (&q)[4], int lane, size_t row0, int ncol, int ncol_end) {
    constexpr int ES = F32 ? 4 : 2;
    constexpr int CPV = 16 / ES;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2), j = lane & 3;
        const int col = ncol + j * CPV;
        const size_t grow = row0 + r;
        q[it] = make_uint4(0u, 0u, 0u, 0u);
        if (col < ncol_end && (int64_t)grow < p.rows_valid)
            q[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.aux) + (grow * p.N + col) * ES));
    }
}

struct EpiCtx {
    unsigned char* stg;      // this warp's staging
    float* bias_s;           // this warp's bias slice
    uint64_t* tfull;
    uint64_t* tempty;
    uint32_t tmem_base;
    int warp, lane;
    uint32_t rank;
    int pair_id, num_pairs;
};

template <int MODE, bool F32, int EPI>
__device__ __forceinline__ void epilogue_role(const GemmParams& p, const EpiCtx& c) {
    constexpr int U = F32 ? 16 : 32;            // columns per unit
    constexpr int ES = F32 ? 4 : 2;
    constexpr int CPV = 16 / ES;                // columns per 16-byte vector
    constexpr bool HAS_BIAS = EPI == AB_EPI_BIAS || EPI == AB_EPI_BIAS_ACT;
    constexpr bool HAS_AUX = EPI == AB_EPI_DACT || EPI == AB_EPI_ADD;
    const int lane = c.lane;
    const int quarter = c.warp & 3;                 // TMEM lane quarter this warp may read
    const int sub = (c.warp - EPI_WARP0) >> 2;      // which 64-column group of the tile this warp owns
    const int N = p.N, bn = p.bn;
    const int g_col0 = sub * GW;
    const int g_cols = g_col0 < bn ? min(GW, bn - g_col0) : 0;
    const int total_tiles = total_tiles_of<MODE>(p);
    unsigned char* stg = c.stg;
    const bool has_drop = (EPI == AB_EPI_BIAS_ACT || EPI == AB_EPI_DACT) && p.drop_seed != nullptr;
    const uint32_t drop_s0 = has_drop ? __ldg(p.drop_seed) : 0u, drop_s1 = has_drop ? __ldg(p.drop_seed + 1) : 0u;
    const float dscale = has_drop ? p.drop_scale : 1.0f;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = c.pair_id; tile < total_tiles; tile += c.num_pairs) {
        const Tile t = decode_tile<MODE>(p, tile);
        const int m_tile = 2 * t.m_pair + (int)c.rank;          // this CTA's 128 rows of the pair tile
        const bool have_acc = t.nk > 0;
        const int n0 = t.n_tile * bn;
        const int ncol_end = min(N, n0 + g_col0 + g_cols);      // columns past the tile (or the matrix) are never touched
        const size_t row0 = (size_t)m_tile * BM + quarter * 32;
        if (HAS_BIAS && g_cols > 0) {
            const int col = n0 + g_col0 + 2 * lane;
            float2 b2 = make_float2(0.f, 0.f);
            if (col < ncol_end) b2 = __ldg(reinterpret_cast<const float2*>(p.bias + (size_t)t.e * N + col));
            *reinterpret_cast<float2*>(c.bias_s + 2 * lane) = b2;
            __syncwarp();
        }
        uint4 nxt[4];
        if (HAS_AUX && g_cols > 0 && n0 + g_col0 < ncol_end) {
            load_aux_unit<F32>(p, nxt, lane, row0, n0 + g_col0, ncol_end);
        }
        if (have_acc) {
            ab_mbar_wait(&c.tfull[acc], acc_phase);
            ab_tc_fence_after();
        }
        const uint32_t t_row = c.tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
        for (int u0 = 0; u0 < g_cols; u0 += U) {
            const int tcol = g_col0 + u0;           // column inside the tile
            const int ncol = n0 + tcol;             // global column
            if (ncol >= ncol_end) break;
            uint4 cur[4];
            if (HAS_AUX) {
#pragma unroll
                for (int it = 0; it < 4; ++it) cur[it] = nxt[it];
                if (u0 + U < g_cols && ncol + U < ncol_end) {
                    load_aux_unit<F32>(p, nxt, lane, row0, ncol + U, ncol_end);
                }
            }
            float f[U];
            {
                uint32_t v[U];
                if (have_acc) {
                    if (U == 32) ab_tmem_ld32(t_row + (uint32_t)tcol, v);
                    else ab_tmem_ld16(t_row + (uint32_t)tcol, v);
                    ab_tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < U; ++i) v[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < U; ++i) f[i] = __uint_as_float(v[i]);
            }
            if (HAS_BIAS) {
#pragma unroll
                for (int i = 0; i < U; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(c.bias_s + u0 + i);      // zeros past the last column
                    f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
                }
            }
            if (EPI == AB_EPI_BIAS_ACT) {
                // the pre-activation (the Linear output, rounded to the activation dtype first) goes out before the
                // activation is applied in place, so only one fragment is live
                if (!F32) round_row_bf16<U>(f);
                stage_and_store<MODE, F32>(p, stg, f, reinterpret_cast<unsigned char*>(p.c2), lane, row0, ncol, ncol_end);
                act_fwd_row<U, F32>(f, p.act, p.drop_seed ? p.drop_scale : 1.0f);
                if (p.drop_seed) dropout_zero_row<U>(f, p.drop_thresh, drop_s0, drop_s1, (uint32_t)row0 + (uint32_t)lane, ncol, N);
            } else if (HAS_AUX) {
                // the tile fetched with coalesced accesses goes through the staging so that each thread reads its own row
#pragma unroll
                for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(stg + stg_off(it * 8 + (lane >> 2), lane & 3)) = cur[it];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) cur[j] = *reinterpret_cast<const uint4*>(stg + stg_off(lane, j));
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 q = cur[j];
                    float a[CPV];
                    if (F32) { a[0] = __uint_as_float(q.x); a[1] = __uint_as_float(q.y); a[2] = __uint_as_float(q.z); a[3] = __uint_as_float(q.w); }
                    else ab_vec16<__nv_bfloat16>::unpack(q, a);
                    if (EPI == AB_EPI_ADD) {
                        // C = acc + aux (aux has C's shape and dtype): a second gradient contribution accumulated here
#pragma unroll
                        for (int i = 0; i < CPV; ++i) f[j * CPV + i] += a[i];
                    } else if (p.act == AB_ACT_GELU) {
#pragma unroll
                        for (int i = 0; i < CPV; i += 2) {
                            const f2 gp = F32 ? f2_mul(gelu_bwd2(f2_pack(a[i], a[i + 1])), f2_bcast(dscale))
                                              : gelu_bwd_fast2(f2_pack(a[i], a[i + 1]), f2_bcast(0.5f * dscale));
                            f2_unpack(f2_mul(f2_pack(f[j * CPV + i], f[j * CPV + i + 1]), gp), f[j * CPV + i], f[j * CPV + i + 1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CPV; ++i) f[j * CPV + i] *= act_bwd(a[i], p.act) * dscale;
                    }
                }
                if (EPI == AB_EPI_DACT && p.drop_seed) dropout_zero_row<U>(f, p.drop_thresh, drop_s0, drop_s1, (uint32_t)row0 + (uint32_t)lane, ncol, N);
            }
            // ---- stage this thread's row, then store 8 rows x 64 B per instruction
            unsigned char* base;
            if (MODE == MODE_TN) base = reinterpret_cast<unsigned char*>(p.cw) + ((size_t)t.e * p.M) * N * 4;
            else if (p.peer_n > 0) {
                // the whole 128-row tile belongs to one source rank: its buffer, shifted so that `grow` indexes it directly
                const int dst = (int)(row0 / (size_t)p.peer_rows);
                base = p.peer_base[dst] + ((int64_t)(p.peer_rank - dst) * p.peer_rows) * (int64_t)N * ES;
            } else base = reinterpret_cast<unsigned char*>(p.c);
            stage_and_store<MODE, F32>(p, stg, f, base, lane, row0, ncol, ncol_end);
        }
        if (have_acc) {
            ab_tc_fence_before();
            __syncwarp();
            if (lane == 0) ab_mbar_arrive_cluster(&c.tempty[acc], 0);     // all of this warp's TMEM reads of the tile are done
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
}

template <int MODE, bool F32>
__device__ __forceinline__ void epilogue_dispatch(const GemmParams& p, const EpiCtx& c) {
    if (MODE == MODE_TN) { epilogue_role<MODE, F32, AB_EPI_NONE>(p, c); return; }
    switch (p.epi) {
        case AB_EPI_BIAS: epilogue_role<MODE, F32, AB_EPI_BIAS>(p, c); break;
        case AB_EPI_BIAS_ACT: epilogue_role<MODE, F32, AB_EPI_BIAS_ACT>(p, c); break;
        case AB_EPI_DACT: epilogue_role<MODE, F32, AB_EPI_DACT>(p, c); break;
        case AB_EPI_ADD: epilogue_role<MODE, F32, AB_EPI_ADD>(p, c); break;
        default: epilogue_role<MODE, F32, AB_EPI_NONE>(p, c); break;
    }
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) grouped_gemm_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                      const __grid_constant__ CUtensorMap tm_b,
                                                                      const GemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ab_smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);      // swizzle-128B atoms need 1024 B alignment
    unsigned char* stg_base = smem + (size_t)NSTAGE * STAGE_BYTES;
    float* bias_base = reinterpret_cast<float*>(stg_base + (size_t)EPI_WARPS * WARP_STG_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(bias_base) + (size_t)EPI_WARPS * WARP_BIAS_BYTES);
    uint64_t* full = bars;                       // leader's: TMA bytes of both CTAs
    uint64_t* empty = bars + NSTAGE;             // each CTA's own: multicast commit of the MMAs that read the slot
    uint64_t* tfull = bars + 2 * NSTAGE;         // each CTA's own: multicast commit of the tile's last MMA
    uint64_t* tempty = bars + 2 * NSTAGE + NACC; // leader's: epilogue warps of both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ab_cluster_ctarank();          // 0 = leader
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    if (warp == 0 && lane == 0) {
        ab_prefetch_tmap(&tm_a);
        ab_prefetch_tmap(&tm_b);
        for (int i = 0; i < NSTAGE; ++i) { ab_mbar_init(&full[i], 1); ab_mbar_init(&empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { ab_mbar_init(&tfull[i], 1); ab_mbar_init(&tempty[i], 2 * EPI_WARPS); }
        ab_fence_mbar_init();
    }
    if (warp == 1) {
        ab_tmem_alloc_pair(tmem_slot, 512);
        ab_tmem_relinquish_pair();
    }
    ab_tc_fence_before();
    ab_cluster_sync();           // both CTAs' barriers are initialised before any remote arrive / multicast commit
    ab_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = total_tiles_of<MODE>(p);
    const int bn = p.bn, bh = bn >> 1;          // this CTA stages bh columns of B
    const uint32_t bh_bytes = MODE == MODE_NT ? (uint32_t)bh * BK * 2 : (uint32_t)(bh / 64) * ATOM_BYTES;
    const uint32_t pair_tx = 2u * (A_BYTES + bh_bytes);

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own A rows, own half of B) =================
        // The whole warp walks the schedule (warp-uniform control flow keeps coordinates and addresses in uniform
        // registers); one elected lane issues.  This instruction stream is what feeds the tensor core: it stays lean.
        int stage = 0; uint32_t phase = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            const Tile t = decode_tile<MODE>(p, tile);
            const int b_col0 = t.n_tile * bn + (int)rank * bh;
            // MODE_TN walks the expert's row blocks source by source (no division per block); the other modes walk K
            const int n_src = MODE == MODE_TN ? p.tn_nsrc : 1;
            const int per_src = t.nk / n_src;
            const int m0 = MODE == MODE_TN ? t.m_pair * PM + (int)rank * BM : (2 * t.m_pair + (int)rank) * BM;
            const int b_row0 = MODE == MODE_NN ? t.e * p.K : (MODE == MODE_NT ? t.e * p.N + b_col0 : 0);
            for (int src = 0; src < n_src; ++src) {
                int r0 = MODE == MODE_TN ? (int)(src * p.tn_src_stride) + t.k_begin : 0;     // TN: first row of the block; NT / NN: k
                for (int kk = 0; kk < per_src; ++kk, r0 += BK) {
                    ab_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + (size_t)stage * STAGE_BYTES;
                    unsigned char* sb = sa + A_BYTES;
                    if (ab_elect_one()) {
                        if (rank == 0) ab_mbar_expect_tx(&full[stage], pair_tx);
                        if (MODE == MODE_TN) {
                            ab_tma_load_2d_pair(sa, &tm_a, &full[stage], m0, r0);
                            ab_tma_load_2d_pair(sa + ATOM_BYTES, &tm_a, &full[stage], m0 + 64, r0);
                            ab_tma_load_2d_pair(sb, &tm_b, &full[stage], b_col0, r0);
                            if (bh > 64) ab_tma_load_2d_pair(sb + ATOM_BYTES, &tm_b, &full[stage], b_col0 + 64, r0);
                        } else {
                            ab_tma_load_2d_pair(sa, &tm_a, &full[stage], r0, m0);
                            if (MODE == MODE_NT) {
                                ab_tma_load_2d_pair(sb, &tm_b, &full[stage], r0, b_row0);
                            } else {
                                ab_tma_load_2d_pair(sb, &tm_b, &full[stage], b_col0, b_row0 + r0);
                                if (bh > 64) ab_tma_load_2d_pair(sb + ATOM_BYTES, &tm_b, &full[stage], b_col0 + 64, b_row0 + r0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
        // drain: every slot this CTA filled has been released, i.e. no multicast arrival is still on its way here
        for (int i = 0; i < NSTAGE; ++i) {
            ab_mbar_wait(&empty[stage], phase ^ 1);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) =================
        if (rank == 0) {
            const uint32_t idesc = make_idesc(bn, MODE == MODE_TN ? 1 : 0, MODE == MODE_NT ? 0 : 1);
            // K-major operand: 8-row groups 1024 B apart; MN-major: 64-wide atoms 8 KB apart, 8-k groups 1024 B apart
            const uint32_t smem0 = ab_smem_u32(smem);
            const uint64_t da0 = MODE == MODE_TN ? make_smem_desc(smem0, ATOM_BYTES, 1024) : make_smem_desc(smem0, 0, 1024);
            const uint64_t db0 = MODE == MODE_NT ? make_smem_desc(smem0 + A_BYTES, 0, 1024) : make_smem_desc(smem0 + A_BYTES, ATOM_BYTES, 1024);
            constexpr uint32_t a_step = MODE == MODE_TN ? (16 * 128) >> 4 : 32 >> 4;     // advance 16 k per MMA
            constexpr uint32_t b_step = MODE == MODE_NT ? 32 >> 4 : (16 * 128) >> 4;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
                const Tile t = decode_tile<MODE>(p, tile);
                if (t.nk == 0) continue;            // empty expert (MODE_TN): the epilogue writes zeros without an accumulator
                ab_mbar_wait(&tempty[acc], acc_phase ^ 1);
                ab_tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < t.nk; ++kb) {
                    ab_mbar_wait(&full[stage], phase);
                    ab_tc_fence_after();
                    const uint64_t soff = (uint64_t)((uint32_t)stage * (STAGE_BYTES >> 4));      // the ring stays below 256 KB: no carry out of the address field
                    const uint64_t da = da0 + soff, db = db0 + soff;
                    if (ab_elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            ab_umma_f16_pair(d_tmem, da + (uint64_t)(k * a_step), db + (uint64_t)(k * b_step), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        ab_umma_commit_pair(&empty[stage], 3);              // the slot is free in both CTAs once these MMAs retire
                        if (kb == t.nk - 1) ab_umma_commit_pair(&tfull[acc], 3);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (both CTAs: own 128 rows of the accumulator) =================
        EpiCtx c;
        c.stg = stg_base + (size_t)(warp - EPI_WARP0) * WARP_STG_BYTES;
        c.bias_s = bias_base + (size_t)(warp - EPI_WARP0) * GW;
        c.tfull = tfull; c.tempty = tempty; c.tmem_base = tmem_base;
        c.warp = warp; c.lane = lane; c.rank = rank; c.pair_id = pair_id; c.num_pairs = num_pairs;
        if (p.c_f32) epilogue_dispatch<MODE, true>(p, c);
        else epilogue_dispatch<MODE, false>(p, c);
    }
    ab_tc_fence_before();
    ab_cluster_sync();           // neither CTA leaves (or frees tensor memory) while its partner may still signal it
    if (warp == 1) ab_tmem_dealloc_pair(tmem_base, 512);
}

int pick_bn(int N, bool mn_major_b) {
    // each CTA of the pair stages bn/2 columns: 8-row groups (K-major B) or whole 64-column atoms (MN-major B)
    const int step = mn_major_b ? 128 : 16;
    int best = step, best_cost = 1 << 30;
    for (int bn = step; bn <= 256; bn += step) {
        const int cost = (int)ab_ceil_div(N, bn) * (bn + 48);
        if (cost <= best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

int make_map2(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
    AB_REQUIRE(((uintptr_t)base % 16) == 0 && (inner * 2) % 16 == 0,
               "grouped_gemm: operand base must be 16-byte aligned and the contiguous dimension (%llu) a multiple of 8",
               (unsigned long long)inner);
    uint64_t dims[2] = {inner, outer};
    uint64_t strides[1] = {inner * 2};
    uint32_t box[2] = {box_inner, box_outer};
    return ab_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// CTA pairs that can be resident at once (one per TPC when every TPC is whole); queried once per kernel and device
template <int MODE>
int resident_pairs(int* out) {
    static int cached[64] = {0};
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && cached[dev] > 0) { *out = cached[dev]; return AB_OK; }
    auto k = grouped_gemm_kernel<MODE>;
    AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (unsigned)(ab_num_sms() / 2), 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    AB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, k, &cfg));
    if (n < 1) n = 1;
    if (n > ab_num_sms() / 2) n = ab_num_sms() / 2;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    *out = n;
    return AB_OK;
}

template <int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int64_t max_tiles, cudaStream_t stream) {
    int pairs = 0;
    if (int e = resident_pairs<MODE>(&pairs)) return e;
    if (max_tiles < pairs) pairs = (int)max_tiles;
    if (pairs < 1) pairs = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (unsigned)pairs, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    AB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, grouped_gemm_kernel<MODE>, ta, tb, p));
    return AB_OK;
}

int gemm_rows(int mode, const void* A, const void* W, const float* bias, const void* aux, void* c, void* c2,
              const int32_t* tile_expert, const int32_t* n_rows, int64_t max_rows, int N, int K, int E, int epi, int act,
              int c_dtype, float drop_p, const uint32_t* drop_seed, cudaStream_t stream, const uint64_t* peer_c = nullptr,
              int peer_w = 0, int peer_rank = 0, int64_t peer_rows = 0) {
    AB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "grouped_gemm: dropout probability must be in [0, 1)");
    AB_REQUIRE(drop_p == 0.f || (drop_seed != nullptr && (epi == AB_EPI_BIAS_ACT || epi == AB_EPI_DACT)),
               "grouped_gemm: dropout needs a seed and the bias+act / dact epilogue");
    AB_REQUIRE(max_rows > 0 && max_rows % PM == 0, "grouped_gemm: max_rows must be a positive multiple of %d", PM);
    AB_REQUIRE(N > 0 && K > 0 && E > 0 && N % 8 == 0 && K % 8 == 0, "grouped_gemm: N (%d) and K (%d) must be multiples of 8", N, K);
    AB_REQUIRE(c_dtype == AB_F32 || c_dtype == AB_BF16, "grouped_gemm: bad output dtype");
    AB_REQUIRE(epi >= AB_EPI_NONE && epi <= AB_EPI_ADD, "grouped_gemm: bad epilogue %d", epi);
    AB_REQUIRE((epi != AB_EPI_BIAS && epi != AB_EPI_BIAS_ACT) || bias, "grouped_gemm: bias epilogue without bias");
    AB_REQUIRE(epi != AB_EPI_BIAS_ACT || c2, "grouped_gemm: bias+act epilogue needs the pre-activation output c2");
    AB_REQUIRE((epi != AB_EPI_DACT && epi != AB_EPI_ADD) || aux, "grouped_gemm: dact / add epilogue needs aux");
    AB_REQUIRE(bias == nullptr || ((uintptr_t)bias % 8) == 0, "grouped_gemm: bias must be 8-byte aligned");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.E = E;
    p.bn = pick_bn(N, mode == MODE_NN);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.epi = epi; p.act = act; p.c_f32 = c_dtype == AB_F32;
    p.tile_expert = tile_expert; p.n_rows = n_rows; p.bias = bias; p.aux = aux; p.c = c; p.c2 = c2;
    p.rows_valid = max_rows;
    if (peer_c != nullptr) {
        AB_REQUIRE(peer_w >= 1 && peer_w <= 16 && peer_rank >= 0 && peer_rank < peer_w && peer_rows > 0 && peer_rows % PM == 0 &&
                   (int64_t)peer_w * peer_rows == max_rows && epi != AB_EPI_BIAS_ACT,
                   "ep_grouped_gemm: bad peer table (W=%d rank=%d rows_per_peer=%lld, max_rows=%lld)", peer_w, peer_rank,
                   (long long)peer_rows, (long long)max_rows);
        for (int i = 0; i < peer_w; ++i) {
            AB_REQUIRE(peer_c[i] != 0 && peer_c[i] % 16 == 0, "ep_grouped_gemm: peer buffer %d is null or misaligned", i);
            p.peer_base[i] = reinterpret_cast<unsigned char*>(peer_c[i]);
        }
        p.peer_n = peer_w; p.peer_rows = (int)peer_rows; p.peer_rank = peer_rank;
        c = reinterpret_cast<void*>(peer_c[peer_rank]);          // for the alignment check below
        p.c = c;
    }
    if (drop_p > 0.f) {
        p.drop_seed = drop_seed;
        p.drop_thresh = (uint32_t)((double)drop_p * 65536.0 + 0.5);
        p.drop_scale = 1.0f / (1.0f - drop_p);
    }
    AB_REQUIRE(((uintptr_t)c % 16) == 0 && (c2 == nullptr || ((uintptr_t)c2 % 16) == 0) && (aux == nullptr || ((uintptr_t)aux % 16) == 0),
               "grouped_gemm: output / aux pointers must be 16-byte aligned");
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)K, (uint64_t)max_rows, BK, BM)) return e;
    const int64_t max_tiles = (max_rows / PM) * p.num_n_tiles;
    if (mode == MODE_NT) {
        if (int e = make_map2(&tb, W, (uint64_t)K, (uint64_t)E * N, BK, (uint32_t)(p.bn / 2))) return e;
        return launch<MODE_NT>(ta, tb, p, max_tiles, stream);
    }
    if (int e = make_map2(&tb, W, (uint64_t)N, (uint64_t)E * K, 64, BK)) return e;
    return launch<MODE_NN>(ta, tb, p, max_tiles, stream);
}

}  // namespace

extern "C" int ab_gemm_row_tile(void) { return PM; }



extern "C" int ab_grouped_gemm_nt(const void* A, const void* W, const float* bias, const void* aux, void* c, void* c2,
                                  const int32_t* tile_expert, const int32_t* n_rows, int64_t max_rows, int N, int K, int E,
                                  int epi, int act, int c_dtype, float drop_p, const uint32_t* drop_seed, cudaStream_t stream) {
    return gemm_rows(MODE_NT, A, W, bias, aux, c, c2, tile_expert, n_rows, max_rows, N, K, E, epi, act, c_dtype, drop_p, drop_seed, stream);
}

extern "C" int ab_grouped_gemm_nn(const void* A, const void* W, const float* bias, const void* aux, void* c, void* c2,
                                  const int32_t* tile_expert, const int32_t* n_rows, int64_t max_rows, int N, int K, int E,
                                  int epi, int act, int c_dtype, float drop_p, const uint32_t* drop_seed, cudaStream_t stream) {
    return gemm_rows(MODE_NN, A, W, bias, aux, c, c2, tile_expert, n_rows, max_rows, N, K, E, epi, act, c_dtype, drop_p, drop_seed, stream);
}

// expert parallelism: the same GEMMs with the result rows written straight into the SOURCE ranks' buffers over NVLink (the
// return exchange of the expert-parallel MoE fused into the epilogue); see the ab_ep_* section of the header
extern "C" int ab_ep_grouped_gemm_nt(const void* A, const void* W, const float* bias, const void* aux, const uint64_t* peer_c, int peer_w,
                                     int rank, int64_t rows_per_peer, const int32_t* tile_expert, const int32_t* n_rows,
                                     int64_t max_rows, int N, int K, int E, int epi, int act, int c_dtype, cudaStream_t stream) {
    AB_REQUIRE(peer_c != nullptr, "ep_grouped_gemm_nt: no peer table");
    return gemm_rows(MODE_NT, A, W, bias, aux, nullptr, nullptr, tile_expert, n_rows, max_rows, N, K, E, epi, act, c_dtype, 0.f, nullptr,
                     stream, peer_c, peer_w, rank, rows_per_peer);
}

extern "C" int ab_ep_grouped_gemm_nn(const void* A, const void* W, const float* bias, const void* aux, const uint64_t* peer_c, int peer_w,
                                     int rank, int64_t rows_per_peer, const int32_t* tile_expert, const int32_t* n_rows,
                                     int64_t max_rows, int N, int K, int E, int epi, int act, int c_dtype, cudaStream_t stream) {
    AB_REQUIRE(peer_c != nullptr, "ep_grouped_gemm_nn: no peer table");
    return gemm_rows(MODE_NN, A, W, bias, aux, nullptr, nullptr, tile_expert, n_rows, max_rows, N, K, E, epi, act, c_dtype, 0.f, nullptr,
                     stream, peer_c, peer_w, rank, rows_per_peer);
}

namespace {
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int nsplit);
// Few output tiles and many source blocks (expert parallelism with one or two local experts): the contraction is cut over
// the source blocks so that every CTA pair has work; slices = the largest divisor of nsrc that still fits one wave.
int grouped_tn_slices(int M, int N, int E, int nsrc) {
    const int64_t tiles = (int64_t)E * ab_ceil_div(M, PM) * ab_ceil_div(N, pick_bn(N, true));
    const int pairs = ab_num_sms() / 2;
    // time ~ waves(s) / s; a finer cut must win at least 12 % (it also costs the sum of the partial products)
    int best = 1;
    double best_cost = (double)ab_ceil_div(tiles, pairs);
    for (int s = 2; s <= nsrc && s <= 4; ++s) {
        if (nsrc % s) continue;
        const double cost = (double)ab_ceil_div(tiles * s, pairs) / s;
        if (cost < 0.88 * best_cost) { best = s; best_cost = cost; }
    }
    return best;
}
}  // namespace

extern "C" size_t ab_grouped_gemm_tn_workspace_bytes(int M, int N, int E, int nsrc) {
    const int slices = nsrc > 1 ? grouped_tn_slices(M, N, E, nsrc) : 1;
    return slices > 1 ? (size_t)slices * E * M * N * sizeof(float) : 0;
}

extern "C" int ab_grouped_gemm_tn(const void* A, const void* Bm, float* Cw, const int32_t* seg_off, int64_t max_rows, int M,
                                  int N, int E, int nsrc, int64_t src_stride, void* ws, size_t ws_bytes, cudaStream_t stream) {
    AB_REQUIRE(nsrc >= 1 && (nsrc == 1 || (src_stride > 0 && src_stride % BK == 0 && nsrc * src_stride <= max_rows)),
               "grouped_gemm_tn: bad source blocking nsrc=%d stride=%lld", nsrc, (long long)src_stride);
    AB_REQUIRE(max_rows > 0 && max_rows % BK == 0, "grouped_gemm_tn: max_rows must be a positive multiple of %d", BK);
    AB_REQUIRE(M > 0 && N > 0 && E > 0 && M % 8 == 0 && N % 8 == 0, "grouped_gemm_tn: M (%d) and N (%d) must be multiples of 8", M, N);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.M = M; p.E = E; p.K = 0;
    p.bn = pick_bn(N, true);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.num_m_tiles = (int)ab_ceil_div(M, PM);
    p.seg_off = seg_off; p.c_f32 = 1; p.epi = AB_EPI_NONE; p.cw = Cw;
    p.tn_nsrc = nsrc; p.tn_src_stride = src_stride;
    AB_REQUIRE(((uintptr_t)Cw % 16) == 0 && ((uintptr_t)ws % 16) == 0, "grouped_gemm_tn: output / workspace must be 16-byte aligned");
    int slices = nsrc > 1 ? grouped_tn_slices(M, N, E, nsrc) : 1;
    if (slices > 1 && (ws == nullptr || ws_bytes < (size_t)slices * E * M * N * sizeof(float))) slices = 1;     // no workspace: unsplit
    if (slices > 1) {
        p.tn_e_real = E; p.E = E * slices; p.tn_nsrc = nsrc / slices;
        p.cw = reinterpret_cast<float*>(ws);                 // [slice][expert][M][N] partial products
    }
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)M, (uint64_t)max_rows, 64, BK)) return e;
    if (int e = make_map2(&tb, Bm, (uint64_t)N, (uint64_t)max_rows, 64, BK)) return e;
    if (int e = launch<MODE_TN>(ta, tb, p, (int64_t)p.E * p.num_m_tiles * p.num_n_tiles, stream)) return e;
    if (slices > 1) {
        const int64_t n = (int64_t)E * M * N;
        splitk_reduce_kernel<<<(unsigned)ab_ceil_div(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(ws), Cw, n, slices);
        AB_LAUNCH_CHECK();
    }
    return AB_OK;
}

// ---- dense GEMMs on the same kernel (one "expert", plain row-major matrices): the SSM layer's projections
//      in_proj_x | in_proj_z, x_param_proj (+ dt_proj_head folded in), out_proj (core.py:366-367, 376-383, 397) and their
//      autograd.  Rows need not be a multiple of the tile: TMA zero-fills what it reads past the matrix, stores are masked.
namespace {
int dense_rows(int mode, const void* A, const void* W, const float* bias, const void* aux, void* c, int64_t S, int N, int K, int epi,
               int c_dtype, cudaStream_t stream) {
    AB_REQUIRE(S > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "dense_gemm: S (%lld) must be positive, N (%d) and K (%d) multiples of 8", (long long)S, N, K);
    AB_REQUIRE(c_dtype == AB_F32 || c_dtype == AB_BF16, "dense_gemm: bad output dtype");
    AB_REQUIRE(epi == AB_EPI_NONE || epi == AB_EPI_BIAS || epi == AB_EPI_ADD, "dense_gemm: epilogue must be none, bias or add");
    AB_REQUIRE(epi != AB_EPI_BIAS || bias, "dense_gemm: bias epilogue without bias");
    AB_REQUIRE(epi != AB_EPI_ADD || aux, "dense_gemm: add epilogue without addend");
    AB_REQUIRE(((uintptr_t)c % 16) == 0 && (aux == nullptr || ((uintptr_t)aux % 16) == 0), "dense_gemm: output / addend must be 16-byte aligned");
    AB_REQUIRE(bias == nullptr || ((uintptr_t)bias % 8) == 0, "dense_gemm: bias must be 8-byte aligned");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.E = 1;
    p.bn = pick_bn(N, mode == MODE_NN);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.epi = epi; p.c_f32 = c_dtype == AB_F32;
    p.bias = bias; p.aux = aux; p.c = c;
    p.rows_valid = S;
    p.dense_m_tiles = (int)ab_ceil_div(S, PM);
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)K, (uint64_t)S, BK, BM)) return e;
    const int64_t tiles = (int64_t)p.dense_m_tiles * p.num_n_tiles;
    if (mode == MODE_NT) {
        if (int e = make_map2(&tb, W, (uint64_t)K, (uint64_t)N, BK, (uint32_t)(p.bn / 2))) return e;
        return launch<MODE_NT>(ta, tb, p, tiles, stream);
    }
    if (int e = make_map2(&tb, W, (uint64_t)N, (uint64_t)K, 64, BK)) return e;
    return launch<MODE_NN>(ta, tb, p, tiles, stream);
}
}  // namespace

extern "C" int ab_dense_gemm_nt(const void* A, const void* W, const float* bias, const void* aux, void* C, int64_t S, int N, int K,
                                int epi, int c_dtype, cudaStream_t stream) {
    return dense_rows(MODE_NT, A, W, bias, aux, C, S, N, K, epi, c_dtype, stream);
}

extern "C" int ab_dense_gemm_nn(const void* A, const void* W, const float* bias, const void* aux, void* C, int64_t S, int N, int K,
                                int epi, int c_dtype, cudaStream_t stream) {
    return dense_rows(MODE_NN, A, W, bias, aux, C, S, N, K, epi, c_dtype, stream);
}

// sum of the split-K partial products, fixed order (bitwise reproducible)
namespace {
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int nsplit) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float4 acc = *reinterpret_cast<const float4*>(part + i);
    for (int s = 1; s < nsplit; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(part + (size_t)s * n + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = acc;
}

int dense_tn_nsplit(int64_t S, int M, int N) {
    const int tiles = (int)(ab_ceil_div(M, PM) * ab_ceil_div(N, pick_bn(N, true)));
    int nsplit = tiles > 0 ? (ab_num_sms() / 2) / tiles : 1;        // one wave of (slice, tile) work items over the CTA pairs
    const int64_t kblocks = ab_ceil_div(S, BK);
    if (nsplit > kblocks / 4) nsplit = (int)(kblocks / 4);          // at least 4 blocks of 64 rows per slice
    return nsplit < 2 ? 1 : nsplit;
}
}  // namespace

extern "C" size_t ab_dense_gemm_tn_workspace_bytes(int64_t S, int M, int N) {
    const int nsplit = dense_tn_nsplit(S, M, N);
    return nsplit < 2 ? 0 : (size_t)nsplit * M * N * sizeof(float);
}

extern "C" int ab_dense_gemm_tn(const void* A, const void* Bm, float* Cw, void* ws, size_t ws_bytes, int64_t S, int M, int N,
                                cudaStream_t stream) {
    AB_REQUIRE(S > 0 && M > 0 && N > 0 && M % 8 == 0 && N % 8 == 0, "dense_gemm_tn: S (%lld) must be positive, M (%d) and N (%d) multiples of 8", (long long)S, M, N);
    AB_REQUIRE(((uintptr_t)Cw % 16) == 0 && ((uintptr_t)ws % 16) == 0, "dense_gemm_tn: output / workspace must be 16-byte aligned");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.M = M; p.K = 0;
    p.bn = pick_bn(N, true);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.num_m_tiles = (int)ab_ceil_div(M, PM);
    p.c_f32 = 1; p.epi = AB_EPI_NONE;
    p.tn_nsrc = 1; p.tn_rows = S;
    // the contraction runs over all S rows and the output has only a few tiles: cut S into slices (split-K) so that the
    // whole chip works, partial products into the workspace, then one fixed-order sum
    int nsplit = dense_tn_nsplit(S, M, N);
    if (nsplit >= 2 && (ws == nullptr || ws_bytes < (size_t)nsplit * M * N * sizeof(float))) nsplit = 1;
    const int64_t kblocks = ab_ceil_div(S, BK);
    p.E = nsplit;
    p.tn_split = ab_ceil_div(kblocks, nsplit) * BK;
    p.cw = nsplit > 1 ? reinterpret_cast<float*>(ws) : Cw;
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)M, (uint64_t)S, 64, BK)) return e;
    if (int e = make_map2(&tb, Bm, (uint64_t)N, (uint64_t)S, 64, BK)) return e;
    if (int e = launch<MODE_TN>(ta, tb, p, (int64_t)nsplit * p.num_m_tiles * p.num_n_tiles, stream)) return e;
    if (nsplit > 1) {
        const int64_t n = (int64_t)M * N;
        splitk_reduce_kernel<<<(unsigned)ab_ceil_div(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(ws), Cw, n, nsplit);
        AB_LAUNCH_CHECK();
    }
    return AB_OK;
}
