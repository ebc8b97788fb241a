// Grouped expert GEMM on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA into 128B-swizzled shared memory, warp-specialised and persistent.
// Replaces the per-(slot, expert) nn.Linear calls of core.py:596 (experts[e] = LN, Linear, act, Dropout,
// Linear; core.py:434-442) and their autograd (dgrad / wgrad).
//
//   MODE_NT  C[r,n] = epi(sum_k A[r,k] * W[e,n,k])      A [rows,K] K-major,   W [E*N, K] K-major   (forward)
//   MODE_NN  C[r,n] = epi(sum_k A[r,k] * W[e,k,n])      A [rows,K] K-major,   W [E*K, N] N-major   (dgrad)
//   MODE_TN  Cw[e,m,n] = sum_{r in seg e} A[r,m]*B[r,n] A [rows,M] M-major,   B [rows,N] N-major   (wgrad)
//
// Tile 128 x BN x 64 (BN <= 256 chosen per shape on the host and carried in the TMA maps / the
// instruction descriptor), 4-stage TMA->MMA ring, 2 accumulator stages in TMEM (2 x 256 columns) so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM
// alloc), 2..17 = epilogue: four warps per TMEM lane quarter, each taking one 16-column chunk of a 64-column
// group of the tile (tcgen05.ld -> bias / activation math in registers -> swizzled per-warp staging -> coalesced
// 16-byte global stores; no CTA-wide synchronisation in the epilogue).
// Row tiles (128 permuted rows) belong to one expert; the number of valid row tiles is read from device
// memory (n_rows[0]) so no host synchronisation is needed after the routing plan.
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64, NSTAGE = 4, NACC = 2;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES_MAX = 256 * BK * 2;     // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int ATOM_BYTES = 64 * BK * 2;       // one 64(MN) x 64(K) swizzle-128B MN-major atom = 8 KB
constexpr int EPI_WARP0 = 2, EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + EPI_THREADS; // warp 0 TMA, warp 1 MMA, 16 epilogue warps
constexpr int GW = 64;                        // epilogue column group: 4 chunks of 16, one per warp of a TMEM lane quarter
constexpr int MODE_NT = 0, MODE_NN = 1, MODE_TN = 2;
constexpr int WARP_STG_BYTES = 32 * 64;        // per-warp epilogue staging: 32 rows x 64 B
constexpr size_t SMEM_BYTES = 1024 + (size_t)NSTAGE * STAGE_BYTES + (size_t)EPI_WARPS * WARP_STG_BYTES + 256;

struct GemmParams {
    int N, K, E, M;          // M only for MODE_TN (rows of each expert's output)
    int bn;                  // N tile
    int num_n_tiles, num_m_tiles;
    int epi, act, c_f32;
    int tn_nsrc;             // MODE_TN: an expert's rows are nsrc blocks [src*stride + seg_off[e], src*stride + seg_off[e+1])
    int64_t tn_src_stride;
    // dense use (one "expert", plain row-major matrices; tile_expert / n_rows / seg_off are NULL):
    int64_t rows_valid;      // MODE_NT / MODE_NN: rows >= rows_valid are never stored (the last row tile may be partial)
    int dense_m_tiles;       // MODE_NT / MODE_NN: number of row tiles when n_rows is NULL
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
};

// dense weight gradient: 64-row blocks of slice e of the contraction (the last block of the last slice may be partial: TMA zero-fills)
__device__ __forceinline__ int dense_tn_blocks(const GemmParams& p, int e) {
    const int64_t lo = (int64_t)e * p.tn_split;
    const int64_t hi = lo + p.tn_split < p.tn_rows ? lo + p.tn_split : p.tn_rows;
    return hi > lo ? (int)((hi - lo + BK - 1) / BK) : 0;
}

// ---- math for the epilogues -------------------------------------------------------------------
// erf with |error| < 1.5e-7 (Abramowitz & Stegun 7.1.26): erf(x/sqrt2) from t = 1/(1+p|x|/sqrt2) and g = exp(-x^2/2)
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

// ---- the same activations on column pairs (FFMA2 / FMUL2): the epilogue is instruction-issue bound at small K, and packed
//      arithmetic halves its floating-point instruction count ----------------------------------------------------------
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
template <int U>
__device__ __forceinline__ void act_fwd_row(float (&f)[U], int act) {
    if (act == AB_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < U; i += 2) f2_unpack(gelu_fwd2(f2_pack(f[i], f[i + 1])), f[i], f[i + 1]);
    } else if (act == AB_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < U; ++i) f[i] = fmaxf(f[i], 0.f);
    } else {
#pragma unroll
        for (int i = 0; i < U; ++i) f[i] = act_fwd(f[i], AB_ACT_SILU);
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
// dropout on a row: one hash per 4 columns, the keep flags become multipliers (0 or 1 / (1 - p)) applied pairwise
template <int U>
__device__ __forceinline__ void dropout_row(float (&f)[U], const GemmParams& p, uint32_t grow, int ncol);

// Dropout keep-mask of the expert hidden activation: counter-based hash of the element index and a 64-bit seed, so the
// backward regenerates exactly the forward's mask without storing it.  One hash serves 4 consecutive columns (four
// 16-bit uniforms, drop when < p * 65536), ~6 integer ops per element.
__device__ __forceinline__ void drop_mask4(uint32_t s0, uint32_t s1, uint32_t row, uint32_t col, uint32_t n_cols, uint32_t thresh16,
                                           bool (&keep)[4]) {
    uint32_t x = ((row * n_cols + col) >> 2) * 0x9E3779B1u + s0;
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    uint32_t y = x * 0x27D4EB2Fu + s1;
    y ^= y >> 15;
    keep[0] = (x & 0xffffu) >= thresh16;
    keep[1] = (x >> 16) >= thresh16;
    keep[2] = (y & 0xffffu) >= thresh16;
    keep[3] = (y >> 16) >= thresh16;
}

template <int U>
__device__ __forceinline__ void dropout_row(float (&f)[U], const GemmParams& p, uint32_t grow, int ncol) {
    const uint32_t s0 = __ldg(p.drop_seed), s1 = __ldg(p.drop_seed + 1);
#pragma unroll
    for (int i = 0; i < U; i += 4) {
        bool keep[4];
        drop_mask4(s0, s1, grow, (uint32_t)(ncol + i), (uint32_t)p.N, p.drop_thresh, keep);
        const f2 m0 = f2_pack(keep[0] ? p.drop_scale : 0.f, keep[1] ? p.drop_scale : 0.f);
        const f2 m1 = f2_pack(keep[2] ? p.drop_scale : 0.f, keep[3] ? p.drop_scale : 0.f);
        f2_unpack(f2_mul(f2_pack(f[i], f[i + 1]), m0), f[i], f[i + 1]);
        f2_unpack(f2_mul(f2_pack(f[i + 2], f[i + 3]), m1), f[i + 2], f[i + 3]);
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
// instruction descriptor: bf16 x bf16 -> f32, M = 128
__host__ __device__ inline uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format f32
    d |= 1u << 7;                    // A format bf16
    d |= 1u << 10;                   // B format bf16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

// ---- epilogue -----------------------------------------------------------------------------------
// One warp owns 32 accumulator rows (its TMEM lane quarter) x one 64-column group and walks it in units of 64 output
// bytes per row (32 bf16 or 16 f32 columns): tcgen05.ld -> math -> swizzled per-warp staging -> coalesced 16-byte
// global stores (8 rows x 64 B per instruction).  No CTA-wide synchronisation; everything is compile-time indexed so the
// fragments stay in registers.
__device__ __forceinline__ uint32_t stg_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

template <int MODE, bool F32>
__device__ __forceinline__ void stage_and_store(const GemmParams& p, unsigned char* stg, const float (&src)[F32 ? 16 : 32],
                                                unsigned char* base, int lane, int quarter, int m_tile, int ncol, int ncol_end) {
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
        const size_t grow = (size_t)m_tile * BM + quarter * 32 + r;
        const bool ok = col < ncol_end && (MODE == MODE_TN ? (int)grow < p.M : (int64_t)grow < p.rows_valid);
        const uint4 q = *reinterpret_cast<const uint4*>(stg + stg_off(r, j));
        if (ok) *reinterpret_cast<uint4*>(base + (grow * N + col) * ES) = q;
    }
    __syncwarp();
}

template <int MODE, bool F32>
__device__ __forceinline__ void epilogue_group(const GemmParams& p, unsigned char* stg, uint32_t t_row, bool have_acc, int lane,
                                               int quarter, int m_tile, int e, int n0, int g_col0, int g_cols) {
    constexpr int U = F32 ? 16 : 32;            // columns per unit
    constexpr int ES = F32 ? 4 : 2;
    constexpr int CPV = 16 / ES;                // columns per 16-byte vector
    const int N = p.N;
    const int ncol_end = min(N, n0 + g_col0 + g_cols);    // columns past the tile (or the matrix) are never touched
    for (int u0 = 0; u0 < g_cols; u0 += U) {
        const int tcol = g_col0 + u0;           // column inside the tile
        const int ncol = n0 + tcol;             // global column
        if (ncol >= ncol_end) break;
        float f[U];
        {
            uint32_t v[U];
            if (have_acc) {
                ab_tmem_ld16(t_row + (uint32_t)tcol, v);
                if (U == 32) ab_tmem_ld16(t_row + (uint32_t)tcol + 16u, v + 16);
                ab_tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < U; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < U; ++i) f[i] = __uint_as_float(v[i]);
        }
        if (MODE != MODE_TN) {
            if (p.epi == AB_EPI_BIAS || p.epi == AB_EPI_BIAS_ACT) {
                const float* bp = p.bias + (size_t)e * N + ncol;
#pragma unroll
                for (int i = 0; i < U; i += 4) {
                    if (ncol + i < ncol_end) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
                        f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
                    }
                }
            }
            if (p.epi == AB_EPI_BIAS_ACT) {
                // the pre-activation (the Linear output, rounded to the activation dtype first) goes out before the
                // activation is applied in place, so only one fragment is live
                if (!F32) round_row_bf16<U>(f);
                stage_and_store<MODE, F32>(p, stg, f, reinterpret_cast<unsigned char*>(p.c2), lane, quarter, m_tile, ncol, ncol_end);
                act_fwd_row<U>(f, p.act);
                if (p.drop_seed) dropout_row<U>(f, p, (uint32_t)(m_tile * BM + quarter * 32 + lane), ncol);
            } else if (p.epi == AB_EPI_ADD) {
                // C = acc + aux (aux has C's shape and dtype): accumulate a second gradient contribution in the epilogue
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + (lane >> 2), j = lane & 3;
                    const int col = ncol + j * CPV;
                    const int64_t grow = (int64_t)m_tile * BM + quarter * 32 + r;
                    uint4 q = make_uint4(0u, 0u, 0u, 0u);
                    if (col < ncol_end && grow < p.rows_valid)
                        q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.aux) + ((size_t)grow * N + col) * ES));
                    *reinterpret_cast<uint4*>(stg + stg_off(r, j)) = q;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4*>(stg + stg_off(lane, j));
                    float add[CPV];
                    if (F32) { add[0] = __uint_as_float(q.x); add[1] = __uint_as_float(q.y); add[2] = __uint_as_float(q.z); add[3] = __uint_as_float(q.w); }
                    else ab_vec16<__nv_bfloat16>::unpack(q, add);
#pragma unroll
                    for (int i = 0; i < CPV; ++i) f[j * CPV + i] += add[i];
                }
                __syncwarp();
            } else if (p.epi == AB_EPI_DACT) {
                // saved pre-activation tile: coalesced 16-byte loads -> staging -> each thread reads its own row
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + (lane >> 2), j = lane & 3;
                    const int col = ncol + j * CPV;
                    uint4 q = make_uint4(0u, 0u, 0u, 0u);
                    if (col < ncol_end)
                        q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.aux) +
                                                                 (((size_t)m_tile * BM + quarter * 32 + r) * N + col) * ES));
                    *reinterpret_cast<uint4*>(stg + stg_off(r, j)) = q;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4*>(stg + stg_off(lane, j));
                    float pre[CPV];
                    if (F32) { pre[0] = __uint_as_float(q.x); pre[1] = __uint_as_float(q.y); pre[2] = __uint_as_float(q.z); pre[3] = __uint_as_float(q.w); }
                    else ab_vec16<__nv_bfloat16>::unpack(q, pre);
                    if (p.act == AB_ACT_GELU) {
#pragma unroll
                        for (int i = 0; i < CPV; i += 2) {
                            const f2 d = f2_mul(f2_pack(f[j * CPV + i], f[j * CPV + i + 1]), gelu_bwd2(f2_pack(pre[i], pre[i + 1])));
                            f2_unpack(d, f[j * CPV + i], f[j * CPV + i + 1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CPV; ++i) f[j * CPV + i] *= act_bwd(pre[i], p.act);
                    }
                }
                __syncwarp();
                if (p.drop_seed) dropout_row<U>(f, p, (uint32_t)(m_tile * BM + quarter * 32 + lane), ncol);
            }
        }
        // ---- stage this thread's row, then store 8 rows x 64 B per instruction
        unsigned char* base;
        if (MODE == MODE_TN) base = reinterpret_cast<unsigned char*>(p.cw) + ((size_t)e * p.M) * N * 4;
        else base = reinterpret_cast<unsigned char*>(p.c);
        stage_and_store<MODE, F32>(p, stg, f, base, lane, quarter, m_tile, ncol, ncol_end);
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + (size_t)EPI_WARPS * WARP_STG_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* tfull = bars + 2 * NSTAGE;
    uint64_t* tempty = bars + 2 * NSTAGE + NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ab_prefetch_tmap(&tm_a);
        ab_prefetch_tmap(&tm_b);
        for (int i = 0; i < NSTAGE; ++i) { ab_mbar_init(&full[i], 1); ab_mbar_init(&empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { ab_mbar_init(&tfull[i], 1); ab_mbar_init(&tempty[i], EPI_WARPS); }
        ab_fence_mbar_init();
    }
    if (warp == 1) {
        ab_tmem_alloc(tmem_slot, 512);
        ab_tmem_relinquish();
    }
    ab_tc_fence_before();
    __syncthreads();
    ab_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- tile schedule (identical in every role)
    int total_tiles;
    if (MODE == MODE_TN) total_tiles = p.E * p.num_m_tiles * p.num_n_tiles;
    else total_tiles = (p.n_rows != nullptr ? p.n_rows[0] / BM : p.dense_m_tiles) * p.num_n_tiles;
    const int bn = p.bn;
    const uint32_t b_bytes = MODE == MODE_NT ? (uint32_t)bn * BK * 2 : (uint32_t)(bn / 64) * ATOM_BYTES;
    const uint32_t stage_tx = A_BYTES + b_bytes;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int m_tile, n_tile, e, k_begin, nk;
                if (MODE == MODE_TN) {
                    e = tile / (p.num_m_tiles * p.num_n_tiles);
                    const int rem = tile % (p.num_m_tiles * p.num_n_tiles);
                    m_tile = rem / p.num_n_tiles; n_tile = rem % p.num_n_tiles;
                    k_begin = p.seg_off != nullptr ? p.seg_off[e] : (int)(e * p.tn_split);
                    nk = p.seg_off != nullptr ? p.tn_nsrc * ((p.seg_off[e + 1] - k_begin) / BK) : dense_tn_blocks(p, e);
                } else {
                    m_tile = tile / p.num_n_tiles; n_tile = tile % p.num_n_tiles;
                    e = p.tile_expert != nullptr ? p.tile_expert[m_tile] : 0;
                    k_begin = 0;
                    nk = (p.K + BK - 1) / BK;
                }
                for (int kb = 0; kb < nk; ++kb) {
                    ab_mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + (size_t)stage * STAGE_BYTES;
                    unsigned char* sb = sa + A_BYTES;
                    ab_mbar_expect_tx(&full[stage], stage_tx);
                    if (MODE == MODE_TN) {
                        const int per_src = nk / p.tn_nsrc;
                        const int r0 = (int)((kb / per_src) * p.tn_src_stride) + k_begin + (kb % per_src) * BK;
                        ab_tma_load_2d(sa, &tm_a, &full[stage], m_tile * BM, r0);
                        ab_tma_load_2d(sa + ATOM_BYTES, &tm_a, &full[stage], m_tile * BM + 64, r0);
                        for (int j = 0; j < bn / 64; ++j)
                            ab_tma_load_2d(sb + j * ATOM_BYTES, &tm_b, &full[stage], n_tile * bn + j * 64, r0);
                    } else {
                        ab_tma_load_2d(sa, &tm_a, &full[stage], kb * BK, m_tile * BM);
                        if (MODE == MODE_NT) {
                            ab_tma_load_2d(sb, &tm_b, &full[stage], kb * BK, e * p.N + n_tile * bn);
                        } else {
                            for (int j = 0; j < bn / 64; ++j)
                                ab_tma_load_2d(sb + j * ATOM_BYTES, &tm_b, &full[stage], n_tile * bn + j * 64, e * p.K + kb * BK);
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(bn, MODE == MODE_TN ? 1 : 0, MODE == MODE_NT ? 0 : 1);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int nk;
                if (MODE == MODE_TN) {
                    const int e = tile / (p.num_m_tiles * p.num_n_tiles);
                    nk = p.seg_off != nullptr ? p.tn_nsrc * ((p.seg_off[e + 1] - p.seg_off[e]) / BK) : dense_tn_blocks(p, e);
                    if (nk == 0) continue;          // empty expert: the epilogue writes zeros without an accumulator
                } else {
                    nk = (p.K + BK - 1) / BK;
                }
                ab_mbar_wait(&tempty[acc], acc_phase ^ 1);
                ab_tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < nk; ++kb) {
                    ab_mbar_wait(&full[stage], phase);
                    ab_tc_fence_after();
                    const uint32_t sa = ab_smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
                    // K-major operand: 8-row groups 1024 B apart; MN-major: 64-wide atoms 8 KB apart, 8-k groups 1024 B apart
                    const uint64_t da = MODE == MODE_TN ? make_smem_desc(sa, ATOM_BYTES, 1024) : make_smem_desc(sa, 0, 1024);
                    const uint64_t db = MODE == MODE_NT ? make_smem_desc(sb, 0, 1024) : make_smem_desc(sb, ATOM_BYTES, 1024);
                    const uint32_t a_step = MODE == MODE_TN ? (16 * 128) >> 4 : 32 >> 4;     // advance 16 k per MMA
                    const uint32_t b_step = MODE == MODE_NT ? 32 >> 4 : (16 * 128) >> 4;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        ab_umma_f16(d_tmem, da + (uint64_t)(k * a_step), db + (uint64_t)(k * b_step), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    ab_umma_commit(&empty[stage]);              // smem slot free once these MMAs retire
                    if (kb == nk - 1) ab_umma_commit(&tfull[acc]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue =================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int sub = (warp - EPI_WARP0) >> 2;      // which 64-column group of the tile this warp owns
        unsigned char* stg = stg_base + (size_t)(warp - EPI_WARP0) * WARP_STG_BYTES;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_tile, n_tile, e, nk;
            if (MODE == MODE_TN) {
                e = tile / (p.num_m_tiles * p.num_n_tiles);
                const int rem = tile % (p.num_m_tiles * p.num_n_tiles);
                m_tile = rem / p.num_n_tiles; n_tile = rem % p.num_n_tiles;
                nk = p.seg_off != nullptr ? p.tn_nsrc * ((p.seg_off[e + 1] - p.seg_off[e]) / BK) : dense_tn_blocks(p, e);
            } else {
                m_tile = tile / p.num_n_tiles; n_tile = tile % p.num_n_tiles;
                e = p.tile_expert != nullptr ? p.tile_expert[m_tile] : 0;
                nk = 1;
            }
            const bool have_acc = nk > 0;
            if (have_acc) {
                ab_mbar_wait(&tfull[acc], acc_phase);
                ab_tc_fence_after();
            }
            const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
            const int g_col0 = sub * GW;
            if (g_col0 < bn) {
                const int g_cols = min(GW, bn - g_col0);
                if (p.c_f32) epilogue_group<MODE, true>(p, stg, t_row, have_acc, lane, quarter, m_tile, e, n_tile * bn, g_col0, g_cols);
                else epilogue_group<MODE, false>(p, stg, t_row, have_acc, lane, quarter, m_tile, e, n_tile * bn, g_col0, g_cols);
            }
            if (have_acc) {
                ab_tc_fence_before();
                __syncwarp();
                if (lane == 0) ab_mbar_arrive(&tempty[acc]);       // all of this warp's TMEM reads of the tile are done
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    ab_tc_fence_before();
    __syncthreads();
    if (warp == 1) ab_tmem_dealloc(tmem_base, 512);
}

int pick_bn(int N, bool mn_major_b) {
    const int step = mn_major_b ? 64 : 16;
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

template <int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int64_t max_tiles, cudaStream_t stream) {
    auto k = grouped_gemm_kernel<MODE>;
    AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int64_t grid = ab_num_sms();
    if (max_tiles < grid) grid = max_tiles;
    if (grid < 1) grid = 1;
    k<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, p);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

int gemm_rows(int mode, const void* A, const void* W, const float* bias, const void* aux, void* c, void* c2,
              const int32_t* tile_expert, const int32_t* n_rows, int64_t max_rows, int N, int K, int E, int epi, int act,
              int c_dtype, float drop_p, const uint32_t* drop_seed, cudaStream_t stream) {
    AB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "grouped_gemm: dropout probability must be in [0, 1)");
    AB_REQUIRE(drop_p == 0.f || (drop_seed != nullptr && (epi == AB_EPI_BIAS_ACT || epi == AB_EPI_DACT)),
               "grouped_gemm: dropout needs a seed and the bias+act / dact epilogue");
    AB_REQUIRE(max_rows > 0 && max_rows % BM == 0, "grouped_gemm: max_rows must be a positive multiple of %d", BM);
    AB_REQUIRE(N > 0 && K > 0 && E > 0 && N % 8 == 0 && K % 8 == 0, "grouped_gemm: N (%d) and K (%d) must be multiples of 8", N, K);
    AB_REQUIRE(c_dtype == AB_F32 || c_dtype == AB_BF16, "grouped_gemm: bad output dtype");
    AB_REQUIRE(epi >= AB_EPI_NONE && epi <= AB_EPI_ADD, "grouped_gemm: bad epilogue %d", epi);
    AB_REQUIRE((epi != AB_EPI_BIAS && epi != AB_EPI_BIAS_ACT) || bias, "grouped_gemm: bias epilogue without bias");
    AB_REQUIRE(epi != AB_EPI_BIAS_ACT || c2, "grouped_gemm: bias+act epilogue needs the pre-activation output c2");
    AB_REQUIRE((epi != AB_EPI_DACT && epi != AB_EPI_ADD) || aux, "grouped_gemm: dact / add epilogue needs aux");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.E = E;
    p.bn = pick_bn(N, mode == MODE_NN);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.epi = epi; p.act = act; p.c_f32 = c_dtype == AB_F32;
    p.tile_expert = tile_expert; p.n_rows = n_rows; p.bias = bias; p.aux = aux; p.c = c; p.c2 = c2;
    p.rows_valid = max_rows;
    if (drop_p > 0.f) {
        p.drop_seed = drop_seed;
        p.drop_thresh = (uint32_t)((double)drop_p * 65536.0 + 0.5);
        p.drop_scale = 1.0f / (1.0f - drop_p);
    }
    AB_REQUIRE(((uintptr_t)c % 16) == 0 && (c2 == nullptr || ((uintptr_t)c2 % 16) == 0) && (aux == nullptr || ((uintptr_t)aux % 16) == 0),
               "grouped_gemm: output / aux pointers must be 16-byte aligned");
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)K, (uint64_t)max_rows, BK, BM)) return e;
    const int64_t max_tiles = (max_rows / BM) * p.num_n_tiles;
    if (mode == MODE_NT) {
        if (int e = make_map2(&tb, W, (uint64_t)K, (uint64_t)E * N, BK, (uint32_t)p.bn)) return e;
        return launch<MODE_NT>(ta, tb, p, max_tiles, stream);
    }
    if (int e = make_map2(&tb, W, (uint64_t)N, (uint64_t)E * K, 64, BK)) return e;
    return launch<MODE_NN>(ta, tb, p, max_tiles, stream);
}

}  // namespace

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

extern "C" int ab_grouped_gemm_tn(const void* A, const void* Bm, float* Cw, const int32_t* seg_off, int64_t max_rows, int M,
                                  int N, int E, int nsrc, int64_t src_stride, cudaStream_t stream) {
    AB_REQUIRE(nsrc >= 1 && (nsrc == 1 || (src_stride > 0 && src_stride % BK == 0 && nsrc * src_stride <= max_rows)),
               "grouped_gemm_tn: bad source blocking nsrc=%d stride=%lld", nsrc, (long long)src_stride);
    AB_REQUIRE(max_rows > 0 && max_rows % BM == 0, "grouped_gemm_tn: max_rows must be a positive multiple of %d", BM);
    AB_REQUIRE(M > 0 && N > 0 && E > 0 && M % 8 == 0 && N % 8 == 0, "grouped_gemm_tn: M (%d) and N (%d) must be multiples of 8", M, N);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.M = M; p.E = E; p.K = 0;
    p.bn = pick_bn(N, true);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.num_m_tiles = (int)ab_ceil_div(M, BM);
    p.seg_off = seg_off; p.c_f32 = 1; p.epi = AB_EPI_NONE; p.cw = Cw;
    p.tn_nsrc = nsrc; p.tn_src_stride = src_stride;
    AB_REQUIRE(((uintptr_t)Cw % 16) == 0, "grouped_gemm_tn: output pointer must be 16-byte aligned");
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)M, (uint64_t)max_rows, 64, BK)) return e;
    if (int e = make_map2(&tb, Bm, (uint64_t)N, (uint64_t)max_rows, 64, BK)) return e;
    return launch<MODE_TN>(ta, tb, p, (int64_t)E * p.num_m_tiles * p.num_n_tiles, stream);
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
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.E = 1;
    p.bn = pick_bn(N, mode == MODE_NN);
    p.num_n_tiles = (int)ab_ceil_div(N, p.bn);
    p.epi = epi; p.c_f32 = c_dtype == AB_F32;
    p.bias = bias; p.aux = aux; p.c = c;
    p.rows_valid = S;
    p.dense_m_tiles = (int)ab_ceil_div(S, BM);
    CUtensorMap ta, tb;
    if (int e = make_map2(&ta, A, (uint64_t)K, (uint64_t)S, BK, BM)) return e;
    const int64_t tiles = (int64_t)p.dense_m_tiles * p.num_n_tiles;
    if (mode == MODE_NT) {
        if (int e = make_map2(&tb, W, (uint64_t)K, (uint64_t)N, BK, (uint32_t)p.bn)) return e;
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
}  // namespace

extern "C" size_t ab_dense_gemm_tn_workspace_bytes(int64_t S, int M, int N) {
    const int tiles = (int)(ab_ceil_div(M, BM) * ab_ceil_div(N, 256));
    int nsplit = tiles > 0 ? ab_num_sms() / tiles : 1;            // one wave of (slice, tile) work items
    const int64_t kblocks = ab_ceil_div(S, BK);
    if (nsplit > kblocks / 4) nsplit = (int)(kblocks / 4);          // at least 4 blocks of 64 rows per slice
    if (nsplit < 2) return 0;
    return (size_t)nsplit * M * N * sizeof(float);
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
    p.num_m_tiles = (int)ab_ceil_div(M, BM);
    p.c_f32 = 1; p.epi = AB_EPI_NONE;
    p.tn_nsrc = 1; p.tn_rows = S;
    // the contraction runs over all S rows and the output has only a few tiles: cut S into slices (split-K) so that the
    // whole chip works, partial products into the workspace, then one fixed-order sum
    const size_t need = ab_dense_gemm_tn_workspace_bytes(S, M, N);
    int nsplit = (int)(need / ((size_t)M * N * sizeof(float)));
    if (nsplit >= 2 && (ws == nullptr || ws_bytes < need)) nsplit = 1;
    if (nsplit < 2) nsplit = 1;
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
