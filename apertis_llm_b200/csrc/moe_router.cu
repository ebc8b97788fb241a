// MoE router: LayerNorm + Linear(E) + noise + softmax + top-K + normalised weights + aux-loss sums,
// and its backward.  Replaces core.py:480-492, 499-505, 524-529 and their autograd.
//
// Forward: one warp per token.  The row is read twice (mean, then centred second moment + the E dot
// products in the same pass; the second read hits L1), LayerNorm affine is folded into the router
// weight once per CTA (G[e,d] = ln_w[d]*Wr[e,d] in shared memory; c[e] = sum_d ln_b[d]*Wr[e,d] + br[e]).
// Lanes 0..E-1 then own one expert each for softmax / top-K.  Ties in top-K go to the lower expert id.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int MAX_K = 8;

// round-trip through the autocast dtype (round to nearest even, like torch's casts)
__device__ __forceinline__ float ab_round_to(float v, int quant) {
    if (quant == AB_ROUTER_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (quant == AB_ROUTER_FP16) return __half2float(__float2half_rn(v));
    return v;
}

template <typename T>
__device__ __forceinline__ void row_load(const T* row, int i, float* f) {   // vector i of the row
    ab_vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(row) + i), f);
}

// softmax / top-K / weights on lanes (lane e < E holds logit e).  Returns the gate in g; lane k < K ends up holding
// slot k's expert id and probability (my_idx, my_p) -- no per-thread arrays, so nothing spills to local memory.
__device__ __forceinline__ void select_topk(float logit, int lane, int E, int K, float& g, float& lse, int& my_idx, float& my_p,
                                            float& den, unsigned& sel_mask) {
    const float lv = lane < E ? logit : -INFINITY;
    const float m = ab_warp_max(lv);
    const float ex = lane < E ? expf(lv - m) : 0.f;
    const float sum = ab_warp_sum(ex);
    g = ex / sum;
    lse = m + logf(sum);
    float cand = lane < E ? g : -INFINITY;
    my_idx = 0; my_p = 0.f; den = 0.f; sel_mask = 0u;
    for (int k = 0; k < K; ++k) {
        const float best = ab_warp_max(cand);
        const unsigned ball = __ballot_sync(0xffffffffu, cand == best);
        const int win = __ffs(ball) - 1;               // lowest expert id among equals
        if (lane == k) { my_idx = win; my_p = best; }
        den += best;                                   // summed in slot order like torch.sum over K
        sel_mask |= 1u << win;
        if (lane == win) cand = -INFINITY;
    }
    den += 1e-6f;
}

__device__ __forceinline__ void write_selection(int s, int lane, int E, int K, float g, float lse, int my_idx, float my_p, float den,
                                                float* gates, int32_t* idx, float* probs, float* w, float* lse_out) {
    if (lane < E) gates[(size_t)s * E + lane] = g;
    if (lane < K) {
        idx[(size_t)s * K + lane] = my_idx;
        probs[(size_t)s * K + lane] = my_p;
        w[(size_t)s * K + lane] = my_p / den;
    }
    if (lane == 0) lse_out[s] = lse;
}

// aux partial layout per CTA: [2E+1] = sum gates[e], count[e], sum lse^2
template <typename T, int EM>
__global__ void __launch_bounds__(WARPS * 32) router_fwd_kernel(const T* __restrict__ x, const float* __restrict__ ln_w,
                                                                const float* __restrict__ ln_b, float eps,
                                                                const float* __restrict__ Wr, const float* __restrict__ br,
                                                                const float* __restrict__ noise,
                                                                const float* __restrict__ noise_scale, float* __restrict__ lclean,
                                                                float* __restrict__ logits, float* __restrict__ gates,
                                                                int32_t* __restrict__ idx, float* __restrict__ probs,
                                                                float* __restrict__ w, float* __restrict__ lse_out,
                                                                float* __restrict__ stats, float* __restrict__ part, int S,
                                                                int Dm, int E, int K, int quant) {
    constexpr int V = ab_vec16<T>::N;
    extern __shared__ float sm[];
    float* G = sm;                 // [E][Dm]
    float* cvec = G + (size_t)E * Dm;   // [E]
    float* red = cvec + 32;        // [WARPS][2E+1]
    float* lnw_s = red + WARPS * (2 * E + 1);      // [Dm], [Dm]: LayerNorm affine (quant modes only)
    float* lnb_s = lnw_s + Dm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (quant == AB_ROUTER_EXACT) {
        for (int e = 0; e < E; ++e)
            for (int d = tid; d < Dm; d += blockDim.x) G[(size_t)e * Dm + d] = ln_w[d] * Wr[(size_t)e * Dm + d];
        for (int e = warp; e < E; e += WARPS) {
            float acc = 0.f;
            for (int d = lane; d < Dm; d += 32) acc = fmaf(ln_b[d], Wr[(size_t)e * Dm + d], acc);
            acc = ab_warp_sum(acc);
            if (lane == 0) cvec[e] = acc + br[e];
        }
    } else {
        // autocast emulation (core.py:482 under torch.autocast): the router Linear sees its input, weight and bias rounded
        // to the autocast dtype, accumulates in fp32 and rounds its output once
        for (int e = 0; e < E; ++e)
            for (int d = tid; d < Dm; d += blockDim.x) G[(size_t)e * Dm + d] = ab_round_to(Wr[(size_t)e * Dm + d], quant);
        for (int d = tid; d < Dm; d += blockDim.x) { lnw_s[d] = ln_w[d]; lnb_s[d] = ln_b[d]; }
        if (tid < E) cvec[tid] = ab_round_to(br[tid], quant);
    }
    __syncthreads();
    const int nvec = Dm / V;
    float a_g = 0.f, a_cnt = 0.f, a_l2 = 0.f;
    for (int s = blockIdx.x * WARPS + warp; s < S; s += gridDim.x * WARPS) {
        const T* row = x + (size_t)s * Dm;
        float sum = 0.f;
        for (int i = lane; i < nvec; i += 32) {
            float f[V];
            row_load<T>(row, i, f);
#pragma unroll
            for (int v = 0; v < V; ++v) sum += f[v];
        }
        const float mean = ab_warp_sum(sum) / (float)Dm;
        float var = 0.f;
        float dot[EM];
#pragma unroll
        for (int e = 0; e < EM; ++e) dot[e] = 0.f;
        for (int i = lane; i < nvec; i += 32) {
            float f[V];
            row_load<T>(row, i, f);
#pragma unroll
            for (int v = 0; v < V; ++v) { f[v] -= mean; var = fmaf(f[v], f[v], var); }
#pragma unroll
            for (int e = 0; e < EM; ++e) {
                if (e < E) {
#pragma unroll
                    for (int v4 = 0; v4 < V; v4 += 4) {
                        const float4 gq = *reinterpret_cast<const float4*>(G + (size_t)e * Dm + i * V + v4);
                        dot[e] = fmaf(f[v4], gq.x, dot[e]); dot[e] = fmaf(f[v4 + 1], gq.y, dot[e]);
                        dot[e] = fmaf(f[v4 + 2], gq.z, dot[e]); dot[e] = fmaf(f[v4 + 3], gq.w, dot[e]);
                    }
                }
            }
        }
        var = ab_warp_sum(var) / (float)Dm;
        const float rstd = rsqrtf(var + eps);
        if (quant != AB_ROUTER_EXACT) {
            // third pass over the (L1-resident) row: the normalised row rounded element by element, then the dot products
#pragma unroll
            for (int e = 0; e < EM; ++e) dot[e] = 0.f;
            for (int i = lane; i < nvec; i += 32) {
                float f[V];
                row_load<T>(row, i, f);
#pragma unroll
                for (int v = 0; v < V; ++v) f[v] = ab_round_to(fmaf((f[v] - mean) * rstd, lnw_s[i * V + v], lnb_s[i * V + v]), quant);
#pragma unroll
                for (int e = 0; e < EM; ++e) {
                    if (e < E) {
#pragma unroll
                        for (int v = 0; v < V; ++v) dot[e] = fmaf(f[v], G[(size_t)e * Dm + i * V + v], dot[e]);
                    }
                }
            }
        }
        float mylogit = 0.f;
#pragma unroll
        for (int e = 0; e < EM; ++e) {
            if (e < E) {
                const float t = ab_warp_sum(dot[e]);
                if (lane == e) mylogit = t;
            }
        }
        float lc = 0.f;
        if (lane < E) {
            lc = quant == AB_ROUTER_EXACT ? fmaf(rstd, mylogit, cvec[lane]) : ab_round_to(mylogit + cvec[lane], quant);
            mylogit = lc;
            if (noise) mylogit = fmaf(noise[(size_t)s * E + lane], noise_scale[lane], lc);
            if (lclean) lclean[(size_t)s * E + lane] = lc;
            if (logits) logits[(size_t)s * E + lane] = mylogit;
        }
        if (lane == 0) { stats[2 * (size_t)s] = mean; stats[2 * (size_t)s + 1] = rstd; }
        float g, lse, my_p, den;
        int my_idx;
        unsigned sel_mask;
        select_topk(mylogit, lane, E, K, g, lse, my_idx, my_p, den, sel_mask);
        write_selection(s, lane, E, K, g, lse, my_idx, my_p, den, gates, idx, probs, w, lse_out);
        if (lane < E) {
            a_g += g;
            a_cnt += (sel_mask >> lane) & 1u ? 1.f : 0.f;
        }
        if (lane == 0) a_l2 = fmaf(lse, lse, a_l2);
    }
    const int NA = 2 * E + 1;
    if (lane < E) { red[warp * NA + lane] = a_g; red[warp * NA + E + lane] = a_cnt; }
    if (lane == 0) red[warp * NA + 2 * E] = a_l2;
    __syncthreads();
    if (tid < NA) {
        float s2 = 0.f;
        for (int wv = 0; wv < WARPS; ++wv) s2 += red[wv * NA + tid];
        part[(size_t)blockIdx.x * NA + tid] = s2;
    }
}

// Same forward with the token's row held in registers (hidden sizes up to 32 * V * NV): one global read of the row, the
// E dot products read G as 16-byte vectors, and for E <= 8 the E partial sums of the 32 lanes are reduced by recursive
// halving (10 shuffles instead of 5 per expert).  The launcher picks this kernel whenever the row fits.
template <int EM>
__device__ __forceinline__ float reduce_dots_to_lane(float (&dot)[EM], int lane, int E) {
    float mine = 0.f;
    if (EM == 8) {
        // halve the vector while doubling the lanes per value: after xor 16 / 8 / 4 a lane holds ONE expert's partial sum
        float v4[4], v2[2], v1;
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float keep = h16 ? dot[4 + j] : dot[j], send = h16 ? dot[j] : dot[4 + j];
            v4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float keep = h8 ? v4[2 + j] : v4[j], send = h8 ? v4[j] : v4[2 + j];
            v2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const float keep = h4 ? v2[1] : v2[0], send = h4 ? v2[0] : v2[1];
            v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        // lanes with bits (4,3,2) = (e2,e1,e0) hold expert e's total: hand it to lane e
        const int src = ((lane >> 2) & 1) << 4 | ((lane >> 1) & 1) << 3 | (lane & 1) << 2;
        mine = __shfl_sync(0xffffffffu, v1, src);
        if (lane >= E) mine = 0.f;
    } else {
#pragma unroll
        for (int e = 0; e < EM; ++e) {
            if (e < E) {
                const float t = ab_warp_sum(dot[e]);
                if (lane == e) mine = t;
            }
        }
    }
    return mine;
}

template <typename T, int EM, int NV>
__global__ void __launch_bounds__(WARPS * 32, (NV * ab_vec16<T>::N <= 24 && EM <= 8) ? 4 : 1) router_fwd_reg_kernel(const T* __restrict__ x, const float* __restrict__ ln_w,
                                                                    const float* __restrict__ ln_b, float eps,
                                                                    const float* __restrict__ Wr, const float* __restrict__ br,
                                                                    const float* __restrict__ noise,
                                                                    const float* __restrict__ noise_scale, float* __restrict__ lclean,
                                                                    float* __restrict__ logits, float* __restrict__ gates,
                                                                    int32_t* __restrict__ idx, float* __restrict__ probs,
                                                                    float* __restrict__ w, float* __restrict__ lse_out,
                                                                    float* __restrict__ stats, float* __restrict__ part, int S,
                                                                    int Dm, int E, int K, int quant) {
    constexpr int V = ab_vec16<T>::N;
    extern __shared__ float sm[];
    float* G = sm;                 // [E][Dm]
    float* cvec = G + (size_t)E * Dm;   // [E]
    float* red = cvec + 32;        // [WARPS][2E+1]
    float* lnw_s = red + WARPS * (2 * E + 1);      // [Dm], [Dm]: LayerNorm affine (quant modes only)
    float* lnb_s = lnw_s + Dm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (quant == AB_ROUTER_EXACT) {
        for (int i = tid; i < E * Dm; i += blockDim.x) G[i] = ln_w[i % Dm] * Wr[i];
        for (int e = warp; e < E; e += WARPS) {
            float acc = 0.f;
            for (int d = lane; d < Dm; d += 32) acc = fmaf(ln_b[d], Wr[(size_t)e * Dm + d], acc);
            acc = ab_warp_sum(acc);
            if (lane == 0) cvec[e] = acc + br[e];
        }
    } else {
        for (int i = tid; i < E * Dm; i += blockDim.x) G[i] = ab_round_to(Wr[i], quant);
        for (int d = tid; d < Dm; d += blockDim.x) { lnw_s[d] = ln_w[d]; lnb_s[d] = ln_b[d]; }
        if (tid < E) cvec[tid] = ab_round_to(br[tid], quant);
    }
    __syncthreads();
    const int nvec = Dm / V;
    const float inv_dm = 1.0f / (float)Dm;
    float a_g = 0.f, a_cnt = 0.f, a_l2 = 0.f;
    for (int s = blockIdx.x * WARPS + warp; s < S; s += gridDim.x * WARPS) {
        const T* row = x + (size_t)s * Dm;
        float f[NV][V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int iv = i * 32 + lane;
            if (iv < nvec) {
                row_load<T>(row, iv, f[i]);
#pragma unroll
                for (int v = 0; v < V; ++v) sum += f[i][v];
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) f[i][v] = 0.f;
            }
        }
        const float mean = ab_warp_sum(sum) * inv_dm;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i * 32 + lane < nvec) {
#pragma unroll
                for (int v = 0; v < V; ++v) { f[i][v] -= mean; var = fmaf(f[i][v], f[i][v], var); }
            }
        }
        var = ab_warp_sum(var) * inv_dm;
        const float rstd = rsqrtf(var + eps);
        float dot[EM];
#pragma unroll
        for (int e = 0; e < EM; ++e) dot[e] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int iv = i * 32 + lane;
            if (iv < nvec) {
                if (quant != AB_ROUTER_EXACT) {
                    // autocast emulation: the normalised row rounded element by element before the Linear
#pragma unroll
                    for (int v4 = 0; v4 < V; v4 += 4) {
                        const float4 lw = *reinterpret_cast<const float4*>(lnw_s + iv * V + v4);
                        const float4 lb = *reinterpret_cast<const float4*>(lnb_s + iv * V + v4);
                        f[i][v4] = ab_round_to(fmaf(f[i][v4] * rstd, lw.x, lb.x), quant);
                        f[i][v4 + 1] = ab_round_to(fmaf(f[i][v4 + 1] * rstd, lw.y, lb.y), quant);
                        f[i][v4 + 2] = ab_round_to(fmaf(f[i][v4 + 2] * rstd, lw.z, lb.z), quant);
                        f[i][v4 + 3] = ab_round_to(fmaf(f[i][v4 + 3] * rstd, lw.w, lb.w), quant);
                    }
                }
#pragma unroll
                for (int e = 0; e < EM; ++e) {
                    if (e < E) {
#pragma unroll
                        for (int v4 = 0; v4 < V; v4 += 4) {
                            const float4 gq = *reinterpret_cast<const float4*>(G + (size_t)e * Dm + iv * V + v4);
                            dot[e] = fmaf(f[i][v4], gq.x, dot[e]); dot[e] = fmaf(f[i][v4 + 1], gq.y, dot[e]);
                            dot[e] = fmaf(f[i][v4 + 2], gq.z, dot[e]); dot[e] = fmaf(f[i][v4 + 3], gq.w, dot[e]);
                        }
                    }
                }
            }
        }
        float mylogit = reduce_dots_to_lane<EM>(dot, lane, E);
        float lc = 0.f;
        if (lane < E) {
            lc = quant == AB_ROUTER_EXACT ? fmaf(rstd, mylogit, cvec[lane]) : ab_round_to(mylogit + cvec[lane], quant);
            mylogit = lc;
            if (noise) mylogit = fmaf(noise[(size_t)s * E + lane], noise_scale[lane], lc);
            if (lclean) lclean[(size_t)s * E + lane] = lc;
            if (logits) logits[(size_t)s * E + lane] = mylogit;
        }
        if (lane == 0) { stats[2 * (size_t)s] = mean; stats[2 * (size_t)s + 1] = rstd; }
        float g, lse, my_p, den;
        int my_idx;
        unsigned sel_mask;
        select_topk(mylogit, lane, E, K, g, lse, my_idx, my_p, den, sel_mask);
        write_selection(s, lane, E, K, g, lse, my_idx, my_p, den, gates, idx, probs, w, lse_out);
        if (lane < E) {
            a_g += g;
            a_cnt += (sel_mask >> lane) & 1u ? 1.f : 0.f;
        }
        if (lane == 0) a_l2 = fmaf(lse, lse, a_l2);
    }
    const int NA = 2 * E + 1;
    if (lane < E) { red[warp * NA + lane] = a_g; red[warp * NA + E + lane] = a_cnt; }
    if (lane == 0) red[warp * NA + 2 * E] = a_l2;
    __syncthreads();
    if (tid < NA) {
        float s2 = 0.f;
        for (int wv = 0; wv < WARPS; ++wv) s2 += red[wv * NA + tid];
        part[(size_t)blockIdx.x * NA + tid] = s2;
    }
}

// out[c] = sum_r part[r][c]; one warp per column, fixed order -> deterministic
__global__ void reduce_rows_kernel(const float* __restrict__ part, int nrows, int ncols, float* __restrict__ out) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= ncols) return;
    float s = 0.f;
    for (int r = lane; r < nrows; r += 32) s += part[(size_t)r * ncols + c];
    s = ab_warp_sum(s);
    if (lane == 0) out[c] = s;
}

__global__ void __launch_bounds__(WARPS * 32) topk_from_logits_kernel(const float* __restrict__ logits, float* __restrict__ gates,
                                                                      int32_t* __restrict__ idx, float* __restrict__ probs,
                                                                      float* __restrict__ w, float* __restrict__ lse_out, int S,
                                                                      int E, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = blockIdx.x * WARPS + warp; s < S; s += gridDim.x * WARPS) {
        const float l = lane < E ? logits[(size_t)s * E + lane] : 0.f;
        float g, lse, my_p, den;
        int my_idx;
        unsigned sel_mask;
        select_topk(l, lane, E, K, g, lse, my_idx, my_p, den, sel_mask);
        write_selection(s, lane, E, K, g, lse, my_idx, my_p, den, gates, idx, probs, w, lse_out);
    }
}

// ---------------------------------------------------------------------------------------------
// backward, kernel B: per token d logits and dx
// part layout per CTA: [2E] = sum dlogit[e] (-> dbr), sum dlogit[e]*noise[s,e] (-> dnoise_scale)
// ---------------------------------------------------------------------------------------------
template <typename T, int EM>
__global__ void __launch_bounds__(WARPS * 32) router_bwd_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                                const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                const float* __restrict__ Wr, const float* __restrict__ br,
                                                                const float* __restrict__ gates, const int32_t* __restrict__ idx,
                                                                const float* __restrict__ probs, const float* __restrict__ lse,
                                                                const float* __restrict__ lclean, const float* __restrict__ noise,
                                                                const float* __restrict__ f, const float* __restrict__ scal,
                                                                const float* __restrict__ dw_row, const float* __restrict__ dxrow,
                                                                const int32_t* __restrict__ row_of, T* __restrict__ dx,
                                                                float* __restrict__ dlogits, float* __restrict__ part, int S,
                                                                int Dm, int E, int K) {
    constexpr int V = ab_vec16<T>::N;
    extern __shared__ float sm[];
    float* Wsm = sm;                         // [E][Dm] raw router weight
    float* gam = Wsm + (size_t)E * Dm;       // [Dm]
    float* GS = gam + Dm;                    // [32]  sum_d ln_w[d]*Wr[e,d]
    float* cvec = GS + 32;                   // [32]
    float* red = cvec + 32;                  // [WARPS][2E]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < E * Dm; i += blockDim.x) Wsm[i] = Wr[i];
    for (int i = tid; i < Dm; i += blockDim.x) gam[i] = ln_w[i];
    for (int e = warp; e < E; e += WARPS) {
        float a1 = 0.f, a2 = 0.f;
        for (int d = lane; d < Dm; d += 32) {
            a1 = fmaf(ln_w[d], Wr[(size_t)e * Dm + d], a1);
            a2 = fmaf(ln_b[d], Wr[(size_t)e * Dm + d], a2);
        }
        a1 = ab_warp_sum(a1); a2 = ab_warp_sum(a2);
        if (lane == 0) { GS[e] = a1; cvec[e] = a2 + br[e]; }
    }
    __syncthreads();
    const float g_lb = scal[0], g_rz = scal[1];
    const int nvec = Dm / V;
    float a_db = 0.f, a_dn = 0.f;
    for (int s = blockIdx.x * WARPS + warp; s < S; s += gridDim.x * WARPS) {
        // ---- d weights -> d probs (lane k < K holds slot k)
        int my_r = -1, my_i = 0;
        float my_dw = 0.f, my_pk = 0.f;
        if (lane < K) {
            my_r = row_of[(size_t)s * K + lane];
            my_i = idx[(size_t)s * K + lane];
            my_pk = probs[(size_t)s * K + lane];
            my_dw = my_r >= 0 ? dw_row[my_r] : 0.f;
        }
        float den = 0.f;
        for (int k = 0; k < K; ++k) den += __shfl_sync(0xffffffffu, my_pk, k);    // slot order, as in the forward
        den += 1e-6f;
        const float dot = ab_warp_sum(my_dw * my_pk);
        const float inv = 1.f / den;
        const float my_dp = my_dw * inv - dot * inv * inv;       // d prob of slot `lane`
        float dgate = 0.f, g = 0.f;
        for (int k = 0; k < K; ++k) {
            const int ik = __shfl_sync(0xffffffffu, my_i, k);
            const float dpk = __shfl_sync(0xffffffffu, my_dp, k);
            if (lane == ik) dgate += dpk;
        }
        if (lane < E) {
            g = gates[(size_t)s * E + lane];
            dgate = fmaf(g_lb, f[lane], dgate);
        } else {
            dgate = 0.f;
        }
        const float gd = ab_warp_sum(g * dgate);
        float dl = 0.f;
        if (lane < E) {
            dl = g * (dgate - gd) + g_rz * 2.f * lse[s] * g;
            dlogits[(size_t)s * E + lane] = dl;
            a_db += dl;
            if (noise) a_dn = fmaf(dl, noise[(size_t)s * E + lane], a_dn);
        }
        // ---- LayerNorm backward constants:  m1 = mean_d(dxhat), m2 = mean_d(dxhat * xhat)
        const float mean = stats[2 * (size_t)s], rstd = stats[2 * (size_t)s + 1];
        float t1 = 0.f, t2 = 0.f;
        if (lane < E) {
            t1 = dl * GS[lane];
            t2 = dl * (lclean[(size_t)s * E + lane] - cvec[lane]);    // sum_d G[e,d]*xhat[d]
        }
        const float m1 = ab_warp_sum(t1) / (float)Dm;
        const float m2 = ab_warp_sum(t2) / (float)Dm;
        float dle[EM];
#pragma unroll
        for (int e = 0; e < EM; ++e) dle[e] = __shfl_sync(0xffffffffu, dl, e);
        int rk[MAX_K];                                   // rows of the K slots, broadcast once (the vector loop below diverges)
#pragma unroll
        for (int k = 0; k < MAX_K; ++k) { const int r = __shfl_sync(0xffffffffu, my_r, k); rk[k] = k < K ? r : -1; }
        const T* row = x + (size_t)s * Dm;
        T* orow = dx + (size_t)s * Dm;
        for (int i = lane; i < nvec; i += 32) {
            float fx[V], o[V];
            row_load<T>(row, i, fx);
            float dxn[V];
#pragma unroll
            for (int v = 0; v < V; ++v) dxn[v] = 0.f;
#pragma unroll
            for (int e = 0; e < EM; ++e) {
                if (e < E) {
#pragma unroll
                    for (int v4 = 0; v4 < V; v4 += 4) {
                        const float4 wq = *reinterpret_cast<const float4*>(Wsm + (size_t)e * Dm + i * V + v4);
                        dxn[v4] = fmaf(dle[e], wq.x, dxn[v4]); dxn[v4 + 1] = fmaf(dle[e], wq.y, dxn[v4 + 1]);
                        dxn[v4 + 2] = fmaf(dle[e], wq.z, dxn[v4 + 2]); dxn[v4 + 3] = fmaf(dle[e], wq.w, dxn[v4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int v4 = 0; v4 < V; v4 += 4) {
                const float4 gq = *reinterpret_cast<const float4*>(gam + i * V + v4);
                const float gg[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float xh = (fx[v4 + u] - mean) * rstd;
                    o[v4 + u] = rstd * (dxn[v4 + u] * gg[u] - m1 - xh * m2);
                }
            }
#pragma unroll
            for (int k = 0; k < MAX_K; ++k) {
                const int rkk = rk[k];
                if (rkk >= 0) {
                    const float* er = dxrow + (size_t)rkk * Dm + i * V;
#pragma unroll
                    for (int v4 = 0; v4 < V; v4 += 4) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(er + v4));
                        o[v4] += q.x; o[v4 + 1] += q.y; o[v4 + 2] += q.z; o[v4 + 3] += q.w;
                    }
                }
            }
            *(reinterpret_cast<uint4*>(orow) + i) = ab_vec16<T>::pack(o);
        }
    }
    const int NA = 2 * E;
    if (lane < E) { red[warp * NA + lane] = a_db; red[warp * NA + E + lane] = a_dn; }
    __syncthreads();
    if (tid < NA) {
        float s2 = 0.f;
        for (int wv = 0; wv < WARPS; ++wv) s2 += red[wv * NA + tid];
        part[(size_t)blockIdx.x * NA + tid] = s2;
    }
}

// kernel C: Q[e,d] = sum_s dlogits[s,e] * xhat[s,d]; thread per column d, token blocks of QB tokens
constexpr int qb_for(int em) { return 2048 / em; }     // tokens per partial block (smem [QB][EM] floats)
int em_for(int E) { return E <= 8 ? 8 : (E <= 16 ? 16 : 32); }
template <typename T, int EM>
__global__ void __launch_bounds__(256) router_q_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                       const float* __restrict__ dlogits, float* __restrict__ qpart, int S,
                                                       int Dm, int E) {
    constexpr int QB = qb_for(EM);
    __shared__ float sdl[QB * EM];
    __shared__ float sst[QB * 2];
    const int s0 = blockIdx.y * QB;
    const int ns = min(QB, S - s0);
    for (int i = threadIdx.x; i < ns * E; i += blockDim.x) sdl[i] = dlogits[(size_t)s0 * E + i];
    for (int i = threadIdx.x; i < ns * 2; i += blockDim.x) sst[i] = stats[(size_t)s0 * 2 + i];
    __syncthreads();
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= Dm) return;
    float acc[EM];
#pragma unroll
    for (int e = 0; e < EM; ++e) acc[e] = 0.f;
    // 8 rows in flight per thread: the loop is a stream of independent, coalesced loads, not a dependent chain
    for (int j0 = 0; j0 < ns; j0 += 8) {
        float xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = j0 + u < ns ? ab_to_float(x[(size_t)(s0 + j0 + u) * Dm + d]) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            if (j < ns) {
                const float xh = (xv[u] - sst[2 * j]) * sst[2 * j + 1];
#pragma unroll
                for (int e = 0; e < EM; ++e)
                    if (e < E) acc[e] = fmaf(sdl[j * E + e], xh, acc[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < EM; ++e)
        if (e < E) qpart[((size_t)blockIdx.y * E + e) * Dm + d] = acc[e];
}

// kernel D: reduce Q partials and derive all router parameter grads.  block = (32 columns, E experts)
//   dbr[e] = sum dlogits;  dWr[e,d] = ln_w[d]*Q[e,d] + ln_b[d]*dbr[e]
//   dln_w[d] = sum_e Wr[e,d]*Q[e,d];  dln_b[d] = sum_e Wr[e,d]*dbr[e]
__global__ void router_param_kernel(const float* __restrict__ qpart, int nqb, const float* __restrict__ part, int nparts,
                                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, const float* __restrict__ Wr,
                                    float* __restrict__ dWr, float* __restrict__ dbr, float* __restrict__ dln_w,
                                    float* __restrict__ dln_b, float* __restrict__ dnoise_scale, int Dm, int E) {
    __shared__ float sdb[64];
    __shared__ float sgw[32][33], sgb[32][33];
    __shared__ float spart[16][64];
    const int tx = threadIdx.x, e = threadIdx.y;
    const int tid = e * 32 + tx;
    const int nthr = 32 * E;
    // column sums of the per-CTA partials: up to 16 thread groups add interleaved rows, then one thread per column adds
    // the groups in index order (fixed order -> deterministic)
    {
        const int ncol = 2 * E;
        const int ngrp = min(16, nthr / ncol);
        const int col = tid % ncol, grp = tid / ncol;
        if (grp < ngrp) {
            float s = 0.f;
            for (int r0 = grp; r0 < nparts; r0 += ngrp * 4) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = r0 + u * ngrp < nparts ? part[(size_t)(r0 + u * ngrp) * ncol + col] : 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) s += v[u];
            }
            spart[grp][col] = s;
        }
        __syncthreads();
        if (tid < ncol) {
            float s = 0.f;
            for (int g2 = 0; g2 < ngrp; ++g2) s += spart[g2][tid];
            sdb[tid] = s;
            if (blockIdx.x == 0) {
                if (tid < E) dbr[tid] = s;
                else if (dnoise_scale) dnoise_scale[tid - E] = s;
            }
        }
    }
    __syncthreads();
    const int d = blockIdx.x * 32 + tx;
    float gw = 0.f, gb = 0.f;
    if (d < Dm) {
        float q = 0.f;
        for (int r0 = 0; r0 < nqb; r0 += 8) {
            float qv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) qv[u] = r0 + u < nqb ? qpart[((size_t)(r0 + u) * E + e) * Dm + d] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) q += qv[u];
        }
        const float wv = Wr[(size_t)e * Dm + d];
        dWr[(size_t)e * Dm + d] = fmaf(ln_w[d], q, ln_b[d] * sdb[e]);
        gw = wv * q;
        gb = wv * sdb[e];
    }
    sgw[e][tx] = gw;
    sgb[e][tx] = gb;
    __syncthreads();
    if (e == 0 && d < Dm) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < E; ++i) { a += sgw[i][tx]; b += sgb[i][tx]; }
        dln_w[d] = a;
        dln_b[d] = b;
    }
}

int router_grid(int S) {
    const int want = (int)ab_ceil_div(S, WARPS);
    const int cap = ab_num_sms() * 4;
    return want < cap ? want : cap;
}

struct FwdWs { size_t part_off, total; int grid; };
FwdWs fwd_ws(int S, int E) {
    FwdWs w;
    w.grid = router_grid(S);
    w.part_off = 0;
    w.total = ab_round_up((int64_t)w.grid * (2 * E + 1) * sizeof(float), 256);
    return w;
}
struct BwdWs { size_t part_off, dl_off, q_off, total; int grid, nqb; };
BwdWs bwd_ws(int S, int Dm, int E) {
    BwdWs w;
    w.grid = router_grid(S);
    w.nqb = (int)ab_ceil_div(S, qb_for(em_for(E)));
    size_t o = 0;
    w.part_off = o; o += ab_round_up((int64_t)w.grid * 2 * E * sizeof(float), 256);
    w.dl_off = o; o += ab_round_up((int64_t)S * E * sizeof(float), 256);
    w.q_off = o; o += ab_round_up((int64_t)w.nqb * E * Dm * sizeof(float), 256);
    w.total = o;
    return w;
}

int check_router(int S, int Dm, int E, int K, int dtype) {
    AB_REQUIRE(S > 0 && Dm > 0, "moe_router: empty shape S=%d Dm=%d", S, Dm);
    AB_REQUIRE(E >= 1 && E <= 32, "moe_router: num_experts must be in [1,32] (got %d)", E);
    AB_REQUIRE(K >= 1 && K <= MAX_K && K <= E, "moe_router: experts_per_token must be in [1,%d] and <= E (got %d)", MAX_K, K);
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "moe_router: bad dtype %d", dtype);
    const int V = dtype == AB_F32 ? 4 : 8;
    AB_REQUIRE(Dm % V == 0, "moe_router: hidden size %d must be a multiple of %d", Dm, V);
    return AB_OK;
}

}  // namespace

extern "C" size_t ab_moe_router_workspace_bytes(int S, int Dm, int E) { (void)Dm; return fwd_ws(S, E).total; }

extern "C" int ab_moe_router_fwd(const void* x, const float* ln_w, const float* ln_b, float eps, const float* Wr,
                                 const float* br, const float* noise, const float* noise_scale, float* lclean, float* logits,
                                 float* gates, int32_t* idx, float* probs, float* w, float* lse, float* stats_out, float* aux,
                                 void* ws, size_t ws_bytes, int S, int Dm, int E, int K, int dtype, int quant, cudaStream_t stream) {
    if (int e = check_router(S, Dm, E, K, dtype)) return e;
    const FwdWs wl = fwd_ws(S, E);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "moe_router_fwd: workspace too small");
    AB_REQUIRE(noise == nullptr || noise_scale != nullptr, "moe_router_fwd: noise without noise_scale");
    AB_REQUIRE(quant == AB_ROUTER_EXACT || quant == AB_ROUTER_BF16 || quant == AB_ROUTER_FP16, "moe_router_fwd: bad logit rounding mode %d", quant);
    const size_t smem = ((size_t)E * Dm + 32 + WARPS * (2 * E + 1) + 2 * (size_t)Dm) * sizeof(float);
    float* part = (float*)ws;
    const int V = dtype == AB_F32 ? 4 : 8;
    const int nv = (int)ab_ceil_div(Dm, 32 * V);          // 16-byte vectors of the row per lane
#define AB_ROUTER_LAUNCH(KERNEL)                                                                                        \
    {                                                                                                                    \
        auto k = KERNEL;                                                                                                 \
        AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        k<<<wl.grid, WARPS * 32, smem, stream>>>((const TT*)x, ln_w, ln_b, eps, Wr, br, noise, noise_scale, lclean,     \
                                                 logits, gates, idx, probs, w, lse, stats_out, part, S, Dm, E, K, quant); \
    }
    // the row-in-registers kernel whenever the row fits (hidden size <= 2048 fp32 / 4096 bf16), the streaming one otherwise
#define AB_ROUTER_FWD(TT_, EMV)                                                                                         \
    {                                                                                                                    \
        using TT = TT_;                                                                                                  \
        if (nv <= 4) AB_ROUTER_LAUNCH((router_fwd_reg_kernel<TT, EMV, 4>))                                               \
        else if (nv <= 6) AB_ROUTER_LAUNCH((router_fwd_reg_kernel<TT, EMV, 6>))                                          \
        else if (nv <= 8) AB_ROUTER_LAUNCH((router_fwd_reg_kernel<TT, EMV, 8>))                                          \
        else if (nv <= 16 && EMV <= 8) AB_ROUTER_LAUNCH((router_fwd_reg_kernel<TT, EMV, 16>))                            \
        else AB_ROUTER_LAUNCH((router_fwd_kernel<TT, EMV>))                                                              \
    }
    if (dtype == AB_F32) {
        if (E <= 8) AB_ROUTER_FWD(float, 8) else if (E <= 16) AB_ROUTER_FWD(float, 16) else AB_ROUTER_FWD(float, 32)
    } else {
        if (E <= 8) AB_ROUTER_FWD(__nv_bfloat16, 8) else if (E <= 16) AB_ROUTER_FWD(__nv_bfloat16, 16) else AB_ROUTER_FWD(__nv_bfloat16, 32)
    }
#undef AB_ROUTER_LAUNCH
#undef AB_ROUTER_FWD
    AB_LAUNCH_CHECK();
    reduce_rows_kernel<<<(unsigned)ab_ceil_div(2 * E + 1, 4), 128, 0, stream>>>(part, wl.grid, 2 * E + 1, aux);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_moe_topk_from_logits(const float* logits, float* gates, int32_t* idx, float* probs, float* w, float* lse,
                                       int S, int E, int K, cudaStream_t stream) {
    if (int e = check_router(S, 8, E, K, AB_F32)) return e;
    topk_from_logits_kernel<<<router_grid(S), WARPS * 32, 0, stream>>>(logits, gates, idx, probs, w, lse, S, E, K);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" size_t ab_moe_router_bwd_workspace_bytes(int S, int Dm, int E) { return bwd_ws(S, Dm, E).total; }

extern "C" int ab_moe_router_bwd(const void* x, const float* stats, const float* ln_w, const float* ln_b, const float* Wr,
                                 const float* br, const float* gates, const int32_t* idx, const float* probs, const float* lse,
                                 const float* lclean, const float* noise, const float* f, const float* scal,
                                 const float* dw_row, const float* dxrow, const int32_t* row_of, void* dx, float* dWr,
                                 float* dbr, float* dln_w, float* dln_b, float* dnoise_scale, void* ws, size_t ws_bytes, int S,
                                 int Dm, int E, int K, int dtype, cudaStream_t stream) {
    if (int e = check_router(S, Dm, E, K, dtype)) return e;
    AB_REQUIRE(Dm % 4 == 0, "moe_router_bwd: hidden size must be a multiple of 4");
    const BwdWs wl = bwd_ws(S, Dm, E);
    AB_REQUIRE(ws && ws_bytes >= wl.total, "moe_router_bwd: workspace too small");
    unsigned char* w8 = (unsigned char*)ws;
    float* part = (float*)(w8 + wl.part_off);
    float* dl = (float*)(w8 + wl.dl_off);
    float* qpart = (float*)(w8 + wl.q_off);
    const size_t smem = ((size_t)E * Dm + Dm + 64 + WARPS * 2 * E) * sizeof(float);
    dim3 qgrid((unsigned)ab_ceil_div(Dm, 256), wl.nqb);
#define AB_ROUTER_BWD(TT, EMV)                                                                                          \
    {                                                                                                                    \
        auto k = router_bwd_kernel<TT, EMV>;                                                                             \
        AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        k<<<wl.grid, WARPS * 32, smem, stream>>>((const TT*)x, stats, ln_w, ln_b, Wr, br, gates, idx, probs, lse,       \
                                                 lclean, noise, f, scal, dw_row, dxrow, row_of, (TT*)dx, dl, part, S,   \
                                                 Dm, E, K);                                                              \
        AB_LAUNCH_CHECK();                                                                                               \
        router_q_kernel<TT, EMV><<<qgrid, 256, 0, stream>>>((const TT*)x, stats, dl, qpart, S, Dm, E);                  \
    }
    if (dtype == AB_F32) {
        if (E <= 8) AB_ROUTER_BWD(float, 8) else if (E <= 16) AB_ROUTER_BWD(float, 16) else AB_ROUTER_BWD(float, 32)
    } else {
        if (E <= 8) AB_ROUTER_BWD(__nv_bfloat16, 8) else if (E <= 16) AB_ROUTER_BWD(__nv_bfloat16, 16) else AB_ROUTER_BWD(__nv_bfloat16, 32)
    }
#undef AB_ROUTER_BWD
    AB_LAUNCH_CHECK();
    router_param_kernel<<<(unsigned)ab_ceil_div(Dm, 32), dim3(32, E), 0, stream>>>(qpart, wl.nqb, part, wl.grid, ln_w, ln_b, Wr, dWr,
                                                                                  dbr, dln_w, dln_b, dnoise_scale, Dm, E);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
