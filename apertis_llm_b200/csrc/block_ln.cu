// Pre-norm LayerNorm of the block wrappers (ApertisAttention.pre_norm / ApertisFeedForward.pre_norm,
// core.py:694-695, 887-888), forward and backward.  SURVEY.md section 8(f) row 1: these bracket every
// hot-path call and are pure HBM round trips, so they get a vectorised warp-per-row kernel and a
// deterministic two-stage reduction for the affine gradients instead of the stock ATen kernels.
#include "common.cuh"

namespace {

constexpr int RB = 256;     // rows per partial block of the affine-gradient reduction

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float* f) {
    if constexpr (sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    }
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float* f) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
        uint2 r; r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, float eps, TY* __restrict__ y,
                                                     float* __restrict__ stats, int S, int Dm) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
        const TX* row = x + (size_t)s * Dm;
        float sum = 0.f;
        for (int d = lane * 4; d < Dm; d += 128) {
            float f[4];
            ld4<TX>(row + d, f);
            sum += (f[0] + f[1]) + (f[2] + f[3]);
        }
        const float mean = ab_warp_sum(sum) / (float)Dm;
        float var = 0.f;
        for (int d = lane * 4; d < Dm; d += 128) {
            float f[4];
            ld4<TX>(row + d, f);
#pragma unroll
            for (int v = 0; v < 4; ++v) { const float c = f[v] - mean; var = fmaf(c, c, var); }
        }
        const float rstd = rsqrtf(ab_warp_sum(var) / (float)Dm + eps);
        if (lane == 0) { stats[2 * (size_t)s] = mean; stats[2 * (size_t)s + 1] = rstd; }
        TY* orow = y + (size_t)s * Dm;
        for (int d = lane * 4; d < Dm; d += 128) {
            float f[4], o[4];
            ld4<TX>(row + d, f);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
            const float4 bv = __ldg(reinterpret_cast<const float4*>(b + d));
            o[0] = fmaf((f[0] - mean) * rstd, wv.x, bv.x);
            o[1] = fmaf((f[1] - mean) * rstd, wv.y, bv.y);
            o[2] = fmaf((f[2] - mean) * rstd, wv.z, bv.z);
            o[3] = fmaf((f[3] - mean) * rstd, wv.w, bv.w);
            st4<TY>(orow + d, o);
        }
    }
}

// dx = rstd * (dy*w - mean_d(dy*w) - xhat * mean_d(dy*w*xhat)) [+ dres]
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) ln_bwd_rows_kernel(const TG* __restrict__ dy, const TX* __restrict__ x,
                                                          const float* __restrict__ stats, const float* __restrict__ w,
                                                          const TX* __restrict__ dres, TX* __restrict__ dx, int S, int Dm) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
        const float mean = stats[2 * (size_t)s], rstd = stats[2 * (size_t)s + 1];
        const TX* xr = x + (size_t)s * Dm;
        const TG* gr = dy + (size_t)s * Dm;
        float a1 = 0.f, a2 = 0.f;
        for (int d = lane * 4; d < Dm; d += 128) {
            float fx[4], fg[4];
            ld4<TX>(xr + d, fx);
            ld4<TG>(gr + d, fg);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float dh = fg[v] * ww[v];
                a1 += dh;
                a2 = fmaf(dh, (fx[v] - mean) * rstd, a2);
            }
        }
        const float m1 = ab_warp_sum(a1) / (float)Dm, m2 = ab_warp_sum(a2) / (float)Dm;
        TX* orow = dx + (size_t)s * Dm;
        for (int d = lane * 4; d < Dm; d += 128) {
            float fx[4], fg[4], o[4], r[4] = {0.f, 0.f, 0.f, 0.f};
            ld4<TX>(xr + d, fx);
            ld4<TG>(gr + d, fg);
            if (dres) ld4<TX>(dres + (size_t)s * Dm + d, r);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) o[v] = fmaf(rstd, fg[v] * ww[v] - m1 - (fx[v] - mean) * rstd * m2, r[v]);
            st4<TX>(orow + d, o);
        }
    }
}

// Row pass and column pass in one kernel (hidden sizes up to NIT * 128): a lane always owns the same columns, so the
// products dy * xhat and dy of its rows accumulate in registers while the row pass streams; the warps of a CTA then add
// their sums in warp order through shared memory and the CTA writes one partial row of `part` -- the second read of dy
// and x by ln_bwd_cols_kernel disappears.  Fixed order -> deterministic.
template <typename TX, typename TG, int NIT>
__global__ void __launch_bounds__(256) ln_bwd_fused_kernel(const TG* __restrict__ dy, const TX* __restrict__ x,
                                                           const float* __restrict__ stats, const float* __restrict__ w,
                                                           const TX* __restrict__ dres, TX* __restrict__ dx,
                                                           float* __restrict__ part, int S, int Dm) {
    __shared__ float sacc[2][NIT * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float gw[NIT][4], gb[NIT][4];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { gw[it][v] = 0.f; gb[it][v] = 0.f; }
    }
    for (int s = blockIdx.x * wpb + warp; s < S; s += gridDim.x * wpb) {
        const float mean = stats[2 * (size_t)s], rstd = stats[2 * (size_t)s + 1];
        const TX* xr = x + (size_t)s * Dm;
        const TG* gr = dy + (size_t)s * Dm;
        float fx[NIT][4], fg[NIT][4];
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int d = it * 128 + lane * 4;
#pragma unroll
            for (int v = 0; v < 4; ++v) { fx[it][v] = 0.f; fg[it][v] = 0.f; }
            if (d < Dm) {
                ld4<TX>(xr + d, fx[it]);
                ld4<TG>(gr + d, fg[it]);
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int d = it * 128 + lane * 4;
            if (d < Dm) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
                const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    fx[it][v] = (fx[it][v] - mean) * rstd;            // xhat
                    gw[it][v] = fmaf(fg[it][v], fx[it][v], gw[it][v]);
                    gb[it][v] += fg[it][v];
                    fg[it][v] *= ww[v];                               // d xhat
                    a1 += fg[it][v];
                    a2 = fmaf(fg[it][v], fx[it][v], a2);
                }
            }
        }
        const float m1 = ab_warp_sum(a1) / (float)Dm, m2 = ab_warp_sum(a2) / (float)Dm;
        TX* orow = dx + (size_t)s * Dm;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int d = it * 128 + lane * 4;
            if (d < Dm) {
                float o[4], r[4] = {0.f, 0.f, 0.f, 0.f};
                if (dres) ld4<TX>(dres + (size_t)s * Dm + d, r);
#pragma unroll
                for (int v = 0; v < 4; ++v) o[v] = fmaf(rstd, fg[it][v] - m1 - fx[it][v] * m2, r[v]);
                st4<TX>(orow + d, o);
            }
        }
    }
    for (int wv = 0; wv < wpb; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int i = it * 128 + lane * 4 + v;
                    sacc[0][i] = wv ? sacc[0][i] + gw[it][v] : gw[it][v];
                    sacc[1][i] = wv ? sacc[1][i] + gb[it][v] : gb[it][v];
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 2 * Dm; i += blockDim.x) {
        const int q = i / Dm, d = i % Dm;
        part[((size_t)blockIdx.x * 2 + q) * Dm + d] = sacc[q][d];
    }
}

// grid (ceil(S/RB), ceil(Dm/128)); block 256 = 8 warps over rows x (32 lanes x 4 columns)
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) ln_bwd_cols_kernel(const TG* __restrict__ dy, const TX* __restrict__ x,
                                                          const float* __restrict__ stats, float* __restrict__ part, int S,
                                                          int Dm) {
    __shared__ float red[8][2][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.y * 128 + lane * 4;
    const int s0 = blockIdx.x * RB;
    const int s1 = min(S, s0 + RB);
    float gw[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
    if (d < Dm) {
        for (int s = s0 + warp; s < s1; s += 8) {
            const float mean = stats[2 * (size_t)s], rstd = stats[2 * (size_t)s + 1];
            float fx[4], fg[4];
            ld4<TX>(x + (size_t)s * Dm + d, fx);
            ld4<TG>(dy + (size_t)s * Dm + d, fg);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                gw[v] = fmaf(fg[v], (fx[v] - mean) * rstd, gw[v]);
                gb[v] += fg[v];
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) { red[warp][0][lane * 4 + v] = gw[v]; red[warp][1][lane * 4 + v] = gb[v]; }
    __syncthreads();
    const int q = threadIdx.x >> 7, c = threadIdx.x & 127;
    float s2 = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s2 += red[wv][q][c];
    const int dd = blockIdx.y * 128 + c;
    if (dd < Dm) part[((size_t)blockIdx.x * 2 + q) * Dm + dd] = s2;
}

// one warp per output column (2*Dm of them), fixed order
__global__ void ln_param_reduce_kernel(const float* __restrict__ part, int nblocks, int Dm, float* __restrict__ dw,
                                       float* __restrict__ db) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= 2 * Dm) return;
    const int q = j / Dm, d = j % Dm;
    float s = 0.f;
    for (int r = lane; r < nblocks; r += 32) s += part[((size_t)r * 2 + q) * Dm + d];
    s = ab_warp_sum(s);
    if (lane == 0) (q == 0 ? dw : db)[d] = s;
}

int grid_rows(int S) {
    const int64_t want = ab_ceil_div(S, 8);
    const int64_t cap = (int64_t)ab_num_sms() * 8;
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int ab_layernorm_fwd(const void* x, const float* w, const float* b, float eps, void* y, float* stats, int S, int Dm,
                                int x_dtype, int y_dtype, cudaStream_t stream) {
    AB_REQUIRE(S > 0 && Dm > 0 && Dm % 4 == 0, "layernorm_fwd: need S > 0 and hidden size a multiple of 4 (S=%d Dm=%d)", S, Dm);
    const int g = grid_rows(S);
#define AB_LNF(TX, TY) ln_fwd_kernel<TX, TY><<<g, 256, 0, stream>>>((const TX*)x, w, b, eps, (TY*)y, stats, S, Dm)
    if (x_dtype == AB_F32 && y_dtype == AB_F32) AB_LNF(float, float);
    else if (x_dtype == AB_F32 && y_dtype == AB_BF16) AB_LNF(float, __nv_bfloat16);
    else if (x_dtype == AB_BF16 && y_dtype == AB_BF16) AB_LNF(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == AB_BF16 && y_dtype == AB_F32) AB_LNF(__nv_bfloat16, float);
    else AB_REQUIRE(false, "layernorm_fwd: bad dtypes");
#undef AB_LNF
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" size_t ab_layernorm_bwd_workspace_bytes(int S, int Dm) {
    const int64_t rows = ab_ceil_div(S, RB) > grid_rows(S) ? ab_ceil_div(S, RB) : grid_rows(S);     // partial rows of either path
    return (size_t)ab_round_up(rows * 2 * (int64_t)Dm * sizeof(float), 256);
}

extern "C" int ab_layernorm_bwd(const void* dy, const void* x, const float* stats, const float* w, const void* dres, void* dx,
                                float* dw, float* db, void* ws, size_t ws_bytes, int S, int Dm, int x_dtype, int dy_dtype,
                                cudaStream_t stream) {
    AB_REQUIRE(S > 0 && Dm > 0 && Dm % 4 == 0, "layernorm_bwd: need S > 0 and hidden size a multiple of 4");
    AB_REQUIRE(ws && ws_bytes >= ab_layernorm_bwd_workspace_bytes(S, Dm), "layernorm_bwd: workspace too small");
    int g = grid_rows(S);
    const int nit = (int)ab_ceil_div(Dm, 128);
    const bool fused = nit <= 8;                     // column sums fit in registers next to the row pass
    if (fused && g > 2 * ab_num_sms()) g = 2 * ab_num_sms();      // two CTAs per SM are resident: one wave, few partial rows
    const int nblocks = fused ? g : (int)ab_ceil_div(S, RB);
    dim3 cgrid(nblocks, (unsigned)ab_ceil_div(Dm, 128));
    float* part = (float*)ws;
#define AB_LNB_F(TX, TG, NIT) \
    ln_bwd_fused_kernel<TX, TG, NIT><<<g, 256, 0, stream>>>((const TG*)dy, (const TX*)x, stats, w, (const TX*)dres, (TX*)dx, part, S, Dm)
#define AB_LNB(TX, TG)                                                                                                     \
    {                                                                                                                      \
        if (fused) {                                                                                                       \
            if (nit <= 2) AB_LNB_F(TX, TG, 2); else if (nit <= 4) AB_LNB_F(TX, TG, 4);                                     \
            else if (nit <= 6) AB_LNB_F(TX, TG, 6); else AB_LNB_F(TX, TG, 8);                                              \
        } else {                                                                                                           \
            ln_bwd_rows_kernel<TX, TG><<<g, 256, 0, stream>>>((const TG*)dy, (const TX*)x, stats, w, (const TX*)dres, (TX*)dx, S, Dm); \
            ln_bwd_cols_kernel<TX, TG><<<cgrid, 256, 0, stream>>>((const TG*)dy, (const TX*)x, stats, part, S, Dm);       \
        }                                                                                                                  \
    }
    if (x_dtype == AB_F32 && dy_dtype == AB_F32) AB_LNB(float, float)
    else if (x_dtype == AB_F32 && dy_dtype == AB_BF16) AB_LNB(float, __nv_bfloat16)
    else if (x_dtype == AB_BF16 && dy_dtype == AB_BF16) AB_LNB(__nv_bfloat16, __nv_bfloat16)
    else if (x_dtype == AB_BF16 && dy_dtype == AB_F32) AB_LNB(__nv_bfloat16, float)
    else AB_REQUIRE(false, "layernorm_bwd: bad dtypes");
#undef AB_LNB
#undef AB_LNB_F
    AB_LAUNCH_CHECK();
    ln_param_reduce_kernel<<<(unsigned)ab_ceil_div(2 * Dm, 8), 256, 0, stream>>>(part, nblocks, Dm, dw, db);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

// ---------------------------------------------------------------------------------------------
// output dropout + residual add of the block wrappers (core.py:836-837, 918-919):  out = dropout(sub) + res
// ---------------------------------------------------------------------------------------------
namespace {

// out[i] = (keep ? sub[i]*scale : 0) + (res ? res[i] : 0); 4 elements per thread
template <typename TS, typename TR>
__global__ void __launch_bounds__(256) dropout_add_kernel(const TS* __restrict__ sub, const TR* __restrict__ res, TR* __restrict__ out,
                                                          const uint32_t* __restrict__ seed, uint32_t thresh, float scale, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float a[4], r[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
    ld4<TS>(sub + i, a);
    if (res) ld4<TR>(res + i, r);
    uint32_t s0 = 0, s1 = 0;
    if (seed) { s0 = __ldg(seed); s1 = __ldg(seed + 1); }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const bool keep = !seed || ab_out_keep(s0, s1, (uint64_t)(i + v), thresh);
        o[v] = (keep ? a[v] * scale : 0.f) + r[v];
    }
    st4<TR>(out + i, o);
}

}  // namespace

extern "C" int ab_dropout_add(const void* sub, const void* res, void* out, float p, const uint32_t* seed, int64_t n, int sub_dtype,
                              int out_dtype, cudaStream_t stream) {
    AB_REQUIRE(n > 0 && n % 4 == 0, "dropout_add: element count must be a positive multiple of 4");
    AB_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || seed), "dropout_add: p must be in [0,1) and needs a seed when > 0");
    const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
    const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    const uint32_t* sd = p > 0.f ? seed : nullptr;
    const unsigned grid = (unsigned)ab_ceil_div(n / 4, 256);
#define AB_DA(TS, TR) dropout_add_kernel<TS, TR><<<grid, 256, 0, stream>>>((const TS*)sub, (const TR*)res, (TR*)out, sd, thresh, scale, n)
    if (sub_dtype == AB_F32 && out_dtype == AB_F32) AB_DA(float, float);
    else if (sub_dtype == AB_BF16 && out_dtype == AB_F32) AB_DA(__nv_bfloat16, float);
    else if (sub_dtype == AB_BF16 && out_dtype == AB_BF16) AB_DA(__nv_bfloat16, __nv_bfloat16);
    else if (sub_dtype == AB_F32 && out_dtype == AB_BF16) AB_DA(float, __nv_bfloat16);
    else AB_REQUIRE(false, "dropout_add: bad dtypes");
#undef AB_DA
    AB_LAUNCH_CHECK();
    return AB_OK;
}
