// Causal depthwise conv1d (k = 4) + SiLU, forward and recompute-based backward.
// Replaces core.py:368-375 (two transposes + nn.Conv1d(groups=Di, padding=3)[:, :, :L] + F.silu) and
// their autograd.  Layout is channels-last [B, L, Di]: a thread owns one 16-byte channel vector and
// walks TT consecutive tokens with a sliding window in registers, so every global access is a
// coalesced 128-bit load/store and each input row is fetched from HBM once (halo rows hit L1/L2).
#include "common.cuh"

namespace {

constexpr int KC = 4;
constexpr int TT = 16;       // tokens per thread
constexpr int ROWS_PER_CTA = 8;   // thread rows (token runs) per CTA

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, bool ok, float* f) {
    constexpr int V = ab_vec16<T>::N;
    if (ok) {
        uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        ab_vec16<T>::unpack(r, f);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = 0.f;
    }
}

// grid.x = ceil(ncv / blockDim.x), grid.y = ceil(L / (TT*ROWS_PER_CTA)), grid.z = B ; block = (cvx, ROWS_PER_CTA)
template <typename T>
__global__ void __launch_bounds__(256) conv_silu_fwd_kernel(const T* __restrict__ xp, int64_t xp_stride,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            T* __restrict__ xa, int L, int Di) {
    constexpr int V = ab_vec16<T>::N;
    const int cv = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = cv * V;
    if (c0 >= Di) return;
    const int b = blockIdx.z;
    const int t0 = (blockIdx.y * ROWS_PER_CTA + threadIdx.y) * TT;
    if (t0 >= L) return;
    float wr[V][KC], br[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(c0 + i) * KC));
        wr[i][0] = w4.x; wr[i][1] = w4.y; wr[i][2] = w4.z; wr[i][3] = w4.w;
        br[i] = __ldg(bias + c0 + i);
    }
    const T* xrow = xp + (size_t)b * L * xp_stride + c0;
    T* orow = xa + ((size_t)b * L) * Di + c0;
    float win[KC - 1][V];   // x[t-3], x[t-2], x[t-1]
#pragma unroll
    for (int j = 0; j < KC - 1; ++j) {
        const int t = t0 - (KC - 1) + j;
        load_vec<T>(xrow + (size_t)t * xp_stride, t >= 0, win[j]);
    }
#pragma unroll 4
    for (int i = 0; i < TT; ++i) {
        const int t = t0 + i;
        if (t >= L) break;
        float cur[V], o[V];
        load_vec<T>(xrow + (size_t)t * xp_stride, true, cur);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float acc = br[v];
            acc = fmaf(wr[v][0], win[0][v], acc);
            acc = fmaf(wr[v][1], win[1][v], acc);
            acc = fmaf(wr[v][2], win[2][v], acc);
            acc = fmaf(wr[v][3], cur[v], acc);
            o[v] = acc * ab_sigmoid(acc);
            win[0][v] = win[1][v]; win[1][v] = win[2][v]; win[2][v] = cur[v];
        }
        *reinterpret_cast<uint4*>(orow + (size_t)t * Di) = ab_vec16<T>::pack(o);
    }
}

// Backward.  For the owned tokens t0..t0+TT-1 the thread needs dxc at t..t+3, i.e. xc (hence xp rows
// t-3..t+3) and dxa rows t..t+3.  It streams u = t0 .. t0+TT+2 computing dxc[u] once each, keeps a
// 4-deep window of dxc and a 7-deep window of xp rows is avoided by re-deriving: dxp[t] needs
// dxc[t..t+3]; dw[c,j] += dxc[u]*xp[u-3+j] only for owned u (so every (u, j) pair is counted once).
template <typename T>
__global__ void __launch_bounds__(256) conv_silu_bwd_kernel(const T* __restrict__ xp, int64_t xp_stride,
                                                            const T* __restrict__ dxa, const float* __restrict__ w,
                                                            const float* __restrict__ bias, T* __restrict__ dxp, int64_t dxs,
                                                            float* __restrict__ part, int L, int Di, int n_part_rows) {
    constexpr int V = ab_vec16<T>::N;
    const int cv = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = cv * V;
    const int b = blockIdx.z;
    const int t0 = (blockIdx.y * ROWS_PER_CTA + threadIdx.y) * TT;
    const bool active = (c0 < Di) && (t0 < L);
    float gw[V][KC], gb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { gb[v] = 0.f; gw[v][0] = gw[v][1] = gw[v][2] = gw[v][3] = 0.f; }
    if (active) {
        float wr[V][KC], br[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(c0 + i) * KC));
            wr[i][0] = w4.x; wr[i][1] = w4.y; wr[i][2] = w4.z; wr[i][3] = w4.w;
            br[i] = __ldg(bias + c0 + i);
        }
        const T* xrow = xp + (size_t)b * L * xp_stride + c0;
        const T* grow = dxa + ((size_t)b * L) * Di + c0;
        T* orow = dxp + ((size_t)b * L) * dxs + c0;
        float win[KC - 1][V];      // xp[u-3], xp[u-2], xp[u-1]
#pragma unroll
        for (int j = 0; j < KC - 1; ++j) {
            const int t = t0 - (KC - 1) + j;
            load_vec<T>(xrow + (size_t)t * xp_stride, t >= 0, win[j]);
        }
        float dwin[KC - 1][V];     // dxc[u-3], dxc[u-2], dxc[u-1]
#pragma unroll
        for (int j = 0; j < KC - 1; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) dwin[j][v] = 0.f;
        const int u_end = min(t0 + TT + KC - 1, L + KC - 1);
        for (int u = t0; u < u_end; ++u) {
            float cur[V], g[V], dxc[V];
            const bool in_seq = u < L;
            load_vec<T>(xrow + (size_t)u * xp_stride, in_seq, cur);
            load_vec<T>(grow + (size_t)u * Di, in_seq, g);
            const bool owned = in_seq && (u < t0 + TT);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float acc = br[v];
                acc = fmaf(wr[v][0], win[0][v], acc);
                acc = fmaf(wr[v][1], win[1][v], acc);
                acc = fmaf(wr[v][2], win[2][v], acc);
                acc = fmaf(wr[v][3], cur[v], acc);
                const float s = ab_sigmoid(acc);
                // d silu(x)/dx = s * (1 + x * (1 - s))
                dxc[v] = in_seq ? g[v] * s * fmaf(acc, 1.f - s, 1.f) : 0.f;
                if (owned) {
                    gb[v] += dxc[v];
                    gw[v][0] = fmaf(dxc[v], win[0][v], gw[v][0]);
                    gw[v][1] = fmaf(dxc[v], win[1][v], gw[v][1]);
                    gw[v][2] = fmaf(dxc[v], win[2][v], gw[v][2]);
                    gw[v][3] = fmaf(dxc[v], cur[v], gw[v][3]);
                }
            }
            // dxp[t] for t = u-3: sum_j w[j] * dxc[t + 3 - j] = w3*dxc[u-3] + w2*dxc[u-2] + w1*dxc[u-1] + w0*dxc[u]
            const int t = u - (KC - 1);
            if (t >= t0 && t < L) {
                float o[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float acc = wr[v][3] * dwin[0][v];
                    acc = fmaf(wr[v][2], dwin[1][v], acc);
                    acc = fmaf(wr[v][1], dwin[2][v], acc);
                    acc = fmaf(wr[v][0], dxc[v], acc);
                    o[v] = acc;
                }
                *reinterpret_cast<uint4*>(orow + (size_t)t * dxs) = ab_vec16<T>::pack(o);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                win[0][v] = win[1][v]; win[1][v] = win[2][v]; win[2][v] = cur[v];
                dwin[0][v] = dwin[1][v]; dwin[1][v] = dwin[2][v]; dwin[2][v] = dxc[v];
            }
        }
    }
    // CTA reduction over the ROWS_PER_CTA thread rows, then one partial row per CTA: part[prow][c][5]
    extern __shared__ float sred[];   // [ROWS_PER_CTA][blockDim.x * V * 5]
    const int per_row = blockDim.x * V * 5;
    float* mine = sred + threadIdx.y * per_row + threadIdx.x * V * 5;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        mine[v * 5 + 0] = gw[v][0]; mine[v * 5 + 1] = gw[v][1]; mine[v * 5 + 2] = gw[v][2]; mine[v * 5 + 3] = gw[v][3];
        mine[v * 5 + 4] = gb[v];
    }
    __syncthreads();
    const int prow = blockIdx.z * gridDim.y + blockIdx.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < per_row; i += blockDim.x * blockDim.y) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS_PER_CTA; ++r) s += sred[r * per_row + i];
        const int c = blockIdx.x * blockDim.x * V + i / 5;
        if (c < Di) part[((size_t)prow * Di + c) * 5 + (i % 5)] = s;
    }
    (void)n_part_rows;
}

// out[c*5+q] = sum over partial rows, fixed order (deterministic)
__global__ void conv_reduce_kernel(const float* __restrict__ part, int n_rows, int Di, float* __restrict__ dw,
                                   float* __restrict__ dbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Di * 5) return;
    float s = 0.f;
    for (int r = 0; r < n_rows; ++r) s += part[(size_t)r * Di * 5 + i];
    const int c = i / 5, q = i % 5;
    if (q < 4) dw[c * KC + q] = s; else dbias[c] = s;
}

template <typename T>
int launch_fwd(const void* xp, int64_t xs, const float* w, const float* bias, void* xa, int B, int L, int Di, cudaStream_t st) {
    constexpr int V = ab_vec16<T>::N;
    const int ncv = Di / V;
    const int bx = ncv >= 32 ? 32 : ncv;
    dim3 block(bx, ROWS_PER_CTA);
    dim3 grid((unsigned)ab_ceil_div(ncv, bx), (unsigned)ab_ceil_div(L, TT * ROWS_PER_CTA), B);
    conv_silu_fwd_kernel<T><<<grid, block, 0, st>>>((const T*)xp, xs, w, bias, (T*)xa, L, Di);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

template <typename T>
int launch_bwd(const void* xp, int64_t xs, const void* dxa, const float* w, const float* bias, void* dxp, int64_t dxs, float* dw,
               float* dbias, float* part, int B, int L, int Di, cudaStream_t st) {
    constexpr int V = ab_vec16<T>::N;
    const int ncv = Di / V;
    const int bx = ncv >= 32 ? 32 : ncv;
    dim3 block(bx, ROWS_PER_CTA);
    dim3 grid((unsigned)ab_ceil_div(ncv, bx), (unsigned)ab_ceil_div(L, TT * ROWS_PER_CTA), B);
    const size_t smem = (size_t)ROWS_PER_CTA * bx * V * 5 * sizeof(float);
    const int n_part = grid.y * grid.z;
    conv_silu_bwd_kernel<T><<<grid, block, smem, st>>>((const T*)xp, xs, (const T*)dxa, w, bias, (T*)dxp, dxs, part, L, Di, n_part);
    AB_LAUNCH_CHECK();
    conv_reduce_kernel<<<(unsigned)ab_ceil_div(Di * 5, 128), 128, 0, st>>>(part, n_part, Di, dw, dbias);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

int check_args(int B, int L, int Di, int Kc, int64_t xs, int dtype) {
    AB_REQUIRE(Kc == KC, "causal_conv1d: only ssm_conv_kernel == 4 is supported (got %d)", Kc);
    AB_REQUIRE(B > 0 && L > 0 && Di > 0, "causal_conv1d: empty shape B=%d L=%d Di=%d", B, L, Di);
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "causal_conv1d: bad dtype %d", dtype);
    const int V = dtype == AB_F32 ? 4 : 8;
    AB_REQUIRE(Di % V == 0 && xs % V == 0, "causal_conv1d: Di (%d) and row stride (%lld) must be multiples of %d", Di, (long long)xs, V);
    return AB_OK;
}

}  // namespace

extern "C" int ab_causal_conv1d_silu_fwd(const void* xp, int64_t xp_stride, const float* w, const float* bias, void* xa,
                                         int B, int L, int Di, int Kc, int dtype, cudaStream_t stream) {
    if (int e = check_args(B, L, Di, Kc, xp_stride, dtype)) return e;
    return dtype == AB_F32 ? launch_fwd<float>(xp, xp_stride, w, bias, xa, B, L, Di, stream)
                           : launch_fwd<__nv_bfloat16>(xp, xp_stride, w, bias, xa, B, L, Di, stream);
}

extern "C" size_t ab_causal_conv1d_silu_bwd_workspace_bytes(int B, int L, int Di) {
    return (size_t)B * ab_ceil_div(L, TT * ROWS_PER_CTA) * Di * 5 * sizeof(float);
}

extern "C" int ab_causal_conv1d_silu_bwd(const void* xp, int64_t xp_stride, const void* dxa, const float* w,
                                         const float* bias, void* dxp, int64_t dxp_stride, float* dw, float* dbias, void* ws,
                                         size_t ws_bytes, int B, int L, int Di, int Kc, int dtype, cudaStream_t stream) {
    if (int e = check_args(B, L, Di, Kc, xp_stride, dtype)) return e;
    if (int e = check_args(B, L, Di, Kc, dxp_stride, dtype)) return e;
    AB_REQUIRE(dxp_stride >= Di, "causal_conv1d_bwd: dxp row stride smaller than the row");
    AB_REQUIRE(ws_bytes >= ab_causal_conv1d_silu_bwd_workspace_bytes(B, L, Di), "causal_conv1d_bwd: workspace too small");
    return dtype == AB_F32
               ? launch_bwd<float>(xp, xp_stride, dxa, w, bias, dxp, dxp_stride, dw, dbias, (float*)ws, B, L, Di, stream)
               : launch_bwd<__nv_bfloat16>(xp, xp_stride, dxa, w, bias, dxp, dxp_stride, dw, dbias, (float*)ws, B, L, Di, stream);
}
