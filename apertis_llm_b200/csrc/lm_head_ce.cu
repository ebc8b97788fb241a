// Shifted cross-entropy of the language-model head (core.py:1412-1460: logits[..., :-1, :] against labels[..., 1:],
// nn.CrossEntropyLoss(ignore_index=-100), mean over the counted positions), forward and backward, on logits that the
// library's own GEMM produced (ab_dense_gemm_nt of the hidden states with the tied embedding matrix).  The reference goes
// through a contiguous fp32 copy of the shifted logits, log_softmax, nll_loss and their autograd: five passes over [S, V];
// here the forward reads the logits once (online max / sum per row, one CTA per row), the backward writes d logits once.
#include "common.cuh"

namespace {

constexpr int CE_THREADS = 256;

template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

// (m, s) pairs of an online log-sum-exp: sum_i exp(x_i) = s * exp(m)
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    if (mm == -INFINITY) { m = mm; s = 0.f; return; }
    s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
    m = mm;
}

__device__ __forceinline__ void block_lse(float& m, float& s, float* sm /*[2 * 32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    if (lane == 0) { sm[warp] = m; sm[32 + warp] = s; }
    __syncthreads();
    if (warp == 0) {
        float mw = lane < nw ? sm[lane] : -INFINITY, sw = lane < nw ? sm[32 + lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, mw, o), s2 = __shfl_xor_sync(0xffffffffu, sw, o);
            lse_merge(mw, sw, m2, s2);
        }
        if (lane == 0) { sm[0] = mw; sm[32] = sw; }
    }
    __syncthreads();
    m = sm[0]; s = sm[32];
}

// grid: B * (L - 1) rows.  lse[r], row_loss[r] (0 for ignored positions), row_valid[r]
template <typename T>
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                            float* __restrict__ lse, float* __restrict__ row_loss,
                                                            float* __restrict__ row_valid, int L, int V, int64_t ignore_index) {
    __shared__ float sm[64];
    constexpr int N = Vec16<T>::N;
    const int r = blockIdx.x;
    const int b = r / (L - 1), l = r % (L - 1);
    const T* row = logits + ((size_t)b * L + l) * V;
    const int64_t y = labels[(size_t)b * L + l + 1];
    if (y != ignore_index && (y < 0 || y >= V)) __trap();      // torch raises a device assert for an out-of-range class
    float m = -INFINITY, s = 0.f;
    const int nvec = V / N;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[N];
        ab_vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(row) + i), f);
        float vm = f[0];
#pragma unroll
        for (int v = 1; v < N; ++v) vm = fmaxf(vm, f[v]);
        const float mm = fmaxf(m, vm);
        float acc = 0.f;
#pragma unroll
        for (int v = 0; v < N; ++v) acc += __expf(f[v] - mm);
        s = s * __expf(m - mm) + acc;
        m = mm;
    }
    for (int i = nvec * N + threadIdx.x; i < V; i += blockDim.x) lse_merge(m, s, ab_to_float(row[i]), 1.f);
    block_lse(m, s, sm);
    if (threadIdx.x == 0) {
        const float z = m + logf(s);
        lse[r] = z;
        const bool valid = y != ignore_index;
        row_valid[r] = valid ? 1.f : 0.f;
        row_loss[r] = valid ? z - ab_to_float(row[y]) : 0.f;
    }
}

// out[0] = sum of row_loss, out[1] = number of counted rows; one CTA, fixed order (bitwise reproducible)
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_valid, int n,
                                                         float* __restrict__ out) {
    __shared__ double sa[32], sb[32];
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { a += (double)row_loss[i]; c += (double)row_valid[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tc = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += sa[w]; tc += sb[w]; }
        out[0] = (float)ta;
        out[1] = (float)tc;
    }
}

// grid: B * L rows of d logits [B, L, V]: (softmax - onehot) * scale for the counted positions, zero elsewhere (the last
// position of every sequence and the ignored labels); scale[0] = upstream gradient / number of counted rows
template <typename T>
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                            const float* __restrict__ lse, const float* __restrict__ scale,
                                                            T* __restrict__ dlogits, int L, int V, int64_t ignore_index) {
    constexpr int N = Vec16<T>::N;
    const int row_id = blockIdx.x;
    const int b = row_id / L, l = row_id % L;
    T* out = dlogits + (size_t)row_id * V;
    const int nvec = V / N;
    const int64_t y = l < L - 1 ? labels[(size_t)b * L + l + 1] : ignore_index;
    if (l == L - 1 || y == ignore_index) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(out)[i] = z;
        for (int i = nvec * N + threadIdx.x; i < V; i += blockDim.x) out[i] = ab_from_float<T>(0.f);
        return;
    }
    const T* row = logits + (size_t)row_id * V;
    const float z = lse[(size_t)b * (L - 1) + l], sc = scale[0];
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[N];
        ab_vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(row) + i), f);
#pragma unroll
        for (int v = 0; v < N; ++v) {
            const float p = __expf(f[v] - z);
            f[v] = (p - ((int64_t)(i * N + v) == y ? 1.f : 0.f)) * sc;
        }
        reinterpret_cast<uint4*>(out)[i] = ab_vec16<T>::pack(f);
    }
    for (int i = nvec * N + threadIdx.x; i < V; i += blockDim.x) {
        const float p = __expf(ab_to_float(row[i]) - z);
        out[i] = ab_from_float<T>((p - ((int64_t)i == y ? 1.f : 0.f)) * sc);
    }
}

}  // namespace

extern "C" int ab_shifted_ce_fwd(const void* logits, const int64_t* labels, float* lse, float* row_loss, float* row_valid,
                                 float* sums, int B, int L, int V, int64_t ignore_index, int dtype, cudaStream_t stream) {
    AB_REQUIRE(B > 0 && L > 1 && V > 0, "shifted_ce_fwd: needs B > 0, L > 1, V > 0 (got %d, %d, %d)", B, L, V);
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "shifted_ce_fwd: bad dtype");
    AB_REQUIRE(((uintptr_t)logits % 16) == 0 && ((size_t)V * (dtype == AB_F32 ? 4 : 2)) % 16 == 0,
               "shifted_ce_fwd: logits rows must be 16-byte aligned (V a multiple of %d)", dtype == AB_F32 ? 4 : 8);
    const int n = B * (L - 1);
    if (dtype == AB_F32) ce_fwd_kernel<float><<<n, CE_THREADS, 0, stream>>>((const float*)logits, labels, lse, row_loss, row_valid, L, V, ignore_index);
    else ce_fwd_kernel<__nv_bfloat16><<<n, CE_THREADS, 0, stream>>>((const __nv_bfloat16*)logits, labels, lse, row_loss, row_valid, L, V, ignore_index);
    AB_LAUNCH_CHECK();
    ce_reduce_kernel<<<1, 1024, 0, stream>>>(row_loss, row_valid, n, sums);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_shifted_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* scale, void* dlogits,
                                 int B, int L, int V, int64_t ignore_index, int dtype, cudaStream_t stream) {
    AB_REQUIRE(B > 0 && L > 1 && V > 0, "shifted_ce_bwd: needs B > 0, L > 1, V > 0 (got %d, %d, %d)", B, L, V);
    AB_REQUIRE(dtype == AB_F32 || dtype == AB_BF16, "shifted_ce_bwd: bad dtype");
    AB_REQUIRE(((uintptr_t)logits % 16) == 0 && ((uintptr_t)dlogits % 16) == 0 && ((size_t)V * (dtype == AB_F32 ? 4 : 2)) % 16 == 0,
               "shifted_ce_bwd: rows must be 16-byte aligned (V a multiple of %d)", dtype == AB_F32 ? 4 : 8);
    if (dtype == AB_F32) ce_bwd_kernel<float><<<B * L, CE_THREADS, 0, stream>>>((const float*)logits, labels, lse, scale, (float*)dlogits, L, V, ignore_index);
    else ce_bwd_kernel<__nv_bfloat16><<<B * L, CE_THREADS, 0, stream>>>((const __nv_bfloat16*)logits, labels, lse, scale, (__nv_bfloat16*)dlogits, L, V, ignore_index);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
