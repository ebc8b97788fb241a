// Weight plumbing of the fused SSM parameter projection (core.py:376-383).
//
// The reference computes  p = xa Wp^T, splits it into (dtf [R] | B [Di] | C [Di]) and then  dt = dtf Wdt^T + b.  Since
// dtf feeds nothing else, dt = xa (Wdt Wp[0:R])^T + b exactly: one GEMM with the stacked weight
//     Wcat [(Hp + 2 Di), Di] = [ Wdt Wp[0:R]  (H rows) ; 0  (Hp - H rows) ; Wp[R : R + 2 Di] ]
// produces [dt (without bias) | pad | B | C] rows that the scan kernels read in place (Hp = H rounded up to 8 keeps every
// column block 16-byte aligned).  These two kernels build Wcat from the module's parameters and map its gradient back.
#include "common.cuh"

namespace {

template <typename T>
__global__ void dt_compose_fwd_kernel(const float* __restrict__ Wp, const float* __restrict__ Wdt, T* __restrict__ Wcat, int H, int Hp, int R, int Di) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (c >= Di) return;
    float v = 0.f;
    if (row < H) {
        for (int r = 0; r < R; ++r) v = fmaf(Wdt[row * R + r], Wp[(size_t)r * Di + c], v);
    } else if (row >= Hp) {
        v = Wp[(size_t)(R + row - Hp) * Di + c];
    }
    Wcat[(size_t)row * Di + c] = ab_from_float<T>(v);
}

// dWp[r, c] = sum_h Wdt[h, r] dWcat[h, c] (r < R);  dWp[R + i, c] = dWcat[Hp + i, c]
__global__ void dt_compose_bwd_wp_kernel(const float* __restrict__ dWcat, const float* __restrict__ Wdt, float* __restrict__ dWp, int H, int Hp, int R, int Di) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (c >= Di) return;
    float v = 0.f;
    if (row < R) {
        for (int h = 0; h < H; ++h) v = fmaf(Wdt[h * R + row], dWcat[(size_t)h * Di + c], v);
    } else {
        v = dWcat[(size_t)(Hp + row - R) * Di + c];
    }
    dWp[(size_t)row * Di + c] = v;
}

// dWdt[h, r] = sum_c dWcat[h, c] Wp[r, c]: one warp per (h, r), fixed-order lane partials + butterfly (reproducible)
__global__ void dt_compose_bwd_wdt_kernel(const float* __restrict__ dWcat, const float* __restrict__ Wp, float* __restrict__ dWdt, int H, int R, int Di) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= H * R) return;
    const int h = w / R, r = w % R;
    float v = 0.f;
    for (int c = lane; c < Di; c += 32) v = fmaf(dWcat[(size_t)h * Di + c], Wp[(size_t)r * Di + c], v);
    v = ab_warp_sum(v);
    if (lane == 0) dWdt[w] = v;
}

}  // namespace

extern "C" int ab_dt_compose_fwd(const float* Wp, const float* Wdt, void* Wcat, int H, int Hp, int R, int Di, int out_dtype,
                                 cudaStream_t stream) {
    AB_REQUIRE(Wp && Wdt && Wcat && H > 0 && Hp >= H && R > 0 && Di > 0, "dt_compose_fwd: bad arguments");
    AB_REQUIRE(out_dtype == AB_F32 || out_dtype == AB_BF16, "dt_compose_fwd: bad dtype");
    dim3 grid((unsigned)ab_ceil_div(Di, 128), (unsigned)(Hp + 2 * Di));
    if (out_dtype == AB_F32) dt_compose_fwd_kernel<float><<<grid, 128, 0, stream>>>(Wp, Wdt, (float*)Wcat, H, Hp, R, Di);
    else dt_compose_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(Wp, Wdt, (__nv_bfloat16*)Wcat, H, Hp, R, Di);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_dt_compose_bwd(const float* dWcat, const float* Wp, const float* Wdt, float* dWp, float* dWdt, int H, int Hp,
                                 int R, int Di, cudaStream_t stream) {
    AB_REQUIRE(dWcat && Wp && Wdt && dWp && dWdt && H > 0 && Hp >= H && R > 0 && Di > 0, "dt_compose_bwd: bad arguments");
    dim3 grid((unsigned)ab_ceil_div(Di, 128), (unsigned)(R + 2 * Di));
    dt_compose_bwd_wp_kernel<<<grid, 128, 0, stream>>>(dWcat, Wdt, dWp, H, Hp, R, Di);
    AB_LAUNCH_CHECK();
    dt_compose_bwd_wdt_kernel<<<(unsigned)ab_ceil_div((int64_t)H * R, 4), 128, 0, stream>>>(dWcat, Wp, dWdt, H, R, Di);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
