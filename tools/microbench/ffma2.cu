// Microbenchmark: fp32 FMA throughput with scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2.cu -o ffma2 && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
constexpr int ITERS = 4096, CH = 8;
__global__ void k_scalar(float* out, float s) {
    float a[2 * CH];
    for (int i = 0; i < 2 * CH; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) a[i] = fmaf(a[i], s, 0.5f);
    float r = 0; for (int i = 0; i < 2 * CH; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_packed(float* out, float s) {
    unsigned long long a[CH];
    const unsigned long long s2 = pk(s, s), c2 = pk(0.5f, 0.5f);
    for (int i = 0; i < CH; ++i) a[i] = pk(threadIdx.x + i, threadIdx.x + i + 0.5f);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma2(a[i], s2, c2);
    float r = 0; for (int i = 0; i < CH; ++i) { float x, y; upk(a[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 8, 1024>>>(out, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 148.0 * 8 * 1024 * ITERS * 2 * CH * 2;
        printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 1024>>>(out, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    return 0;
}
