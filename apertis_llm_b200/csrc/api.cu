// Library plumbing: error text, device check, TMA tensor-map encoding, dtype helpers.
#include <stdarg.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace

void ab_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ab_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int ab_encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* base, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            ab_set_error("cuTensorMapEncodeTiled driver entry point not available (%s)", cudaGetErrorString(e));
            return AB_ERR_CUDA;
        }
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t gdims[5];
    cuuint64_t gstr[5];
    cuuint32_t gbox[5], estr[5];
    for (uint32_t i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
    for (uint32_t i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = g_encode(map, dt, rank, const_cast<void*>(base), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // a thread that has not touched the runtime yet (autograd's backward thread served from the caching allocator):
        // bind the primary context and try again
        cudaFree(nullptr);
        r = g_encode(map, dt, rank, const_cast<void*>(base), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        ab_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %u dims [%llu,%llu,%llu] box [%u,%u,%u] stride0 %llu base %p",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
        return AB_ERR_CUDA;
    }
    return AB_OK;
}

namespace {

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        *reinterpret_cast<uint4*>(dst + i) = ab_vec16<__nv_bfloat16>::pack(f);
    } else {
        for (int64_t k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
}

// hi = bf16(x), lo = bf16(x - hi); dst row r (3*cols wide) = which==0 ? [hi|hi|lo] : [hi|lo|hi]
__global__ void split3_cols_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int64_t cols,
                                   int which) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i % cols;
    const float x = src[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    __nv_bfloat16* d = dst + r * 3 * cols;
    d[c] = hi;
    d[cols + c] = which == 0 ? hi : lo;
    d[2 * cols + c] = which == 0 ? lo : hi;
}

// row-stacked variant: group g occupies src rows [off[g], off[g+1]); dst rows [3*off[g], 3*off[g+1]) hold
// which==0 ? [hi; hi; lo] : [hi; lo; hi] (each block len rows).  off == NULL -> uniform groups of rows_per_group.
__global__ void split3_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, const int32_t* __restrict__ off,
                                   int G, int64_t rows_per_group, int64_t cols, int which) {
    const int g = blockIdx.y;
    const int64_t r0 = off ? off[g] : g * rows_per_group;
    const int64_t len = off ? (off[g + 1] - off[g]) : rows_per_group;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i % cols;
        const float x = src[(r0 + r) * cols + c];
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        __nv_bfloat16* d = dst + 3 * r0 * cols;
        d[r * cols + c] = hi;
        d[(len + r) * cols + c] = which == 0 ? hi : lo;
        d[(2 * len + r) * cols + c] = which == 0 ? lo : hi;
    }
    (void)G;
}

}  // namespace

extern "C" int ab_version(void) { return 200; }      // 2xx: round-2 ABI (256-row GEMM tile, ab_ep_*, ab_shifted_ce_*, wgrad workspace)

extern "C" int ab_device_check(int device) {
    cudaDeviceProp prop;
    AB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        ab_set_error("apertis_b200 kernels are built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
        return AB_ERR_UNSUPPORTED;
    }
    return AB_OK;
}

extern "C" int ab_last_error(char* buf, size_t n) {
    if (!buf || !n) return (int)strlen(g_err);
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
    return (int)strlen(buf);
}

extern "C" int ab_cast_f32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t stream) {
    if (n <= 0) return AB_OK;
    AB_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "cast: pointers must be 16-byte aligned");
    cast_kernel<<<(unsigned)ab_ceil_div(ab_ceil_div(n, 8), 256), 256, 0, stream>>>(src, (__nv_bfloat16*)dst, n);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_split_f32_to_bf16x3(const float* src, void* dst, int64_t rows, int64_t cols, int which, cudaStream_t stream) {
    if (rows * cols <= 0) return AB_OK;
    split3_cols_kernel<<<(unsigned)ab_ceil_div(rows * cols, 256), 256, 0, stream>>>(src, (__nv_bfloat16*)dst, rows, cols, which);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" int ab_split_f32_to_bf16x3_rows(const float* src, void* dst, const int32_t* seg_off, int G, int64_t rows_per_group,
                                           int64_t cols, int which, cudaStream_t stream) {
    if (G <= 0 || cols <= 0) return AB_OK;
    dim3 grid(256, G);
    split3_rows_kernel<<<grid, 256, 0, stream>>>(src, (__nv_bfloat16*)dst, seg_off, G, rows_per_group, cols, which);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
