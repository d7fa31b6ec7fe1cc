// Shared helpers for the mridc_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mridc_b200.h"

namespace mrb {

// ---- error plumbing --------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local long long g_launches;

#define MRB_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            mrb::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

#define MRB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            mrb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MRB_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define MRB_LAUNCHED()                                                                              \
    do {                                                                                            \
        ++mrb::g_launches;                                                                          \
        cudaError_t _e = cudaPeekAtLastError();                                                     \
        if (_e != cudaSuccess) {                                                                    \
            (void)cudaGetLastError();                                                               \
            mrb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MRB_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

// ---- complex helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int device_sm_count();
size_t device_max_smem_optin();

// Mask descriptor shared by the fused DC kernels (see include/mridc_b200.h).
struct MaskDesc {
    const void* ptr;
    int dtype;   // MRB_MASK_U8 / MRB_MASK_F32
    int bstride; // 0 if broadcast over batch, else mask_h_eff * W
    int hstride; // 0 if broadcast over H, else W
};

__device__ __forceinline__ float mask_value(const MaskDesc& m, int b, int h, int w) {
    long long i = (long long)b * m.bstride + (long long)h * m.hstride + w;
    return m.dtype == MRB_MASK_F32 ? ((const float*)m.ptr)[i] : (float)((const unsigned char*)m.ptr)[i];
}

}  // namespace mrb
