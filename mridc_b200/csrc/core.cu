// Library plumbing + element-wise / coil-reduction primitives.
// Reference behaviour: mridc/collections/common/parts/utils.py:96-272.
#include <stdarg.h>

#include "common.cuh"

namespace mrb {

static thread_local char g_err[512] = "";
thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) cached = v;
        cached_dev = dev;
    }
    return cached;
}

size_t device_max_smem_optin() {
    static thread_local int cached_dev = -1;
    static thread_local size_t cached = 227 * 1024;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0)
            cached = (size_t)v;
        cached_dev = dev;
    }
    return cached;
}

// ---- kernels ---------------------------------------------------------------------------------------
struct Strided6 {
    long long shape[6];
    long long xs[6];
    long long ys[6];
};

__global__ void complex_mul_kernel(const float2* __restrict__ x, const float2* __restrict__ y,
                                   float2* __restrict__ out, long long total, Strided6 d, int conj_y) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long r = t, xo = 0, yo = 0;
#pragma unroll
        for (int i = 5; i >= 0; --i) {
            long long c = r % d.shape[i];
            r /= d.shape[i];
            xo += c * d.xs[i];
            yo += c * d.ys[i];
        }
        float2 a = x[xo], b = y[yo];
        // same operation order as utils.py:115-116 (re = xr*yr - xi*yi ; im = xr*yi + xi*yr)
        if (conj_y) b.y = -b.y;
        out[t] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
}

__global__ void complex_conj_kernel(const float2* __restrict__ x, float2* __restrict__ out, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        float2 a = x[t];
        out[t] = make_float2(a.x, -a.y);
    }
}

__global__ void complex_abs_kernel(const float2* __restrict__ x, float* __restrict__ out, long long n, int squared) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        float2 a = x[t];
        float s = a.x * a.x + a.y * a.y;
        out[t] = squared ? s : sqrtf(s);
    }
}

// x [outer, C, inner] -> out [outer, inner]; MODE 0: real rss, 1: complex rss, 2: sense combine with S
template <int MODE>
__global__ void coil_reduce_kernel(const void* __restrict__ xv, const void* __restrict__ sv, void* __restrict__ ov,
                                   long long outer, int C, long long inner) {
    const long long total = outer * inner;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long o = t / inner, i = t - o * inner;
        long long base = o * C * inner + i;
        if (MODE == 0) {
            const float* x = (const float*)xv;
            float acc = 0.f;
            for (int c = 0; c < C; ++c) {
                float v = x[base + c * inner];
                acc += v * v;
            }
            ((float*)ov)[t] = sqrtf(acc);
        } else if (MODE == 1) {
            const float2* x = (const float2*)xv;
            float acc = 0.f;
            for (int c = 0; c < C; ++c) {
                float2 v = x[base + c * inner];
                acc += v.x * v.x + v.y * v.y;
            }
            ((float*)ov)[t] = sqrtf(acc);
        } else {
            const float2* x = (const float2*)xv;
            const float2* s = (const float2*)sv;
            float2 acc = make_float2(0.f, 0.f);
            for (int c = 0; c < C; ++c) {
                float2 v = x[base + c * inner], m = s[base + c * inner];
                // complex_mul(x, conj(S)): re = xr*sr - xi*(-si) ; im = xr*(-si) + xi*sr  (utils.py:248)
                acc.x += v.x * m.x - v.y * (-m.y);
                acc.y += v.x * (-m.y) + v.y * m.x;
            }
            ((float2*)ov)[t] = acc;
        }
    }
}

// BaseSensitivityModel.divide_root_sum_of_squares (reconstruction/models/base.py:826-840): x [outer, C, inner] complex
// -> x / sqrt(sum_c |x_c|^2); the second read of x hits L1/L2
__global__ void divide_rss_kernel(const float2* __restrict__ x, float2* __restrict__ out, long long outer, int C,
                                  long long inner) {
    const long long total = outer * inner;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long o = t / inner, i = t - o * inner;
        const long long base = o * C * inner + i;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float2 v = x[base + c * inner];
            acc += v.x * v.x + v.y * v.y;
        }
        const float r = sqrtf(acc);
        for (int c = 0; c < C; ++c) {
            const float2 v = x[base + c * inner];
            out[base + c * inner] = make_float2(__fdiv_rn(v.x, r), __fdiv_rn(v.y, r));
        }
    }
}

static inline unsigned grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    long long cap = (long long)device_sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace mrb

using namespace mrb;

extern "C" const char* mrb_last_error(void) { return g_err; }
extern "C" int mrb_version(void) { return 100; }
extern "C" long long mrb_launch_count(void) { return g_launches; }
extern "C" void mrb_reset_launch_count(void) { g_launches = 0; }
extern "C" void mrb_add_launch_count(long long n) { mrb::g_launches += n; }

extern "C" int mrb_complex_mul(const void* x, const void* y, void* out, int ndim, const long long* shape,
                               const long long* xstride, const long long* ystride, int conj_y, void* stream) {
    MRB_REQUIRE(x && y && out, MRB_EINVAL, "mrb_complex_mul: null pointer");
    MRB_REQUIRE(ndim >= 0 && ndim <= 6, MRB_EINVAL, "mrb_complex_mul: ndim %d > 6", ndim);
    Strided6 d;
    long long total = 1;
    for (int i = 0; i < 6; ++i) {
        int src = i - (6 - ndim);
        d.shape[i] = src >= 0 ? shape[src] : 1;
        d.xs[i] = src >= 0 ? xstride[src] : 0;
        d.ys[i] = src >= 0 ? ystride[src] : 0;
        MRB_REQUIRE(d.shape[i] >= 0, MRB_EINVAL, "mrb_complex_mul: negative extent");
        total *= d.shape[i];
    }
    if (total == 0) return MRB_OK;
    complex_mul_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (const float2*)y,
                                                                              (float2*)out, total, d, conj_y);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_complex_conj(const void* x, void* out, long long n, void* stream) {
    MRB_REQUIRE(x && out && n >= 0, MRB_EINVAL, "mrb_complex_conj: bad argument");
    if (n == 0) return MRB_OK;
    complex_conj_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (float2*)out, n);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_complex_abs(const void* x, void* out, long long n, int squared, void* stream) {
    MRB_REQUIRE(x && out && n >= 0, MRB_EINVAL, "mrb_complex_abs: bad argument");
    if (n == 0) return MRB_OK;
    complex_abs_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (float*)out, n, squared);
    MRB_LAUNCHED();
    return MRB_OK;
}

template <int MODE>
static int coil_reduce(const void* x, const void* S, void* out, long long outer, int C, long long inner,
                       void* stream, const char* who) {
    MRB_REQUIRE(x && out && (MODE != 2 || S), MRB_EINVAL, "%s: null pointer", who);
    MRB_REQUIRE(outer >= 0 && inner >= 0 && C >= 0, MRB_EINVAL, "%s: negative extent", who);
    if (outer * inner == 0) return MRB_OK;
    coil_reduce_kernel<MODE><<<grid_for(outer * inner, 256), 256, 0, (cudaStream_t)stream>>>(x, S, out, outer, C, inner);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_divide_rss(const void* x, void* out, long long outer, int C, long long inner, void* stream) {
    MRB_REQUIRE(x && out, MRB_EINVAL, "mrb_divide_rss: null pointer");
    MRB_REQUIRE(outer >= 0 && inner >= 0 && C >= 0, MRB_EINVAL, "mrb_divide_rss: negative extent");
    if (outer * inner * C == 0) return MRB_OK;
    divide_rss_kernel<<<grid_for(outer * inner, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (float2*)out, outer,
                                                                                       C, inner);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_rss(const void* x, void* out, long long outer, int C, long long inner, void* stream) {
    return coil_reduce<0>(x, nullptr, out, outer, C, inner, stream, "mrb_rss");
}
extern "C" int mrb_rss_complex(const void* x, void* out, long long outer, int C, long long inner, void* stream) {
    return coil_reduce<1>(x, nullptr, out, outer, C, inner, stream, "mrb_rss_complex");
}
extern "C" int mrb_sense_combine(const void* x, const void* S, void* out, long long outer, int C, long long inner,
                                 void* stream) {
    return coil_reduce<2>(x, S, out, outer, C, inner, stream, "mrb_sense_combine");
}
