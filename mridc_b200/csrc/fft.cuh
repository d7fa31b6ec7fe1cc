// Block-cooperative mixed-radix Stockham FFT over lines resident in shared memory.
//
// Every transform on the hot path (standalone fft2/ifft2 and the fused data-consistency kernels) runs
// through block_fft<>: `nlines` independent length-n complex lines live in shared memory (line l at
// base + l*ls), all threads of the CTA share the butterflies of one autosort stage, two buffers ping-pong
// between stages.  Radices 4, 2, 3, 5 are register butterflies; any other prime factor is handled by a
// generic O(p^2) shared-to-shared stage so every length works (the reference's tests use 3, 6, 10 ...).
// Twiddles come from an exact (double-precision generated) table exp(-2*pi*i*k/n) staged in shared memory.
#pragma once
#include "common.cuh"

namespace mrb {

constexpr int kMaxStages = 20;
constexpr int kMaxFftLen = 8192;

struct FftPlan {
    int n;
    int nstages;
    int ls;  // shared-memory line stride in float2 (n or n+1: keeps column-strip loads conflict free)
    int radix[kMaxStages];
    const float2* tw;  // device, [n]
};

// host: builds (and caches per device) the plan for length n. Returns MRB_* code.
int get_fft_plan(int n, FftPlan* plan);

__host__ __device__ inline size_t fft_smem_bytes(const FftPlan& p, int nlines) {
    return (size_t)(2 * (size_t)nlines * p.ls + p.n) * sizeof(float2);
}

// Stage twiddle table into shared memory (call once per CTA, then __syncthreads()).
__device__ __forceinline__ void load_twiddles(float2* tw_s, const FftPlan& p) {
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) tw_s[i] = p.tw[i];
}

template <bool INV>
__device__ __forceinline__ float2 twd(const float2* tw_s, int idx) {
    float2 w = tw_s[idx];
    if (INV) w.y = -w.y;
    return w;
}

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <int R, bool INV>
__device__ __forceinline__ void butterfly(float2* v) {
    if (R == 2) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if (R == 4) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    } else if (R == 3) {
        const float s = 0.86602540378443864676f;
        float2 t1 = cadd(v[1], v[2]);
        float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        float2 t3 = mul_mi<INV>(cscale(csub(v[1], v[2]), s));
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, t3);
        v[2] = csub(t2, t3);
    } else if (R == 5) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        float2 p1 = cadd(v[1], v[4]), m1 = csub(v[1], v[4]);
        float2 p2 = cadd(v[2], v[3]), m2 = csub(v[2], v[3]);
        float2 a0 = v[0];
        float2 A = make_float2(a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y);
        float2 B = make_float2(a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y);
        float2 U = mul_mi<INV>(make_float2(s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y));
        float2 V = mul_mi<INV>(make_float2(s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y));
        v[0] = make_float2(a0.x + p1.x + p2.x, a0.y + p1.y + p2.y);
        v[1] = cadd(A, U);
        v[4] = csub(A, U);
        v[2] = cadd(B, V);
        v[3] = csub(B, V);
    }
}

template <int R, bool INV>
__device__ __forceinline__ void fft_stage(const float2* __restrict__ src, float2* __restrict__ dst, int nlines,
                                          int n, int ls, int Ns, const float2* __restrict__ tw_s) {
    const int nb = n / R;
    const int total = nlines * nb;
    const int twmul = n / (Ns * R);
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int line = t / nb;
        const int j = t - line * nb;
        const int k = j % Ns;
        const float2* s = src + (size_t)line * ls;
        float2* d = dst + (size_t)line * ls;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = s[j + r * nb];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], twd<INV>(tw_s, r * k * twmul));
        }
        butterfly<R, INV>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) d[base + r * Ns] = v[r];
    }
}

// Any radix (used for prime factors other than 2, 3, 5): one thread per output element.
template <bool INV>
__device__ __forceinline__ void fft_stage_generic(const float2* __restrict__ src, float2* __restrict__ dst,
                                                  int nlines, int n, int ls, int Ns, int R,
                                                  const float2* __restrict__ tw_s) {
    const int nb = n / R;
    const int total = nlines * n;
    const int twmul = n / (Ns * R);
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int line = t / n;
        const int o = t - line * n;
        const int j = o / R;
        const int q = o - j * R;
        const int k = j % Ns;
        const float2* s = src + (size_t)line * ls;
        float2 acc = make_float2(0.f, 0.f);
        for (int r = 0; r < R; ++r) {
            long long e = (long long)r * k * twmul + (long long)((r * q) % R) * nb;
            float2 w = twd<INV>(tw_s, (int)(e % n));
            float2 x = s[j + r * nb];
            acc.x += x.x * w.x - x.y * w.y;
            acc.y += x.x * w.y + x.y * w.x;
        }
        dst[(size_t)line * ls + (j - k) * R + k + q * Ns] = acc;
    }
}

// Which buffer the caller must fill so that the result of block_fft lands in A.
__device__ __forceinline__ float2* fft_start_buf(const FftPlan& p, float2* A, float2* B) {
    return (p.nstages & 1) ? B : A;
}

// All threads of the CTA must call this. Input in fft_start_buf(), output in A.  Ends with __syncthreads().
template <bool INV>
__device__ __forceinline__ void block_fft(float2* A, float2* B, int nlines, const FftPlan& p,
                                          const float2* tw_s) {
    float2* src = fft_start_buf(p, A, B);
    float2* dst = (src == A) ? B : A;
    int Ns = 1;
    __syncthreads();
    for (int s = 0; s < p.nstages; ++s) {
        const int R = p.radix[s];
        switch (R) {
            case 4: fft_stage<4, INV>(src, dst, nlines, p.n, p.ls, Ns, tw_s); break;
            case 2: fft_stage<2, INV>(src, dst, nlines, p.n, p.ls, Ns, tw_s); break;
            case 3: fft_stage<3, INV>(src, dst, nlines, p.n, p.ls, Ns, tw_s); break;
            case 5: fft_stage<5, INV>(src, dst, nlines, p.n, p.ls, Ns, tw_s); break;
            default: fft_stage_generic<INV>(src, dst, nlines, p.n, p.ls, Ns, R, tw_s); break;
        }
        Ns *= R;
        float2* t = src;
        src = dst;
        dst = t;
        __syncthreads();
    }
}

}  // namespace mrb
