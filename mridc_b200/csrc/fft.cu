// Standalone FFT entry points: mrb_fft1d_c2c, mrb_fft2_c2c, mrb_roll.
// Reference behaviour: mridc/collections/common/parts/fft.py:13-88 (fft2), :91-166 (ifft2), :169-240 (roll).
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include <stdlib.h>

#include "fft.cuh"

namespace mrb {

// ---- plan cache ------------------------------------------------------------------------------------
static std::mutex g_plan_mu;
static std::map<std::pair<int, int>, FftPlan> g_plans;  // (device, n) -> plan

int get_fft_plan(int n, FftPlan* out) {
    MRB_REQUIRE(n >= 1, MRB_EINVAL, "fft length must be >= 1 (got %d)", n);
    MRB_REQUIRE(n <= kMaxFftLen, MRB_EUNSUPPORTED, "fft length %d exceeds the shared-memory engine limit %d", n,
                kMaxFftLen);
    int dev = 0;
    MRB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find({dev, n});
    if (it != g_plans.end()) {
        *out = it->second;
        return MRB_OK;
    }
    FftPlan p;
    memset(&p, 0, sizeof(p));
    p.n = n;
    p.ls = (n % 2 == 0) ? n + 1 : n;
    // factorise: odd primes first (largest generic ones first), then 5, 3, then 4s and a final 2.
    int m = n;
    std::vector<int> f4, f2, f3, f5, fg;
    while (m % 4 == 0) { f4.push_back(4); m /= 4; }
    while (m % 2 == 0) { f2.push_back(2); m /= 2; }
    while (m % 3 == 0) { f3.push_back(3); m /= 3; }
    while (m % 5 == 0) { f5.push_back(5); m /= 5; }
    for (int q = 7; (long long)q * q <= m; q += 2)
        while (m % q == 0) { fg.push_back(q); m /= q; }
    if (m > 1) fg.push_back(m);
    std::vector<int> all;
    for (int v : fg) all.push_back(v);
    for (int v : f5) all.push_back(v);
    for (int v : f3) all.push_back(v);
    for (int v : f2) all.push_back(v);
    for (int v : f4) all.push_back(v);
    if (n == 320) all = {5, 8, 8};   // fixed-plan fast paths of fft.cuh (3 stages: input buffer parity must match)
    if (n == 640) all = {10, 8, 8};
    MRB_REQUIRE((int)all.size() <= kMaxStages, MRB_EUNSUPPORTED, "too many FFT stages for n=%d", n);
    p.nstages = (int)all.size();
    for (int i = 0; i < p.nstages; ++i) p.radix[i] = all[i];
    std::vector<float2> tw(n);
    for (int k = 0; k < n; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2* d = nullptr;
    MRB_CUDA(cudaMalloc(&d, sizeof(float2) * n));
    MRB_CUDA(cudaMemcpy(d, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    p.tw = d;
    g_plans[{dev, n}] = p;
    *out = p;
    return MRB_OK;
}

// ---- kernels ---------------------------------------------------------------------------------------
// Contiguous lines (inner == 1): CTA handles LPB consecutive lines of length n.
template <bool INV>
__global__ void fft1d_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long nlines_total,
                                  int lpb, FftPlan p, int in_rot, int out_rot, float scale) {
    extern __shared__ float2 smem[];
    float2* A = smem;
    float2* B = A + (size_t)lpb * p.ls;
    float2* tw_s = B + (size_t)lpb * p.ls;
    load_twiddles(tw_s, p);
    const long long line0 = (long long)blockIdx.x * lpb;
    const int nl = (int)min((long long)lpb, nlines_total - line0);
    float2* S = fft_start_buf(p, A, B);
    const int n = p.n;
    for (int t = threadIdx.x; t < nl * n; t += blockDim.x) {
        int l = t / n, j = t - l * n;
        int src = j + in_rot;
        if (src >= n) src -= n;
        S[(size_t)l * p.ls + j] = in[(line0 + l) * n + src];
    }
    block_fft<INV>(A, B, nl, p, tw_s);
    for (int t = threadIdx.x; t < nl * n; t += blockDim.x) {
        int l = t / n, d = t - l * n;  // d = storage position, k = logical output index
        int k = d - out_rot;
        if (k < 0) k += n;
        float2 v = A[(size_t)l * p.ls + k];
        out[(line0 + l) * n + d] = cscale(v, scale);
    }
}

// Strided lines (inner > 1): CTA handles one `outer` index and a strip of TI inner positions.
template <bool INV>
__global__ void fft1d_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long inner, int ti,
                                  FftPlan p, int in_rot, int out_rot, float scale) {
    extern __shared__ float2 smem[];
    float2* A = smem;
    float2* B = A + (size_t)ti * p.ls;
    float2* tw_s = B + (size_t)ti * p.ls;
    load_twiddles(tw_s, p);
    const int n = p.n;
    const long long i0 = (long long)blockIdx.x * ti;
    const int nl = (int)min((long long)ti, inner - i0);
    const float2* gin = in + (long long)blockIdx.y * n * inner + i0;
    float2* gout = out + (long long)blockIdx.y * n * inner + i0;
    float2* S = fft_start_buf(p, A, B);
    for (int t = threadIdx.x; t < nl * n; t += blockDim.x) {
        int j = t / nl, i = t - j * nl;
        int src = j + in_rot;
        if (src >= n) src -= n;
        S[(size_t)i * p.ls + j] = gin[(long long)src * inner + i];
    }
    block_fft<INV>(A, B, nl, p, tw_s);
    for (int t = threadIdx.x; t < nl * n; t += blockDim.x) {
        int d = t / nl, i = t - d * nl;
        int k = d - out_rot;
        if (k < 0) k += n;
        gout[(long long)d * inner + i] = cscale(A[(size_t)i * p.ls + k], scale);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// 320-point fast path (both fastMRI geometries): 16 lines per CTA of 320 threads, register-resident 16 x 20 transform
// with one shared-memory transpose (the building blocks of the DC gradient kernel, fft.cuh r320::).
//   ROWS (inner == 1): thread (line l = tid / 20, t = tid % 20) -- a line's 20 threads read 160 contiguous bytes per step.
//   COLS (inner > 1):  thread (t = tid / 16, column l = tid % 16) -- 16 adjacent columns = 128 contiguous bytes per row.
// Centring needs no index rotation: x'[j] = x[j + N/2] multiplies the output by (-1)^k and out[d] = X[d - N/2] equals
// a (-1)^j modulation of the input, so the rotations of fft1d_launch (0 or N/2 each) become two thread-constant signs
// (j = 20*n1 + t has the parity of t, k = k1 + 16*k2 the parity of k1).
// In-place use is fine: a CTA reads its 16 lines into registers before the first barrier and writes only those lines.
// ---------------------------------------------------------------------------------------------------------------
namespace r320 {
constexpr int FN = 320, FN1 = 16, FN2 = 20, FXS = 22, FLINES = 16, FTHREADS = 320;

template <bool INV, bool COLS>
__global__ void __launch_bounds__(FTHREADS, 3) fft320_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                             long long nlines, long long inner, const float2* __restrict__ tw,
                                                             int in_sign, int out_sign, float scale) {
    // line stride of the exchange buffer: ROWS wants 16-byte aligned rows read with 128-bit loads by lanes that walk k1
    // (356 = 16*22 + 4), COLS has lanes walking the 16 lines with 64-bit accesses (odd stride: 2 banks per lane)
    constexpr int CS = COLS ? FN1 * FXS + 1 : FN1 * FXS + 4;
    extern __shared__ float2 smem[];
    float2* xch = smem;               // [FLINES][CS]
    float2* tw1_s = xch + FLINES * CS;  // [k1][t]  w320^(t*k1)
    const int tid = threadIdx.x;
    const int l = COLS ? (tid & 15) : tid / FN2, t = COLS ? (tid >> 4) : tid - (tid / FN2) * FN2;
    const long long lid = (long long)blockIdx.x * FLINES + l;  // line (ROWS) or column (COLS) index
    const bool valid = COLS ? lid < inner : lid < nlines;
    const long long base = COLS ? (long long)blockIdx.y * FN * inner + lid : lid * FN;
    const long long stride = COLS ? inner : 1;
    cx v[FN1];
    if (valid) {
        const float2* g = in + base + (long long)t * stride;
        float2 x[FN1];
#pragma unroll
        for (int n1 = 0; n1 < FN1; ++n1) x[n1] = __ldg(g + (long long)(FN2 * n1) * stride);
        const float sg = (in_sign && (t & 1)) ? -1.f : 1.f;
#pragma unroll
        for (int n1 = 0; n1 < FN1; ++n1) v[n1] = pk(x[n1].x * sg, x[n1].y * sg);
    }
    {
        const int a = tid / FN2, b = tid - a * FN2;
        tw1_s[tid] = __ldg(&tw[a * b]);  // tid = 20*k1 + t
    }
    __syncthreads();
    float2* xl = xch + (size_t)l * CS;
    if (valid) {
        dft16<INV>(v);
        xl[t] = upk(v[0]);
#pragma unroll
        for (int k1 = 1; k1 < FN1; ++k1) {
            const float2 w = tw1_s[k1 * FN2 + t];
            xl[k1 * FXS + t] = mulw<INV>(upk(v[k1]), w.x, w.y);
        }
    }
    __syncthreads();
    // pass 2: 16 lines x 16 k1 = 256 threads, 20-point DFT over n2, store X[k1 + 16*k2]
    if (tid < FLINES * FN1) {
        const int l2 = COLS ? (tid & 15) : (tid >> 4), k1 = COLS ? (tid >> 4) : (tid & 15);
        const long long lid2 = (long long)blockIdx.x * FLINES + l2;
        const bool valid2 = COLS ? lid2 < inner : lid2 < nlines;
        if (valid2) {
            cx u[FN2];
            const float2* row = xch + (size_t)l2 * CS + k1 * FXS;
            if (COLS) {
#pragma unroll
                for (int i = 0; i < FN2; ++i) u[i] = pk(row[i]);
            } else {
                const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll
                for (int i = 0; i < FN2 / 2; ++i) {
                    const float4 q = row4[i];
                    u[2 * i] = pk(q.x, q.y);
                    u[2 * i + 1] = pk(q.z, q.w);
                }
            }
            dft20<INV>(u);
            const float sc = (out_sign && (k1 & 1)) ? -scale : scale;
            const long long base2 = COLS ? (long long)blockIdx.y * FN * inner + lid2 : lid2 * FN;
            float2* o = out + base2 + (long long)k1 * stride;
#pragma unroll
            for (int k2 = 0; k2 < FN2; ++k2) {
                const float2 r = upk(u[k2]);
                o[(long long)(FN1 * k2) * stride] = make_float2(r.x * sc, r.y * sc);
            }
        }
    }
}

template <bool INV, bool COLS>
static int launch_fft320(const float2* in, float2* out, long long outer, long long inner, const float2* tw, int in_sign,
                         int out_sign, float scale, cudaStream_t st) {
    constexpr int CS = COLS ? FN1 * FXS + 1 : FN1 * FXS + 4;
    const size_t smem = (size_t)(FLINES * CS + FN) * sizeof(float2);
    auto k = fft320_kernel<INV, COLS>;
    MRB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (COLS) {
        const long long gx = (inner + FLINES - 1) / FLINES;
        k<<<dim3((unsigned)gx, (unsigned)outer), FTHREADS, smem, st>>>(in, out, 0, inner, tw, in_sign, out_sign, scale);
    } else {
        const long long gx = (outer + FLINES - 1) / FLINES;
        k<<<(unsigned)gx, FTHREADS, smem, st>>>(in, out, outer, 1, tw, in_sign, out_sign, scale);
    }
    MRB_LAUNCHED();
    return MRB_OK;
}

// 640-point transform along a strided axis (the H = 640 columns of the brain geometry): 8 adjacent columns per CTA,
// fft640_cols_cta (two 320-point transforms + one radix-2 combine).  grid (inner / 8 rounded up, outer)
template <bool INV>
__global__ void __launch_bounds__(FTHREADS, 2) fft640_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                  long long inner, const float2* __restrict__ tw320,
                                                                  const float2* __restrict__ tw640, int in_mod, int out_sign,
                                                                  float scale) {
    extern __shared__ float2 smem[];
    float2* xch = smem;
    float2* tw1_s = xch + CTA_LINES * (CTA_N1 * CTA_XS + 1);
    float2* buf = tw1_s + CTA_N;
    const long long c0 = (long long)blockIdx.x * 8;
    const long long base = (long long)blockIdx.y * (2 * CTA_N) * inner + c0;
    const int ncols = (int)(inner - c0 < 8 ? inner - c0 : 8);
    auto ld = [&](int cl, int n) { return __ldg(in + base + (long long)n * inner + cl); };
    auto st = [&](int cl, int k, float2 v) { out[base + (long long)k * inner + cl] = v; };
    fft640_cols_cta<INV>(xch, tw1_s, buf, tw320, tw640, ncols, in_mod, out_sign, scale, ld, st);
}
inline size_t fft640_smem() { return (size_t)(CTA_LINES * (CTA_N1 * CTA_XS + 1) + CTA_N + CTA_LINES * (CTA_N + 1)) * sizeof(float2); }

template <bool INV>
static int launch_fft640_cols(const float2* in, float2* out, long long outer, long long inner, const float2* tw320,
                              const float2* tw640, int in_mod, int out_sign, float scale, cudaStream_t st) {
    auto k = fft640_cols_kernel<INV>;
    MRB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft640_smem()));
    k<<<dim3((unsigned)((inner + 7) / 8), (unsigned)outer), FTHREADS, fft640_smem(), st>>>(in, out, inner, tw320, tw640, in_mod,
                                                                                          out_sign, scale);
    MRB_LAUNCHED();
    return MRB_OK;
}
}  // namespace r320

static int choose_lines(const FftPlan& p, int want_min_threads_work, long long avail_lines, size_t max_smem) {
    // lines per CTA: enough butterflies for 256 threads, bounded by shared memory and by what exists.
    int lines = 1;
    while (lines < 16 && (long long)lines * (p.n / 4 + 1) < want_min_threads_work) lines *= 2;
    while (lines > 1 && fft_smem_bytes(p, lines) > max_smem) lines /= 2;
    if ((long long)lines > avail_lines) lines = (int)avail_lines;
    return lines < 1 ? 1 : lines;
}

int fft1d_launch(const float2* in, float2* out, long long outer, int n, long long inner, int inverse, int in_rot,
                 int out_rot, float scale, cudaStream_t st) {
    if (outer == 0 || inner == 0 || n == 0) return MRB_OK;
    if (inner > 1 && outer > 65535) {
        // the strided-axis kernels put `outer` in gridDim.y: more leading elements than that go in chunks
        for (long long o = 0; o < outer; o += 65535) {
            const long long cnt = outer - o < 65535 ? outer - o : 65535;
            const int rcc = fft1d_launch(in + o * n * inner, out + o * n * inner, cnt, n, inner, inverse, in_rot, out_rot, scale, st);
            if (rcc) return rcc;
        }
        return MRB_OK;
    }
    FftPlan p;
    int rc = get_fft_plan(n, &p);
    if (rc) return rc;
    MRB_REQUIRE(in_rot >= 0 && in_rot < n && out_rot >= 0 && out_rot < n, MRB_EINVAL, "rotation out of range");
    const size_t max_smem = device_max_smem_optin();
    MRB_REQUIRE(fft_smem_bytes(p, 1) <= max_smem, MRB_EUNSUPPORTED, "fft length %d does not fit shared memory", n);
    if (n == r320::FN && (in_rot == 0 || in_rot == n / 2) && (out_rot == 0 || out_rot == n / 2) &&
        !getenv("MRIDC_B200_FFT_STOCKHAM")) {
        const long long gx = inner == 1 ? (outer + 15) / 16 : (inner + 15) / 16;
        if (gx <= 2147483647LL && (inner == 1 || outer <= 65535)) {
            // x'[j] = x[j + N/2] -> (-1)^k on the output; out[d] = X[d - N/2] -> (-1)^j on the input
            const int in_sign = out_rot != 0, out_sign = in_rot != 0;
            if (inner == 1)
                return inverse ? r320::launch_fft320<true, false>(in, out, outer, 1, p.tw, in_sign, out_sign, scale, st)
                               : r320::launch_fft320<false, false>(in, out, outer, 1, p.tw, in_sign, out_sign, scale, st);
            return inverse ? r320::launch_fft320<true, true>(in, out, outer, inner, p.tw, in_sign, out_sign, scale, st)
                           : r320::launch_fft320<false, true>(in, out, outer, inner, p.tw, in_sign, out_sign, scale, st);
        }
    }
    if (n == 2 * r320::FN && inner > 1 && (in_rot == 0 || in_rot == n / 2) && (out_rot == 0 || out_rot == n / 2) &&
        outer <= 65535 && (inner + 7) / 8 <= 2147483647LL && !getenv("MRIDC_B200_FFT_STOCKHAM")) {
        FftPlan p320;
        int rc = get_fft_plan(r320::FN, &p320);
        if (rc) return rc;
        return inverse ? r320::launch_fft640_cols<true>(in, out, outer, inner, p320.tw, p.tw, out_rot != 0, in_rot != 0, scale, st)
                       : r320::launch_fft640_cols<false>(in, out, outer, inner, p320.tw, p.tw, out_rot != 0, in_rot != 0, scale, st);
    }
    const int threads = 256;
    if (inner == 1) {
        int lpb = choose_lines(p, 512, outer, max_smem < 98304 ? max_smem : 98304);
        size_t smem = fft_smem_bytes(p, lpb);
        long long grid = (outer + lpb - 1) / lpb;
        MRB_REQUIRE(grid <= 2147483647LL, MRB_EUNSUPPORTED, "too many FFT lines");
        auto k = inverse ? fft1d_rows_kernel<true> : fft1d_rows_kernel<false>;
        MRB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        k<<<(unsigned)grid, threads, smem, st>>>(in, out, outer, lpb, p, in_rot, out_rot, scale);
        MRB_LAUNCHED();
    } else {
        int ti = 16;
        while (ti > 1 && fft_smem_bytes(p, ti) > (max_smem < 131072 ? max_smem : 131072)) ti /= 2;
        if ((long long)ti > inner) ti = (int)inner;
        size_t smem = fft_smem_bytes(p, ti);
        long long gx = (inner + ti - 1) / ti;
        MRB_REQUIRE(gx <= 2147483647LL && outer <= 65535, MRB_EUNSUPPORTED,
                    "fft1d: outer=%lld exceeds grid.y limit or inner too large", outer);
        auto k = inverse ? fft1d_cols_kernel<true> : fft1d_cols_kernel<false>;
        MRB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        k<<<dim3((unsigned)gx, (unsigned)outer), threads, smem, st>>>(in, out, inner, ti, p, in_rot, out_rot, scale);
        MRB_LAUNCHED();
    }
    return MRB_OK;
}

float norm_scale(int norm, int inverse, double npts) {
    // torch.fft norm semantics (fft.py:77-81 / :155-159)
    if (norm == MRB_NORM_ORTHO) return (float)(1.0 / sqrt(npts));
    if (norm == MRB_NORM_BACKWARD) return inverse ? (float)(1.0 / npts) : 1.0f;
    return inverse ? 1.0f : (float)(1.0 / npts);  // forward
}

// ---- roll ------------------------------------------------------------------------------------------
template <typename T>
__global__ void roll_kernel(const T* __restrict__ in, T* __restrict__ out, long long total, long long n,
                            long long inner, long long shift) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long i = t % inner;
        long long r = t / inner;
        long long j = r % n;
        long long o = r / n;
        long long src = j - shift;
        if (src < 0) src += n;
        out[t] = in[(o * n + src) * inner + i];
    }
}

}  // namespace mrb

using namespace mrb;

extern "C" int mrb_fft1d_c2c(const void* in, void* out, long long outer, int n, long long inner, int inverse,
                             int in_rot, int out_rot, float scale, void* stream) {
    MRB_REQUIRE(in && out, MRB_EINVAL, "mrb_fft1d_c2c: null pointer");
    MRB_REQUIRE(outer >= 0 && inner >= 0 && n >= 0, MRB_EINVAL, "mrb_fft1d_c2c: negative extent");
    return fft1d_launch((const float2*)in, (float2*)out, outer, n, inner, inverse, in_rot, out_rot, scale,
                        (cudaStream_t)stream);
}

extern "C" int mrb_fft2_c2c(const void* in, void* out, long long batch, int H, int W, int inverse, int centered,
                            int norm, void* stream) {
    MRB_REQUIRE(in && out, MRB_EINVAL, "mrb_fft2_c2c: null pointer");
    MRB_REQUIRE(batch >= 0 && H >= 1 && W >= 1, MRB_EINVAL, "mrb_fft2_c2c: bad shape");
    MRB_REQUIRE(norm >= 0 && norm <= 2, MRB_EINVAL, "mrb_fft2_c2c: bad norm %d", norm);
    cudaStream_t st = (cudaStream_t)stream;
    const int rw = centered ? W / 2 : 0, rh = centered ? H / 2 : 0;
    const float s = norm_scale(norm, inverse, (double)H * (double)W);
    // rows (along W), then columns (along H); the scale rides on the second pass.
    int rc = fft1d_launch((const float2*)in, (float2*)out, batch * H, W, 1, inverse, rw, rw, 1.0f, st);
    if (rc) return rc;
    return fft1d_launch((const float2*)out, (float2*)out, batch, H, W, inverse, rh, rh, s, st);
}

extern "C" int mrb_roll(const void* in, void* out, long long outer, long long n, long long inner, int elem_bytes,
                        long long shift, void* stream) {
    MRB_REQUIRE(in && out && in != out, MRB_EINVAL, "mrb_roll: null or aliased pointers");
    MRB_REQUIRE(outer >= 0 && n >= 0 && inner >= 0, MRB_EINVAL, "mrb_roll: negative extent");
    long long total = outer * n * inner;
    if (total == 0) return MRB_OK;
    shift %= n;
    if (shift < 0) shift += n;
    cudaStream_t st = (cudaStream_t)stream;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    switch (elem_bytes) {
        case 1: roll_kernel<uint8_t><<<(unsigned)blocks, threads, 0, st>>>((const uint8_t*)in, (uint8_t*)out, total, n, inner, shift); break;
        case 2: roll_kernel<uint16_t><<<(unsigned)blocks, threads, 0, st>>>((const uint16_t*)in, (uint16_t*)out, total, n, inner, shift); break;
        case 4: roll_kernel<uint32_t><<<(unsigned)blocks, threads, 0, st>>>((const uint32_t*)in, (uint32_t*)out, total, n, inner, shift); break;
        case 8: roll_kernel<uint64_t><<<(unsigned)blocks, threads, 0, st>>>((const uint64_t*)in, (uint64_t*)out, total, n, inner, shift); break;
        case 16: roll_kernel<uint4><<<(unsigned)blocks, threads, 0, st>>>((const uint4*)in, (uint4*)out, total, n, inner, shift); break;
        default: MRB_REQUIRE(false, MRB_EINVAL, "mrb_roll: unsupported element size %d", elem_bytes);
    }
    MRB_LAUNCHED();
    return MRB_OK;
}
