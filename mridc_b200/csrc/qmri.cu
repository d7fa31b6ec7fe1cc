// Quantitative (qRIM / qCIRIM) pointwise kernels around the fused data-consistency operator (sm_100a).
//
//   mrb_megre_signal     SignalForwardModel.MEGRESignalModel / MEGRENoPhaseSignalModel
//                        (mridc/collections/quantitative/models/qrim/utils.py:68-155)
//   mrb_megre_grad       the analytic d/d(R2*, S0) part of analytical_log_likelihood_gradient (qrim/utils.py:236-295)
//                        applied to the coil-combined residual that the fused DC operator (dc.cu) produced per echo
//   mrb_qrim_eta_update  eta + conv_out with the R2* channel clamped at 0 (qrim/qrim_block.py:232-236)
//   mrb_scale_batch      x (or |x|) * gamma[batch index]  (RescaleByMax.reverse, qrim/utils.py:25-28; qcirim.py:287-289)
//
// The reference evaluates every product as a separate fp32 torch op, so the kernels use explicit round-to-nearest
// multiplies / adds in the same association (no FMA contraction); exp / cos / sin are the accurate libdevice versions.
#include "common.cuh"

namespace mrb {

constexpr int QMAX_ECHOES = 16;

struct EchoTimes {
    float neg_te_s[QMAX_ECHOES];  // fl32(-TE * scaling): the Python-side double product the reference hands to torch
    float neg_te[QMAX_ECHOES];    // fl32(-TE)
};

__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }

// one echo's model terms: f = exp(-TE s R2*), c = cos(B0 s (-TE)), sn = sin(B0 s (-TE))  (qrim/utils.py:98-105)
__device__ __forceinline__ void megre_terms(float r2, float b0, float scaling, float neg_te_s, float neg_te, float& f,
                                            float& c, float& sn) {
    f = expf(mulr(neg_te_s, r2));
    const float ph = mulr(mulr(b0, scaling), neg_te);
    c = cosf(ph);
    sn = sinf(ph);
}

// maps [B,HW] x4 (each multiplied by gam[k] first: qrim_block.py:196-199; pass 1 for the plain model) -> pred [B,E,HW,2]
__global__ void megre_signal_kernel(const float* __restrict__ r2m, const float* __restrict__ s0m, const float* __restrict__ b0m,
                                    const float* __restrict__ phm, float4 gam, EchoTimes te, float scaling, int E,
                                    long long HW, long long total, int no_phase, float2* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / HW, p = i - b * HW;
    const float r2 = mulr(r2m[i], gam.x), s0 = mulr(s0m[i], gam.y);
    const float b0 = no_phase ? 0.f : mulr(b0m[i], gam.z), ph = no_phase ? 0.f : mulr(phm[i], gam.w);
    for (int e = 0; e < E; ++e) {
        float2 v;
        if (no_phase) {
            const float f = expf(mulr(te.neg_te_s[e], r2));
            v.x = mulr(s0, f);
            v.y = v.x;
        } else {
            float f, c, sn;
            megre_terms(r2, b0, scaling, te.neg_te_s[e], te.neg_te[e], f, c, sn);
            // S0r*f*c - S0i*f*sn ; S0r*f*sn + S0i*f*c, left-to-right products (qrim/utils.py:110-113)
            const float sf = mulr(s0, f), pf = mulr(ph, f);
            v.x = __fsub_rn(mulr(sf, c), mulr(pf, sn));
            v.y = __fadd_rn(mulr(sf, sn), mulr(pf, c));
        }
        if (v.x != v.x) v.x = 0.f;  // pred[pred != pred] = 0 (:121)
        if (v.y != v.y) v.y = 0.f;
        out[(b * E + e) * HW + p] = v;
    }
}

// d [B*E, 4, HW] (the fused DC operator's output: channels 2,3 = coil-combined residual image of echo e), maps as above
// -> out[b, ch, p] for ch = (R2*_re, S0_re, R2*_im, S0_im) = mean over echoes / divisor, NaN -> 0 when zero_nan;
// out has `out_ch` channels per sample (4, or 8 when it is the front half of the qRIM conv input).
__global__ void megre_grad_kernel(const float* __restrict__ d, const float* __restrict__ r2m, const float* __restrict__ s0m,
                                  const float* __restrict__ b0m, const float* __restrict__ phm, float4 gam, EchoTimes te,
                                  float scaling, int E, long long HW, long long total, float divisor, int zero_nan,
                                  float* __restrict__ out, int out_ch) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / HW, p = i - b * HW;
    const float r2 = mulr(r2m[i], gam.x), s0 = mulr(s0m[i], gam.y), b0 = mulr(b0m[i], gam.z), ph = mulr(phm[i], gam.w);
    float s0r = 0.f, s0i = 0.f, r2r = 0.f, r2i = 0.f;
    for (int e = 0; e < E; ++e) {
        const float* de = d + ((b * E + e) * 4 + 2) * HW + p;
        const float dr = de[0], di = de[HW];
        float f, c, sn;
        megre_terms(r2, b0, scaling, te.neg_te_s[e], te.neg_te[e], f, c, sn);
        const float a0 = mulr(f, c), a1 = mulr(-f, sn);  // S0_part_der (:249-251)
        // R2str_part_der (:253-264): (-TE*s) * f * (S0r c - S0i sn), (-TE*s) * f * (-S0r sn - S0i c)
        const float tf = mulr(te.neg_te_s[e], f);
        const float q0 = mulr(tf, __fsub_rn(mulr(s0, c), mulr(ph, sn)));
        const float q1 = mulr(tf, __fsub_rn(mulr(-s0, sn), mulr(ph, c)));
        // torch.mean over the echo axis = sequential fp32 sum / E (:286-289)
        s0r = __fadd_rn(s0r, __fsub_rn(mulr(dr, a0), mulr(di, a1)));
        s0i = __fadd_rn(s0i, __fadd_rn(mulr(dr, a1), mulr(di, a0)));
        r2r = __fadd_rn(r2r, __fsub_rn(mulr(dr, q0), mulr(di, q1)));
        r2i = __fadd_rn(r2i, __fadd_rn(mulr(dr, q1), mulr(di, q0)));
    }
    const float Ef = (float)E;
    float v[4] = {__fdiv_rn(r2r, Ef), __fdiv_rn(s0r, Ef), __fdiv_rn(r2i, Ef), __fdiv_rn(s0i, Ef)};
    float* o = out + b * out_ch * HW + p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = __fdiv_rn(v[k], divisor);
        if (zero_nan && x != x) x = 0.f;
        o[k * HW] = x;
    }
}

// eta_new = eta + delta; channel 0 (R2*): values < 0 -> 0 (NaN stays NaN, like eta_tmp[eta_tmp < 0] = 0).
// eta lives at channel offset `eta_off` of a buffer with `eta_ch` channels per sample (the qRIM conv input); it is updated
// in place and also written to out [B,4,HW].
__global__ void qrim_eta_update_kernel(float* __restrict__ eta, int eta_ch, int eta_off, const float* __restrict__ delta,
                                       float* __restrict__ out, long long HW, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bc = i / HW, p = i - bc * HW;
    const long long b = bc >> 2;
    const int ch = (int)(bc & 3);
    float* e = eta + (b * eta_ch + eta_off + ch) * HW + p;
    float v = __fadd_rn(*e, delta[i]);
    if (ch == 0 && v < 0.f) v = 0.f;
    *e = v;
    out[i] = v;
}

__global__ void scale_batch_kernel(const float* __restrict__ x, float* __restrict__ out, long long per_batch,
                                   long long total, float4 g, int take_abs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / per_batch;
    const float s = b == 0 ? g.x : b == 1 ? g.y : b == 2 ? g.z : g.w;
    const float v = x[i];
    out[i] = mulr(take_abs ? fabsf(v) : v, s);
}

static int fill_tes(const double* tes, int E, double scaling, EchoTimes& te, const char* who) {
    MRB_REQUIRE(tes && E >= 1 && E <= QMAX_ECHOES, MRB_EINVAL, "%s: 1 <= n_echoes <= %d required (got %d)", who, QMAX_ECHOES, E);
    for (int e = 0; e < QMAX_ECHOES; ++e) te.neg_te_s[e] = te.neg_te[e] = 0.f;
    for (int e = 0; e < E; ++e) {
        te.neg_te_s[e] = (float)(-tes[e] * scaling);
        te.neg_te[e] = (float)(-tes[e]);
    }
    return MRB_OK;
}

static inline unsigned qgrid(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace mrb

using namespace mrb;

extern "C" int mrb_megre_signal(const void* r2, const void* s0, const void* b0, const void* phi, const float* gamma4,
                                const double* tes, int E, double scaling, int B, long long HW, int no_phase, void* out,
                                void* stream) {
    MRB_REQUIRE(r2 && s0 && out && (no_phase || (b0 && phi)), MRB_EINVAL, "mrb_megre_signal: null pointer");
    MRB_REQUIRE(B >= 0 && HW >= 0, MRB_EINVAL, "mrb_megre_signal: negative extent");
    EchoTimes te;
    if (int rc = fill_tes(tes, E, scaling, te, "mrb_megre_signal")) return rc;
    const long long total = (long long)B * HW;
    if (total == 0) return MRB_OK;
    const float4 g = gamma4 ? make_float4(gamma4[0], gamma4[1], gamma4[2], gamma4[3]) : make_float4(1.f, 1.f, 1.f, 1.f);
    megre_signal_kernel<<<qgrid(total), 256, 0, (cudaStream_t)stream>>>((const float*)r2, (const float*)s0, (const float*)b0,
                                                                          (const float*)phi, g, te, (float)scaling, E, HW, total,
                                                                          no_phase, (float2*)out);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_megre_grad(const void* d, const void* r2, const void* s0, const void* b0, const void* phi,
                              const float* gamma4, const double* tes, int E, double scaling, int B, long long HW,
                              float divisor, int zero_nan, void* out, int out_channels, void* stream) {
    MRB_REQUIRE(d && r2 && s0 && b0 && phi && out, MRB_EINVAL, "mrb_megre_grad: null pointer");
    MRB_REQUIRE(B >= 0 && HW >= 0 && out_channels >= 4, MRB_EINVAL, "mrb_megre_grad: bad extent");
    MRB_REQUIRE(divisor != 0.f, MRB_EINVAL, "mrb_megre_grad: divisor must be non-zero");
    EchoTimes te;
    if (int rc = fill_tes(tes, E, scaling, te, "mrb_megre_grad")) return rc;
    const long long total = (long long)B * HW;
    if (total == 0) return MRB_OK;
    const float4 g = gamma4 ? make_float4(gamma4[0], gamma4[1], gamma4[2], gamma4[3]) : make_float4(1.f, 1.f, 1.f, 1.f);
    megre_grad_kernel<<<qgrid(total), 256, 0, (cudaStream_t)stream>>>((const float*)d, (const float*)r2, (const float*)s0,
                                                                        (const float*)b0, (const float*)phi, g, te, (float)scaling,
                                                                        E, HW, total, divisor, zero_nan, (float*)out,
                                                                        out_channels);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_qrim_eta_update(void* eta, int eta_channels, int eta_offset, const void* delta, void* out, int B,
                                   long long HW, void* stream) {
    MRB_REQUIRE(eta && delta && out, MRB_EINVAL, "mrb_qrim_eta_update: null pointer");
    MRB_REQUIRE(B >= 0 && HW >= 0 && eta_offset >= 0 && eta_offset + 4 <= eta_channels, MRB_EINVAL,
                "mrb_qrim_eta_update: bad extent");
    const long long total = (long long)B * 4 * HW;
    if (total == 0) return MRB_OK;
    qrim_eta_update_kernel<<<qgrid(total), 256, 0, (cudaStream_t)stream>>>((float*)eta, eta_channels, eta_offset,
                                                                             (const float*)delta, (float*)out, HW, total);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_scale_batch(const void* x, void* out, int B, long long per_batch, const float* scales, int take_abs,
                               void* stream) {
    MRB_REQUIRE(x && out && scales, MRB_EINVAL, "mrb_scale_batch: null pointer");
    MRB_REQUIRE(B >= 0 && B <= 4 && per_batch >= 0, MRB_EINVAL,
                "mrb_scale_batch: batch %d exceeds the 4 regularisation factors the reference indexes by batch", B);
    const long long total = (long long)B * per_batch;
    if (total == 0) return MRB_OK;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < B; ++b) s[b] = scales[b];
    scale_batch_kernel<<<qgrid(total), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)out, per_batch, total,
                                                                         make_float4(s[0], s[1], s[2], s[3]), take_abs);
    MRB_LAUNCHED();
    return MRB_OK;
}
