// Tensor-core (tcgen05 / TMEM) implicit-GEMM kernels for the RIM regulariser, NHWC fp32 activations.
//
// Reference behaviour: ConvNonlinear (rim/conv_layers.py:36-123, replicate padding) and ConvGRUCell with
// kernel_size 1 (rim/rnn_cells.py:93-127), as used by RIMBlock's time loop (rim/rim_block.py:217-249).
//
// Numerics: the reference is fp32 and the end-to-end tolerance (rel-L2 1e-4) rules out plain TF32 (SURVEY
// section 7: 9.3e-4).  Every product is therefore evaluated as an error-compensated 3xTF32 sum
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,    x_hi = rn_tf32(x), x_lo = rn_tf32(x - x_hi)
// with fp32 accumulation in TMEM (split error and dropped term a_lo*b_lo are both ~2^-24 relative).
//
// Structure: one persistent CTA per SM.  Work item = (128-pixel tile, half of the output channels).  The
// weights of the CTA's channel half (hi and lo, pre-packed in the UMMA K-major SWIZZLE_128B layout) stay
// resident in shared memory for the whole kernel; activations stream through a ring of stages:
//   loader warps (2 groups x 4, alternating segments) gather a [128 px x 32 ch] fp32 chunk per "segment"
//                     (source tensor, tap offset with replicate clamp, channel chunk), split it into hi/lo and
//                     store both in the swizzled layout (no TMA descriptor: the gather does the clamp that TMA's
//                     zero fill cannot); the global loads of a group's next segment are in flight (registers)
//                     while it converts and stores the current one;
//   MMA warp (1 lane) issues 4 k-steps x 3 tcgen05.mma.kind::tf32 per segment into the TMEM accumulator and
//                     tcgen05.commit's the stage back to the loaders / the accumulator to the epilogue;
//   epilogue warps (8) tcgen05.ld the accumulator (one pixel per thread), apply bias + ReLU or the GRU gates,
//                     and write NHWC.  Two TMEM accumulator buffers overlap the epilogue with the next tile.
#include "common.cuh"

namespace mrb {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KC = 32;                        // channels per K chunk = one 128-byte swizzle row
constexpr int CHUNK_BYTES = TILE_M * 128;     // A chunk (hi or lo): 16 KB
constexpr int STAGE_BYTES = 2 * CHUNK_BYTES;  // hi + lo
constexpr int EPI_WARPS = 8;                  // 2 warps per TMEM lane quadrant (each takes half of the channels)
constexpr int LOAD_GROUPS = 2, LOAD_WARPS = 4 * LOAD_GROUPS;  // loader groups alternate segments
constexpr int THREADS = (EPI_WARPS + LOAD_WARPS + 1) * 32;
constexpr int MAX_SEGS = 20;

enum Mode { MODE_CONV_RELU = 0, MODE_GRU = 1, MODE_CONV_NOACT = 2 };

struct Segment {
    short src;     // 0 / 1: which input tensor
    short dy, dx;  // pixel offset of this tap (replicate clamp)
    short c0;      // first channel of the chunk in the source (or im2col chunk index when im2col != 0)
    short wchunk;  // resident weight chunk index
    short dcol;    // accumulator column offset
    short n;       // MMA N for this segment
    short first;   // 1: overwrite the accumulator columns (first contribution)
};

struct Params {
    const float* src[2];     // NHWC inputs [B, H, W, cs]
    int cs[2];               // channels per pixel of each source
    const float* wpack;      // packed weights: [half][nseg_w][hi|lo][n rows x 128 B swizzled]
    const float* bias;       // conv: [Cout]; GRU: b_ih [3*Ch] (may be null)
    const float* hprev;      // GRU: previous hidden state NHWC [P, Ch]
    float* out;              // NHWC [P, Cout]
    int B, H, W;
    long long P;             // B*H*W pixels
    int n_tiles;
    int nseg;
    int wchunk_rows;         // rows (N) of one resident weight chunk
    int n_wchunks;
    int acc_cols;            // TMEM columns of one accumulator buffer
    int tmem_cols;           // allocation (power of two >= 2*acc_cols)
    int cout;                // output channels per pixel (both halves)
    int nhalf;               // output channels handled per item (cout / 2)
    int mode;
    int stages;
    int im2col;              // 1: source 0 is [B,H,W,4] and chunk c0 gathers taps 8*c0 .. 8*c0+7 of a 5x5 window
    int debug;               // profiling switches (mrb_tc_set_debug): 1 skip MMAs, 2 skip global loads, 4 skip epilogue math
    int small_off;           // != 0: the two cross terms (lo*hi, hi*lo) accumulate in columns dcol + small_off, so the
                             // long hi*hi chain sees 3x fewer (truncating) tensor-core accumulations; summed in the epilogue
    Segment seg[MAX_SEGS];
};

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, M = 128, K = 8
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMEM -> registers: 8 consecutive columns of this thread's lane.  The wait is part of the same asm statement so
// that no consumer of v[] can be scheduled before the asynchronous load has landed.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
    d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major A and B, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa, low 13 bits zero).  Used for both split terms so that the
// tensor core's own operand truncation is a no-op: a = hi + lo + O(2^-24 |a|).
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(THREADS, 1) tc_kernel(const Params P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // layout: [weights resident][stages][barriers]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int wbytes_chunk = P.wchunk_rows * 128;
    uint8_t* w_s = smem;                                              // [n_wchunks][hi|lo][rows*128]
    uint8_t* st_s = w_s + (size_t)P.n_wchunks * 2 * wbytes_chunk;     // stages (1024-aligned: chunks are multiples of 1 KB)
    uint64_t* bars = (uint64_t*)(st_s + (size_t)P.stages * STAGE_BYTES);
    uint64_t* full = bars;                  // [stages]
    uint64_t* empty = bars + P.stages;      // [stages]
    uint64_t* acc_full = bars + 2 * P.stages;   // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = blockIdx.x & 1;
    const int first_tile = blockIdx.x >> 1;
    const int tile_stride = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&full[s], 128);  // one loader group fills a stage
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == EPI_WARPS + LOAD_WARPS) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    // resident weights of this half: straight copy (already swizzled by the packer)
    {
        const float4* g = reinterpret_cast<const float4*>(P.wpack) + (size_t)half * P.n_wchunks * 2 * wbytes_chunk / 16;
        float4* s = reinterpret_cast<float4*>(w_s);
        const int n16 = P.n_wchunks * 2 * wbytes_chunk / 16;
        for (int i = threadIdx.x; i < n16; i += THREADS) s[i] = g[i];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= EPI_WARPS && warp < EPI_WARPS + LOAD_WARPS) {
        // ============================== LOADERS ==============================
        const int lw = warp - EPI_WARPS;
        const int grp = lw >> 2;                     // loader group: handles segments grp, grp + 2, ...
        const int lt = (lw & 3) * 32 + lane;         // 0..127 within the group
        const int c16 = lt & 7;                      // 16-byte chunk within the 128-byte row
        const int r0 = lt >> 3;                      // rows r0 + 16*i
        int pb[8], py[8], px[8];
        bool pv[8];
        int coords_tile = -1;

        auto issue_loads = [&](int tile, int sgi, float4(&v)[8]) {
            if (tile != coords_tile) {
                coords_tile = tile;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    long long p = (long long)tile * TILE_M + r0 + 16 * i;
                    pv[i] = p < P.P;
                    long long q = pv[i] ? p : 0;
                    px[i] = (int)(q % P.W);
                    long long t = q / P.W;
                    py[i] = (int)(t % P.H);
                    pb[i] = (int)(t / P.H);
                }
            }
            const Segment sg = P.seg[sgi];
            const float* src = P.src[sg.src];
            const int cs = P.cs[sg.src];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pv[i] && !(P.debug & 2)) {
                    if (P.im2col) {
                        // chunk c16 of im2col row = tap t of the 5x5 window (4 channels = one float4)
                        const int t = sg.c0 * 8 + c16;
                        if (t < 25) {
                            int yy = min(max(py[i] + t / 5 - 2, 0), P.H - 1);
                            int xx = min(max(px[i] + t % 5 - 2, 0), P.W - 1);
                            v[i] = __ldg(reinterpret_cast<const float4*>(src + (((long long)pb[i] * P.H + yy) * P.W + xx) * 4));
                        }
                    } else {
                        int yy = min(max(py[i] + sg.dy, 0), P.H - 1);
                        int xx = min(max(px[i] + sg.dx, 0), P.W - 1);
                        v[i] = __ldg(reinterpret_cast<const float4*>(src + (((long long)pb[i] * P.H + yy) * P.W + xx) * cs +
                                                                     sg.c0 + c16 * 4));
                    }
                }
            }
        };

        int it = 0, tile = first_tile, sgi = grp;
        bool have = tile < P.n_tiles && sgi < P.nseg;
        float4 vc[8], vn[8];
        if (have) issue_loads(tile, sgi, vc);
        while (have) {
            int nsgi = sgi + LOAD_GROUPS, nit = it, ntile = tile;
            if (nsgi >= P.nseg) { nsgi = grp; nit = it + 1; ntile = tile + tile_stride; }
            const bool nhave = ntile < P.n_tiles;
            if (nhave) issue_loads(ntile, nsgi, vn);  // in flight while the current chunk is converted and stored
            const int sglob = it * P.nseg + sgi;
            const int stage = sglob % P.stages;
            const uint32_t phase = (uint32_t)(sglob / P.stages) & 1u;
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a_hi = st_s + (size_t)stage * STAGE_BYTES;
            uint8_t* a_lo = a_hi + CHUNK_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 16 * i;
                float4 hi = make_float4(tf32_rn(vc[i].x), tf32_rn(vc[i].y), tf32_rn(vc[i].z), tf32_rn(vc[i].w));
                float4 lo = make_float4(tf32_rn(vc[i].x - hi.x), tf32_rn(vc[i].y - hi.y), tf32_rn(vc[i].z - hi.z),
                                        tf32_rn(vc[i].w - hi.w));
                const uint32_t o = swz(r, c16);
                if (!(P.debug & 8)) {
                    *reinterpret_cast<float4*>(a_hi + o) = hi;
                    *reinterpret_cast<float4*>(a_lo + o) = lo;
                }
            }
            if (!(P.debug & 16)) fence_proxy_async();
            mbar_arrive(&full[stage]);
#pragma unroll
            for (int i = 0; i < 8; ++i) vc[i] = vn[i];
            it = nit; tile = ntile; sgi = nsgi; have = nhave;
        }
    } else if (warp == EPI_WARPS + LOAD_WARPS) {
        // ============================== MMA ISSUER ==============================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = first_tile; tile < P.n_tiles; tile += tile_stride, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&acc_empty[buf], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + (uint32_t)(buf * P.acc_cols);
                for (int sgi = 0; sgi < P.nseg; ++sgi) {
                    const Segment sg = P.seg[sgi];
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(st_s + (size_t)stage * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + CHUNK_BYTES;
                    const uint32_t b_hi = smem_u32(w_s + (size_t)sg.wchunk * 2 * wbytes_chunk);
                    const uint32_t b_lo = b_hi + wbytes_chunk;
                    const uint64_t dah = make_desc(a_hi), dal = make_desc(a_lo);
                    const uint64_t dbh = make_desc(b_hi), dbl = make_desc(b_lo);
                    const uint32_t idesc = make_idesc(TILE_M, sg.n);
                    const uint32_t d = d_base + (uint32_t)sg.dcol;
                    const uint32_t ds = d + (uint32_t)P.small_off;
                    const bool split = P.small_off != 0;
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) {
                        if (P.debug & 1) break;
                        const uint64_t ko = (uint64_t)(k * 2);  // 32 bytes per k-step, in 16-byte units
                        const bool fresh = sg.first && k == 0;
                        umma_tf32(ds, dal + ko, dbh + ko, idesc, fresh ? 0u : 1u);
                        umma_tf32(ds, dah + ko, dbl + ko, idesc, 1u);
                        umma_tf32(d, dah + ko, dbh + ko, idesc, (fresh && split) ? 0u : 1u);
                    }
                    umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ============================== EPILOGUE ==============================
        const int quad = warp & 3;             // TMEM lane quadrant this warp may access
        const int chalf = warp >> 2;           // which half of this item's channels the warp handles
        const int m = quad * 32 + lane;        // TMEM lane = pixel row of the tile
        const int j_lo = chalf * (P.nhalf / 2), j_hi = j_lo + P.nhalf / 2;
        int it = 0;
        for (int tile = first_tile; tile < P.n_tiles; tile += tile_stride, ++it) {
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&acc_full[buf], acc_phase);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * P.acc_cols);
            const long long p = (long long)tile * TILE_M + m;
            const bool valid = p < P.P;
            const int ch0 = half * P.nhalf;
            if (P.debug & 4) {
            } else if (P.mode == MODE_GRU) {
                const int nh = P.nhalf;  // hidden channels of this half
                for (int j = j_lo; j < j_hi; j += 8) {
                    float xr[8], xz[8], xn[8], hr[8], hz[8], hn[8];
                    tmem_ld8(t0 + j, xr);
                    tmem_ld8(t0 + nh + j, xz);
                    tmem_ld8(t0 + 2 * nh + j, xn);
                    tmem_ld8(t0 + 3 * nh + j, hr);
                    tmem_ld8(t0 + 4 * nh + j, hz);
                    tmem_ld8(t0 + 5 * nh + j, hn);
                    tmem_ld_wait();
                    if (valid) {
                        const int Ch = P.cout;
                        float hp[8];
                        const float4* hpp = reinterpret_cast<const float4*>(P.hprev + p * Ch + ch0 + j);
                        float4 h0 = hpp[0], h1 = hpp[1];
                        hp[0] = h0.x; hp[1] = h0.y; hp[2] = h0.z; hp[3] = h0.w;
                        hp[4] = h1.x; hp[5] = h1.y; hp[6] = h1.z; hp[7] = h1.w;
                        float o[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int c = ch0 + j + q;
                            const float br = P.bias ? P.bias[c] : 0.f;
                            const float bz = P.bias ? P.bias[Ch + c] : 0.f;
                            const float bn = P.bias ? P.bias[2 * Ch + c] : 0.f;
                            // rnn_cells.py:121-125
                            const float r = sigmoid_acc((xr[q] + br) + hr[q]);
                            const float z = sigmoid_acc((xz[q] + bz) + hz[q]);
                            const float n = tanhf((xn[q] + bn) + r * hn[q]);
                            o[q] = n * (1.f - z) + z * hp[q];
                        }
                        float4* op = reinterpret_cast<float4*>(P.out + p * Ch + ch0 + j);
                        op[0] = make_float4(o[0], o[1], o[2], o[3]);
                        op[1] = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            } else {
                for (int j = j_lo; j < j_hi; j += 8) {
                    float a[8];
                    tmem_ld8(t0 + j, a);
                    if (P.small_off) {
                        float a2[8];
                        tmem_ld8(t0 + P.small_off + j, a2);
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[q] += a2[q];
                    }
                    if (valid) {
                        float o[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float v = a[q] + (P.bias ? P.bias[ch0 + j + q] : 0.f);
                            o[q] = (P.mode == MODE_CONV_RELU) ? fmaxf(v, 0.f) : v;
                        }
                        float4* op = reinterpret_cast<float4*>(P.out + p * P.cout + ch0 + j);
                        op[0] = make_float4(o[0], o[1], o[2], o[3]);
                        op[1] = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + LOAD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

// ---- weight packing ----------------------------------------------------------------------------------
// dst chunk layout: [rows x 128 B] SWIZZLE_128B, hi then lo.  Source value for (half, chunk, row n, col k) is
// given by a small descriptor evaluated on the device.
struct PackDesc {
    const float* w;       // conv: [Cout][Cin][k][k]; GRU: w_ih [3Ch][Cx] then w_hh via w2
    const float* w2;
    int mode;             // 0 conv taps (chunk = tap*2 + kchunk), 1 GRU (chunks 0,1 = ih ; 2,3 = hh), 2 im2col 5x5x4
    int cout, cin, ksz;   // conv geometry
    int nhalf;            // output channels per half
    int rows;             // rows per chunk
    int n_chunks;
};

__global__ void pack_weights_kernel(PackDesc D, float* dst) {
    const int total = 2 * D.n_chunks * D.rows * KC;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int k = t % KC;
        int r = (t / KC) % D.rows;
        int ch = (t / (KC * D.rows)) % D.n_chunks;
        int half = t / (KC * D.rows * D.n_chunks);
        float v = 0.f;
        if (D.mode == 0) {
            const int tap = ch >> 1, kc = ch & 1;
            const int co = half * D.nhalf + r, ci = kc * KC + k;
            if (co < D.cout && ci < D.cin) v = D.w[((long long)co * D.cin + ci) * D.ksz * D.ksz + tap];
        } else if (D.mode == 1) {
            // rows: [r gate (nhalf) ; z gate ; n gate] of this half
            const int Ch = D.cout;
            const int g = r / D.nhalf, j = r % D.nhalf;
            const int row = g * Ch + half * D.nhalf + j;
            const float* w = (ch < 2) ? D.w : D.w2;
            const int ci = (ch & 1) * KC + k;
            v = w[(long long)row * D.cin + ci];
        } else {
            // im2col of a 5x5 window over 4 channels: K index = tap*4 + ci, chunk ch covers taps 8*ch .. 8*ch+7
            const int kk = ch * KC + k;
            const int tap = kk >> 2, ci = kk & 3;
            const int co = half * D.nhalf + r;
            if (tap < 25 && co < D.cout) v = D.w[((long long)co * 4 + ci) * 25 + tap];
        }
        const float hi = tf32_rn(v);
        const float lo = tf32_rn(v - hi);
        const size_t chunk_floats = (size_t)D.rows * KC;
        float* base = dst + ((size_t)half * D.n_chunks + ch) * 2 * chunk_floats;
        const uint32_t off = (uint32_t)(r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4);
        base[off / 4] = hi;
        base[chunk_floats + off / 4] = lo;
    }
}

static size_t smem_needed(const Params& P) {
    return 1024 + (size_t)P.n_wchunks * 2 * P.wchunk_rows * 128 + (size_t)P.stages * STAGE_BYTES + 256;
}

static int g_debug = 0;

static int launch(Params& P, cudaStream_t st) {
    P.debug = g_debug;
    const size_t max_smem = device_max_smem_optin();
    P.stages = 4;
    while (P.stages > 2 && smem_needed(P) > max_smem) --P.stages;
    MRB_REQUIRE(smem_needed(P) <= max_smem, MRB_EUNSUPPORTED, "tensor-core conv: weights do not fit shared memory");
    MRB_CUDA(cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    int sms = device_sm_count();
    int grid = sms & ~1;  // pairs of CTAs (one per channel half)
    if (grid > 2 * P.n_tiles) grid = 2 * P.n_tiles;
    tc_kernel<<<grid, THREADS, smem_needed(P), st>>>(P);
    MRB_LAUNCHED();
    return MRB_OK;
}

}  // namespace tc
}  // namespace mrb

using namespace mrb;

extern "C" void mrb_tc_set_debug(int flags) { tc::g_debug = flags; }

extern "C" size_t mrb_tc_packed_floats(int kind, int cout, int cin, int k) {
    // kind 0: conv k x k (cin multiple of 32); 1: GRU 1x1 (cout = hidden, cin = 64 for both inputs); 2: conv 5x5 x 4ch
    if (kind == 0) return (size_t)2 * (k * k * (cin / 32)) * 2 * (cout / 2) * 32;
    if (kind == 1) return (size_t)2 * 4 * 2 * (3 * cout / 2) * 32;
    return (size_t)2 * 4 * 2 * (cout / 2) * 32;
}

extern "C" int mrb_tc_pack_conv(const void* w, void* dst, int cout, int cin, int k, void* stream) {
    MRB_REQUIRE(w && dst, MRB_EINVAL, "mrb_tc_pack_conv: null pointer");
    MRB_REQUIRE(cin == 64 && (cout % 32) == 0 && cout >= 32 && cout <= 256 && (k % 2) == 1 && k * k * 2 <= tc::MAX_SEGS,
                MRB_EUNSUPPORTED, "mrb_tc_pack_conv: need cin == 64, cout multiple of 32, k*k*2 <= %d", tc::MAX_SEGS);
    tc::PackDesc D{(const float*)w, nullptr, 0, cout, cin, k, cout / 2, cout / 2, k * k * 2};
    tc::pack_weights_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(D, (float*)dst);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_tc_pack_gru(const void* w_ih, const void* w_hh, void* dst, int ch, int cx, void* stream) {
    MRB_REQUIRE(w_ih && w_hh && dst, MRB_EINVAL, "mrb_tc_pack_gru: null pointer");
    MRB_REQUIRE(ch == 64 && cx == 64, MRB_EUNSUPPORTED, "mrb_tc_pack_gru: tensor-core GRU needs 64 input and hidden channels");
    tc::PackDesc D{(const float*)w_ih, (const float*)w_hh, 1, ch, cx, 1, ch / 2, 3 * ch / 2, 4};
    tc::pack_weights_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(D, (float*)dst);
    MRB_LAUNCHED();
    return MRB_OK;
}

extern "C" int mrb_tc_pack_conv5x5x4(const void* w, void* dst, int cout, void* stream) {
    MRB_REQUIRE(w && dst, MRB_EINVAL, "mrb_tc_pack_conv5x5x4: null pointer");
    MRB_REQUIRE((cout % 32) == 0 && cout >= 32 && cout <= 256, MRB_EUNSUPPORTED, "mrb_tc_pack_conv5x5x4: bad cout");
    tc::PackDesc D{(const float*)w, nullptr, 2, cout, 4, 5, cout / 2, cout / 2, 4};
    tc::pack_weights_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(D, (float*)dst);
    MRB_LAUNCHED();
    return MRB_OK;
}

static int tc_common(tc::Params& P, int B, int H, int W) {
    MRB_REQUIRE(B >= 1 && H >= 1 && W >= 1, MRB_EINVAL, "tensor-core conv: bad shape");
    P.B = B; P.H = H; P.W = W;
    P.P = (long long)B * H * W;
    long long nt = (P.P + tc::TILE_M - 1) / tc::TILE_M;
    MRB_REQUIRE(nt <= 2147483647LL, MRB_EUNSUPPORTED, "tensor-core conv: too many pixels");
    P.n_tiles = (int)nt;
    return MRB_OK;
}

// ConvNonlinear k x k (dilated, replicate padding), 64 -> cout channels, NHWC, bias + optional ReLU.
extern "C" int mrb_tc_conv_nhwc(const void* x, const void* wpack, const void* bias, void* out, int B, int H, int W,
                                int cout, int k, int dil, int relu, void* stream) {
    MRB_REQUIRE(x && wpack && out, MRB_EINVAL, "mrb_tc_conv_nhwc: null pointer");
    MRB_REQUIRE((cout % 32) == 0 && cout >= 32 && cout <= 256 && (k % 2) == 1 && k * k * 2 <= tc::MAX_SEGS && dil >= 1,
                MRB_EUNSUPPORTED, "mrb_tc_conv_nhwc: unsupported geometry");
    tc::Params P;
    memset(&P, 0, sizeof(P));
    int rc = tc_common(P, B, H, W);
    if (rc) return rc;
    P.src[0] = (const float*)x; P.cs[0] = 64; P.src[1] = nullptr; P.cs[1] = 0;
    P.wpack = (const float*)wpack; P.bias = (const float*)bias; P.out = (float*)out;
    P.cout = cout; P.nhalf = cout / 2;
    P.wchunk_rows = cout / 2; P.n_wchunks = k * k * 2;
    P.acc_cols = cout;  // hi*hi chain + cross-term chain
    P.small_off = cout / 2;
    P.tmem_cols = 32;
    while (P.tmem_cols < 2 * P.acc_cols) P.tmem_cols *= 2;
    P.mode = relu ? tc::MODE_CONV_RELU : tc::MODE_CONV_NOACT;
    P.nseg = k * k * 2;
    const int pad = dil * (k - 1) / 2;
    for (int t = 0; t < k * k; ++t)
        for (int kc = 0; kc < 2; ++kc) {
            tc::Segment& s = P.seg[t * 2 + kc];
            s.src = 0; s.dy = (short)((t / k) * dil - pad); s.dx = (short)((t % k) * dil - pad);
            s.c0 = (short)(kc * 32); s.wchunk = (short)(t * 2 + kc); s.dcol = 0; s.n = (short)(cout / 2);
            s.first = (t == 0 && kc == 0);
        }
    return tc::launch(P, (cudaStream_t)stream);
}

// ConvNonlinear 5x5 over a 4-channel NHWC input (the RIM gradient), cout outputs, bias + ReLU.
extern "C" int mrb_tc_conv5x5x4_nhwc(const void* x, const void* wpack, const void* bias, void* out, int B, int H, int W,
                                     int cout, int relu, void* stream) {
    MRB_REQUIRE(x && wpack && out, MRB_EINVAL, "mrb_tc_conv5x5x4_nhwc: null pointer");
    MRB_REQUIRE((cout % 32) == 0 && cout >= 32 && cout <= 256, MRB_EUNSUPPORTED, "mrb_tc_conv5x5x4_nhwc: bad cout");
    tc::Params P;
    memset(&P, 0, sizeof(P));
    int rc = tc_common(P, B, H, W);
    if (rc) return rc;
    P.src[0] = (const float*)x; P.cs[0] = 4;
    P.wpack = (const float*)wpack; P.bias = (const float*)bias; P.out = (float*)out;
    P.cout = cout; P.nhalf = cout / 2;
    P.wchunk_rows = cout / 2; P.n_wchunks = 4;
    P.acc_cols = cout;
    P.small_off = cout / 2;
    P.tmem_cols = 32;
    while (P.tmem_cols < 2 * P.acc_cols) P.tmem_cols *= 2;
    P.mode = relu ? tc::MODE_CONV_RELU : tc::MODE_CONV_NOACT;
    P.im2col = 1;
    P.nseg = 4;
    for (int c = 0; c < 4; ++c) {
        tc::Segment& s = P.seg[c];
        s.src = 0; s.dy = 0; s.dx = 0; s.c0 = (short)c; s.wchunk = (short)c; s.dcol = 0; s.n = (short)(cout / 2);
        s.first = (c == 0);
    }
    return tc::launch(P, (cudaStream_t)stream);
}

// ConvGRUCell, kernel size 1, 64 input and 64 hidden channels, NHWC.
extern "C" int mrb_tc_gru_nhwc(const void* x, const void* h, const void* wpack, const void* b_ih, void* h_out, int B,
                               int H, int W, int ch, void* stream) {
    MRB_REQUIRE(x && h && wpack && h_out, MRB_EINVAL, "mrb_tc_gru_nhwc: null pointer");
    MRB_REQUIRE(h_out != h, MRB_EINVAL, "mrb_tc_gru_nhwc: h_out must not alias h");
    MRB_REQUIRE(ch == 64, MRB_EUNSUPPORTED, "mrb_tc_gru_nhwc: 64 channels only");
    tc::Params P;
    memset(&P, 0, sizeof(P));
    int rc = tc_common(P, B, H, W);
    if (rc) return rc;
    P.src[0] = (const float*)x; P.cs[0] = 64; P.src[1] = (const float*)h; P.cs[1] = 64;
    P.wpack = (const float*)wpack; P.bias = (const float*)b_ih; P.hprev = (const float*)h; P.out = (float*)h_out;
    P.cout = ch; P.nhalf = ch / 2;
    P.wchunk_rows = 3 * ch / 2; P.n_wchunks = 4;
    P.acc_cols = 6 * (ch / 2);  // x-part (r,z,n) then h-part (r,z,n)
    P.tmem_cols = 512;
    P.mode = tc::MODE_GRU;
    P.nseg = 4;
    for (int i = 0; i < 4; ++i) {
        tc::Segment& s = P.seg[i];
        s.src = (short)(i >> 1); s.dy = 0; s.dx = 0; s.c0 = (short)((i & 1) * 32); s.wchunk = (short)i;
        s.dcol = (short)((i >> 1) * 3 * (ch / 2)); s.n = (short)(3 * ch / 2); s.first = ((i & 1) == 0);
    }
    return tc::launch(P, (cudaStream_t)stream);
}
