// Micro-benchmark of tcgen05.mma.kind::tf32 issue/dependency behaviour (tools/ only; not on the product path).
// One CTA, one issuing thread: `iters` MMAs (M=128, K=8) into `nacc` round-robin accumulators, A from TMEM or smem.
#include "common.cuh"

namespace mrb {
namespace tcmb {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) mb_kernel(int N, int nacc, int iters, int a_in_tmem, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* f = (float*)smem;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) f[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    if (a_in_tmem == 5) {
        // two issuing threads (lane 0 of warps 0 and 1), disjoint accumulators: is the ~45-cycle floor per thread or per pipe?
        __shared__ uint64_t bar5[2];
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar5[0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar5[1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 64) {
            const int w = threadIdx.x >> 5;
            const uint64_t bdesc = make_desc(smem_u32(smem + 16384));
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_t = tb + 480;
            const uint32_t d0 = tb + (uint32_t)(w * 2 * N), d1 = d0 + (uint32_t)N;
            long long t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t}\n" ::"r"(d0),
                    "r"(d1), "r"(a_t), "l"(bdesc), "r"(idesc)
                    : "memory");
            }
            long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar5[w])) : "memory");
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok)
                             : "r"(smem_u32(&bar5[w]))
                             : "memory");
            }
            long long t2 = clock64();
            if (w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else
    if (threadIdx.x == 0) {
        const uint64_t adesc = make_desc(smem_u32(smem));
        const uint64_t bdesc = make_desc(smem_u32(smem + 16384));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_t = tb + 480;  // A operand columns (garbage values, timing only)
        long long t0 = clock64();
        if (a_in_tmem == 3 || a_in_tmem == 4) {
            // like mode 2 but with a tcgen05.commit (mode 3), or commit + fence::after (mode 4), after every 12 MMAs
            __shared__ uint64_t bar2;
            if (true) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
            const uint32_t d0 = tb, d1 = tb + (uint32_t)N;
            for (int i = 0; i < iters; i += 12) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t}\n" ::"r"(d0),
                    "r"(d1), "r"(a_t), "l"(bdesc), "r"(idesc)
                    : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                if (a_in_tmem == 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
        } else if (a_in_tmem == 2 || a_in_tmem == 6) {
            // tight issue loop: 8 unrolled MMAs per iteration, predicate hoisted; mode 2 alternates between two
            // accumulators, mode 6 chains every MMA onto the same accumulator columns (dependent accumulation)
            const uint32_t d0 = tb, d1 = a_in_tmem == 6 ? tb : tb + (uint32_t)N;
            for (int i = 0; i < iters; i += 8) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %3, %4, p;\n\t}\n" ::"r"(d0),
                    "r"(d1), "r"(a_t), "l"(bdesc), "r"(idesc)
                    : "memory");
            }
        } else
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tb + (uint32_t)((i % nacc) * N);
            if (a_in_tmem)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                             "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                             "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u)
                             : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar))
                         : "memory");
        }
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
}  // namespace tcmb
}  // namespace mrb

extern "C" int mrb_tc_microbench(int N, int nacc, int iters, int a_in_tmem, void* out, void* stream) {
    using namespace mrb;
    MRB_REQUIRE(out && N >= 16 && N <= 256 && (N % 16) == 0 && nacc >= 1 && nacc * N <= 448, MRB_EINVAL, "microbench: bad args");
    MRB_CUDA(cudaFuncSetAttribute(tcmb::mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    tcmb::mb_kernel<<<1, 128, 56 * 1024, (cudaStream_t)stream>>>(N, nacc, iters, a_in_tmem, (long long*)out);
    MRB_LAUNCHED();
    return MRB_OK;
}
