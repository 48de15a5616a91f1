// Micro-benchmark (experiments only, not product code): cycles per tcgen05.mma for M128 x N x K16 bf16,
// A from shared memory (SS) or tensor memory (TS), accumulating into one TMEM tile or alternating
// between two.  One CTA per SM, one issuing lane.  Operands are whatever is in smem/TMEM (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 SS, 1 TS, 2 TS with MN-major B (the P V product of attn_tc.cu), 3 TS while 4 other warps stream tcgen05.ld.  n: UMMA N.  alt: number of accumulators cycled through (1, 2 or 4).  kchain: MMAs per commit group.
__global__ void __launch_bounds__(256, 1) k_bench(int mode, int n, int alt, int iters, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode == 2 ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = umma_desc(base, 16, 1024);
        const uint64_t db = mode == 2 ? umma_desc(base + 16384, 16384, 1024) : umma_desc(base + 16384, 16, 1024);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            const uint32_t amask = (uint32_t)alt - 1u, ta = tmem + 448;
            if (mode == 0) {
#pragma unroll 8
                for (int it = 0; it < iters; ++it) mma_ss(tmem + ((uint32_t)it & amask) * (uint32_t)n, da, db, idesc, 1u);
            } else {
#pragma unroll 8
                for (int it = 0; it < iters; ++it) mma_ts(tmem + ((uint32_t)it & amask) * (uint32_t)n, ta, db, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
        long long tt = __shfl_sync(0xffffffffu, t0, 0);
        (void)tt;
    }
    if (warp >= 4 && mode == 3) {
        // 4 warps (one per TMEM lane quarter) read a 128-column tile over and over while the MMAs run
        uint32_t acc = 0;
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        for (int it = 0; it < iters / 4; ++it) {
            uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
            for (int c = 0; c < 128; c += 8) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(ta + c) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
            }
        }
        if (acc == 0x12345u) out[1] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// Queue depth of the tensor pipe as seen by the issuing thread: groups of `group` SS N=256 MMAs separated by a
// dependent ALU chain of `delay` steps (~4 cycles each).  If the delay is hidden, time per group stays group*128.
__global__ void __launch_bounds__(128, 1) k_queue(int group, int delay, int groups, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = umma_desc(base, 16, 1024), db = umma_desc(base + 16384, 16, 1024);
        long long t0 = clock64();
        uint32_t x = (uint32_t)t0;
        for (int g = 0; g < groups; ++g) {
            if (elect_one()) {
#pragma unroll 4
                for (int it = 0; it < group; ++it) mma_ss(tmem, da, db, idesc, 1u);
            }
            __syncwarp();
            for (int d = 0; d < delay; ++d) x = x * 1664525u + 1013904223u;      // dependent chain, ~4-6 cycles per step
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = x; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// The issue pattern of attn_tc.cu's pass 2 without the softmax: per key tile 24 TS MMAs (N=128) into score tile
// j&1, then 24 TS MMAs (N=128) into O whose A operand is (alias=1) the score tile written two tiles earlier —
// the layout the kernel uses — or (alias=0) a separate TMEM region.  Ideal: 48 x 64 cycles per tile.
template <int commits>
__global__ void __launch_bounds__(128, 1) k_attn_pattern(int alias, int tiles, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 66 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t db = umma_desc(base + 16384, 16, 1024);
        const uint32_t idesc_mn = idesc | (1u << 16);
        const long long t0 = clock64();
        for (int j = 0; j < tiles; ++j) {
            const uint32_t s_tile = tmem + (uint32_t)(j & 1) * 128u, o_tile = tmem + 256u;
            const uint32_t p_src = alias ? s_tile : tmem + 384u;       // P(j) lives where S(j) was (alias) or elsewhere
            if (alias == 2) {                                          // as attn_tc.cu: V MN-major, 16-key steps of 2048 B, two 64-column chunks 16 KB apart
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < 24; ++i)
                        mma_ts(o_tile, s_tile + (uint32_t)((i & 7) >> 1) * 32u + (uint32_t)(i & 1) * 8u,
                               umma_desc(base + (uint32_t)(i & 7) * 2048u, 16384, 1024), idesc_mn, (j | i) ? 1u : 0u);
#pragma unroll
                    for (int i = 0; i < 24; ++i)
                        mma_ts(s_tile, tmem + 384u + (uint32_t)(i & 7) * 8u,
                               umma_desc(base + 32768 + (uint32_t)((i & 7) >> 2) * 16384u + (uint32_t)(i & 3) * 32u, 16, 1024), idesc, i ? 1u : 0u);
                }
                __syncwarp();
                continue;
            }
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    mma_ts(o_tile, p_src + (uint32_t)(i & 7) * 8u, db, idesc, (j | i) ? 1u : 0u);   // P(j) V(j)
                    if (commits && (i % commits) == commits - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                }
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    mma_ts(s_tile, tmem + 384u + (uint32_t)(i & 7) * 8u, db, idesc, i ? 1u : 0u); // S(j+2) over the same tile
                    if (commits && (i % commits) == commits - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// The same pass-2 issue pattern while other agents use the SM as in attn_tc.cu: (flags & 1) one warp streams bulk
// copies global -> shared memory (what the K/V ring does: 128 KB per key tile in the kernel), (flags & 2) 16 warps
// read and write 32-column chunks of the score tiles (tcgen05.ld / tcgen05.st, what the softmax warps do).
// out[0] = cycles of the MMA stream, out[1] = bytes copied while it ran, out[2] = ld+st chunk pairs done.
__global__ void __launch_bounds__(576, 1) k_attn_traffic(int flags, int tiles, int pace, const unsigned char* gsrc, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar, cbar[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        done = 0;
    }
    for (int i = threadIdx.x; i < 66 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_mn = idesc | (1u << 16);
        const long long t0 = clock64();
        for (int j = 0; j < tiles; ++j) {
            const uint32_t s_tile = tmem + (uint32_t)(j & 1) * 128u, o_tile = tmem + 256u;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 24; ++i)
                    mma_ts(o_tile, s_tile + (uint32_t)((i & 7) >> 1) * 32u + (uint32_t)(i & 1) * 8u,
                           umma_desc(base + (uint32_t)(i & 7) * 2048u, 16384, 1024), idesc_mn, (j | i) ? 1u : 0u);
#pragma unroll
                for (int i = 0; i < 24; ++i)
                    mma_ts(s_tile, tmem + 384u + (uint32_t)(i & 7) * 8u,
                           umma_desc(base + 32768 + (uint32_t)((i & 7) >> 2) * 16384u + (uint32_t)(i & 3) * 32u, 16, 1024), idesc, i ? 1u : 0u);
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t1 = clock64();
        done = 1;
        if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    } else if (warp == 1) {
        if ((flags & 1) && lane == 0) {
            // 4 copies of 32 KB in flight into a 128 KB region above the operands; optionally paced to `pace` cycles per copy
            long long bytes = 0;
            uint32_t ph[4] = {0, 0, 0, 0};
            const unsigned char* src = gsrc + (size_t)blockIdx.x * (1u << 20);
            int it = 0;
            long long next = clock64();
            for (int i = 0; i < 4; ++i, ++it) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[i])), "r"(32768u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(base + 66 * 1024 + i * 32768), "l"(src + (size_t)(it & 31) * 32768), "r"(32768u), "r"(smem_u32(&cbar[i])) : "memory");
            }
            while (!done) {
                const int i = it & 3;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(ok) : "r"(smem_u32(&cbar[i])), "r"(ph[i]) : "memory");
                ph[i] ^= 1;
                bytes += 32768;
                if (pace) { next += pace; while (clock64() < next) {} }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[i])), "r"(32768u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(base + 66 * 1024 + i * 32768), "l"(src + (size_t)(it & 31) * 32768), "r"(32768u), "r"(smem_u32(&cbar[i])) : "memory");
                ++it;
            }
            for (int i = 0; i < 4; ++i) {            // drain
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(ok) : "r"(smem_u32(&cbar[(it + i) & 3])), "r"(ph[(it + i) & 3]) : "memory");
            }
            if (blockIdx.x == 0) out[1] = bytes;
        }
    } else if (flags & 2) {
        const int g = warp - 2;                                    // 16 warps: lane quarter warp & 3, 64-column slice g >> 2
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g >> 2) * 64u;
        long long n = 0;
        uint32_t acc = 0;
        while (!done) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(ta + c * 32) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) { acc += r[i]; r[i] = (r[i] & 0x3fff3fffu) | 0x3c003c00u; }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                             ::"r"(ta + c * 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                               "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (pace) { const long long t = clock64(); while (clock64() - t < pace / 4) {} }
            ++n;
        }
        if (acc == 0x12345u) out[3] = acc;
        if (blockIdx.x == 0 && threadIdx.x == 64) out[2] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int iters = 4096;
    for (int warm = 0; warm < 2; ++warm)
        for (int mode = 0; mode < 4; ++mode)
            for (int n : {64, 128, 256})
                for (int alt : {1, 2, 4}) {
                    if (alt * n > 256) continue;
                    if (mode >= 2 && (alt != 1 || n == 256)) continue;
                    k_bench<<<148, 256, 66 * 1024>>>(mode, n, alt, iters, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long cyc = 0;
                    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
                    if (warm)
                        printf("%-13s N=%3d accumulators=%d : %7.1f cycles per MMA (ideal %d) %s\n", mode == 0 ? "SS" : mode == 1 ? "TS" : mode == 2 ? "TS B=MN-major" : "TS + 4w LDTM", n, alt,
                               (double)cyc / iters, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    auto run_pattern = [&](auto kern, int commits) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
        for (int alias : {0, 1, 2}) {
            long long cyc = 0;
            for (int rep = 0; rep < 2; ++rep) {
                kern<<<148, 128, 66 * 1024>>>(alias, 256, d_out);
                cudaDeviceSynchronize();
                cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
            }
            printf("attention pattern, P %s, a commit every %2d MMAs: %7.1f cycles per key tile (ideal 3072)\n",
                   alias == 2 ? "in place, V MN-major stepping as attn_tc.cu" : alias ? "in place over S" : "in its own columns", commits, (double)cyc / 256);
        }
    };
    run_pattern(k_attn_pattern<0>, 0);
    run_pattern(k_attn_pattern<24>, 24);
    run_pattern(k_attn_pattern<8>, 8);
    run_pattern(k_attn_pattern<4>, 4);
    run_pattern(k_attn_pattern<1>, 1);
    cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    for (int group : {1, 4, 12})
        for (int delay : {0, 25, 50, 100, 200, 400}) {
            const int groups = 1024;
            long long cyc = 0, base_cyc = 0;
            for (int rep = 0; rep < 2; ++rep) {
                k_queue<<<148, 128, 66 * 1024>>>(group, delay, groups, d_out);
                cudaDeviceSynchronize();
                cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
            }
            k_queue<<<148, 128, 66 * 1024>>>(0, delay, groups, d_out);     // the delay chain alone
            cudaDeviceSynchronize();
            cudaMemcpy(&base_cyc, d_out, 8, cudaMemcpyDeviceToHost);
            printf("queue: %2d MMAs (N=256, %4d cycles) then a %6.1f-cycle chain : %7.1f cycles per group\n", group, group * 128,
                   (double)base_cyc / groups, (double)cyc / groups);
        }
    {
        unsigned char* gsrc;
        cudaMalloc(&gsrc, (size_t)148 << 20);
        cudaMemset(gsrc, 0x3c, (size_t)148 << 20);
        cudaFuncSetAttribute(k_attn_traffic, cudaFuncAttributeMaxDynamicSharedMemorySize, (66 + 128 + 2) * 1024);
        for (int pace : {0, 768})
            for (int flags : {0, 1, 2, 3}) {
                long long h[4] = {0, 0, 0, 0};
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemset(d_out, 0, 32);
                    k_attn_traffic<<<148, 576, (66 + 128 + 2) * 1024>>>(flags, 256, pace, gsrc, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) printf("traffic kernel: %s\n", cudaGetErrorString(e));
                    cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
                }
                printf("attention pattern + %s%s%s (%s): %7.1f cycles per key tile (ideal 3072), %.0f KB copied and %.1f ld/st chunk pairs per warp per tile\n",
                       flags & 1 ? "bulk copies into smem" : "", flags == 3 ? " + " : "", flags & 2 ? "16 warps of tcgen05.ld/st" : (flags ? "" : "nothing"),
                       pace ? "paced like the kernel" : "flat out", (double)h[0] / 256, (double)h[1] / 256 / 1024, 2.0 * (double)h[2] / 256);
            }
    }
    return 0;
}
