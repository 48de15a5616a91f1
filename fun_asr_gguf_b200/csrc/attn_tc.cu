// Flash-style masked attention on tcgen05 / TMEM / TMA (model_definition.py:68-90
// for the SAN-M encoder, :132-145 for the adaptor and CTC blocks).
//
//   scores = (q * d_k^-0.5) k^T + (m-1)*10000 ; softmax over keys ; times v
//
// In fp32 the additive -10000 makes a masked key's weight exactly zero, so masked keys are never
// visited: kv_len[b] keys are attended (valid frames for encoder/adaptor, every physical frame for
// the unmasked CTC head — SURVEY F7).
//
// Precision: q (pre-scaled by d_k^-0.5 * log2 e in the producing GEMM's epilogue), k, v and the
// softmax weights p are all carried as bf16 hi/lo planes and every product is hi*hi + hi*lo + lo*hi
// into an fp32 TMEM accumulator, like the projections.  The softmax itself is fp32.
//
// Two passes over the keys instead of an online rescale: pass 1 only takes the row maxima of S,
// pass 2 recomputes S, forms p = exp2(S - max) and accumulates O = sum p v in TMEM with no
// correction step.  That costs one extra QK^T (tensor time 9 instead of 6 units) but removes the
// TMEM read-modify-write of O and every dependency between a tile's softmax and the previous PV.
//
// One CTA per SM, persistent over (segment, head, 128-query tile) items:
//   warp 0      TMA: Q planes once per item; 64-key K / V tiles into a 4-slot ring (pass 1: four K tiles in
//               flight; pass 2: K in slots 0-1, freed as soon as S is issued, V in slots 2-3, freed after PV)
//   warp 1      TMEM allocator + MMA issuer: S = Q K^T (M128 x N64, K-major both) into a double-buffered
//               TMEM tile, O += P V (M128 x N d_k, P from smem K-major, V from smem MN-major: no transpose)
//   warps 2..5  softmax: thread = query row = TMEM lane; tcgen05.ld S, max / exp2 / sum, split p into
//               hi/lo planes written to smem in the UMMA SWIZZLE_128B layout; final 1/l scaling and store
#include "kernels.h"
#include "tc_ptx.cuh"

namespace fa {

namespace {

constexpr int QT = 128;          // queries per item (UMMA M)
constexpr int KT = 64;           // keys per tile (UMMA N for S, K extent for PV)
constexpr int kAttThreads = 192;

template <int DK> struct ACfg {
    static constexpr int kChunks = DK / 64;                 // 128-byte column chunks per head row
    static constexpr int kQBytes = 2 * kChunks * QT * 128;   // planes x chunks x rows x 128 B
    static constexpr int kKBytes = 2 * kChunks * KT * 128;   // one 64-key tile of K (or V), both planes
    static constexpr int kSlots = 4;                         // pass 1: four K tiles in flight; pass 2: K in slots 0-1, V in 2-3
    static constexpr int kPBytes = 2 * QT * 128;             // planes x rows x (64 keys * 2 B)
    static constexpr int kSmemBytes = kQBytes + kSlots * kKBytes + kPBytes + 1024 + 256;
    static constexpr uint32_t kTmemCols = 256;                      // S0 [0,64) | S1 [64,128) | O [128,128+DK)
    static constexpr uint32_t kOCol = 128;
};

// 2^x for x <= 0 on the MUFU pipe (2 ulp); denormal results flush to zero, which is what a vanishing
// softmax weight should do.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnParams {
    int batch, frames, heads, d_model, ld;     // ld = row stride (elements) of the qkv planes
    const int* kv_len;
    float* ctx;
    __nv_bfloat16* ctx_hi;
    __nv_bfloat16* ctx_lo;
    int ldo;
};

template <int DK>
__global__ void __launch_bounds__(kAttThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, AttnParams p) {
    using C = ACfg<DK>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t q_base = (raw + 1023u) & ~1023u;
    const uint32_t kv_base = q_base + C::kQBytes;
    const uint32_t p_base = kv_base + C::kSlots * C::kKBytes;
    const uint32_t bars = p_base + C::kPBytes;
    const uint32_t bar_qfull = bars, bar_qempty = bars + 8;
    const uint32_t bar_kvfull = bars + 16, bar_kvempty = bars + 48;       // [4] each: one pair per ring slot
    const uint32_t bar_sfull = bars + 80, bar_sempty = bars + 96;         // [2] each
    const uint32_t bar_pfull = bars + 112, bar_pempty = bars + 120;
    const uint32_t bar_ofull = bars + 128, bar_oempty = bars + 136;
    const uint32_t tmem_slot = bars + 144;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    unsigned char* p_ptr = smem_raw + (p_base - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tiles = (p.frames + QT - 1) / QT;
    const int items = p.batch * p.heads * q_tiles;

    if (threadIdx.x == 0) {
        mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(bar_kvfull + 8 * s, 1); mbar_init(bar_kvempty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, 4); }
        mbar_init(bar_pfull, 4); mbar_init(bar_pempty, 1);
        mbar_init(bar_ofull, 1); mbar_init(bar_oempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // item -> (segment b, head h, query tile qt); key tiles n
    auto decode = [&](int item, int& b, int& h, int& qt, int& n, int& klen) {
        qt = item % q_tiles;
        const int bh = item / q_tiles;
        h = bh % p.heads;
        b = bh / p.heads;
        klen = p.kv_len ? p.kv_len[b] : p.frames;
        n = (klen + KT - 1) / KT;
    };

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            uint32_t item_it = 0, fills = 0;          // bit s of `fills`: parity of how often slot s has been filled
            // one 64-key tile of K (which = 1) or V (which = 2), both planes, into ring slot `slot`
            auto load_tile = [&](int slot, int which, int h, int row) {
                mbar_wait(bar_kvempty + 8 * slot, ((fills >> slot) & 1) ^ 1);
                const uint32_t full = bar_kvfull + 8 * slot, sb = kv_base + slot * C::kKBytes;
                mbar_arrive_expect_tx(full, C::kKBytes);
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int c = 0; c < C::kChunks; ++c)
                        tma_load_3d(sb + (pl * C::kChunks + c) * KT * 128, &map_kv, full, which * p.d_model + h * DK + c * 64, row, pl);
                fills ^= 1u << slot;
            };
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_it) {
                int b, h, qt, n, klen;
                decode(item, b, h, qt, n, klen);
                const int row0 = b * p.frames;
                mbar_wait(bar_qempty, (item_it & 1) ^ 1);
                mbar_arrive_expect_tx(bar_qfull, C::kQBytes);
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int c = 0; c < C::kChunks; ++c)
                        tma_load_3d(q_base + (pl * C::kChunks + c) * QT * 128, &map_q, bar_qfull, h * DK + c * 64,
                                    row0 + qt * QT, pl);
                for (int j = 0; j < n; ++j) load_tile(j & 3, 1, h, row0 + j * KT);             // pass 1: K only
                for (int j = 0; j < n; ++j) {                                                  // pass 2: K then V
                    load_tile(j & 1, 1, h, row0 + j * KT);
                    load_tile(2 + (j & 1), 2, h, row0 + j * KT);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t kIdescS = umma_idesc_bf16(QT, KT, false);
            constexpr uint32_t kIdescO = umma_idesc_bf16(QT, DK, true);
            uint32_t item_it = 0, uses = 0, s_it = 0, p_it = 0;   // bit s of `uses`: parity of how often slot s has been consumed
            // S[sbuf] = Q K^T over the K planes in `slot`: lo*hi, hi*lo, then hi*hi
            auto issue_s = [&](int slot, uint32_t sbuf) {
                const uint32_t stage_base = kv_base + slot * C::kKBytes;
                const uint32_t tmem_s = tmem_base + sbuf * KT;
                uint32_t accum = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const int qa = term == 0 ? 1 : 0, kb = term == 1 ? 1 : 0;     // plane of Q, plane of K
#pragma unroll
                    for (int ks = 0; ks < DK / 16; ++ks) {
                        const int c = ks >> 2, o = (ks & 3) * 32;
                        const uint64_t da = umma_desc(q_base + (qa * C::kChunks + c) * QT * 128 + o, 16, 1024);
                        const uint64_t db = umma_desc(stage_base + (kb * C::kChunks + c) * KT * 128 + o, 16, 1024);
                        tc_mma(tmem_s, da, db, kIdescS, accum);
                        accum = 1;
                    }
                }
            };
            // wait for K in `slot` and a free S buffer, issue S, signal the softmax warps, free the slot
            auto do_s = [&](int slot) {
                const uint32_t sbuf = s_it & 1;
                mbar_wait(bar_kvfull + 8 * slot, (uses >> slot) & 1);
                mbar_wait(bar_sempty + 8 * sbuf, ((s_it >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_s(slot, sbuf);
                tc_commit(bar_sfull + 8 * sbuf);
                tc_commit(bar_kvempty + 8 * slot);
                uses ^= 1u << slot;
                ++s_it;
            };
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_it) {
                int b, h, qt, n, klen;
                decode(item, b, h, qt, n, klen);
                mbar_wait(bar_qfull, item_it & 1);
                tc_fence_after();
                for (int j = 0; j < n; ++j) do_s(j & 3);            // ---- pass 1: row maxima
                // ---- pass 2: S runs one tile ahead of PV
                do_s(0);
                if (n == 1) tc_commit(bar_qempty);                  // last read of Q: the next item's Q may land
                mbar_wait(bar_oempty, (item_it & 1) ^ 1);           // previous item's O has been read out
                tc_fence_after();
                for (int j = 0; j < n; ++j, ++p_it) {
                    if (j + 1 < n) {
                        do_s((j + 1) & 1);
                        if (j + 2 == n) tc_commit(bar_qempty);
                    }
                    const int vslot = 2 + (j & 1);
                    const uint32_t v_base = kv_base + vslot * C::kKBytes;
                    mbar_wait(bar_kvfull + 8 * vslot, (uses >> vslot) & 1);
                    mbar_wait(bar_pfull, p_it & 1);
                    tc_fence_after();
                    const uint32_t tmem_o = tmem_base + C::kOCol;
                    uint32_t accum = j > 0 ? 1u : 0u;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const int pa = term == 0 ? 1 : 0, vb = term == 1 ? 1 : 0;  // plane of P, plane of V
#pragma unroll
                        for (int ks = 0; ks < KT / 16; ++ks) {
                            // A = P [128 q][64 keys] K-major; B = V [16 keys][DK] MN-major: 64-column chunks
                            // KT*128 B apart (LBO), 8-key groups 1024 B apart (SBO)
                            const uint64_t da = umma_desc(p_base + pa * QT * 128 + ks * 32, 16, 1024);
                            const uint64_t db = umma_desc(v_base + vb * C::kChunks * KT * 128 + ks * 16 * 128, KT * 128, 1024);
                            tc_mma(tmem_o, da, db, kIdescO, accum);
                            accum = 1;
                        }
                    }
                    tc_commit(bar_pempty);
                    tc_commit(bar_kvempty + 8 * vslot);
                    uses ^= 1u << vslot;
                }
                tc_commit(bar_ofull);
            }
        }
        __syncwarp();
    } else {
        // ================================================================== softmax + epilogue
        const int lane_grp = warp & 3;
        const int r = lane_grp * 32 + lane;                         // query row in the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
        uint32_t item_it = 0, s_it = 0, p_it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_it) {
            int b, h, qt, n, klen;
            decode(item, b, h, qt, n, klen);
            float mx = -INFINITY;
            // ---- pass 1
            for (int j = 0; j < n; ++j, ++s_it) {
                const uint32_t sbuf = s_it & 1;
                mbar_wait(bar_sfull + 8 * sbuf, (s_it >> 1) & 1);
                tc_fence_after();
                uint32_t s0[32], s1[32];
                tc_ld32(tmem_base + lane_addr + sbuf * KT, s0);
                tc_ld32(tmem_base + lane_addr + sbuf * KT + 32, s1);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sempty + 8 * sbuf);
                const int kbase = j * KT;
                if (kbase + KT <= klen) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (kbase + i < klen) mx = fmaxf(mx, __uint_as_float(s0[i]));
                        if (kbase + 32 + i < klen) mx = fmaxf(mx, __uint_as_float(s1[i]));
                    }
                }
            }
            // ---- pass 2
            float lsum = 0.f;
            for (int j = 0; j < n; ++j, ++s_it, ++p_it) {
                const uint32_t sbuf = s_it & 1;
                mbar_wait(bar_sfull + 8 * sbuf, (s_it >> 1) & 1);
                tc_fence_after();
                uint32_t s0[32], s1[32];
                tc_ld32(tmem_base + lane_addr + sbuf * KT, s0);
                tc_ld32(tmem_base + lane_addr + sbuf * KT + 32, s1);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sempty + 8 * sbuf);
                const int kbase = j * KT;
                const bool full_tile = kbase + KT <= klen;
                uint32_t hi[32], lo[32];                             // 64 keys x bf16, packed in pairs
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int half = i >> 4, w = (i & 15) * 2;       // word i covers keys 2i, 2i+1
                    const uint32_t* src = half ? s1 : s0;
                    float p0 = fast_exp2(__uint_as_float(src[w]) - mx);
                    float p1 = fast_exp2(__uint_as_float(src[w + 1]) - mx);
                    if (!full_tile) {
                        if (kbase + 2 * i >= klen) p0 = 0.f;
                        if (kbase + 2 * i + 1 >= klen) p1 = 0.f;
                    }
                    lsum += p0 + p1;
                    split_bf16x2(p0, p1, hi[i], lo[i]);
                }
                mbar_wait(bar_pempty, (p_it & 1) ^ 1);               // PV of the previous tile has consumed P
                // row r of the K-major SWIZZLE_128B tile: 16-byte chunk c lives at chunk c ^ (r & 7)
                unsigned char* prow_hi = p_ptr + r * 128;
                unsigned char* prow_lo = p_ptr + QT * 128 + r * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int pc = (c ^ (r & 7)) * 16;
                    *reinterpret_cast<uint4*>(prow_hi + pc) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    *reinterpret_cast<uint4*>(prow_lo + pc) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
                fence_async_smem();                                  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pfull);
            }
            // ---- epilogue: O / l
            mbar_wait(bar_ofull, item_it & 1);
            tc_fence_after();
            const float inv = 1.0f / lsum;
            const int row = qt * QT + r;
            const int64_t grow = (int64_t)b * p.frames + row;
#pragma unroll 1
            for (int c = 0; c < DK / 32; ++c) {
                uint32_t o[32];
                tc_ld32(tmem_base + lane_addr + C::kOCol + c * 32, o);
                tc_wait_ld();
                if (row < p.frames) {
                    const int64_t off = grow * p.ldo + h * DK + c * 32;
                    if (p.ctx) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(p.ctx + off + 4 * i) =
                                make_float4(__uint_as_float(o[4 * i]) * inv, __uint_as_float(o[4 * i + 1]) * inv,
                                            __uint_as_float(o[4 * i + 2]) * inv, __uint_as_float(o[4 * i + 3]) * inv);
                    }
                    if (p.ctx_hi) {
                        uint32_t hw[16], lw[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            split_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv, hw[i], lw[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            reinterpret_cast<uint4*>(p.ctx_hi + off)[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
                            if (p.ctx_lo)
                                reinterpret_cast<uint4*>(p.ctx_lo + off)[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_oempty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

int g_att_sms = 0;

}  // namespace

void attention_tc_init_device() {
    FA_CUDA(cudaFuncSetAttribute(k_attention_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<128>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute(k_attention_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<64>::kSmemBytes));
    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    FA_CUDA(cudaDeviceGetAttribute(&g_att_sms, cudaDevAttrMultiProcessorCount, dev));
}

void launch_attention_tc(Planes qkv, int64_t plane_stride, int ld, int d_model, int batch, int frames, int heads, int dk,
                         const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st) {
    FA_REQUIRE(dk == 64 || dk == 128, "attention head width must be 64 or 128");
    FA_REQUIRE(heads * dk == d_model, "heads * d_k must equal the model width");
    FA_REQUIRE(ldo % 8 == 0 && ld % 8 == 0, "attention strides must be multiples of 8");
    const int rows = batch * frames;
    // one tensor per use: the q map fetches 128-row boxes, the k/v map 64-row boxes
    const TcOperand mq = tc_make_operand(qkv.hi, rows, ld, ld, plane_stride, 2, QT);
    const TcOperand mkv = tc_make_operand(qkv.hi, rows, ld, ld, plane_stride, 2, KT);
    AttnParams p{};
    p.batch = batch; p.frames = frames; p.heads = heads; p.d_model = d_model; p.ld = ld; p.kv_len = kv_len;
    p.ctx = ctx_f32; p.ctx_hi = ctx_pl.hi; p.ctx_lo = ctx_pl.lo; p.ldo = ldo;
    const int items = batch * heads * cdiv(frames, QT);
    const int grid = items < g_att_sms ? items : g_att_sms;
    prof_note_work(4.0 * batch * heads * (double)frames * frames * dk, 0.0);
    if (dk == 128) {
        FA_LAUNCH(k_attention_tc<128>, grid, kAttThreads, ACfg<128>::kSmemBytes, st, mq.map, mkv.map, p);
    } else {
        FA_LAUNCH(k_attention_tc<64>, grid, kAttThreads, ACfg<64>::kSmemBytes, st, mq.map, mkv.map, p);
    }
}

}  // namespace fa
