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
// No online rescale.  The softmax is invariant to the shift subtracted from the scores, so the shift only has to keep
// exp2 in range, and an item (segment, head, 128-query tile) runs in one of two modes:
//   fast   the shift of a row is the maximum of its exact scores over key tile 0 (the first 128 keys).  One pass:
//          S = three products, p = exp2(S - shift), O += P V.  A later key may beat that shift; the weights then
//          exceed 1, which fp32 carries up to 2^127.  A row whose sum of weights passes 2^60 marks the item ...
//   exact  ... and marked items are repeated after the CTA's other items (every role walks the same sequence, a
//          CTA-wide barrier in between) with the shift taken over ALL keys by a first pass that runs ONE product,
//          q_hi k_hi^T, and loads only the hi plane of K.  Tensor work per key tile: 1 + 3 + 3 units.
// Measured on one B200 (tools/attn_bench.py, 32 x 4 heads x 1001 frames): exact everywhere 200 us, fast 181 us; the
// repeat needs a key that outscores the first 128 by 60 in log2 units, which the parity cases never produce and a
// kernel test plants on purpose.
//
// Shapes follow what the tensor core was measured to sustain on B200 (tools/umma_bench.cu, cycles per
// M128 x N x K16 bf16 instruction): A from shared memory 83 / 97 / 161 for N = 64 / 128 / 256, A from
// tensor memory 65 / 72 / 136.  So both left operands live in TENSOR MEMORY (tcgen05.mma with A from
// TMEM) — the query tile is written there once per item, the softmax weights once per key tile,
// straight from the registers that produced them, IN PLACE over the scores they came from — and a key
// tile is 128 keys, so every instruction is N = 128.
//
// One CTA per SM, persistent over (segment, head, 128-query tile) items:
//   warp 0       TMA: one plane of a 128-key K or V tile per ring slot (128 KB of slots), in the order the MMA warp
//                consumes them; pass 1 loads the hi plane of K only
//   warp 1       TMEM allocator + MMA issuer (all lanes walk the loop, one elected lane issues):
//                S = Q K^T (M128 x N128) into two score tiles, O += P V (M128 x N d_k, V from smem MN-major:
//                no transpose).  The tensor pipe executes in issue order, which is what lets S(j+2) reuse
//                the tile P(j) lives in without a barrier: it is issued after P(j) V(j).
//   warps 2..17  softmax: 4 groups of 4 warps.  Every key tile is shared by all sixteen warps: group g owns keys
//                32 g .. 32 g + 31 of each tile (a warp can only touch its own quarter of the TMEM lanes, so the split
//                is by columns): thread = query row = TMEM lane; tcgen05.ld 32 scores, exp2 / sum, split p into hi/lo
//                planes, tcgen05.st over the same 32 columns.  (Two groups per tile, 64 keys per thread on alternate
//                tiles, has fewer hand-offs but twice the latency from "scores complete" to "weights written", which is
//                exposed at the head and the tail of every item: 188 -> 182 us per encoder launch.)  The groups exchange
//                row maxima once per item and their row sums are added by the epilogue.
//   warps 18..21 query loader (global -> registers -> TMEM, next item's Q while the current item finishes
//                its PV products) and output epilogue: O / l -> bf16 hi/lo planes, staged in shared memory as the
//                source tiles of bulk tensor stores, so that O is handed back after four TMEM loads and no warp
//                waits for global stores (fp32 rows, the kernel tests' other output form, go straight from registers)
// TMEM columns: score/weight tile 0 at 0, tile 1 at 128, O at 256, Q at 384 (hi d_k/2 | lo d_k/2).
// Inside a score tile, keys 32c .. 32c+31 become: hi pairs in columns 32c .. +15, lo pairs in 32c+16 .. +31.
#include "kernels.h"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace fa {

namespace {

constexpr int QT = 128;          // queries per item (UMMA M)
constexpr int KT = 128;          // keys per tile (UMMA N for S, K extent for PV)
constexpr int kAttThreads = 704;
// K/V ring: 128 KB of one-plane tiles (4 slots at d_k = 128, 8 at d_k = 64).  At d_k = 128 a ring of 4, 5 or 6 slots
// measured the same; the other 64 KB stage the output.  At d_k = 64 a key tile is half the tensor time, so the same ring
// in BYTES is what keeps the loads as far ahead in time (4 slots there: 236 -> 249 us).
constexpr int kRedoWords = 64;       // one bit per item of a CTA's sequence (2048 items per CTA)
constexpr bool kAttnTiming = false;   // tuning aid: set true, rebuild, run with FUNASR_B200_ATTN_TIMING=1 (per-cause wait cycles of the MMA warp)
constexpr bool kAttnTrace = false;    // tuning aid: clock64 stamps of CTA 0's roles per item (set true, rebuild, FUNASR_B200_ATTN_TIMING=1)

template <int DK> struct ACfg {
    static constexpr int kChunks = DK / 64;                  // 128-byte column chunks per head row
    static constexpr int kSlotBytes = kChunks * KT * 128;    // one plane of a 128-key tile of K (or V)
    static constexpr int kSlots = 4 * 32768 / kSlotBytes;
    static constexpr int kStageBytes = 4 * 16384;           // per loader warp: its 32 rows of O as bf16 hi/lo planes, 64-column boxes
    static constexpr int kSmemBytes = kSlots * kSlotBytes + 16 * QT * 4 /*row sums and maxima*/ + kStageBytes + kRedoWords * 4 + 1024 + 256;
    static constexpr uint32_t kTmemCols = 512;
    static constexpr uint32_t kSCol = 0, kOCol = 256, kQCol = 384;
    static constexpr uint32_t kQPlaneCols = DK / 2;          // packed bf16 pairs
};

// 2^x on the MUFU pipe (2 ulp); denormal results flush to zero, which is what a vanishing softmax
// weight should do.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct AttnParams {
    int batch, frames, heads, d_model, ld;     // ld = row stride (elements) of the qkv planes
    const __nv_bfloat16* q_hi;                 // plane pointers of the fused q|k|v matrix (q at column 0)
    const __nv_bfloat16* q_lo;
    const int* kv_len;
    float* ctx;
    __nv_bfloat16* ctx_hi;
    __nv_bfloat16* ctx_lo;
    int ldo;
    int force_exact;                           // FUNASR_B200_ATTENTION_SHIFT=exact: every item in the exact (two-pass) mode
    int items;                                 // (segment, head, query tile) items of the launch
    // packed rows (kernels.h Packing; null for the uniform [batch][frames] layout): one entry per 128-query tile
    // {first row of the segment, its length, tile index, segment}; item = tile * heads + head
    const int4* tile_tab;
    const float* key_bias;                     // [batch] added to the score of each segment's last key (Packing::last_key_bias)
    long long* dbg;                            // tuning aid (FUNASR_B200_ATTN_TIMING): cycles the MMA warp waits, by cause
};

// PK: packed rows (kernels.h Packing).  A template parameter so that the uniform-layout instance — every launch of a
// batch of equal-length segments — carries none of the packed path's code (item search, straddling output warps, the
// CTC head's key multiplicity).
template <int DK, bool PK>
__global__ void __launch_bounds__(kAttThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_out, AttnParams p) {
    using C = ACfg<DK>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t kv_base = (raw + 1023u) & ~1023u;
    const uint32_t l_base = kv_base + C::kSlots * C::kSlotBytes;              // float[2 item parities][4 groups][128]: row sums, then row maxima
    const uint32_t stage_base = l_base + 16 * QT * 4;                      // 1024-aligned: tensor-store source tiles (SWIZZLE_128B)
    const uint32_t redo_base = stage_base + C::kStageBytes;                      // uint32[kRedoWords]: items to repeat with the exact shift
    const uint32_t bars = redo_base + kRedoWords * 4;
    const uint32_t bar_kvfull = bars, bar_kvempty = bars + 8 * C::kSlots;   // [C::kSlots] each
    const uint32_t bar_sfull = bars + 16 * C::kSlots, bar_sempty = bar_sfull + 16;   // [2] each
    const uint32_t bar_pfull = bar_sempty + 16;                          // [2]
    const uint32_t bar_qfull = bar_pfull + 16, bar_qempty = bar_qfull + 8;   // hi plane of Q in TMEM (all pass 1 needs)
    const uint32_t bar_ofull = bar_qempty + 8, bar_oempty = bar_ofull + 8;
    const uint32_t bar_lfull = bar_oempty + 8;                           // [2]
    const uint32_t bar_qlofull = bar_lfull + 16;                         // lo plane of Q in TMEM (first needed by pass 2)
    const uint32_t tmem_slot = bar_qlofull + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    float* l_smem = reinterpret_cast<float*>(smem_raw + (l_base - raw));
    float* mx_smem = l_smem + 8 * QT;
    volatile uint32_t* redo_bits = reinterpret_cast<volatile uint32_t*>(smem_raw + (redo_base - raw));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q_tiles = (p.frames + QT - 1) / QT;
    const int items = p.items;

    if (threadIdx.x < kRedoWords) redo_bits[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kSlots; ++s) { mbar_init(bar_kvfull + 8 * s, 1); mbar_init(bar_kvempty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, 16);
            mbar_init(bar_pfull + 8 * s, 16);
            mbar_init(bar_lfull + 8 * s, 16);
        }
        mbar_init(bar_qfull, 4); mbar_init(bar_qlofull, 4); mbar_init(bar_qempty, 1);
        mbar_init(bar_ofull, 1); mbar_init(bar_oempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);    // provably warp-uniform
    grid_dependency_wait();                       // everything above overlapped the previous kernel's tail

    // item -> (segment b, head h, query tile qt); key tiles n; first row of the segment and its query count
    auto decode = [&](int item, int& b, int& h, int& qt, int& n, int& klen, int& row0, int& qlen) {
        if constexpr (PK) {
            const int t = item / p.heads;
            h = item - t * p.heads;
            const int4 e = __ldg(p.tile_tab + t);
            row0 = e.x; klen = e.y; qt = e.z; b = e.w;
            qlen = klen;
        } else {
            qt = item % q_tiles;
            const int bh = item / q_tiles;
            h = bh % p.heads;
            b = bh / p.heads;
            klen = p.kv_len ? p.kv_len[b] : p.frames;
            qlen = p.frames;
            row0 = b * p.frames;
        }
        n = (klen + KT - 1) / KT;
    };

    // A CTA's work is a sequence of (item, mode): first every item of its stride in the fast mode (the shift of a row
    // is the maximum of its exact scores over the FIRST key tile only; any shift gives the same softmax as long as
    // nothing overflows, and a row whose sum of weights exceeds 2^60 marks its item), then — after a CTA-wide barrier —
    // the marked items again in the exact mode (pass 1 over all keys for the shift).  Every role walks the same
    // sequence, so the barrier phases stay in step.
    // (written as a pair of macros rather than a lambda taking a lambda: the role bodies keep their state in
    // registers and are instantiated once)
    auto trace = [&](uint32_t it, int ev) {
        if (kAttnTrace && p.dbg && blockIdx.x == 0 && lane == 0 && it < 8) {
            long long t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
            p.dbg[it * 32 + ev] = t;
        }
    };
#define FA_ATT_SEQUENCE_BEGIN                                                                              \
    for (int phase = 0; phase < 2; ++phase) {                                                              \
        if (phase) asm volatile("bar.sync 2, %0;" ::"n"(kAttThreads) : "memory");                          \
        int k = 0;                                                                                         \
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {                                \
            if (phase && !((redo_bits[k >> 5] >> (k & 31)) & 1u)) continue;                                \
            const bool exact = phase != 0 || p.force_exact != 0;
#define FA_ATT_SEQUENCE_END \
        }                   \
    }

    if (warp == 0) {
        // ================================================================== TMA producer
        uint32_t cnt = 0;                             // slots produced so far: slot = cnt % kSlots
        // plane `pl` of the 128-key tile of K (which = 1) or V (which = 2) starting at `row`, into the next ring slot
        auto load_plane = [&](int which, int pl, int h, int row) {
            const uint32_t slot = cnt % C::kSlots, ph = (cnt / C::kSlots) & 1;
            mbar_wait(bar_kvempty + 8 * slot, ph ^ 1);
            const uint32_t full = bar_kvfull + 8 * slot, sb = kv_base + slot * C::kSlotBytes;
            if (elect_one()) {
                mbar_arrive_expect_tx(full, C::kSlotBytes);
#pragma unroll
                for (int c = 0; c < C::kChunks; ++c)
                    tma_load_3d(sb + c * KT * 128, &map_kv, full, which * p.d_model + h * DK + c * 64, row, pl);
            }
            __syncwarp();
            ++cnt;
        };
        FA_ATT_SEQUENCE_BEGIN
            int b, h, qt, n, klen, row0, qlen;
            decode(item, b, h, qt, n, klen, row0, qlen);
            if (exact)
                for (int j = 0; j < n; ++j) load_plane(1, 0, h, row0 + j * KT);                // pass 1: K hi
            for (int j = 0; j < 2 && j < n; ++j) {                                             // pass 2: K0, K1, then V(j), K(j+2)
                load_plane(1, 0, h, row0 + j * KT);
                load_plane(1, 1, h, row0 + j * KT);
            }
            for (int j = 0; j < n; ++j) {
                load_plane(2, 0, h, row0 + j * KT);
                load_plane(2, 1, h, row0 + j * KT);
                if (j + 2 < n) {
                    load_plane(1, 0, h, row0 + (j + 2) * KT);
                    load_plane(1, 1, h, row0 + (j + 2) * KT);
                }
            }
        FA_ATT_SEQUENCE_END
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        // The pipe queues only about four instructions ahead of the issuing thread (256 cycles of N = 128 work), so
        // nothing slow may sit between the last instruction of one group of products and the first of the next:
        // ring slots advance by increment-and-wrap, a wait that succeeds costs one instruction, and in the steady
        // state of pass 2 the barriers of the next group are polled in the middle of issuing the current one,
        // while the queue is full and the thread would be blocked anyway.
        constexpr uint32_t kIdescS = umma_idesc_bf16(QT, KT, false);
        constexpr uint32_t kIdescO = umma_idesc_bf16(QT, DK, true);
        const uint32_t tmem_q = tmem_base + C::kQCol, tmem_o = tmem_base + C::kOCol;
        uint32_t item_it = 0;
        uint32_t slot = 0, slot_ph = 0;                             // next ring slot and its fill parity
        uint32_t se0 = 0, se1 = 0, pf0 = 0, pf1 = 0;                // per score tile: waits so far on "scores consumed" / "weights ready"
        long long dbg_c[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tph = 0;
        auto tic = [&]() { if (kAttnTiming && p.dbg) tph = clock64(); };
        auto toc = [&](int cat) { if (kAttnTiming && p.dbg) dbg_c[cat] += clock64() - tph; };
        auto twait = [&](int cat, uint32_t bar, uint32_t parity) {           // mbar_wait, timed when p.dbg is set
            if (!kAttnTiming || !p.dbg) { mbar_wait(bar, parity); return; }
            const long long t0 = clock64();            // try_wait itself may block for a while: time all of it
            mbar_wait(bar, parity);
            dbg_c[cat] += clock64() - t0;
        };
        const long long dbg_t0 = kAttnTiming ? clock64() : 0;
        auto take = [&](uint32_t& s_out, uint32_t& ph_out) {
            s_out = slot; ph_out = slot_ph;
            if (++slot == C::kSlots) { slot = 0; slot_ph ^= 1; }
        };
        // one product Q(plane) K(plane)^T into the score tile: DK/16 instructions
        auto s_term = [&](uint32_t tile, uint32_t qcol, uint32_t kpl, uint32_t first_accum) {
#pragma unroll
            for (int ks = 0; ks < DK / 16; ++ks)
                tc_mma_ts(tile, qcol + ks * 8, umma_desc(kpl + (ks >> 2) * KT * 128 + (ks & 3) * 32, 16, 1024), kIdescS,
                          ks ? 1u : first_accum);
        };
        // one product P(plane) V(plane) into O: KT/16 instructions.  A = P [128 q][16 keys] from TMEM (hi pairs at +0,
        // lo pairs at +16 of each 32-column group); B = V [16 keys][DK] MN-major: 64-column chunks KT*128 B apart
        // (LBO), 8-key groups 1024 B apart (SBO)
        auto pv_term = [&](uint32_t tile, uint32_t p_off, uint32_t vpl, uint32_t first_accum) {
#pragma unroll
            for (int ks = 0; ks < KT / 16; ++ks)
                tc_mma_ts(tmem_o, tile + (ks >> 1) * 32 + p_off + (ks & 1) * 8, umma_desc(vpl + ks * 16 * 128, KT * 128, 1024), kIdescO,
                          ks ? 1u : first_accum);
        };
        FA_ATT_SEQUENCE_BEGIN
            int b, h, qt, n, klen, row0, qlen;
            decode(item, b, h, qt, n, klen, row0, qlen);
            trace(item_it, 0);
            twait(0, bar_qfull, item_it & 1);
            trace(item_it, 1);                       // this item's Q (hi plane) is in TMEM
            // ---- pass 1 (exact mode only): shift = max of the hi*hi scores.  Tiles 0, 1 follow the previous item's P V
            // in issue order; later tiles wait until the softmax group has read the scores they overwrite.
            for (int j = 0; exact && j < n; ++j) {
                const uint32_t buf = j & 1;
                uint32_t s0, p0;
                take(s0, p0);
                twait(1, bar_kvfull + 8 * s0, p0);
                if (j >= 2) { uint32_t& se = buf ? se1 : se0; twait(2, bar_sempty + 8 * buf, se & 1); ++se; }
                tc_fence_after();
                tic();
                if (elect_one()) {
                    s_term(tmem_base + C::kSCol + buf * KT, tmem_q, kv_base + s0 * C::kSlotBytes, 0u);
                    tc_commit(bar_sfull + 8 * buf);
                    tc_commit(bar_kvempty + 8 * s0);
                }
                __syncwarp();
                toc(12);
            }
            // ---- pass 2.  S(0), S(1) wait for the last pass-1 scores of their tile to be read (exact mode; in the fast
            // mode they follow the previous item's P V in issue order); after that S(j+2) follows P(j) V(j) in issue
            // order and needs no barrier of its own.
            uint32_t ka, kpa, kb, kpb;                              // ring slots of the S about to be issued (K hi, K lo)
            twait(15, bar_qlofull, item_it & 1);                    // the lo plane of Q landed during pass 1
            trace(item_it, 2);
            for (int j = 0; j < 2 && j < n; ++j) {
                const uint32_t buf = j;
                take(ka, kpa); take(kb, kpb);
                twait(3, bar_kvfull + 8 * ka, kpa);
                twait(3, bar_kvfull + 8 * kb, kpb);
                if (exact) { uint32_t& se = buf ? se1 : se0; twait(4, bar_sempty + 8 * buf, se & 1); ++se; }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t tile = tmem_base + C::kSCol + buf * KT;
                    s_term(tile, tmem_q + C::kQPlaneCols, kv_base + ka * C::kSlotBytes, 0u);
                    s_term(tile, tmem_q, kv_base + kb * C::kSlotBytes, 1u);
                    s_term(tile, tmem_q, kv_base + ka * C::kSlotBytes, 1u);
                    tc_commit(bar_sfull + 8 * buf);
                    tc_commit(bar_kvempty + 8 * ka);
                    tc_commit(bar_kvempty + 8 * kb);
                    if (j == n - 1) tc_commit(bar_qempty);          // last read of Q: the next item's Q may land
                }
                __syncwarp();
            }
            // barriers of P(0) V(0): V hi, V lo, the weights, a drained O
            uint32_t va, vpa, vb, vpb;
            take(va, vpa); take(vb, vpb);
            trace(item_it, 3);
            twait(5, bar_oempty, (item_it & 1) ^ 1);
            trace(item_it, 4);
            twait(6, bar_kvfull + 8 * va, vpa);
            twait(6, bar_kvfull + 8 * vb, vpb);
            trace(item_it, 5);
            twait(7, bar_pfull, pf0 & 1);
            trace(item_it, 6);
            ++pf0;
            for (int j = 0; j < n; ++j) {
                const uint32_t buf = j & 1;
                const uint32_t tile = tmem_base + C::kSCol + buf * KT;
                const bool s_next = j + 2 < n, pv_next = j + 1 < n;
                // ---- P(j) V(j); in its middle, poll the K slots of S(j+2)
                bool k_ready = true;
                if (s_next) { take(ka, kpa); take(kb, kpb); }
                tc_fence_after();
                tic();
                if (elect_one()) pv_term(tile, 16, kv_base + va * C::kSlotBytes, j ? 1u : 0u);
                __syncwarp();
                if (s_next) k_ready = mbar_test(bar_kvfull + 8 * ka, kpa) & mbar_test(bar_kvfull + 8 * kb, kpb);
                if (elect_one()) {
                    pv_term(tile, 0, kv_base + vb * C::kSlotBytes, 1u);
                    pv_term(tile, 0, kv_base + va * C::kSlotBytes, 1u);
                    tc_commit(bar_kvempty + 8 * va);
                    tc_commit(bar_kvempty + 8 * vb);
                    if (j == n - 1) tc_commit(bar_ofull);
                }
                __syncwarp();
                toc(13);
                // ---- S(j+2) over the tile P(j) was in; in its middle, poll the barriers of P(j+1) V(j+1)
                bool v_ready = true;
                const uint32_t nbuf = buf ^ 1;
                uint32_t& pfn = nbuf ? pf1 : pf0;
                if (pv_next) { take(va, vpa); take(vb, vpb); }
                if (s_next) {
                    if (!k_ready) { twait(8, bar_kvfull + 8 * ka, kpa); twait(8, bar_kvfull + 8 * kb, kpb); }
                    tc_fence_after();
                    tic();
                    if (elect_one()) s_term(tile, tmem_q + C::kQPlaneCols, kv_base + ka * C::kSlotBytes, 0u);
                    __syncwarp();
                    v_ready = mbar_test(bar_kvfull + 8 * va, vpa) & mbar_test(bar_kvfull + 8 * vb, vpb) &
                              mbar_test(bar_pfull + 8 * nbuf, pfn & 1);
                    if (elect_one()) {
                        s_term(tile, tmem_q, kv_base + kb * C::kSlotBytes, 1u);
                        s_term(tile, tmem_q, kv_base + ka * C::kSlotBytes, 1u);
                        tc_commit(bar_sfull + 8 * buf);
                        tc_commit(bar_kvempty + 8 * ka);
                        tc_commit(bar_kvempty + 8 * kb);
                        if (j + 3 == n) tc_commit(bar_qempty);
                    }
                    __syncwarp();
                    toc(14);
                } else {
                    v_ready = false;
                }
                if (pv_next) {
                    if (!v_ready) {
                        twait(9, bar_kvfull + 8 * va, vpa);
                        twait(9, bar_kvfull + 8 * vb, vpb);
                        twait(10, bar_pfull + 8 * nbuf, pfn & 1);
                    }
                    ++pfn;
                }
            }
            trace(item_it, 7);
            ++item_it;
        FA_ATT_SEQUENCE_END
        if (kAttnTiming && p.dbg && lane == 0) {
            dbg_c[11] = clock64() - dbg_t0;
            for (int i = 0; i < 16; ++i) p.dbg[blockIdx.x * 16 + i] = dbg_c[i];
        }
    } else if (warp < 18) {
        // ================================================================== softmax (4 groups of 4 warps)
        // Every key tile is shared by all sixteen warps: group g4 owns keys 32*g4 .. +31 of each tile, so the latency from
        // "scores complete" to "weights written" is that of 32 columns per thread on four warps per SM sub-partition.
        const int g4 = (warp - 2) >> 2;                             // group 0..3
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                          // query row in the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t tmem_s = tmem_base + lane_addr + C::kSCol + g4 * 32;     // + (j & 1) * KT for key tile j
        uint32_t item_it = 0, su0 = 0, su1 = 0;                      // waits so far on "scores ready" of score tile 0 / 1
        // row maximum of this thread's 32 keys of the score tile at `ts` (keys kbase .. kbase+31, those < klen)
        auto tile_max = [&](uint32_t ts, int kbase, int klen, float mx) {
            uint32_t s[32];
            tc_ld32(ts, s);
            tc_wait_ld();
            if (kbase + 32 <= klen) {
                float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};      // four short chains instead of one of 16
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        m4[a] = fmaxf(m4[a], fmaxf(__uint_as_float(s[i + 2 * a]), __uint_as_float(s[i + 2 * a + 1])));
                }
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (kbase + i < klen) mx = fmaxf(mx, __uint_as_float(s[i]));
            }
            return mx;
        };
        FA_ATT_SEQUENCE_BEGIN
            int b, h, qt, n, klen, row0, qlen;
            decode(item, b, h, qt, n, klen, row0, qlen);
            float mx = -INFINITY;
            float* mxb = mx_smem + (item_it & 1) * 4 * QT;           // buffers alternate by item parity
            if (warp == 2) trace(item_it, 16);
            if (!exact) {
                // ---- fast mode: the shift is the row maximum over key tile 0, taken from the exact scores pass 2 is
                // about to turn into weights
                mbar_wait(bar_sfull, su0 & 1);                       // S(0); the pass-2 loop waits on it again (at once)
                if (warp == 2) trace(item_it, 17);
                tc_fence_after();
                mx = tile_max(tmem_s, g4 * 32, klen, mx);
                if (warp == 2) trace(item_it, 18);
                mxb[g4 * QT + r] = mx;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                mx = fmaxf(fmaxf(mxb[r], mxb[QT + r]), fmaxf(mxb[2 * QT + r], mxb[3 * QT + r]));
                if (warp == 2) trace(item_it, 19);
            }
            // ---- pass 1 (exact mode)
            for (int j = 0; exact && j < n; ++j) {
                uint32_t& su = (j & 1) ? su1 : su0;
                mbar_wait(bar_sfull + 8 * (j & 1), su & 1);
                ++su;
                tc_fence_after();
                mx = tile_max(tmem_s + (j & 1) * KT, j * KT + g4 * 32, klen, mx);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sempty + 8 * (j & 1));    // these scores may be overwritten
            }
            if (exact) {
                // the four groups saw disjoint keys: exchange the row maxima
                mxb[g4 * QT + r] = mx;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                mx = fmaxf(fmaxf(mxb[r], mxb[QT + r]), fmaxf(mxb[2 * QT + r], mxb[3 * QT + r]));
            }
            // ---- pass 2: scores -> weights in place
            float lsum = 0.f;
            for (int j = 0; j < n; ++j) {
                uint32_t& su = (j & 1) ? su1 : su0;
                mbar_wait(bar_sfull + 8 * (j & 1), su & 1);
                ++su;
                tc_fence_after();
                const int kbase = j * KT + g4 * 32;
                const uint32_t ts = tmem_s + (j & 1) * KT;
                {
                    uint32_t s[32];
                    tc_ld32(ts, s);
                    tc_wait_ld();
                    if (PK && p.key_bias && j == n - 1 && klen - 1 >= kbase && klen - 1 < kbase + 32) {   // warp-uniform, once per item
                        // the segment's last key stands for n_pad identical keys: weight n_pad * exp2(s) = exp2(s + log2 n_pad)
                        const int at = klen - 1 - kbase;
                        const float kb = p.key_bias[b];
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == at) s[i] = __float_as_uint(__uint_as_float(s[i]) + kb);
                    }
                    const bool full_chunk = kbase + 32 <= klen;
                    uint32_t hi[16], lo[16];                         // 32 keys x bf16, packed in pairs (even key in the low half)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = fast_exp2(__uint_as_float(s[2 * i]) - mx);
                        float p1 = fast_exp2(__uint_as_float(s[2 * i + 1]) - mx);
                        if (!full_chunk) {
                            if (kbase + 2 * i >= klen) p0 = 0.f;
                            if (kbase + 2 * i + 1 >= klen) p1 = 0.f;
                        }
                        lsum += p0 + p1;
                        split_bf16x2(p0, p1, hi[i], lo[i]);
                    }
                    tc_st16(ts, hi);
                    tc_st16(ts + 16, lo);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pfull + 8 * (j & 1));
                if (warp == 2 && j < 2) trace(item_it, 20);
            }
            // ---- hand this group's row sums to the epilogue warps
            l_smem[((item_it & 1) * 4 + g4) * QT + r] = lsum;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_lfull + 8 * (item_it & 1));
            ++item_it;
        FA_ATT_SEQUENCE_END
    } else {
        // ================================================================== query loader + output epilogue
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        // Q rows of one item, global -> registers -> TMEM columns kQCol..: word c of a plane = elements 2c, 2c+1.
        // The hi plane of the NEXT item is fetched into registers before the current item's last S product has
        // retired (the loads do not touch TMEM), so that once Q's columns are free only the stores are left before
        // pass 1 can start; the lo plane, which pass 2 needs some thousand cycles later, follows.
        constexpr int kQW = DK / 2;                                 // 32-bit words per row per plane
        auto q_src = [&](int item, int pl, bool& ok) -> const uint4* {
            int b, h, qt, n, klen, row0, qlen;
            decode(item, b, h, qt, n, klen, row0, qlen);
            const int row = qt * QT + r;
            ok = row < qlen;
            const int64_t off = ((int64_t)row0 + row) * p.ld + h * DK;
            return reinterpret_cast<const uint4*>((pl ? p.q_lo : p.q_hi) + off);
        };
        auto q_fetch = [&](int item, int pl, uint32_t (&w)[kQW]) {
            bool ok;
            const uint4* src = q_src(item, pl, ok);
#pragma unroll
            for (int i = 0; i < kQW / 4; ++i) {
                const uint4 v = ok ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        };
        auto q_store = [&](int pl, const uint32_t (&w)[kQW], uint32_t bar) {
#pragma unroll
            for (int g = 0; g < kQW / 32; ++g) {
                tc_st32(tmem_base + lane_addr + C::kQCol + pl * C::kQPlaneCols + g * 32,
                        *reinterpret_cast<const uint32_t(*)[32]>(&w[g * 32]));
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        uint32_t item_it = 0;
        uint32_t qw[kQW];
        // this role looks one item ahead, so it walks the two phases of the sequence (see FA_ATT_SEQUENCE_BEGIN) by hand
        for (int second = 0; second < 2; ++second) {
            if (second) { bulk_wait_all0(); asm volatile("bar.sync 2, %0;" ::"n"(kAttThreads) : "memory"); }   // a repeated item rewrites its rows
            const bool exact = second || p.force_exact != 0;
            auto in_phase = [&](int kk) { return !second || (((redo_bits[kk >> 5] >> (kk & 31)) & 1u) != 0u); };
            auto advance = [&](int& it, int& kk) { do { it += gridDim.x; ++kk; } while (it < items && !in_phase(kk)); };
            int item = blockIdx.x, k = 0;
            if (item < items && !in_phase(0)) advance(item, k);
            if (item >= items) continue;
            // Q of the phase's first item: nothing is reading Q's columns (kernel start, or every role is past phase 0)
            q_fetch(item, 0, qw);
            q_store(0, qw, bar_qfull);
            q_fetch(item, 1, qw);
            q_store(1, qw, bar_qlofull);
            while (item < items) {
            int b, h, qt, n, klen, seg_row0, qlen;
            decode(item, b, h, qt, n, klen, seg_row0, qlen);
            int next = item, next_k = k;
            advance(next, next_k);
            if (next < items) {
                q_fetch(next, 0, qw);                               // in flight while this item's S products finish
                mbar_wait(bar_qempty, item_it & 1);                 // every S product of this item has retired
                if (warp == 18) trace(item_it, 8);
                tc_fence_after();
                q_store(0, qw, bar_qfull);
                if (warp == 18) trace(item_it, 9);
                q_fetch(next, 1, qw);
                q_store(1, qw, bar_qlofull);
                if (warp == 18) trace(item_it, 10);
            }
            // ---- epilogue: O / l
            mbar_wait(bar_ofull, item_it & 1);
            if (warp == 18) trace(item_it, 11);
            mbar_wait(bar_lfull + 8 * (item_it & 1), (item_it >> 1) & 1);
            tc_fence_after();
            const float* lb = l_smem + (item_it & 1) * 4 * QT + r;
            const float l_row = (lb[0] + lb[QT]) + (lb[2 * QT] + lb[3 * QT]);
            const float inv = 1.0f / l_row;
            // fast mode: a weight above 2^60 means a key outside tile 0 beat the shift by more than 60 (log2 units); the
            // sums may then have overflowed, so the item is repeated with the exact shift (its output is overwritten)
            if (!exact && __any_sync(0xffffffffu, !(l_row <= 0x1p60f)) && lane == 0)
                atomicOr(const_cast<uint32_t*>(redo_bits) + (k >> 5), 1u << (k & 31));
            const int row = qt * QT + r;
            const int64_t grow = (int64_t)seg_row0 + row;
            // O leaves TMEM 64 columns at a time; the accumulator is handed back to the MMA warp as soon as the last
            // columns are in registers, before they are scaled, split and stored (the next item's first P V is only
            // some three thousand cycles behind this item's last one)
            auto emit = [&](const uint32_t (&o)[32], int c) {
                if (row >= qlen) return;
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
            };
            if (p.ctx_hi && !p.ctx) {
                // The engine's form of the output: bf16 hi/lo planes.  The warp turns its 32 rows of O into planes 32
                // columns at a time and parks them in shared memory, laid out as the source tiles of bulk tensor stores
                // (64-column boxes, 128-byte rows, SWIZZLE_128B); O goes back to the MMA warp as soon as the last columns
                // are in registers, and no warp ever waits for a global store — the stores of all CTAs arrive at the
                // memory system in the same microsecond (the items of a launch run in lockstep) and used to hold the
                // hand-back for ~6 k cycles of every item.
                unsigned char* stg = smem_raw + (stage_base - raw) + (warp - 18) * 16384;
                const uint32_t stg_u32 = stage_base + (warp - 18) * 16384;
                // packed rows: the store map cannot clip at the end of a segment (the next segment's rows follow), so a
                // warp whose 32 rows straddle the end writes its valid rows with plain stores (once per segment and head)
                const int wrow0 = qt * QT + quarter * 32;
                const bool direct = PK && wrow0 + 32 > qlen;
                if (lane == 0) bulk_wait_read0();                    // the previous item's stores have read the tiles (an item ago)
                __syncwarp();
#pragma unroll 1
                for (int c = 0; c < DK / 32; ++c) {
                    uint32_t o[32];
                    tc_ld32(tmem_base + lane_addr + C::kOCol + c * 32, o);
                    tc_wait_ld();
                    if (c + 1 == DK / 32) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_oempty);
                        if (warp == 18) trace(item_it, 13);
                    }
                    uint32_t hw[16], lw[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        split_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv, hw[i], lw[i]);
                    if (direct) {
                        if (row < qlen) {
                            const int64_t off = grow * p.ldo + h * DK + c * 32;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                reinterpret_cast<uint4*>(p.ctx_hi + off)[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
                                if (p.ctx_lo)
                                    reinterpret_cast<uint4*>(p.ctx_lo + off)[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
                            }
                        }
                        continue;
                    }
                    unsigned char* box = stg + (c >> 1) * 8192;          // hi tile, then lo tile, of this 64-column box
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int off = lane * 128 + (((4 * (c & 1) + i) ^ (lane & 7)) << 4);
                        *reinterpret_cast<uint4*>(box + off) = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
                        *reinterpret_cast<uint4*>(box + 4096 + off) = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0 && !direct && wrow0 < qlen) {          // uniform layout: rows past the segment's end are clipped by the map
                    const int mrow = PK ? seg_row0 + wrow0 : wrow0, mseg = PK ? 0 : b;
#pragma unroll
                    for (int bx = 0; bx < DK / 64; ++bx)     // one store per 64-column box: the map's box spans both planes (hi tile, lo tile)
                        tma_store_4d(&map_out, stg_u32 + bx * 8192, h * DK + bx * 64, mrow, mseg, 0);
                    bulk_commit();
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < DK / 32; c += 2) {
                uint32_t o0[32], o1[32];
                tc_ld32(tmem_base + lane_addr + C::kOCol + c * 32, o0);
                tc_ld32(tmem_base + lane_addr + C::kOCol + c * 32 + 32, o1);
                tc_wait_ld();
                if (c + 2 == DK / 32) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_oempty);
                    if (warp == 18) trace(item_it, 13);
                }
                emit(o0, c);
                emit(o1, c + 1);
            }
            }
            if (warp == 18) trace(item_it, 14);
            ++item_it;
            item = next;
            k = next_k;
            }
        }
        bulk_wait_all0();                                           // this thread's bulk stores have been written
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

#undef FA_ATT_SEQUENCE_BEGIN
#undef FA_ATT_SEQUENCE_END

int g_att_sms = 0;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_att_encode = nullptr;
std::once_flag g_att_once;

// Store map of the output planes: {model column, frame of the segment, segment, plane}, boxes of 64 columns x 32 rows.
// Frames past the end of a segment are clipped by the map.
CUtensorMap att_out_map(Planes out, int ldo, int d_model, int frames, int batch) {
    FA_REQUIRE(g_att_encode != nullptr, "attention_tc_init_device() has not run");
    FA_REQUIRE((reinterpret_cast<uintptr_t>(out.hi) & 15) == 0, "TMA base must be 16-byte aligned");
    const int64_t plane_stride = out.lo ? (out.lo - out.hi) : (int64_t)batch * frames * ldo;
    FA_REQUIRE(plane_stride > 0 && plane_stride % 8 == 0, "attention output planes: lo must follow hi at a multiple of 8 elements");
    const cuuint64_t dims[4] = {(cuuint64_t)d_model, (cuuint64_t)frames, (cuuint64_t)batch, (cuuint64_t)(out.lo ? 2 : 1)};
    const cuuint64_t strides[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)frames * ldo * 2, (cuuint64_t)plane_stride * 2};
    const cuuint32_t box[4] = {64, 32, 1, (cuuint32_t)(out.lo ? 2 : 1)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = g_att_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out.hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled (attention output) failed with code " + std::to_string((int)r));
    return m;
}

}  // namespace

void attention_tc_init_device() {
    std::call_once(g_att_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        FA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        FA_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_att_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    FA_CUDA(cudaFuncSetAttribute((k_attention_tc<128, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<128>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute((k_attention_tc<64, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<64>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute((k_attention_tc<128, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<128>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute((k_attention_tc<64, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<64>::kSmemBytes));
    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    FA_CUDA(cudaDeviceGetAttribute(&g_att_sms, cudaDevAttrMultiProcessorCount, dev));
}

void launch_attention_tc(Planes qkv, int64_t plane_stride, int ld, int d_model, int batch, int frames, int heads, int dk,
                         const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st, const Packing* pk) {
    FA_REQUIRE(dk == 64 || dk == 128, "attention head width must be 64 or 128");
    FA_REQUIRE(!pk || kv_len, "packed attention needs the per-segment lengths");
    FA_REQUIRE(heads * dk == d_model, "heads * d_k must equal the model width");
    FA_REQUIRE(ldo % 8 == 0 && ld % 8 == 0, "attention strides must be multiples of 8");
    FA_REQUIRE(qkv.lo == qkv.hi + plane_stride, "attention expects the lo plane `plane_stride` elements after the hi plane");
    const int rows = pk ? pk->total_rows : batch * frames;        // rows past the end are zero-filled by the load map
    const TcOperand mkv = tc_make_operand(qkv.hi, rows, ld, ld, plane_stride, 2, KT);     // 128-row boxes of k and v
    AttnParams p{};
    p.batch = batch; p.frames = frames; p.heads = heads; p.d_model = d_model; p.ld = ld; p.kv_len = kv_len;
    p.q_hi = qkv.hi; p.q_lo = qkv.lo;
    p.ctx = ctx_f32; p.ctx_hi = ctx_pl.hi; p.ctx_lo = ctx_pl.lo; p.ldo = ldo;
    {
        const char* sh = getenv("FUNASR_B200_ATTENTION_SHIFT");      // comparison aid, read at every launch
        p.force_exact = (sh && !strcmp(sh, "exact")) ? 1 : 0;
    }
    CUtensorMap map_out;
    memset(&map_out, 0, sizeof(map_out));
    if (ctx_pl.hi && !ctx_f32) map_out = pk ? att_out_map(ctx_pl, ldo, d_model, pk->total_rows, 1) : att_out_map(ctx_pl, ldo, d_model, frames, batch);
    const int items = pk ? heads * pk->total_tiles : batch * heads * cdiv(frames, QT);
    p.items = items;
    if (pk) { p.tile_tab = pk->tile_tab; p.key_bias = pk->last_key_bias; }
    int grid = items < g_att_sms ? items : g_att_sms;
    if (const char* e = getenv("FUNASR_B200_ATT_SMS")) { const int v = atoi(e); if (v > 0 && v < grid) grid = v; }   // tuning aid: fewer CTAs
    // the repeat bitmap holds one bit per item of a CTA's sequence; more would spill into the barriers behind it
    FA_REQUIRE(cdiv(items, grid) <= 32 * kRedoWords, "attention: too many items per CTA for the repeat bitmap");
    static const bool timing = (kAttnTiming || kAttnTrace) && getenv("FUNASR_B200_ATTN_TIMING") != nullptr;
    long long* dbg = nullptr;
    if (timing) {
        FA_CUDA(cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(long long)));
        FA_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(long long), st));
        p.dbg = dbg;
    }
    prof_note_work(pk ? 4.0 * heads * pk->sum_len_sq * dk : 4.0 * batch * heads * (double)frames * frames * dk, 0.0);
    if (g_prof_on) prof_note_tag(dk == 128 ? "dk128" : "dk64");
    if (dk == 128) {
        if (pk) FA_LAUNCH((k_attention_tc<128, true>), grid, kAttThreads, ACfg<128>::kSmemBytes, st, mkv.map, map_out, p);
        else FA_LAUNCH((k_attention_tc<128, false>), grid, kAttThreads, ACfg<128>::kSmemBytes, st, mkv.map, map_out, p);
    } else {
        if (pk) FA_LAUNCH((k_attention_tc<64, true>), grid, kAttThreads, ACfg<64>::kSmemBytes, st, mkv.map, map_out, p);
        else FA_LAUNCH((k_attention_tc<64, false>), grid, kAttThreads, ACfg<64>::kSmemBytes, st, mkv.map, map_out, p);
    }
    if (timing) {       // tuning aid only: synchronises
        std::vector<long long> h((size_t)grid * 16);
        FA_CUDA(cudaStreamSynchronize(st));
        FA_CUDA(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        FA_CUDA(cudaFree(dbg));
        if (kAttnTrace) {   // events: MMA warp 0 item start, 1 Q hi, 2 Q lo, 3 S(0) S(1) issued, 4 O drained, 5 first V, 6 first P, 7 item issued;
                            // loader 8 Q free, 9/10 next Q hi/lo stored, 11 O complete, 13 O handed back, 14 output stored;
                            // softmax warp 2: 16 item start, 17 S(0) seen, 18 tile maximum, 19 shift exchanged, 20 P(0) written
            for (int it = 0; it < 8; ++it) {
                fprintf(stderr, "trace item %d:", it);
                for (int e = 0; e < 32; ++e) fprintf(stderr, " %d=%lld", e, h[it * 32 + e] ? h[it * 32 + e] - h[0] : -1LL);
                fprintf(stderr, "\n");
            }
            return;
        }
        double c[16] = {0};
        for (int b = 0; b < grid; ++b)
            for (int i = 0; i < 16; ++i) c[i] += (double)h[(size_t)b * 16 + i] / grid;
        static const char* names[16] = {"qfull", "p1 K", "p1 sempty", "p2 S01 K", "p2 S01 sempty", "oempty", "first V", "first P",
                                        "loop K", "loop V", "loop P", "total", "issue p1 S", "issue PV", "issue S", "qlo"};
        fprintf(stderr, "attention MMA-warp waits (mean cycles per CTA, %d items over %d CTAs, dk %d):", items, grid, dk);
        for (int i = 0; i < 16; ++i) fprintf(stderr, " %s=%.0f", names[i], c[i]);
        fprintf(stderr, "\n");
    }
}

}  // namespace fa
