// tcgen05 / TMEM / TMA GEMM for every dense projection on the path (SURVEY §8a rows a6, a9, a11,
// a13, a14):  C[M][N] = A[M][K] * W[N][K]^T + bias, with ReLU / residual / bf16-plane / fused
// vocabulary-argmax epilogues.
//
// Precision.  The reference computes these nn.Linear layers in fp32 and the contract is token-exact
// CTC ids, so operands are carried as bf16 hi/lo planes (v = hi + lo, 16 mantissa bits) and each
// K-block issues three MMA groups into the same fp32 TMEM accumulator:
//     A_lo*W_hi + A_hi*W_lo + A_hi*W_hi          ("bf16x3", NP = 2)
// The planes of one K-block are loaded once (4 tiles) and reused by the 3 groups, so the kernel
// moves 4/3 of the bytes of a plain bf16 GEMM per MMA, not 2x.  NP = 1 is the plain bf16 mode.
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps {K, rows, plane}, SWIZZLE_128B) -> smem ring
//   warp 1      TMEM allocator + MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M128 x N256 x K16
//   warps 2..9  epilogue: tcgen05.ld 32x32b -> registers -> bias/ReLU/residual -> swizzled smem transpose ->
//               row-contiguous global stores (fp32 / bf16 planes), or a running argmax for the vocabulary
// Two 256-column accumulators (all 512 TMEM columns) are ping-ponged so the epilogue of tile i
// overlaps the main loop of tile i+1.  Pipelines: smem full/empty mbarriers (TMA <-> MMA) and
// TMEM full/empty mbarriers (MMA <-> epilogue); tcgen05.commit signals both.
//
// k_gemm_tc2 is the product kernel: the same roles on a CTA PAIR (cluster of 2, tcgen05 cta_group::2).
// A pair owns a 256 x 256 output tile; each CTA loads its own 128 rows of A and its own half
// (128 rows) of the W tile, the leader CTA issues M256 x N256 x K16 MMAs that read both CTAs'
// shared memory, and each CTA's TMEM receives its 128 rows of the result.  Per SM that is 64 KB per
// K-block instead of 96 KB (less L2->SM traffic and operand fetch per MMA cycle), which also makes room
// for a third stage.  The tail of the tile list (the last, partial wave) is cut into 128-column half
// tiles so that a GEMM of 3.4 waves costs 3.5, not 4.  k_gemm_tc (one CTA per tile) is kept as the
// comparison kernel (FUNASR_B200_GEMM=1cta).
#include "kernels.h"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace fa {

namespace {

constexpr int BM = kTcBlockM, BN = kTcBlockN, BK = kTcBlockK;
constexpr int kATileBytes = BM * BK * 2;         // 16 KB
constexpr int kThreads = 320;                   // TMA warp, MMA warp, 8 epilogue warps
constexpr int kStagingBytes = 8 * 4096;         // one 4 KB transpose tile per epilogue warp
constexpr uint32_t kTmemCols = 512;

// TN: columns of a tile of the single-CTA kernel.  256 is the throughput shape; 128 and 64 spread a small M (single-
// segment calls: 8 row tiles) over more SMs and shorten the K loop, which is what the latency of such a call is made of.
template <int NP, int TN = BN> struct Cfg {
    static constexpr int kWBytes = TN * BK * 2;                              // 32 / 16 / 8 KB
    static constexpr int kStageBytes = NP * (kATileBytes + kWBytes);        // TN = 256: 96 KB (NP=2) / 48 KB (NP=1)
    static constexpr int kStages = 192 * 1024 / kStageBytes > 6 ? 6 : 192 * 1024 / kStageBytes;     // 2 / 4 at TN = 256
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// CTA-pair kernel: per CTA and stage, NP planes of (A 128 x 64 + W-half up to 128 x 64)
// F8: the e4m3 kernel's main loop is 2 k cycles per K = 512 tile, so its throughput is the epilogue's: SIXTEEN epilogue warps
// (two 32-column chunks each instead of four: half the dependent chain per warp, twice the loads and stores in flight),
// paid for with one operand stage (5 x 32 KB + 16 x 4 KB staging tiles)
template <int NP, bool F8 = false> struct Cfg2 {
    static constexpr int kHalfWBytes = (BN / 2) * BK * 2;                    // 16 KB: this CTA's 128 rows of the W tile
    static constexpr int kStageBytes = NP * (kATileBytes + kHalfWBytes);      // 64 KB (NP=2) / 32 KB (NP=1)
    static constexpr int kStages = NP == 2 ? 3 : (F8 ? 5 : 6);
    static constexpr int kEpiWarps = F8 ? 16 : 8;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kStagingBytes = kEpiWarps * 4096;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Work list of the pair kernel: `full_units` 256-column tiles, then the remaining tiles cut into
// `1 << split_log2` column parts each (the partial last wave).
struct Sched {
    int m_tiles, n_tiles, band;      // in 256 x 256 pair tiles
    int full_units, total_units, split_log2;
};

struct EpiParams {
    const float* bias;
    const float* resid;
    float* out;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    float* amax_val;
    int32_t* amax_idx;
    int32_t* cand_run_max;                       // vocabulary candidate lists (kernels.h: VocabCand)
    int32_t* cand_count;
    int2* cand_list;
    const float* cand_bound2;
    int cand_cap;
    const int32_t* gate;                         // launch does nothing when *gate == 0
    const float* col_scale;                      // fp8 mode: per-output-channel weight scale, applied to the accumulator before the bias
    uint8_t* out_f8;                             // fp8 mode: e4m3 output [M][ld8] (the next projection's A operand)
    int ld8;
    int ldr, ldc, ldp, relu, n_slots;            // n_slots: 128-column groups of N (fused argmax partials)
    float pl_col_scale;
    int pl_col_scale_end, f32_col_begin;
    int reduce_add;                              // the residual IS the fp32 output (x += ...): added by a tensor reduction, never read
    long long* dbg;                              // tuning aid (-DFA_GEMM_TIMING, FUNASR_B200_GEMM_TIMING=1): per-phase cycles
};

#ifdef FA_GEMM_TIMING
#define FA_GT(...) __VA_ARGS__
#else
#define FA_GT(...)
#endif

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart.
// Field layout as in cute::UMMA::SmemDescriptor (start>>4 @0, LBO>>4 @16, SBO>>4 @32, version=1 @46, layout @61).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 @17, M>>4 @24 (built per kernel).

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int band, int& mt, int& nt) {
    // bands of `band` m-tiles; inside a band the n index is outermost so that the CTAs running together
    // share a handful of W tiles and one A band, both L2-resident.
    const int per_band = band * n_tiles;
    const int b = tile / per_band, r = tile - b * per_band;
    const int rows = min(band, m_tiles - b * band);
    nt = r / rows;
    mt = b * band + (r - nt * rows);
}

// order-preserving int encoding of a float (atomicMax on signed ints); same as rowops.cu
__device__ __forceinline__ int32_t f2ord_dev(float f) {
    const int32_t b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f_dev(int32_t o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

// Epilogue of one output unit: rows row0..row0+127 (this CTA's TMEM lanes), columns n0..n0+nw-1 held
// in TMEM columns tmem_acc..tmem_acc+nw-1.  Run by the 8 epilogue warps of a CTA.
// Warp e owns TMEM lanes 32*(e&3).. (the hardware's lane-quarter rule: warp id % 4) and column chunks
// 4*(e>>2) .. +3 of the accumulator.  A thread holds one row of a 32-column chunk in registers;
// everything that touches global memory goes through a 4 KB per-warp staging tile (XOR-swizzled,
// conflict-free both ways) so that loads and stores are row-contiguous: 8 lanes cover one 128-byte
// line instead of 32 lanes touching 32 different lines.
template <int kEpiWarps = 8>
__device__ __forceinline__ void epilogue_unit(const EpiParams& ep, const CUtensorMap* map_out, const CUtensorMap* map_pl,
                                              unsigned char* stg, bool& store_pending, int e, int lane, int m, int n, int row0,
                                              int n0, int nw, uint32_t tmem_acc, uint32_t bar_ready, uint32_t ready_parity) {
    // epilogue warps are warps 2..: warp & 3 == (e + 2) & 3 picks the TMEM lane quarter; e >> 2 the group of column chunks
    constexpr int kCpw = 32 / kEpiWarps;                         // 32-column chunks per warp: 4 (8 warps) or 2 (16 warps)
    constexpr bool kPrefetch = kEpiWarps == 8;                   // next chunk's residual one chunk ahead (16 warps: the other warps hide it)
    const int quarter = (e + 2) & 3, half = (e >> 2) * kCpw / 4; // `half`: which 128 columns (the fused-argmax slots; 8 warps only)
    const int sub = lane >> 3, q8 = lane & 7;                    // fp32 staging: row sub+4i, 16-byte chunk q8
    const int row_base = row0 + quarter * 32;
    const int row = row_base + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    bool any = false;
    // residual rows of a chunk, row-contiguous (8 lanes cover one 128-byte line); requested one chunk ahead of
    // their use — the first chunk's before the accumulator is even complete — so the loads never stall the tile
    auto load_resid = [&](int c, float4 (&rv)[8]) {
        const int col0 = n0 + c * 32;
        const bool live = ep.resid && !ep.reduce_add && c * 32 < nw && col0 < n;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = row_base + sub + 4 * i, cq = col0 + q8 * 4;
            rv[i] = (live && rr < m && cq < n) ? *reinterpret_cast<const float4*>(ep.resid + (int64_t)rr * ep.ldr + cq)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    const uint32_t stg_s = smem_u32(stg);
    // the staging tile is the source of asynchronous tensor stores: before it is written again, the lane that
    // issued the last store waits until that store has read it (it has had a whole chunk of arithmetic to do so)
    auto stg_acquire = [&]() {
        if (store_pending) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            store_pending = false;
        }
    };
    float4 rv[8], rv_next[kPrefetch ? 8 : 1];
    const int c_first = (e >> 2) * kCpw;
    FA_GT(long long gt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long g0 = clock64(); long long g1;)
    load_resid(c_first, rv);
    mbar_wait(bar_ready, ready_parity);                          // the accumulator is complete
    tc_fence_after();
    FA_GT(g1 = clock64(); gt[0] += g1 - g0; g0 = g1;)
#pragma unroll 1
    for (int cc = 0; cc < kCpw; ++cc) {
        const int c = c_first + cc;
        const int col0 = n0 + c * 32;
        if (c * 32 >= nw || col0 >= n) break;                    // warp-uniform
        any = true;
        if constexpr (kPrefetch) { if (cc < kCpw - 1) load_resid(c + 1, rv_next); }
        // bias of the chunk's 32 columns: the same 8 vectors for every lane (broadcast loads)
        float bias[32];
        if (col0 + 32 <= n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j);
                bias[4 * j] = b4.x; bias[4 * j + 1] = b4.y; bias[4 * j + 2] = b4.z; bias[4 * j + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) bias[j] = col0 + j < n ? __ldg(ep.bias + col0 + j) : 0.f;
        }
        uint32_t r[32];
        FA_GT(g1 = clock64(); gt[1] += g1 - g0; g0 = g1;)        // resid request + bias loads
        tc_ld32(taddr + c * 32, r);
        tc_wait_ld();
        FA_GT(g1 = clock64(); gt[2] += g1 - g0; g0 = g1;)        // TMEM load
        if (ep.col_scale) {                                     // warp-uniform
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 s4 = col0 + 32 <= n ? __ldg(reinterpret_cast<const float4*>(ep.col_scale + col0) + j)
                                                 : make_float4(1.f, 1.f, 1.f, 1.f);
                r[4 * j] = __float_as_uint(__fmul_rn(__uint_as_float(r[4 * j]), s4.x));
                r[4 * j + 1] = __float_as_uint(__fmul_rn(__uint_as_float(r[4 * j + 1]), s4.y));
                r[4 * j + 2] = __float_as_uint(__fmul_rn(__uint_as_float(r[4 * j + 2]), s4.z));
                r[4 * j + 3] = __float_as_uint(__fmul_rn(__uint_as_float(r[4 * j + 3]), s4.w));
            }
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(__uint_as_float(r[j]), bias[j]);
        if (ep.amax_val) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < n && v[j] > best) { best = v[j]; best_i = col0 + j; }
            continue;
        }
        if (ep.cand_list) {
            // `best` is this warp's running maximum of the row, seeded from the row's global running maximum (possibly
            // stale, hence lower: the test below is then looser, never tighter).  Every column within bound2 of it
            // can still be the true argmax and is appended with its approximate logit, so that the rescoring kernel
            // can drop the entries the final maximum rules out.
            if (cc == 0 && row < m) best = ord2f_dev(__ldcg(ep.cand_run_max + row));     // L2: other SMs raise it
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < n) cm = fmaxf(cm, v[j]);
            best = fmaxf(best, cm);
            if (row < m) {
                const float thr = best - ep.cand_bound2[row];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < n && v[j] >= thr) {
                        const int pos = atomicAdd(ep.cand_count + row, 1);
                        if (pos < ep.cand_cap) ep.cand_list[(int64_t)row * ep.cand_cap + pos] = make_int2(col0 + j, __float_as_int(v[j]));
                    }
                }
            }
            continue;
        }
        if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.resid && !ep.reduce_add) {
            FA_GT(g1 = clock64(); gt[3] += g1 - g0; g0 = g1;)    // bias add etc.
            stg_acquire();
            FA_GT(g1 = clock64(); gt[4] += g1 - g0; g0 = g1;)    // wait: previous store has read the staging tile
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = sub + 4 * i;
                *reinterpret_cast<float4*>(stg + rr * 128 + ((q8 ^ (rr & 7)) * 16)) = rv[i];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(stg + lane * 128 + ((j ^ (lane & 7)) * 16));
                v[4 * j + 0] = __fadd_rn(t.x, v[4 * j + 0]);
                v[4 * j + 1] = __fadd_rn(t.y, v[4 * j + 1]);
                v[4 * j + 2] = __fadd_rn(t.z, v[4 * j + 2]);
                v[4 * j + 3] = __fadd_rn(t.w, v[4 * j + 3]);
            }
            __syncwarp();
            FA_GT(g1 = clock64(); gt[5] += g1 - g0; g0 = g1;)    // residual through the staging tile (incl. waiting for its loads)
        }
        // Outputs leave through the staging tile as tensor stores: the tile is written in the layout the tensor
        // map's swizzle mode expects (128-byte rows / SWIZZLE_128B for fp32, 64-byte rows / SWIZZLE_64B for a
        // bf16 plane — also what keeps the 16-byte shared stores conflict-free), one lane issues the store, and
        // rows or columns past the matrix edge are clipped by the map.
        if (ep.out && col0 >= ep.f32_col_begin) {
            stg_acquire();
            FA_GT(g1 = clock64(); gt[4] += g1 - g0; g0 = g1;)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) * 16)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                // x += tile: the residual stream is updated in place by the memory system (no residual read in this
                // kernel; same bits as adding in registers: one fp32 round-to-nearest add per element, each element once)
                if (ep.reduce_add) tma_reduce_add_2d(map_out, stg_s, col0, row_base);
                else tma_store_2d(map_out, stg_s, col0, row_base);
                bulk_commit();
            }
            store_pending = true;
            FA_GT(g1 = clock64(); gt[6] += g1 - g0; g0 = g1;)    // fp32 store: smem write, fence, issue
        }
        if (ep.out_f8) {
            // e4m3 (round to nearest even, saturating at +-448): a thread owns 32 consecutive bytes of its row per chunk; the
            // warp's chunks are adjacent columns, so they are staged side by side and leave as ONE tensor store after the last
            // (map_pl is the e4m3 map in this case; rows and columns past the matrix edge are clipped by it)
            uint32_t w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = f32x4_to_e4m3(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (cc == 0) stg_acquire();
            unsigned char* d8 = stg + lane * (kCpw * 32) + cc * 32;
            *reinterpret_cast<uint4*>(d8) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            *reinterpret_cast<uint4*>(d8 + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
        }
        if (ep.out_hi) {
            if (col0 < ep.pl_col_scale_end) {                    // warp-uniform: the q columns of a fused q|k|v projection
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(v[j], ep.pl_col_scale);
            }
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
            stg_acquire();
            // hi rows in the first 2 KB (64 B per row), lo rows in the second
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16);
                *reinterpret_cast<uint4*>(stg + o) = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
                *reinterpret_cast<uint4*>(stg + 2048 + o) = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(map_pl, stg_s, col0, row_base, 0);      // ONE store: the map's box spans both planes (hi tile, lo tile)
                bulk_commit();
            }
            store_pending = true;
            FA_GT(g1 = clock64(); gt[7] += g1 - g0; g0 = g1;)    // plane store: split, (acquire), smem write, fence, issue
        }
        if constexpr (kPrefetch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rv[i] = rv_next[i];
        } else {
            if (cc < kCpw - 1) load_resid(c + 1, rv);            // after this chunk's stores have been issued
        }
    }
    if (ep.out_f8 && any) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(map_pl, stg_s, n0 + c_first * 32, row_base);
            bulk_commit();
        }
        store_pending = true;
    }
    FA_GT(if (ep.dbg && lane == 0) { for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(ep.dbg) + i, (unsigned long long)gt[i]);
                                      atomicAdd(reinterpret_cast<unsigned long long*>(ep.dbg) + 8, 1ull); })
    if (ep.cand_list && any && row < m) atomicMax(ep.cand_run_max + row, f2ord_dev(best));
    if (kEpiWarps == 8 && ep.amax_val && half * 128 < nw) {
        // one partial slot per (row, 128-column group): the two column halves of a 256-wide tile live in
        // different warps, and a half tile of the tail is exactly one group.  A group that lies wholly past N
        // (the last tile of the vocabulary) still gets its slot written — (-inf, none) — because the combine
        // kernel reads every slot.
        if (row < m) {
            const int slot = (n0 >> 7) + half;
            ep.amax_val[(int64_t)row * ep.n_slots + slot] = best;
            ep.amax_idx[(int64_t)row * ep.n_slots + slot] = best_i;
        }
    }
}

template <int NP, int TN = BN>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
          const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_pl, int m, int n, int k,
          int band, EpiParams ep) {
    using C = Cfg<NP, TN>;
    constexpr uint32_t kIdescT = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles_base = (raw + 1023u) & ~1023u;                      // SWIZZLE_128B wants 1024 B alignment
    const uint32_t stg_base = tiles_base + C::kStages * C::kStageBytes;
    const uint32_t bars = stg_base + kStagingBytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages;
    const uint32_t bar_tfull = bars + 16 * C::kStages, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m_tiles = (m + BM - 1) / BM, n_tiles = (n + TN - 1) / TN;
    int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (k + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);    // provably warp-uniform
    grid_dependency_wait();                       // everything above overlapped the previous kernel's tail
    if (ep.gate && __ldcg(ep.gate) == 0) num_tiles = 0;                   // gated launch with nothing to do

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int mt, nt;
                tile_coords(tile, m_tiles, n_tiles, band, mt, nt);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sbase = tiles_base + stage * C::kStageBytes;
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_arrive_expect_tx(full, C::kStageBytes);
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        tma_load_3d(sbase + p * kATileBytes, &map_a, full, kb * BK, mt * BM, p);
                        tma_load_3d(sbase + NP * kATileBytes + p * C::kWBytes, &map_w, full, kb * BK, nt * TN, p);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sbase = tiles_base + stage * C::kStageBytes;
                    const uint32_t a_hi = sbase, a_lo = sbase + kATileBytes;
                    const uint32_t w_hi = sbase + NP * kATileBytes, w_lo = w_hi + C::kWBytes;
                    uint32_t accum = kb > 0 ? 1u : 0u;
                    if constexpr (NP == 2) {
                        // small cross terms first, the dominant hi*hi term last
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            tc_mma(tmem_d, umma_desc_sw128(a_lo + ks * 32), umma_desc_sw128(w_hi + ks * 32), kIdescT, accum);
                            accum = 1u;
                        }
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks)
                            tc_mma(tmem_d, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(w_lo + ks * 32), kIdescT, 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        tc_mma(tmem_d, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(w_hi + ks * 32), kIdescT, accum);
                        accum = 1u;
                    }
                    tc_commit(bar_empty + 8 * stage);                 // smem stage free once these MMAs retire
                    if (kb == k_blocks - 1) tc_commit(bar_tfull + 8 * acc);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int e = warp - 2;
        unsigned char* stg = smem_raw + (stg_base - raw) + e * 4096;
        bool store_pending = false;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int mt, nt;
            tile_coords(tile, m_tiles, n_tiles, band, mt, nt);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            epilogue_unit(ep, &map_out, &map_pl, stg, store_pending, e, lane, m, n, mt * BM, nt * TN, TN, tmem_base + acc * BN,
                          bar_tfull + 8 * acc, acc_phase);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        if (lane == 0) bulk_wait_read0();                  // the staging tile must outlive the last store's read
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}


// ------------------------------------------------------------------------------------ CTA-pair kernel
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same offset in the pair's leader CTA (rank 0): the CTA rank within the
// pair is bit 24 of a shared-window address (the Sm100MmaPeerBitMask convention)
__device__ __forceinline__ uint32_t leader_addr(uint32_t a) { return a & 0xFEFFFFFFu; }

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// the same on e4m3 operands (kind::f8f6f4: K = 32 per instruction, i.e. the same 32 bytes of a 128-byte swizzled row)
__device__ __forceinline__ void tc_mma_pair_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    // relaxed: nothing in memory is being published (the TMEM reads are complete after tcgen05.wait::ld); a
    // release at cluster scope would first drain every global store this warp has in flight
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

__device__ __forceinline__ void unit_coords(int u, const Sched& s, int& mt, int& n0, int& nw) {
    int tile = u, part = 0, parts_log2 = 0;
    if (u >= s.full_units) {
        const int r = u - s.full_units;
        parts_log2 = s.split_log2;
        tile = s.full_units + (r >> parts_log2);
        part = r & ((1 << parts_log2) - 1);
    }
    int nt;
    tile_coords(tile, s.m_tiles, s.n_tiles, s.band, mt, nt);
    nw = BN >> parts_log2;
    n0 = nt * BN + part * nw;
}

// F8: one plane of e4m3 bytes per operand (the reference's int8 graph variant, 02-Quantize-ONNX.py:41-44, as W8A8
// floating point): a K-block is 128 elements = the same 128-byte rows, tcgen05.mma.kind::f8f6f4 takes K = 32 per
// instruction, the per-output-channel weight scale is applied in the epilogue.  Everything else is the NP = 1 kernel.
template <int NP, bool F8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Cfg2<NP, F8>::kThreads), 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
           const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_pl, int m, int n, int k,
           Sched sched, EpiParams ep) {
    using C = Cfg2<NP, F8>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles_base = (raw + 1023u) & ~1023u;
    const uint32_t stg_base = tiles_base + C::kStages * C::kStageBytes;
    const uint32_t bars = stg_base + C::kStagingBytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages;
    const uint32_t bar_tfull = bars + 16 * C::kStages, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();                 // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    constexpr int kBKe = F8 ? 2 * BK : BK;                    // elements per K-block: 128 bytes of a row either way
    static_assert(!F8 || NP == 1, "the fp8 kernel has one plane per operand");
    const int k_blocks = (k + kBKe - 1) / kBKe;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);       // used in the leader only: one arrive.expect_tx for both CTAs' bytes
            mbar_init(bar_empty + 8 * s, 1);      // one multicast commit per use, in each CTA
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);      // multicast commit, in each CTA
            mbar_init(bar_tempty + 8 * a, 2 * C::kEpiWarps);    // used in the leader only: the epilogue warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                           // both CTAs' barriers initialised, both TMEM allocations done
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);    // provably warp-uniform
    grid_dependency_wait();                       // everything above overlapped the previous kernel's tail
    if (ep.gate && __ldcg(ep.gate) == 0) sched.total_units = 0;           // gated launch with nothing to do

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = pair; u < sched.total_units; u += num_pairs) {
                int mt, n0, nw;
                unit_coords(u, sched, mt, n0, nw);
                const int w_rows = nw >> 1;                               // this CTA's share of the W tile
                const int a_row = (mt * 2 + rank) * BM, w_row = n0 + rank * w_rows;
                const uint32_t cta_bytes = NP * (kATileBytes + w_rows * BK * 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sbase = tiles_base + stage * C::kStageBytes;
                    const uint32_t full = leader_addr(bar_full + 8 * stage);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * cta_bytes);
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            tma_load_3d_pair(sbase + p * kATileBytes, &map_a, full, kb * kBKe, a_row, p);
                            const uint32_t wdst = sbase + NP * kATileBytes + p * C::kHalfWBytes;
                            for (int r = 0; r < w_rows; r += 64)
                                tma_load_3d_pair(wdst + r * BK * 2, &map_w, full, kb * kBKe, w_row + r, p);
                        }
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = pair; u < sched.total_units; u += num_pairs, ++it) {
                int mt, n0, nw;
                unit_coords(u, sched, mt, n0, nw);
                // D = f32, A = B = bf16, K-major both, N = nw, M = 256 (128 rows in each CTA)
                // (F8: A = B = e4m3, format code 0)
                const uint32_t idesc = (1u << 4) | (F8 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(nw >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                FA_GT(long long m0 = clock64();)
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);       // both CTAs' epilogues have drained this accumulator
                FA_GT(if (ep.dbg && lane == 0) { atomicAdd(reinterpret_cast<unsigned long long*>(ep.dbg) + 9, (unsigned long long)(clock64() - m0));
                                                  atomicAdd(reinterpret_cast<unsigned long long*>(ep.dbg) + 10, 1ull); })
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    FA_GT(long long f0 = clock64();)
                    mbar_wait(bar_full + 8 * stage, phase);
                    FA_GT(if (ep.dbg && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(ep.dbg) + 11, (unsigned long long)(clock64() - f0));)
                    tc_fence_after();
                    const uint32_t sbase = tiles_base + stage * C::kStageBytes;
                    const uint32_t a_hi = sbase, a_lo = sbase + kATileBytes;
                    const uint32_t w_hi = sbase + NP * kATileBytes, w_lo = w_hi + C::kHalfWBytes;
                    if (elect_one()) {
                        uint32_t accum = kb > 0 ? 1u : 0u;
                        if constexpr (NP == 2) {
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                tc_mma_pair(tmem_d, umma_desc_sw128(a_lo + ks * 32), umma_desc_sw128(w_hi + ks * 32), idesc, accum);
                                accum = 1u;
                            }
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks)
                                tc_mma_pair(tmem_d, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(w_lo + ks * 32), idesc, 1u);
                        }
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            if constexpr (F8) tc_mma_pair_f8(tmem_d, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(w_hi + ks * 32), idesc, accum);
                            else tc_mma_pair(tmem_d, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(w_hi + ks * 32), idesc, accum);
                            accum = 1u;
                        }
                        tc_commit_pair(bar_empty + 8 * stage);            // frees this stage in both CTAs
                        if (kb == k_blocks - 1) tc_commit_pair(bar_tfull + 8 * acc);
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps in each CTA)
        const int e = warp - 2;
        unsigned char* stg = smem_raw + (stg_base - raw) + e * 4096;
        bool store_pending = false;
        int it = 0;
        for (int u = pair; u < sched.total_units; u += num_pairs, ++it) {
            int mt, n0, nw;
            unit_coords(u, sched, mt, n0, nw);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            epilogue_unit<C::kEpiWarps>(ep, &map_out, &map_pl, stg, store_pending, e, lane, m, n, (mt * 2 + rank) * BM, n0, nw,
                                        tmem_base + acc * BN, bar_tfull + 8 * acc, acc_phase);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_addr(bar_tempty + 8 * acc));
        }
        if (lane == 0) bulk_wait_read0();                  // the staging tile must outlive the last store's read
    }

    tc_fence_before();
    cluster_sync_all();                           // the peer's smem and TMEM stay alive until the leader's last MMA retired
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------ A-resident pair kernel
// The vocabulary projection in its one-product form (NP = 1, K <= 512, candidate-list epilogue, nothing stored) is
// bound by the L2 -> SM feed in k_gemm_tc2: 32 KB per K-block per SM against 512 tensor cycles is 64 B/clk/SM,
// above what the L2 delivers to all SMs at once.  Here a pair keeps its 256 x K block of A in shared memory
// (128 KB per CTA) and walks a run of consecutive 256-column tiles of W, so only W streams: 16 KB per K-block per
// SM.  Work items are (256-row block, run of n-tiles); consecutive items are different row blocks of the same
// run, so the pairs running together read the same W tiles (L2) and a row is almost never in two pairs at once,
// which keeps the running maxima of the candidate lists tight.  With a K this short the epilogue (one pass over
// 256 x 256 approximate logits per 4096 tensor cycles) is the other limiter: 16 epilogue warps per CTA, two
// 32-column chunks each, TMEM loads one chunk ahead, the tile's bias staged in shared memory before the
// accumulator is ready, and the append path entered only by lanes whose chunk maximum passes the row's test.
struct SchedAR {
    int m_tiles, n_tiles, groups, group_tiles, items;
};
struct CfgAR {
    static constexpr int kMaxKBlocks = 8;
    static constexpr int kABytes = kMaxKBlocks * kATileBytes;                 // 128 KB
    static constexpr int kHalfWBytes = (BN / 2) * BK * 2;                     // 16 KB
    static constexpr int kStages = 5;
    static constexpr int kEpiWarps = 16, kChunks = 2;                         // 16 warps x 2 chunks x 32 columns = 4 quarters x 256
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kBiasBytes = kEpiWarps * kChunks * 32 * 4;           // 4 KB
    static constexpr int kSmemBytes = kABytes + kStages * kHalfWBytes + kBiasBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// One 32-column chunk of approximate logits (already biased): raise the running maximum, append what passes.
__device__ __forceinline__ void cand_chunk(const EpiParams& ep, const float (&v)[32], int col0, int n, int row, bool row_live,
                                           float bound2, float& best) {
    float cm = -INFINITY;
    if (col0 + 32 <= n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) cm = fmaxf(cm, v[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < n) cm = fmaxf(cm, v[j]);
    }
    best = fmaxf(best, cm);
    const float thr = best - bound2;
    if (row_live && cm >= thr) {                 // rare per lane: this chunk holds a column within bound2 of the maximum so far
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (col0 + j < n && v[j] >= thr) {
                int pos;
                asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(pos) : "l"(ep.cand_count + row) : "memory");
                if (pos < ep.cand_cap) ep.cand_list[(int64_t)row * ep.cand_cap + pos] = make_int2(col0 + j, __float_as_int(v[j]));
            }
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CfgAR::kThreads, 1)
k_gemm_tc2_ar(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int m, int n, int k,
              SchedAR sched, EpiParams ep) {
    using C = CfgAR;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t a_base = (raw + 1023u) & ~1023u;
    const uint32_t w_base = a_base + C::kABytes;
    const uint32_t bias_base = w_base + C::kStages * C::kHalfWBytes;
    const uint32_t bars = bias_base + C::kBiasBytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages;
    const uint32_t bar_tfull = bars + 16 * C::kStages, bar_tempty = bar_tfull + 16;
    const uint32_t bar_afull = bar_tempty + 16, bar_aempty = bar_afull + 8;
    const uint32_t tmem_slot = bar_aempty + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int k_blocks = (k + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 2 * C::kEpiWarps);
        }
        mbar_init(bar_afull, 1);                  // leader only: one arrive.expect_tx for both CTAs' A blocks
        mbar_init(bar_aempty, 1);                 // multicast commit after the last MMA of an item, in each CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
    grid_dependency_wait();

    auto item_coords = [&](int item, int& mt, int& nt0, int& nt1) {
        const int g = item / sched.m_tiles;
        mt = item - g * sched.m_tiles;
        nt0 = g * sched.group_tiles;
        nt1 = min(sched.n_tiles, nt0 + sched.group_tiles);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0, a_phase = 0;
        for (int item = pair; item < sched.items; item += num_pairs, a_phase ^= 1) {
            int mt, nt0, nt1;
            item_coords(item, mt, nt0, nt1);
            const int a_row = (mt * 2 + rank) * BM;
            mbar_wait(bar_aempty, a_phase ^ 1);                       // the previous item's MMAs have read A
            if (elect_one()) {
                if (rank == 0) mbar_arrive_expect_tx(bar_afull, 2u * (uint32_t)k_blocks * kATileBytes);
                const uint32_t afull = leader_addr(bar_afull);
                for (int kb = 0; kb < k_blocks; ++kb)
                    tma_load_3d_pair(a_base + kb * kATileBytes, &map_a, afull, kb * BK, a_row, 0);
            }
            __syncwarp();
            for (int nt = nt0; nt < nt1; ++nt) {
                const int w_row = nt * BN + rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t wdst = w_base + stage * C::kHalfWBytes;
                    const uint32_t full = leader_addr(bar_full + 8 * stage);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * C::kHalfWBytes);
                        tma_load_3d_pair(wdst, &map_w, full, kb * BK, w_row, 0);
                        tma_load_3d_pair(wdst + 64 * BK * 2, &map_w, full, kb * BK, w_row + 64, 0);
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            int it = 0;
            for (int item = pair; item < sched.items; item += num_pairs, a_phase ^= 1) {
                int mt, nt0, nt1;
                item_coords(item, mt, nt0, nt1);
                mbar_wait(bar_afull, a_phase);
                for (int nt = nt0; nt < nt1; ++nt, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t a_t = a_base + kb * kATileBytes, w_t = w_base + stage * C::kHalfWBytes;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks)
                                tc_mma_pair(tmem_d, umma_desc_sw128(a_t + ks * 32), umma_desc_sw128(w_t + ks * 32), idesc,
                                            (kb | ks) ? 1u : 0u);
                            tc_commit_pair(bar_empty + 8 * stage);
                            if (kb == k_blocks - 1) {
                                tc_commit_pair(bar_tfull + 8 * acc);
                                if (nt == nt1 - 1) tc_commit_pair(bar_aempty);        // A may be overwritten
                            }
                        }
                        __syncwarp();
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (16 warps in each CTA)
        const int e = warp - 2;
        const int quarter = (e + 2) & 3, c_first = (e >> 2) * C::kChunks;    // TMEM lane quarter = warp id % 4
        float* wbias = reinterpret_cast<float*>(smem_raw + (bias_base - raw)) + e * (C::kChunks * 32);
        int it = 0;
        for (int item = pair; item < sched.items; item += num_pairs) {
            int mt, nt0, nt1;
            item_coords(item, mt, nt0, nt1);
            const int row = (mt * 2 + rank) * BM + quarter * 32 + lane;
            const bool row_live = row < m;
            // the row's running maximum lives in a register for the whole run of tiles; other pairs (other runs of the
            // same rows) may have raised the global one meanwhile: a stale value only lengthens the list
            float best = row_live ? ord2f_dev(__ldcg(ep.cand_run_max + row)) : -INFINITY;
            const float bound2 = row_live ? ep.cand_bound2[row] : 0.f;
            const uint32_t taddr_row = ((uint32_t)(quarter * 32) << 16) + c_first * 32;
            for (int nt = nt0; nt < nt1; ++nt, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                const int col_w = nt * BN + c_first * 32;                    // this warp's 64 columns
                // bias of those columns into the warp's shared slice while the accumulator is still being computed
                __syncwarp();
                {
                    const int c = col_w + 2 * lane;
                    float2 b2v;
                    b2v.x = c < n ? __ldg(ep.bias + c) : 0.f;
                    b2v.y = c + 1 < n ? __ldg(ep.bias + c + 1) : 0.f;
                    *reinterpret_cast<float2*>(wbias + 2 * lane) = b2v;
                }
                __syncwarp();
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN + taddr_row;
                uint32_t r0[32], r1[32];
                tc_ld32(taddr, r0);
                tc_wait_ld();
                tc_ld32(taddr + 32, r1);                                     // in flight while chunk 0 is examined
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(wbias + 4 * j);
                    v[4 * j] = __fadd_rn(__uint_as_float(r0[4 * j]), b4.x);
                    v[4 * j + 1] = __fadd_rn(__uint_as_float(r0[4 * j + 1]), b4.y);
                    v[4 * j + 2] = __fadd_rn(__uint_as_float(r0[4 * j + 2]), b4.z);
                    v[4 * j + 3] = __fadd_rn(__uint_as_float(r0[4 * j + 3]), b4.w);
                }
                if (col_w < n) cand_chunk(ep, v, col_w, n, row, row_live, bound2, best);
                tc_wait_ld();
                // the accumulator has been read: hand it back before the second chunk is examined
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_addr(bar_tempty + 8 * acc));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(wbias + 32 + 4 * j);
                    v[4 * j] = __fadd_rn(__uint_as_float(r1[4 * j]), b4.x);
                    v[4 * j + 1] = __fadd_rn(__uint_as_float(r1[4 * j + 1]), b4.y);
                    v[4 * j + 2] = __fadd_rn(__uint_as_float(r1[4 * j + 2]), b4.z);
                    v[4 * j + 3] = __fadd_rn(__uint_as_float(r1[4 * j + 3]), b4.w);
                }
                if (col_w + 32 < n) cand_chunk(ep, v, col_w + 32, n, row, row_live, bound2, best);
            }
            if (row_live) atomicMax(ep.cand_run_max + row, f2ord_dev(best));
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;
int g_num_sms = 0;
int g_num_pairs = 0;      // CTA pairs (clusters of 2) that can be co-resident with the pair kernel's footprint

}  // namespace

void tc_init_device() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        FA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        FA_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    FA_CUDA(cudaFuncSetAttribute(k_gemm_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute(k_gemm_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute((k_gemm_tc<2, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, (Cfg<2, 128>::kSmemBytes)));
    FA_CUDA(cudaFuncSetAttribute((k_gemm_tc<2, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (Cfg<2, 64>::kSmemBytes)));
    FA_CUDA(cudaFuncSetAttribute(k_gemm_tc2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<1>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute(k_gemm_tc2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<2>::kSmemBytes));
    FA_CUDA(cudaFuncSetAttribute((k_gemm_tc2<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (Cfg2<1, true>::kSmemBytes)));
    FA_CUDA(cudaFuncSetAttribute(k_gemm_tc2_ar, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgAR::kSmemBytes));
    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    FA_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    // how many clusters of 2 fit at once (one CTA per SM by shared memory): 74 on a full B200
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(g_num_sms & ~1)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg2<2>::kSmemBytes;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int clusters = 0;
    FA_CUDA(cudaOccupancyMaxActiveClusters(&clusters, k_gemm_tc2<2>, &cfg));
    g_num_pairs = clusters < g_num_sms / 2 ? clusters : g_num_sms / 2;
    FA_REQUIRE(g_num_pairs >= 1, "no CTA pair of the tcgen05 GEMM fits on this device");
}

TcOperand tc_make_operand(const __nv_bfloat16* base, int rows, int k, int64_t row_stride_elems,
                          int64_t plane_stride_elems, int planes, int box_rows) {
    FA_REQUIRE(g_encode != nullptr, "tc_init_device() has not run");
    FA_REQUIRE(row_stride_elems % 8 == 0 && plane_stride_elems % 8 == 0, "TMA strides must be multiples of 16 bytes");
    FA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    TcOperand op;
    op.rows = rows; op.k = k; op.planes = planes;
    const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2, (cuuint64_t)plane_stride_elems * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(&op.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return op;
}

TcOperand tc_make_operand_f8(const uint8_t* base, int rows, int k, int64_t row_stride_bytes, int box_rows) {
    FA_REQUIRE(g_encode != nullptr, "tc_init_device() has not run");
    FA_REQUIRE(row_stride_bytes % 16 == 0, "TMA strides must be multiples of 16 bytes");
    FA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    TcOperand op;
    op.rows = rows; op.k = k; op.planes = 1;
    const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 1};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride_bytes, (cuuint64_t)row_stride_bytes * rows};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * BK), (cuuint32_t)box_rows, 1};      // 128 e4m3 = 128 bytes
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(&op.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled (e4m3 operand) failed with code " + std::to_string((int)r));
    op.map64 = op.map;
    op.has64 = box_rows == 64;
    return op;
}

namespace {
// tensor map of an epilogue output: `rank`-D, 32 x 32 boxes, rows of one box 128 B (fp32) or 64 B (bf16) wide
CUtensorMap make_store_map(CUtensorMapDataType dt, int elem_bytes, void* base, int rank, const cuuint64_t* dims,
                           const cuuint64_t* strides_bytes, CUtensorMapSwizzle swz, int box_cols = 32, int box_planes = 1) {
    FA_REQUIRE(g_encode != nullptr, "tc_init_device() has not run");
    FA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    (void)elem_bytes;
    CUtensorMap map;
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, 32, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(&map, dt, (cuuint32_t)rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled (store map) failed with code " + std::to_string((int)r));
    return map;
}
}  // namespace

TcOperand tc_make_weight(const __nv_bfloat16* base, int rows, int k, int64_t plane_stride_elems, int planes) {
    TcOperand op = tc_make_operand(base, rows, k, k, plane_stride_elems, planes, BN);
    op.map64 = tc_make_operand(base, rows, k, k, plane_stride_elems, planes, 64).map;
    op.map128 = tc_make_operand(base, rows, k, k, plane_stride_elems, planes, 128).map;
    op.has64 = true;
    return op;
}

int tc_num_pairs() { return g_num_pairs; }
int tc_argmax_tiles(int n) { return 2 * cdiv(n, BN); }   // one partial slot per 128 columns

namespace {
// 0 = choose by size, 1 = one CTA per tile, 2 = CTA pairs.  FUNASR_B200_GEMM=1cta|2cta is a test and
// comparison aid; read at every launch so a test can flip it.
int gemm_kernel_override() {
    const char* s = getenv("FUNASR_B200_GEMM");
    if (!s) return 0;
    if (!strcmp(s, "1cta")) return 1;
    if (!strcmp(s, "2cta")) return 2;
    return 0;
}
bool gemm_ar_enabled() {
    const char* s = getenv("FUNASR_B200_GEMM_AR");
    return !(s && s[0] == '0');
}
}  // namespace

void launch_gemm_tc(const TcOperand& a, const TcOperand& w, int m, int n, int k, int n_planes, const Epilogue& e,
                    cudaStream_t st) {
    FA_REQUIRE(a.k == k && w.k == k && a.rows >= m && w.rows == n, "tc gemm operand shapes do not match");
    FA_REQUIRE(n_planes == 1 || n_planes == 2, "tc gemm supports 1 or 2 planes");
    FA_REQUIRE(a.planes >= n_planes && w.planes >= n_planes, "operand lacks the requested planes");
    EpiParams ep{};
    ep.bias = e.bias; ep.resid = e.resid; ep.out = e.out_f32; ep.out_hi = e.out_pl.hi; ep.out_lo = e.out_pl.lo;
    ep.amax_val = e.amax_val; ep.amax_idx = e.amax_idx;
    ep.cand_run_max = e.cand.run_max; ep.cand_count = e.cand.count; ep.cand_list = e.cand.list; ep.cand_bound2 = e.cand.bound2;
    ep.cand_cap = e.cand.cap; ep.gate = e.gate;
    ep.col_scale = e.col_scale; ep.out_f8 = e.out_f8; ep.ld8 = e.ld8;
#ifdef FA_GEMM_TIMING
    struct DbgGuard {
        long long* d = nullptr; cudaStream_t st; int m, n, k, np; bool resid;
        ~DbgGuard() {
            if (!d) return;
            unsigned long long h[12];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            cudaFree(d);
            const double u = h[8] ? (double)h[8] : 1.0;
            fprintf(stderr, "gemm timing m=%d n=%d k=%d planes=%d resid=%d: per epilogue-warp tile (cycles): wait_acc %.0f | req+bias %.0f tmem_ld %.0f "
                            "arith %.0f acquire %.0f resid_stage %.0f f32_store %.0f plane_store %.0f | MMA warp per tile: wait_drained %.0f wait_operands %.0f\n",
                    m, n, k, np, (int)resid, h[0] / u, h[1] / u, h[2] / u, h[3] / u, h[4] / u, h[5] / u, h[6] / u, h[7] / u,
                    h[10] ? (double)h[9] / h[10] : 0.0, h[10] ? (double)h[11] / h[10] : 0.0);
        }
    } dbg_guard;
    if (getenv("FUNASR_B200_GEMM_TIMING")) {
        FA_CUDA(cudaMalloc(&dbg_guard.d, 12 * sizeof(long long)));
        FA_CUDA(cudaMemsetAsync(dbg_guard.d, 0, 12 * sizeof(long long), st));
        dbg_guard.st = st; dbg_guard.m = m; dbg_guard.n = n; dbg_guard.k = k; dbg_guard.np = e.f8 ? 0 : n_planes; dbg_guard.resid = e.resid != nullptr;
        ep.dbg = dbg_guard.d;
    }
#endif
    FA_REQUIRE(!ep.out_f8 || (n % 32 == 0 && e.ld8 % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.out_f8) & 15) == 0),
               "e4m3 output needs N % 32 == 0 and 16-byte aligned rows");
    FA_REQUIRE(!ep.col_scale || (reinterpret_cast<uintptr_t>(ep.col_scale) & 15) == 0, "column scales must be 16-byte aligned");
    ep.ldr = e.ldr; ep.ldc = e.ldc; ep.ldp = e.ldp; ep.relu = e.relu ? 1 : 0; ep.n_slots = tc_argmax_tiles(n);
    ep.pl_col_scale = e.pl_col_scale; ep.pl_col_scale_end = e.pl_col_scale_end; ep.f32_col_begin = e.f32_col_begin;
    FA_REQUIRE(e.pl_col_scale_end % 32 == 0 && e.f32_col_begin % 32 == 0, "column ranges of the epilogue must be multiples of 32");
    FA_REQUIRE(!ep.out || (e.ldc % 4 == 0), "fp32 output stride must be a multiple of 4");
    FA_REQUIRE(!ep.resid || (e.ldr % 4 == 0), "residual stride must be a multiple of 4");
    FA_REQUIRE(!ep.out_hi || (e.ldp % 8 == 0), "plane output stride must be a multiple of 8");
    FA_REQUIRE(ep.bias != nullptr, "the tcgen05 GEMM epilogue expects a bias vector");
    FA_REQUIRE((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0, "bias must be 16-byte aligned");
    FA_REQUIRE(!ep.out || n % 4 == 0, "fp32 output needs N % 4 == 0");
    FA_REQUIRE(!ep.out_hi || n % 8 == 0, "plane output needs N % 8 == 0");
    // store maps of the epilogue outputs (unused ones are placeholders the kernel never touches)
    CUtensorMap map_out = a.map, map_pl = a.map;
    if (ep.out) {
        const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
        const cuuint64_t strides[1] = {(cuuint64_t)e.ldc * 4};
        map_out = make_store_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ep.out, 2, dims, strides, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (ep.out_hi) {
        const int planes = ep.out_lo ? 2 : 1;
        const int64_t plane_stride = ep.out_lo ? (ep.out_lo - ep.out_hi) : (int64_t)m * e.ldp;
        FA_REQUIRE(plane_stride > 0 && plane_stride % 8 == 0, "output planes must be hi then lo, a multiple of 8 elements apart");
        const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)m, (cuuint64_t)planes};
        const cuuint64_t strides[2] = {(cuuint64_t)e.ldp * 2, (cuuint64_t)plane_stride * 2};
        map_pl = make_store_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep.out_hi, 3, dims, strides, CU_TENSOR_MAP_SWIZZLE_64B, 32, planes);
    }
    // A gated launch whose gate is closed executes nothing: it is booked at 0 FLOP under its own key (the gate is
    // only known on the device; the second-chance vocabulary pass is the one gated launch and its gate is closed
    // unless a candidate list overflowed).
    if (ep.out_f8) {
        FA_REQUIRE(!ep.out_hi && !ep.out, "a launch writes e4m3 or the other output forms, not both (they share the staging tile)");
        FA_REQUIRE(e.f8, "the e4m3 output form belongs to the fp8 projection kernel");
        const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
        const cuuint64_t strides[1] = {(cuuint64_t)e.ld8};
        map_pl = make_store_map(CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, ep.out_f8, 2, dims, strides, CU_TENSOR_MAP_SWIZZLE_NONE,
                                32 * (32 / Cfg2<1, true>::kEpiWarps));      // the chunks of one epilogue warp side by side
    }
    prof_note_work(e.gate ? 0.0 : 2.0 * m * (double)n * k, 0.0);
    if (g_prof_on) {
        char tag[64];
        if (e.gate) snprintf(tag, sizeof tag, "gated");
        else snprintf(tag, sizeof tag, "n%d_k%d", n, k);
        prof_note_tag(tag);
    }
    // CTA pairs (256 x 256 tiles) once there is at least a full wave of them; below that the 128-row
    // tiles of the single-CTA kernel spread a small M over twice as many SMs.
    const int pair_tiles = cdiv(m, 2 * BM) * cdiv(n, BN);
    if (e.f8) {
        // e4m3 operands: always the CTA-pair kernel (a short M simply runs fewer pairs)
        FA_REQUIRE(a.has64 == false && w.has64, "fp8 gemm: A needs 128-row boxes, W 64-row boxes");
        Sched s{};
        s.m_tiles = cdiv(m, 2 * BM); s.n_tiles = cdiv(n, BN); s.band = 8;
        const int pairs = pair_tiles < g_num_pairs ? pair_tiles : g_num_pairs;
        s.full_units = pair_tiles / pairs * pairs;
        const int rest = pair_tiles - s.full_units;
        s.split_log2 = (rest > 0 && 2 * rest <= pairs) ? 1 : 0;
        s.total_units = s.full_units + (rest << s.split_log2);
        if (g_prof_on) { char tag[64]; snprintf(tag, sizeof tag, "fp8_n%d_k%d", n, k); prof_note_tag(tag); }
        FA_REQUIRE(!ep.amax_val && !ep.cand_list, "the fp8 projection kernel has no vocabulary-argmax epilogue");
        FA_LAUNCH((k_gemm_tc2<1, true>), 2 * pairs, (Cfg2<1, true>::kThreads), (Cfg2<1, true>::kSmemBytes), st, a.map, w.map, map_out,
                  map_pl, m, n, k, s, ep);
        return;
    }
    const int force = gemm_kernel_override();
    // one-product projections that store nothing (the vocabulary argmax epilogues) and whose K fits: A stays in
    // shared memory, only W streams.  FUNASR_B200_GEMM_AR=0 falls back to the general pair kernel (comparison aid).
    const bool no_stores = !ep.out && !ep.out_hi && !ep.resid && ep.cand_list;
    if (w.has64 && force != 1 && n_planes == 1 && no_stores && cdiv(k, BK) <= CfgAR::kMaxKBlocks &&
        pair_tiles >= 2 * g_num_pairs && gemm_ar_enabled()) {
        SchedAR s{};
        s.m_tiles = cdiv(m, 2 * BM); s.n_tiles = cdiv(n, BN);
        // runs of n-tiles: as few (long) as keep the last round of items full; an item also pays ~1.5 tile times for its A block
        double best_cost = 0.0;
        for (int g = 1; g <= 32 && g <= s.n_tiles; ++g) {
            const int gt = cdiv(s.n_tiles, g), items = s.m_tiles * cdiv(s.n_tiles, gt);
            const double cost = (double)cdiv(items, g_num_pairs) * (gt + 1.5);
            if (g == 1 || cost < best_cost) { best_cost = cost; s.group_tiles = gt; }
        }
        s.groups = cdiv(s.n_tiles, s.group_tiles);
        s.items = s.m_tiles * s.groups;
        const int pairs = s.items < g_num_pairs ? s.items : g_num_pairs;
        FA_LAUNCH(k_gemm_tc2_ar, 2 * pairs, CfgAR::kThreads, CfgAR::kSmemBytes, st, a.map, w.map64, m, n, k, s, ep);
        return;
    }
    if (w.has64 && force != 1 && (force == 2 || pair_tiles >= g_num_pairs)) {
        Sched s{};
        s.m_tiles = cdiv(m, 2 * BM); s.n_tiles = cdiv(n, BN); s.band = 8;
        int pairs = pair_tiles < g_num_pairs ? pair_tiles : g_num_pairs;
        if (const char* pe = getenv("FUNASR_B200_GEMM_PAIRS")) { const int v = atoi(pe); if (v > 0 && v < pairs) pairs = v; }   // tuning aid
        s.full_units = pair_tiles / pairs * pairs;
        const int rest = pair_tiles - s.full_units;
        s.split_log2 = (rest > 0 && 2 * rest <= pairs) ? 1 : 0;      // a last wave at most half full is cut into half tiles
        s.total_units = s.full_units + (rest << s.split_log2);
        if (n_planes == 2) {
            FA_LAUNCH(k_gemm_tc2<2>, 2 * pairs, kThreads, Cfg2<2>::kSmemBytes, st, a.map, w.map64, map_out, map_pl, m, n, k, s, ep);
        } else {
            FA_LAUNCH(k_gemm_tc2<1>, 2 * pairs, kThreads, Cfg2<1>::kSmemBytes, st, a.map, w.map64, map_out, map_pl, m, n, k, s, ep);
        }
        return;
    }
    const int band = 16;
    // Small M (calls of one to four segments): the latency of the call is the sum of ~310 dependent projections, each
    // one round of tiles, so pick the tile width that makes that round shortest: K blocks x 12 instructions at 48 / 64 /
    // 128 cycles (N = 64 / 128 / 256, tools/umma_bench) plus an epilogue of ~60 cycles per column, times the rounds the
    // tiles need on the SMs there are.  FUNASR_B200_GEMM_TN=256 keeps the wide tiles (comparison aid).
    if (n_planes == 2 && w.has64 && !ep.amax_val && !ep.cand_list) {
        const char* tn_env = getenv("FUNASR_B200_GEMM_TN");
        int best_tn = BN;
        double best_cost = 0.0;
        const int tns[3] = {256, 128, 64}, cyc[3] = {128, 64, 48};
        for (int i = 0; i < 3; ++i) {
            const int t = cdiv(m, BM) * cdiv(n, tns[i]);
            const double cost = (double)cdiv(t, g_num_sms) * (cdiv(k, BK) * 12.0 * cyc[i] + 60.0 * tns[i]);
            if (i == 0 || cost < best_cost) { best_cost = cost; best_tn = tns[i]; }
        }
        if (tn_env) best_tn = atoi(tn_env);
        if (best_tn == 128 || best_tn == 64) {
            const int t = cdiv(m, BM) * cdiv(n, best_tn);
            const int g = t < g_num_sms ? t : g_num_sms;
            if (best_tn == 128) {
                FA_LAUNCH((k_gemm_tc<2, 128>), g, kThreads, (Cfg<2, 128>::kSmemBytes), st, a.map, w.map128, map_out, map_pl, m, n, k, band, ep);
            } else {
                FA_LAUNCH((k_gemm_tc<2, 64>), g, kThreads, (Cfg<2, 64>::kSmemBytes), st, a.map, w.map64, map_out, map_pl, m, n, k, band, ep);
            }
            return;
        }
    }
    const int tiles = cdiv(m, BM) * cdiv(n, BN);
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    if (n_planes == 2) {
        FA_LAUNCH(k_gemm_tc<2>, grid, kThreads, Cfg<2>::kSmemBytes, st, a.map, w.map, map_out, map_pl, m, n, k, band, ep);
    } else {
        FA_LAUNCH(k_gemm_tc<1>, grid, kThreads, Cfg<1>::kSmemBytes, st, a.map, w.map, map_out, map_pl, m, n, k, band, ep);
    }
}

}  // namespace fa
