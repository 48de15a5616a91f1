// Front-end DFT on the tensor cores (STFT_Process, model_definition.py:244-256).
//
// The windowed 400-point DFT of every frame is a [frames x 400] x [400 x 402] product.  k_fbank (frontend.cu) does it
// in fp32 on the CUDA cores at ~40 % of the FMA pipe, 3 ms for 32 x 60 s; here it is three tcgen05 products of fp16
// hi/lo planes into fp32 TMEM accumulators:
//     y_lo t_hi + y_hi t_lo + y_hi t_hi,   y = 2^10 * signal,  t = 2^8 * table
// fp16 planes carry 22 mantissa bits (the scaling keeps the lo planes out of the subnormal range for every sample
// and table entry that matters), so the dropped lo*lo term is 2^-22 per product — below the rounding a 400-term
// fp32 sum accumulates anyway (the reference's own conv1d included).
//
//   k_frame_planes   audio -> mean-removed, pre-emphasised, masked signal (exactly k_fbank's arithmetic), scaled and
//                    split into planes, with the 200-sample centre padding materialised: frame f then starts at
//                    element 160 f, so consecutive rows of the A operand simply OVERLAP in memory and one tensor
//                    map {K = 448, frames (stride 160 elements), segment, plane} feeds TMA with no im2col copy
//   k_dft_tc         persistent, one CTA per SM: warp 0 TMA, warp 1 MMA issuer (M128 x N208 x K16, kind::f16), warps
//                    2..5 epilogue.  A tile is 128 frames x 416 table columns (re/im of a bin interleaved: column 2k
//                    and 2k+1) in two 208-column accumulators; per (accumulator, 64-sample K-block) stage: 2 planes of
//                    A (32 KB) + 2 planes of the table slice (52 KB), two stages.  Epilogue: re^2 + im^2, unscaled,
//                    to a [frames][208] power buffer.
//   k_mel_log        80 x 201 HTK mel matrix over each filter's support (ascending bins, like the dense product) and
//                    log(x + 1e-7): the tail of k_fbank as its own small kernel
#include "kernels.h"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <mutex>

namespace fa {

namespace {

constexpr int kFtM = 128;                       // frames per tile (UMMA M)
constexpr int kFtN = 208;                       // table columns per accumulator (UMMA N): 2 x 208 = 416 >= 402
constexpr int kFtK = 448;                       // 400 samples padded to 7 K-blocks of 64 (table rows 400.. are zero)
constexpr int kFtKBlocks = kFtK / 64;
constexpr int kFtATile = kFtM * 128;            // one plane of A: 128 rows x 64 fp16 = 16 KB
constexpr int kFtBTile = kFtN * 128;            // one plane of a table slice: 208 rows x 64 fp16 = 26 KB
constexpr int kFtStage = 2 * kFtATile + 2 * kFtBTile;     // 84 KB
constexpr int kFtStages = 2;
constexpr int kFtSmem = kFtStages * kFtStage + 1024 + 128;
constexpr int kFtThreads = 192;                 // TMA warp, MMA warp, 4 epilogue warps
constexpr float kSigScale = 1024.f, kTabScale = 256.f;
constexpr float kPowUnscale = 1.0f / (kSigScale * kSigScale * kTabScale * kTabScale);     // 2^-36, exact

__global__ void __launch_bounds__(256)
k_frame_planes(const float* __restrict__ audio, int64_t s_phys, const int* __restrict__ n_valid,
               const double* __restrict__ partials, __half* __restrict__ hi, __half* __restrict__ lo, int64_t sp) {
    grid_dependency_wait();
    const int b = blockIdx.y;
    __shared__ float mean_s;
    if (threadIdx.x == 0) {
        // mean over the valid samples: the same fixed-order sum of the partials as k_fbank
        double tot = 0.0;
        for (int k = 0; k < kMeanParts; ++k) tot += partials[b * kMeanParts + k];
        mean_s = __fdiv_rn((float)tot, (float)n_valid[b]);
    }
    __syncthreads();
    const float mean = mean_s;
    const int nv = n_valid[b];
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;     // 8 samples per thread, 16-byte stores
    if (i0 >= sp) return;
    const float* x = audio + (int64_t)b * s_phys;
    const int64_t lim = s_phys < nv ? s_phys : nv;                               // samples at or past it are masked
    const int64_t n0 = i0 - kNfft / 2;                                           // a multiple of 8: the loads are aligned
    float xs[9];                                                                 // x[n0 - 1 .. n0 + 7] - mean (0 outside)
    if (n0 >= 0 && n0 + 8 <= lim && ((reinterpret_cast<uintptr_t>(x + n0) & 15) == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(x + n0), c = *reinterpret_cast<const float4*>(x + n0 + 4);
        xs[0] = n0 > 0 ? __fsub_rn(x[n0 - 1], mean) : 0.f;
        xs[1] = __fsub_rn(a.x, mean); xs[2] = __fsub_rn(a.y, mean); xs[3] = __fsub_rn(a.z, mean); xs[4] = __fsub_rn(a.w, mean);
        xs[5] = __fsub_rn(c.x, mean); xs[6] = __fsub_rn(c.y, mean); xs[7] = __fsub_rn(c.z, mean); xs[8] = __fsub_rn(c.w, mean);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int64_t n = n0 - 1 + k;
            xs[k] = (n >= 0 && n < lim) ? __fsub_rn(x[n], mean) : 0.f;
        }
    }
    __align__(16) __half h8[8], l8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t n = n0 + k;
        float y = 0.f;
        if (n >= 0 && n < lim) y = n == 0 ? xs[k + 1] : __fsub_rn(xs[k + 1], __fmul_rn(0.97f, xs[k]));
        const float ys = y * kSigScale;
        h8[k] = __float2half_rn(ys);
        l8[k] = __float2half_rn(ys - __half2float(h8[k]));
    }
    *reinterpret_cast<uint4*>(hi + (int64_t)b * sp + i0) = *reinterpret_cast<const uint4*>(h8);
    *reinterpret_cast<uint4*>(lo + (int64_t)b * sp + i0) = *reinterpret_cast<const uint4*>(l8);
}

// dft_t: [400][kDftLd] fp32, columns 0..200 cos bins, 201..401 -sin bins.  Output planes [2][416][448] fp16 with
// table column 2k = cos bin k, 2k+1 = sin bin k (k <= 200), everything else zero.
__global__ void __launch_bounds__(256)
k_table_planes(const float* __restrict__ dft_t, __half* __restrict__ hi, __half* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * kFtN * kFtK) return;
    const int col = i / kFtK, n = i - col * kFtK;
    const int bin = col >> 1, part = col & 1;
    float t = 0.f;
    if (bin < kBins && n < kNfft) t = dft_t[(int64_t)n * kDftLd + part * kBins + bin] * kTabScale;
    const __half h = __float2half_rn(t);
    hi[i] = h;
    lo[i] = __float2half_rn(t - __half2float(h));
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__global__ void __launch_bounds__(kFtThreads, 1)
k_dft_tc(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_t, int batch, int t_mel,
         float* __restrict__ power /*[batch * t_mel][kFtN]*/) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles_base = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles_base + kFtStages * kFtStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kFtStages;
    const uint32_t bar_tfull = bars + 16 * kFtStages, bar_tempty = bar_tfull + 8;
    const uint32_t tmem_slot = bar_tempty + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m_tiles = (t_mel + kFtM - 1) / kFtM, tiles = batch * m_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFtStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_t) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
    grid_dependency_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int b = tile / m_tiles, f0 = (tile - b * m_tiles) * kFtM;
            for (int nt = 0; nt < 2; ++nt) {
                for (int kb = 0; kb < kFtKBlocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sb = tiles_base + stage * kFtStage, full = bar_full + 8 * stage;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full, kFtStage);
                        tma_load_4d(sb, &map_y, full, kb * 64, f0, b, 0);
                        tma_load_4d(sb + kFtATile, &map_y, full, kb * 64, f0, b, 1);
                        tma_load_3d(sb + 2 * kFtATile, &map_t, full, kb * 64, nt * kFtN, 0);
                        tma_load_3d(sb + 2 * kFtATile + kFtBTile, &map_t, full, kb * 64, nt * kFtN, 1);
                    }
                    __syncwarp();
                    if (++stage == kFtStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // D = f32, A = B = f16 (format 0), K-major both, N = 208, M = 128
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(kFtN >> 3) << 17) | ((uint32_t)(kFtM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            mbar_wait(bar_tempty, (it & 1) ^ 1);                     // the epilogue has read the previous tile
            tc_fence_after();
            for (int nt = 0; nt < 2; ++nt) {
                const uint32_t tmem_d = tmem_base + nt * 256;
                for (int kb = 0; kb < kFtKBlocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_hi = tiles_base + stage * kFtStage, a_lo = a_hi + kFtATile;
                    const uint32_t t_hi = a_hi + 2 * kFtATile, t_lo = t_hi + kFtBTile;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma(tmem_d, umma_desc(a_lo + ks * 32, 16, 1024), umma_desc(t_hi + ks * 32, 16, 1024), idesc,
                                   (kb | ks) ? 1u : 0u);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma(tmem_d, umma_desc(a_hi + ks * 32, 16, 1024), umma_desc(t_lo + ks * 32, 16, 1024), idesc, 1u);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma(tmem_d, umma_desc(a_hi + ks * 32, 16, 1024), umma_desc(t_hi + ks * 32, 16, 1024), idesc, 1u);
                        tc_commit(bar_empty + 8 * stage);
                        if (nt == 1 && kb == kFtKBlocks - 1) tc_commit(bar_tfull);
                    }
                    __syncwarp();
                    if (++stage == kFtStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: power spectrum of a frame row
        const int quarter = warp & 3;                               // TMEM lane quarter = warp id % 4
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = tile / m_tiles, f0 = (tile - b * m_tiles) * kFtM;
            const int f = f0 + quarter * 32 + lane;
            float* prow = power + ((int64_t)b * t_mel + f) * kFtN;
            mbar_wait(bar_tfull, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll 1
                for (int c = 0; c < kFtN / 16; ++c) {               // 16 columns = 8 bins
                    uint32_t r[16];
                    tc_ld16(tmem_base + lane_addr + nt * 256 + c * 16, r);
                    tc_wait_ld();
                    if (nt == 1 && c == kFtN / 16 - 1) {            // last read of the accumulators: hand them back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty);
                    }
                    float pw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float re = __uint_as_float(r[2 * k]), im = __uint_as_float(r[2 * k + 1]);
                        pw[k] = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)) * kPowUnscale;
                    }
                    if (f < t_mel) {
                        float4* dst = reinterpret_cast<float4*>(prow + nt * (kFtN / 2) + c * 8);
                        dst[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                        dst[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// thread = (frame, mel filter): dot product over the filter's support, bins ascending, then log(x + 1e-7)
__global__ void __launch_bounds__(320)
k_mel_log(const float* __restrict__ power, int64_t frames, const float* __restrict__ melfb_t /*[201][80]*/,
          const int* __restrict__ mel_range, float* __restrict__ logmel) {
    grid_dependency_wait();
    const int j = threadIdx.x % kMels;
    const int64_t f = (int64_t)blockIdx.x * 4 + threadIdx.x / kMels;
    if (f >= frames) return;
    const int k0 = mel_range[2 * j], k1 = mel_range[2 * j + 1];
    const float* p = power + f * kFtN;
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(melfb_t[k * kMels + j], p[k], acc);
    logmel[f * kMels + j] = logf(__fadd_rn(acc, 1e-7f));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_ft_encode = nullptr;
std::once_flag g_ft_once;
int g_ft_sms = 0;

CUtensorMap ft_map(void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    CUtensorMap m;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = g_ft_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled (front end) failed with code " + std::to_string((int)r));
    return m;
}

}  // namespace

void fbank_tc_init_device() {
    std::call_once(g_ft_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        FA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        FA_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_ft_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    FA_CUDA(cudaFuncSetAttribute(k_dft_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmem));
    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    FA_CUDA(cudaDeviceGetAttribute(&g_ft_sms, cudaDevAttrMultiProcessorCount, dev));
}

int64_t fbank_tc_padded_samples(int64_t s_phys) { return (s_phys + kNfft + 64 + 7) / 8 * 8; }
size_t fbank_tc_table_bytes() { return (size_t)2 * 2 * kFtN * kFtK * sizeof(__half); }
int fbank_tc_power_ld() { return kFtN; }

void launch_fbank_table_planes(const float* dft_t, void* table_planes, cudaStream_t st) {
    __half* hi = static_cast<__half*>(table_planes);
    FA_LAUNCH(k_table_planes, cdiv(2 * kFtN * kFtK, 256), 256, 0, st, dft_t, hi, hi + (size_t)2 * kFtN * kFtK);
}

void launch_fbank_tc(const float* audio, int batch, int64_t s_phys, const int* n_valid, const double* partials,
                     const void* table_planes, const float* melfb_t, const int* mel_range, void* y_planes, float* power,
                     float* logmel, int t_mel, cudaStream_t st) {
    FA_REQUIRE(g_ft_encode != nullptr, "fbank_tc_init_device() has not run");
    const int64_t sp = fbank_tc_padded_samples(s_phys);
    __half* yh = static_cast<__half*>(y_planes);
    __half* yl = yh + (size_t)batch * sp;
    FA_LAUNCH(k_frame_planes, dim3(cdiv(sp, 256 * 8), batch), 256, 0, st, audio, s_phys, n_valid, partials, yh, yl, sp);
    // A: {K = 448 samples, frames with a stride of one hop, segments, planes}: rows overlap in memory
    const cuuint64_t ydims[4] = {(cuuint64_t)kFtK, (cuuint64_t)t_mel, (cuuint64_t)batch, 2};
    const cuuint64_t ystr[3] = {(cuuint64_t)kHop * 2, (cuuint64_t)sp * 2, (cuuint64_t)batch * sp * 2};
    const cuuint32_t ybox[4] = {64, (cuuint32_t)kFtM, 1, 1};
    const CUtensorMap map_y = ft_map(yh, 4, ydims, ystr, ybox);
    const cuuint64_t tdims[3] = {(cuuint64_t)kFtK, (cuuint64_t)(2 * kFtN), 2};
    const cuuint64_t tstr[2] = {(cuuint64_t)kFtK * 2, (cuuint64_t)2 * kFtN * kFtK * 2};
    const cuuint32_t tbox[3] = {64, (cuuint32_t)kFtN, 1};
    const CUtensorMap map_t = ft_map(const_cast<void*>(table_planes), 3, tdims, tstr, tbox);
    const int tiles = batch * cdiv(t_mel, kFtM);
    prof_note_work(2.0 * batch * (double)t_mel * kNfft * (2 * kBins), 0.0);
    FA_LAUNCH(k_dft_tc, tiles < g_ft_sms ? tiles : g_ft_sms, kFtThreads, kFtSmem, st, map_y, map_t, batch, t_mel, power);
    const int64_t frames = (int64_t)batch * t_mel;
    FA_LAUNCH(k_mel_log, cdiv(frames, 4), 320, 0, st, power, frames, melfb_t, mel_range, logmel);
}

}  // namespace fa
