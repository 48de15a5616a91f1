// Front end: waveform -> masked, scaled, position-encoded LFR features.
//
// Follows EncoderExportWrapperPaddable.forward steps 0-3 (fun_asr_gguf/model_definition.py:269-311)
// and STFT_Process (model_definition.py:244-256).  The reference has no Kaldi fbank and no CMVN
// (SURVEY F6): it removes the mean of the valid samples, applies a global pre-emphasis, takes a
// 400-point windowed DFT as two stride-160 correlations with fp32 tables, squares, applies an
// 80x201 HTK mel matrix, log(x+1e-7), then stacks 7 frames every 6 with replicate padding.
//
// All of it is HBM-bound except the DFT, which is a [frames x 400] x [400 x 402] product done on
// the CUDA cores in fp32 against the reference's own (inexact-twiddle) tables; an FFT would
// differ from the reference by the tables' rounding error (~1e-4 relative), so it is not used.
#include "kernels.h"

namespace fa {

namespace {

constexpr int kFT = 32;                                  // mel frames per CTA
constexpr int kYLen = kFT * kHop + (kNfft - kHop);       // 5360 samples feed 32 frames
constexpr int kRiLd = 404;                               // re|im row stride in smem
constexpr int kPwLd = 204;
constexpr int kFbankThreads = 416;                       // 13 warps: one DFT output column per thread (402 used)
constexpr size_t kFbankSmem = sizeof(float) * (kYLen + kFT * kRiLd + kFT * kPwLd);

__global__ void __launch_bounds__(256)
k_segment_sums(const float* __restrict__ audio, int64_t s_phys, const int* __restrict__ n_valid,
               double* __restrict__ partials) {
    grid_dependency_wait();
    const int b = blockIdx.y, part = blockIdx.x;
    const int nv = n_valid[b];
    const int64_t chunk = (nv + kMeanParts - 1) / kMeanParts;
    const int64_t lo = part * chunk, hi = min((int64_t)nv, lo + chunk);
    const float* x = audio + (int64_t)b * s_phys;
    double acc = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += (double)x[i];
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        partials[b * kMeanParts + part] = s;
    }
}

__global__ void __launch_bounds__(kFbankThreads)
k_fbank(const float* __restrict__ audio, int64_t s_phys, const int* __restrict__ n_valid,
        const double* __restrict__ partials, const float* __restrict__ dft_t, const float* __restrict__ melfb_t,
        const int* __restrict__ mel_range, float* __restrict__ logmel, int t_mel) {
    grid_dependency_wait();
    extern __shared__ __align__(16) float smem[];
    float* y_s = smem;                       // [kYLen]
    float* ri_s = y_s + kYLen;               // [kFT][kRiLd]
    float* pw_s = ri_s + kFT * kRiLd;        // [kFT][kPwLd]
    const int b = blockIdx.y, f0 = blockIdx.x * kFT, tid = threadIdx.x;
    const int nv = n_valid[b];
    const float* x = audio + (int64_t)b * s_phys;

    // mean over the valid samples (fixed-order sum of the partials => run-to-run identical)
    double tot = 0.0;
    for (int i = 0; i < kMeanParts; ++i) tot += partials[b * kMeanParts + i];
    const float mean = __fdiv_rn((float)tot, (float)nv);

    // y = pre-emphasised, mean-removed, masked signal; zero outside [0, s_phys) (centre padding)
    const int64_t s0 = (int64_t)f0 * kHop - kNfft / 2;
    for (int i = tid; i < kYLen; i += kFbankThreads) {
        const int64_t n = s0 + i;
        float y = 0.f;
        if (n >= 0 && n < s_phys && n < nv) {
            const float a = __fsub_rn(x[n], mean);
            if (n == 0) {
                y = a;
            } else {
                const float ap = __fsub_rn(x[n - 1], mean);       // n-1 < nv holds here
                y = __fsub_rn(a, __fmul_rn(0.97f, ap));
            }
        }
        y_s[i] = y;
    }
    __syncthreads();

    // windowed DFT as a [32 frames x 400] x [400 x 402] product: thread = 4 adjacent output columns (cos bins
    // 0..200, then -sin bins 0..200) x 8 frames.  Per 4 samples that is 4 coalesced 16-byte table loads
    // (requested one step ahead of their use), 8 broadcast 16-byte loads of the signal and 128 FMAs.
    constexpr int kQuads = (2 * kBins + 3) / 4;          // 101 column quads (the table is zero-padded to 416 columns)
    if (tid < kQuads * 4) {
        const int q = tid % kQuads, fg = tid / kQuads;
        float acc[8][4];
#pragma unroll
        for (int f = 0; f < 8; ++f)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[f][c] = 0.f;
        const float4* wq = reinterpret_cast<const float4*>(dft_t) + q;
        constexpr int kLd4 = kDftLd / 4;
        const float* ybase = y_s + fg * 8 * kHop;
        float4 w0 = wq[0], w1 = wq[kLd4], w2 = wq[2 * kLd4], w3 = wq[3 * kLd4];
        for (int n = 0; n < kNfft; n += 4) {
            float4 x0 = w0, x1 = w1, x2 = w2, x3 = w3;
            if (n + 4 < kNfft) {
                x0 = wq[(n + 4) * kLd4]; x1 = wq[(n + 5) * kLd4]; x2 = wq[(n + 6) * kLd4]; x3 = wq[(n + 7) * kLd4];
            }
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const float4 yv = *reinterpret_cast<const float4*>(&ybase[f * kHop + n]);
                acc[f][0] = fmaf(yv.x, w0.x, acc[f][0]); acc[f][1] = fmaf(yv.x, w0.y, acc[f][1]);
                acc[f][2] = fmaf(yv.x, w0.z, acc[f][2]); acc[f][3] = fmaf(yv.x, w0.w, acc[f][3]);
                acc[f][0] = fmaf(yv.y, w1.x, acc[f][0]); acc[f][1] = fmaf(yv.y, w1.y, acc[f][1]);
                acc[f][2] = fmaf(yv.y, w1.z, acc[f][2]); acc[f][3] = fmaf(yv.y, w1.w, acc[f][3]);
                acc[f][0] = fmaf(yv.z, w2.x, acc[f][0]); acc[f][1] = fmaf(yv.z, w2.y, acc[f][1]);
                acc[f][2] = fmaf(yv.z, w2.z, acc[f][2]); acc[f][3] = fmaf(yv.z, w2.w, acc[f][3]);
                acc[f][0] = fmaf(yv.w, w3.x, acc[f][0]); acc[f][1] = fmaf(yv.w, w3.y, acc[f][1]);
                acc[f][2] = fmaf(yv.w, w3.z, acc[f][2]); acc[f][3] = fmaf(yv.w, w3.w, acc[f][3]);
            }
            w0 = x0; w1 = x1; w2 = x2; w3 = x3;
        }
#pragma unroll
        for (int f = 0; f < 8; ++f)
            *reinterpret_cast<float4*>(&ri_s[(fg * 8 + f) * kRiLd + q * 4]) = make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
    }
    __syncthreads();

    for (int i = tid; i < kFT * kBins; i += kFbankThreads) {
        const int f = i / kBins, k = i - f * kBins;
        const float re = ri_s[f * kRiLd + k], im = ri_s[f * kRiLd + kBins + k];
        pw_s[f * kPwLd + k] = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    }
    __syncthreads();

    // mel + log: thread = (mel filter j, group of 8 frames); one filter weight feeds 8 independent accumulators,
    // and only the filter's support is walked (bins in ascending order, like a dense dot product would)
    if (tid < kMels * (kFT / 8)) {
        const int j = tid % kMels, fg = tid / kMels;
        const int k0 = mel_range[2 * j], k1 = mel_range[2 * j + 1];
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const float* p = pw_s + fg * 8 * kPwLd;
        for (int k = k0; k < k1; ++k) {
            const float w = melfb_t[k * kMels + j];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, p[i * kPwLd + k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int f = fg * 8 + i;
            if (f0 + f < t_mel) logmel[((int64_t)b * t_mel + f0 + f) * kMels + j] = logf(__fadd_rn(acc[i], 1e-7f));
        }
    }
}

__global__ void __launch_bounds__(160)
k_lfr_embed(const float* __restrict__ logmel, int t_mel, int t_lfr, const int* __restrict__ n_valid,
            const float* __restrict__ pos_enc, float* __restrict__ x0, float* __restrict__ lfr_raw,
            const int* __restrict__ seg_off) {
    grid_dependency_wait();
    const int t = blockIdx.x, b = blockIdx.y, c4 = threadIdx.x;
    if (c4 >= kDin / 4) return;
    const int nv = n_valid[b];
    const int t_mel_valid = nv / kHop + 1;
    const int t_valid = (t_mel_valid + kLfrN - 1) / kLfrN;
    if (seg_off && t >= t_valid) return;                       // packed layout: padded frames have no row
    const int col = c4 * 4, slot = col / kMels, j = col - slot * kMels;
    int src = t * kLfrN + slot - (kLfrM - 1) / 2;
    src = max(0, min(src, t_mel - 1));
    src = min(src, t_mel_valid - 1);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < t_valid) v = *reinterpret_cast<const float4*>(&logmel[((int64_t)b * t_mel + src) * kMels + j]);
    const int64_t o = (seg_off ? (int64_t)seg_off[b] + t : (int64_t)b * t_lfr + t) * kDin + col;
    if (lfr_raw) *reinterpret_cast<float4*>(&lfr_raw[o]) = v;
    const float s = 22.627416610717773f;      // fp32(512 ** 0.5), model_definition.py:206
    const float4 pe = *reinterpret_cast<const float4*>(&pos_enc[(int64_t)t * kDin + col]);
    float4 r;
    r.x = __fadd_rn(__fmul_rn(v.x, s), pe.x);
    r.y = __fadd_rn(__fmul_rn(v.y, s), pe.y);
    r.z = __fadd_rn(__fmul_rn(v.z, s), pe.z);
    r.w = __fadd_rn(__fmul_rn(v.w, s), pe.w);
    *reinterpret_cast<float4*>(&x0[o]) = r;
}

}  // namespace

void frontend_init_device() {
    FA_CUDA(cudaFuncSetAttribute(k_fbank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFbankSmem));
}

void launch_segment_sums(const float* audio, int batch, int64_t s_phys, const int* n_valid, double* partials,
                         cudaStream_t st) {
    FA_LAUNCH(k_segment_sums, dim3(kMeanParts, batch), 256, 0, st, audio, s_phys, n_valid, partials);
}

void launch_fbank(const float* audio, int batch, int64_t s_phys, const int* n_valid, const double* partials,
                  const float* dft_t, const float* melfb_t, const int* mel_range, float* logmel, int t_mel, cudaStream_t st) {
    FA_LAUNCH(k_fbank, dim3(cdiv(t_mel, kFT), batch), kFbankThreads, kFbankSmem, st, audio, s_phys, n_valid, partials,
              dft_t, melfb_t, mel_range, logmel, t_mel);
}

void launch_lfr_embed(const float* logmel, int batch, int t_mel, int t_lfr, const int* n_valid, const float* pos_enc,
                      float* x0, float* lfr_raw, cudaStream_t st, const int* seg_off) {
    FA_LAUNCH(k_lfr_embed, dim3(t_lfr, batch), 160, 0, st, logmel, t_mel, t_lfr, n_valid, pos_enc, x0, lfr_raw, seg_off);
}

}  // namespace fa
