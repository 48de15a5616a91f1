// Host-callable launchers for every kernel on the path.  One entry per row of SURVEY §8a.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace fa {

// Per-device one-time setup (opt-in shared memory sizes, driver entry points).  Call with the
// device current, once per context.
void frontend_init_device();
void attention_init_device();
void attention_tc_init_device();
void tc_init_device();

// ---------------------------------------------------------------- front end (§8a a1-a4)
// Per-segment sum of the valid samples, as kMeanParts double partials per segment.
constexpr int kMeanParts = 64;
void launch_segment_sums(const float* audio, int batch, int64_t s_phys, const int* n_valid, double* partials,
                         cudaStream_t st);
// mean removal + pre-emphasis + framing + windowed DFT + power + mel + log.  logmel: [B][T_mel][80].
void launch_fbank(const float* audio, int batch, int64_t s_phys, const int* n_valid, const double* partials,
                  const float* dft_t /*[400][kDftLd] cos|sin transposed*/, const float* melfb /*[201][80] transposed*/,
                  const int* mel_range /*[80][2] first / last+1 non-zero bin of each filter*/,
                  float* logmel, int t_mel, cudaStream_t st);
constexpr int kDftLd = 416;   // 402 real columns (201 cos + 201 -sin), padded
// LFR stacking with replicate padding, frame mask, sqrt(512) scale and positional table.
void launch_lfr_embed(const float* logmel, int batch, int t_mel, int t_lfr, const int* n_valid, const float* pos_enc,
                      float* x0 /*[B*T][560]*/, float* lfr_raw /*optional tap [B*T][560]*/, cudaStream_t st);

// ---------------------------------------------------------------- row kernels (§8a a5, a7, a12)
struct Planes {           // bf16 hi/lo planes of an [M][ld] activation; lo may be null (bf16 mode)
    __nv_bfloat16* hi = nullptr;
    __nv_bfloat16* lo = nullptr;
};
// y = LN(x) * gamma + beta (optionally zeroing rows t >= t_valid[b]); writes fp32 and/or planes.
void launch_layernorm(const float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                      const int* t_valid /*nullable*/, int frames, float* y_f32, Planes y_pl, cudaStream_t st);
// FSMN memory block: out = (resid ? resid : 0) + depthwise_conv11(v*m) + v*m.   v has row stride ldv.
void launch_fsmn(const float* v, int ldv, const float* w /*[512][11]*/, const int* t_valid, int batch, int frames,
                 const float* resid, float* out, cudaStream_t st);
// out[b,t,:] = t < keep[b] ? in[b,t,:] : 0
void launch_row_keep(const float* in, float* out, int batch, int frames, int d, const int* keep, cudaStream_t st);
// fp32 -> planes
void launch_split_planes(const float* x, int64_t n, Planes out, cudaStream_t st);
// ids[r] = first argmax over logits[r][0..n)
void launch_argmax_rows(const float* logits, int rows, int n, int ld, int32_t* ids, cudaStream_t st);
// combine per-tile (max, idx) partials written by the fused vocabulary GEMM epilogue
void launch_argmax_combine(const float* pmax, const int32_t* pidx, int rows, int tiles, int32_t* ids, cudaStream_t st);
// greedy collapse: ids [B][T] -> per segment (token, start_frame) pairs, blanks dropped (nano_ctc.py:70-99)
void launch_ctc_collapse(const int32_t* ids, int batch, int frames, int blank, int32_t* tokens, int32_t* starts,
                         int32_t* counts, cudaStream_t st);

// ---------------------------------------------------------------- dense projections (§8a a6, a9, a11, a13, a14)
struct Epilogue {
    const float* bias = nullptr;     // [N]
    const float* resid = nullptr;    // [M][ldr] added after bias
    int ldr = 0;
    bool relu = false;
    float* out_f32 = nullptr;        // [M][ldc]
    int ldc = 0;
    Planes out_pl;                   // [M][ldp] planes
    int ldp = 0;
    float pl_col_scale = 1.f;        // plane outputs of columns < pl_col_scale_end are multiplied by this
    int pl_col_scale_end = 0;        //   (folds d_k^-0.5 * log2 e into the q planes the attention kernel reads)
    int f32_col_begin = 0;           // fp32 output only for columns >= this (multiple of 32)
    // fused vocabulary argmax: per (row, n-tile) partial max/idx instead of the logits
    float* amax_val = nullptr;       // [M][n_tiles]
    int32_t* amax_idx = nullptr;
};
// C = A[M][K] * W[N][K]^T  in fp32 on the CUDA cores (exact-precision mode and on-device arbiter).
void launch_gemm_simt(const float* a, int lda, const float* w, int m, int n, int k, const Epilogue& ep, cudaStream_t st);

// tcgen05 GEMM on bf16 planes. a_planes / w_planes: plane-major [P][rows][K].  n_planes 2 = bf16x3, 1 = bf16.
struct TcOperand {
    CUtensorMap map;        // 3D: {K, rows, planes}, box {64, box_rows, 1}, SWIZZLE_128B
    int rows = 0, k = 0, planes = 0;
    CUtensorMap map64;      // weights only: the same tensor with 64-row boxes (CTA-pair kernel)
    bool has64 = false;
};
TcOperand tc_make_operand(const __nv_bfloat16* base, int rows, int k, int64_t row_stride_elems, int64_t plane_stride_elems,
                          int planes, int box_rows);
// a weight matrix [planes][rows][k], contiguous rows: both box shapes, so either GEMM kernel can read it
TcOperand tc_make_weight(const __nv_bfloat16* base, int rows, int k, int64_t plane_stride_elems, int planes);
constexpr int kTcBlockM = 128, kTcBlockN = 256, kTcBlockK = 64;
void launch_gemm_tc(const TcOperand& a, const TcOperand& w, int m, int n, int k, int n_planes, const Epilogue& ep,
                    cudaStream_t st);
int tc_argmax_tiles(int n);

// ---------------------------------------------------------------- attention (§8a a8)
// q,k,v: fp32 [B*T][ld], head h at column offset h*dk of each pointer.  kv_len[b] keys are attended
// (nullable => all `frames`).  Output ctx [B*T][ldo] fp32 and/or planes.
void launch_attention_simt(const float* q, const float* k, const float* v, int ld, int batch, int frames, int heads,
                           int dk, const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st);

// tcgen05 attention on bf16 planes of a fused [rows][ld] q|k|v matrix (q pre-scaled by d_k^-0.5 * log2 e):
// q at column 0, k at d_model, v at 2*d_model; head h at +h*dk.  Two planes `plane_stride` elements apart.
void launch_attention_tc(Planes qkv, int64_t plane_stride, int ld, int d_model, int batch, int frames, int heads, int dk,
                         const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st);

}  // namespace fa
