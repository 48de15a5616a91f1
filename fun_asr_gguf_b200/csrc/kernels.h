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

// Padding-free ("packed") row layout of a mixed-length batch (BASELINE configs[2], model_definition_paddable path):
// segment b's valid frames are rows seg_off[b] .. seg_off[b] + t_valid[b] - 1 of every activation matrix of the encoder
// and the adaptor, so the projections, LayerNorm, the memory block and the attention queries only ever see rows the
// reference does not define as zero.  All arrays live in device memory and are staged per call.
struct Packing {
    const int* seg_off = nullptr;    // [batch + 1] first packed row of each segment
    // [total_tiles] one entry per 128-query tile, longest segments first (the order attention items are dealt in, so that
    // every CTA's share mixes long and short items): {first packed row of the segment, its length, tile index, segment}
    const int4* tile_tab = nullptr;
    int total_rows = 0, total_tiles = 0, max_len = 0;
    double sum_len_sq = 0.0;         // sum of t_valid^2: attention FLOP accounting
    // CTC head only (SURVEY §7 hard part 2): [batch] log2 of the multiplicity of each segment's LAST key.  The head is
    // unmasked over every physical frame (F7) and all zero-padded frames of a segment are identical rows through every
    // layer, so they are carried as ONE row whose key counts n_pad times: exp2(s + log2 n_pad) = n_pad * exp2(s).
    const float* last_key_bias = nullptr;
};

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
// The same front end with the DFT on the tensor cores (fbank_tc.cu): fp16 hi/lo planes, three products, fp32 TMEM
// accumulators; y_planes holds 2 * batch * fbank_tc_padded_samples(s_phys) halves, power batch * t_mel * fbank_tc_power_ld()
// floats, table_planes fbank_tc_table_bytes() bytes filled once by launch_fbank_table_planes.
void fbank_tc_init_device();
int64_t fbank_tc_padded_samples(int64_t s_phys);
size_t fbank_tc_table_bytes();
int fbank_tc_power_ld();
void launch_fbank_table_planes(const float* dft_t, void* table_planes, cudaStream_t st);
void launch_fbank_tc(const float* audio, int batch, int64_t s_phys, const int* n_valid, const double* partials,
                     const void* table_planes, const float* melfb_t, const int* mel_range, void* y_planes, float* power,
                     float* logmel, int t_mel, cudaStream_t st);
constexpr int kDftLd = 416;   // 402 real columns (201 cos + 201 -sin), padded
// LFR stacking with replicate padding, frame mask, sqrt(512) scale and positional table.
// seg_off != nullptr: only the valid frames are written, frame t of segment b to packed row seg_off[b] + t.
void launch_lfr_embed(const float* logmel, int batch, int t_mel, int t_lfr, const int* n_valid, const float* pos_enc,
                      float* x0 /*[B*T][560]*/, float* lfr_raw /*optional tap [B*T][560]*/, cudaStream_t st,
                      const int* seg_off = nullptr);

// ---------------------------------------------------------------- row kernels (§8a a5, a7, a12)
struct Planes {           // bf16 hi/lo planes of an [M][ld] activation; lo may be null (bf16 mode)
    __nv_bfloat16* hi = nullptr;
    __nv_bfloat16* lo = nullptr;
};
// y = LN(x) * gamma + beta (optionally zeroing rows t >= t_valid[b]); writes fp32 and/or planes.
// seg_off != nullptr ("unpack"): x is packed, the outputs are physical [batch * frames] rows: row (b, t) is
// LN(x[seg_off[b] + t]) for t < t_valid[b] and zero otherwise (rows = batch * frames, t_valid required).
void launch_layernorm(const float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                      const int* t_valid /*nullable*/, int frames, float* y_f32, Planes y_pl, cudaStream_t st,
                      const int* seg_off = nullptr, uint8_t* y_f8 = nullptr);
// FSMN memory block: out = (resid ? resid : 0) + depthwise_conv11(v*m) + v*m.   v has row stride ldv.
// pk != nullptr: packed rows; segment b is rows seg_off[b] .. + t_valid[b] - 1 (`frames` is then ignored).
void launch_fsmn(const float* v, int ldv, const float* w /*[512][11]*/, const int* t_valid, int batch, int frames,
                 const float* resid, float* out, cudaStream_t st, const Packing* pk = nullptr);
// out[b,t,:] = t < keep[b] ? in[b,t,:] : 0
// seg_off != nullptr: `in` is packed (row (b, t) read from seg_off[b] + t), `out` physical
void launch_row_keep(const float* in, float* out, int batch, int frames, int d, const int* keep, cudaStream_t st,
                     const int* seg_off = nullptr);
// CTC-head packing: copy rows of `row_bytes` bytes from the encoder's packed layout (segment b at src_off[b], tv[b] rows)
// to the head's (segment b at dst_off[b], tv[b] rows followed by one all-zero row if the segment has padded frames)
void launch_repack_rows(const void* src, void* dst, int row_bytes, const int* src_off, const int* dst_off, const int* tv,
                        int batch, int max_len, cudaStream_t st);
// ids of the head's packed rows -> physical [batch][frames]: frame t >= tv[b] takes the id of the segment's pad row
// tphys != nullptr (ragged batches): frames t >= tphys[b] get -1
void launch_unpack_ids(const int32_t* packed, int32_t* ids, int batch, int frames, const int* off, const int* tv, cudaStream_t st,
                       const int* tphys = nullptr);
// fp32 -> planes
void launch_split_planes(const float* x, int64_t n, Planes out, cudaStream_t st);
// fp32 -> e4m3 (round to nearest even, saturating)
void launch_to_e4m3(const float* x, int64_t n, uint8_t* out, cudaStream_t st);
// per-output-channel weight quantisation (02-Quantize-ONNX.py:41-44 per_channel=True): scale[r] = max|w[r]| / 448,
// w8[r] = e4m3(w[r] / scale[r])
void launch_quant_rows_e4m3(const float* w, int rows, int k, uint8_t* w8, float* scale, cudaStream_t st);
// ids[r] = first argmax over logits[r][0..n)
void launch_argmax_rows(const float* logits, int rows, int n, int ld, int32_t* ids, cudaStream_t st);
// combine per-tile (max, idx) partials written by the fused vocabulary GEMM epilogue
// only_if_over: write only the rows whose only_if_over[row] > over (second-chance pass of the candidate path)
void launch_argmax_combine(const float* pmax, const int32_t* pidx, int rows, int tiles, int32_t* ids, cudaStream_t st,
                           const int32_t* only_if_over = nullptr, int over = 0);
// greedy collapse: ids [B][T] -> per segment (token, start_frame) pairs, blanks dropped (nano_ctc.py:70-99)
void launch_ctc_collapse(const int32_t* ids, int batch, int frames, int blank, int32_t* tokens, int32_t* starts,
                         int32_t* counts, cudaStream_t st);

// ---------------------------------------------------------------- dense projections (§8a a6, a9, a11, a13, a14)
// Candidate lists of the vocabulary argmax (see launch_vocab_prepare / launch_vocab_rescore).
constexpr int kVocabCandCap = 192;   // list slots per row; a row that needs more is rescored over the whole vocabulary
struct VocabCand {
    int32_t* run_max = nullptr;      // [M] running approximate row maximum, order-preserving int encoding of the float
    int32_t* count = nullptr;        // [M] candidates appended so far (may exceed cap: the row then takes the second-chance pass)
    int32_t* overflowed = nullptr;   // [1] rows whose list overflowed; gates the second-chance pass
    int2* list = nullptr;            // [M][cap] (column, bits of its approximate logit), in no particular order
    const float* bound2 = nullptr;   // [M] twice the worst-case error of a one-product logit of that row
    int cap = 0;
};
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
    // vocabulary candidates (one-product pass): instead of logits or partial maxima, every column whose approximate
    // logit is within bound2[row] of the running approximate row maximum is appended to the row's list
    VocabCand cand;
    // the whole launch is a no-op when *gate == 0 (read on the device: no host round trip)
    const int32_t* gate = nullptr;
    // fp8 speed mode (FA_PREC_FP8): operands are e4m3 bytes; the accumulator is multiplied by the per-output-channel
    // weight scale before the bias, and the output may leave as e4m3 for the next projection
    bool f8 = false;
    const float* col_scale = nullptr; // [N]
    uint8_t* out_f8 = nullptr;        // [M][ld8]
    int ld8 = 0;
};
// C = A[M][K] * W[N][K]^T  in fp32 on the CUDA cores (exact-precision mode and on-device arbiter).
void launch_gemm_simt(const float* a, int lda, const float* w, int m, int n, int k, const Epilogue& ep, cudaStream_t st);

// tcgen05 GEMM on bf16 planes. a_planes / w_planes: plane-major [P][rows][K].  n_planes 2 = bf16x3, 1 = bf16.
struct TcOperand {
    CUtensorMap map;        // 3D: {K, rows, planes}, box {64, box_rows, 1}, SWIZZLE_128B
    int rows = 0, k = 0, planes = 0;
    CUtensorMap map64;      // weights only: the same tensor with 64-row boxes (CTA-pair kernel, narrow tiles of the single-CTA kernel)
    CUtensorMap map128;     // weights only: 128-row boxes (128-column tiles of the single-CTA kernel)
    bool has64 = false;
};
TcOperand tc_make_operand(const __nv_bfloat16* base, int rows, int k, int64_t row_stride_elems, int64_t plane_stride_elems,
                          int planes, int box_rows);
// e4m3 operand [rows][k] bytes (one plane); box_rows 128 for activations, 64 for weights
TcOperand tc_make_operand_f8(const uint8_t* base, int rows, int k, int64_t row_stride_bytes, int box_rows);
// a weight matrix [planes][rows][k], contiguous rows: both box shapes, so either GEMM kernel can read it
TcOperand tc_make_weight(const __nv_bfloat16* base, int rows, int k, int64_t plane_stride_elems, int planes);
constexpr int kTcBlockM = 128, kTcBlockN = 256, kTcBlockK = 64;
void launch_gemm_tc(const TcOperand& a, const TcOperand& w, int m, int n, int k, int n_planes, const Epilogue& ep,
                    cudaStream_t st);
int tc_argmax_tiles(int n);
int tc_num_pairs();      // CTA pairs of the tcgen05 GEMM that run at once (74 on a full B200)

// Vocabulary argmax in two steps (CTC.ctc_lo + torch.argmax, model_definition.py:216-221,337):
//   1. the projection runs with ONE bf16 product per element; its worst-case error against the fp32 logit is
//      b_r = 2^-8 * 1.05 * |x_r|_2 * max_c |w_c|_2, so every column that can hold the true maximum has an approximate
//      logit within 2 b_r of the approximate maximum and lands in the row's candidate list;
//   2. the candidates (a handful per row) are rescored with an fp32 dot product and the first maximal index wins.
// prepare: planes of x, |x_r|, bound2, list reset.  rescore: ids of the rows whose list did not overflow.  The rest
// (none, unless a row has more than kVocabCandCap near-maximal columns) take the second-chance pass: the
// three-product projection with the fused running argmax, launched behind a device-side gate so that it costs a few
// microseconds when no row needs it.
void launch_vocab_prepare(const float* x, int rows, int d, float w_norm_max, Planes x_pl, VocabCand c, cudaStream_t st);
void launch_vocab_rescore(const float* x, const float* w, const float* bias, int rows, int d, int n, VocabCand c, int32_t* ids,
                          cudaStream_t st);
// the three-product projection's partial maxima (one slot per 128 columns) decided by the same fp32 rescoring: every
// column of the slots within the three-product error bound of the row's best is rescored, first maximal index wins
void launch_vocab_rescore_slots(const float* x, const float* w, const float* bias, int rows, int d, int n, float w_norm_max,
                                const float* pmax, int slots, int32_t* ids, cudaStream_t st,
                                const int32_t* only_if_over = nullptr, int over = 0);
// max over rows of |w_row|_2, left in *out (device)
void launch_row_norm_max(const float* w, int rows, int d, float* out, cudaStream_t st);

// ---------------------------------------------------------------- attention (§8a a8)
// q,k,v: fp32 [B*T][ld], head h at column offset h*dk of each pointer.  kv_len[b] keys are attended
// (nullable => all `frames`).  Output ctx [B*T][ldo] fp32 and/or planes.
void launch_attention_simt(const float* q, const float* k, const float* v, int ld, int batch, int frames, int heads,
                           int dk, const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st);

// tcgen05 attention on bf16 planes of a fused [rows][ld] q|k|v matrix (q pre-scaled by d_k^-0.5 * log2 e):
// q at column 0, k at d_model, v at 2*d_model; head h at +h*dk.  Two planes `plane_stride` elements apart.
// pk != nullptr: packed rows (kernels.h Packing); kv_len[b] is then both the key and the query count of segment b.
void launch_attention_tc(Planes qkv, int64_t plane_stride, int ld, int d_model, int batch, int frames, int heads, int dk,
                         const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st,
                         const Packing* pk = nullptr);

}  // namespace fa
