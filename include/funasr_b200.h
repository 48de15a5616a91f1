/* funasr_b200.h — C ABI of the B200-native audio front half of Fun-ASR-Nano.
 *
 * The reference has no FFI for this path: it runs two ONNX Runtime InferenceSession objects
 * from Python (fun_asr_gguf/nano_onnx.py:35-45).  Each entry point below names the reference
 * call it replaces; the Python mirror of the InferenceSession surface that binds them through
 * ctypes is fun_asr_gguf_b200/ort_shim.py, and INTEGRATION.md shows the two-line change a
 * maintainer makes in the reference.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * non-zero code on failure, with the message available from fa_last_error() (thread-local).
 * "host" pointers are CPU memory (pinned memory makes the copies asynchronous); "dev"
 * pointers are CUDA device memory on the context's device.  All tensors are dense row-major.
 * There is no CPU fallback: creating a context without an sm_100 device fails.
 */
#ifndef FUNASR_B200_H
#define FUNASR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_ABI_VERSION 2 /* 2: + fa_front_half_dev, fa_front_half_ragged(_dev), FA_PREC_FP8 (additions only) */

#if defined(__GNUC__)
#define FA_API __attribute__((visibility("default")))
#else
#define FA_API
#endif

typedef struct fa_ctx fa_ctx;

/* arithmetic of the dense projections (attention scores and all row kernels are fp32 in every mode) */
enum fa_precision {
    FA_PREC_FP32 = 0,   /* fp32 FMA on the CUDA cores: exact-precision mode and on-device arbiter      */
    FA_PREC_BF16X3 = 1, /* tcgen05, operands as bf16 hi+lo planes, 3 MMAs per product, fp32 accumulate */
    FA_PREC_BF16 = 2,   /* tcgen05, plain bf16 operands, fp32 accumulate (fast mode, not token-exact)  */
    FA_PREC_FP8 = 3     /* tcgen05 kind::f8f6f4: e4m3 activations, e4m3 weights with a per-output-channel scale (the
                           reference's int8 graph variant, 02-Quantize-ONNX.py:41-44, as W8A8 floating point); the
                           attention products, its output projection, LayerNorm and softmax stay as in FA_PREC_BF16 /
                           fp32 (02-Quantize-ONNX.py:26).  Speed mode with a stated id-mismatch budget.             */
};

FA_API int fa_abi_version(void);
FA_API const char* fa_last_error(void);
FA_API int fa_device_count(int* count);

/* shape helpers — the arithmetic of nano_onnx.py:122-125 and model_definition.py:288-291,317-318 */
FA_API int64_t fa_frames_for_samples(int64_t samples);       /* LFR frames of a segment: ceil((S/160+1)/6) */
FA_API int64_t fa_adaptor_rows_for_samples(int64_t n_valid); /* target_len: rows of adaptor_output kept    */

/* ---- context: replaces onnxruntime.InferenceSession(path, ...) x2 (nano_onnx.py:35-45) ------------
 * One context serves both sessions; weights are loaded once per device. */
FA_API int fa_ctx_create(int device, int max_batch, int64_t max_samples, int precision, fa_ctx** out);
FA_API int fa_ctx_destroy(fa_ctx* ctx);
/* Tensor hand-off by reference state_dict key (HybridSenseVoice.load_weights, model_definition.py:231-238)
 * plus the non-parameter constants const.dft_cos / const.dft_sin (201x400, model_definition.py:252-253),
 * const.mel_fbank (80x201, 01-Export-Encoder-Adaptor-CTC.py:102) and const.pos_enc (T_max x 560,
 * model_definition.py:13-21).  fp32, host memory, copied. */
FA_API int fa_ctx_load_tensor(fa_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
FA_API int fa_ctx_finalize(fa_ctx* ctx);                     /* checks the inventory, builds derived forms */
FA_API int fa_ctx_set_stream(fa_ctx* ctx, void* cuda_stream);/* run on a caller-owned cudaStream_t          */
FA_API int fa_ctx_sync(fa_ctx* ctx);
FA_API int fa_ctx_vocab(fa_ctx* ctx, int* vocab);
FA_API int64_t fa_launch_count(void);                        /* kernels launched by the calling thread so far */
/* per-launch CUDA-event timing of the calling thread's launches, aggregated by kernel, as JSON:
 * {"k_name": {"launches": n, "ms": total, "flops": algorithmic, "bytes": algorithmic}, ...} */
FA_API int fa_prof_begin(void);
FA_API int fa_prof_end(char* json_out, int64_t capacity);

/* ---- encoder session: replaces encoder_sess.run / run_with_ort_values (nano_onnx.py:62,117) -------
 * audio  [batch][samples] fp32, every row zero-padded to the same physical length
 * ilens  [batch] true sample counts (host memory in both variants)
 * enc    [batch][T][512],  rows >= valid frames are zero          ("enc_output")
 * adaptor[batch][T][1024], rows >= target_len are zero             ("adaptor_output")
 * Each row is an independent segment (the reference graph is batch-1). */
FA_API int fa_encode(fa_ctx* ctx, const float* audio_host, int batch, int64_t samples, const int64_t* ilens,
              float* enc_host, float* adaptor_host);
FA_API int fa_encode_dev(fa_ctx* ctx, const float* audio_dev, int batch, int64_t samples, const int64_t* ilens_host,
                  float* enc_dev, float* adaptor_dev);

/* ---- CTC session: replaces ctc_sess.run(None, {"enc_output": enc})[0] (core/decoder.py:27) --------
 * enc [batch][frames][512] -> ids [batch][frames] int32 ("indices"); unmasked over all physical frames. */
FA_API int fa_ctc(fa_ctx* ctx, const float* enc_host, int batch, int frames, int32_t* ids_host);
FA_API int fa_ctc_dev(fa_ctx* ctx, const float* enc_dev, int batch, int frames, int32_t* ids_dev);

/* both graphs back to back, enc staying on the device between them (enc/adaptor may be NULL) */
FA_API int fa_front_half(fa_ctx* ctx, const float* audio_host, int batch, int64_t samples, const int64_t* ilens,
                  float* enc_host, float* adaptor_host, int32_t* ids_host);

/* the same on device pointers, asynchronous on the context's stream (enc_dev / adaptor_dev may be NULL: the results then
 * stay in the context).  Calling the two graphs in ONE entry point lets a mixed-length batch keep its padding-free row
 * layout through the CTC head as well (one row per segment stands for all its zero-padded frames). */
FA_API int fa_front_half_dev(fa_ctx* ctx, const float* audio_dev, int batch, int64_t samples, const int64_t* ilens_host,
                      float* enc_dev, float* adaptor_dev, int32_t* ids_dev);

/* ---- ragged batches: every segment at ITS OWN physical length ------------------------------------------------
 * The unmasked CTC head makes a segment's ids depend on its physical (padded) length, so padding a short
 * window up to a longer neighbour would change them: windows of different lengths could not share a batch.  Here row b of
 * `audio` (row stride `samples`) is computed exactly as the reference computes a segment of ilens[b] valid samples fed at
 * physical length phys[b] (ilens[b] <= phys[b] <= samples; nano_onnx.py:90-99 pads to max(len, 1 s) on the CPU provider):
 * the padding-free row layout never materialises padded frames, and the head counts each segment's own
 * frames(phys[b]) - frames(ilens[b]) zero frames as one key of that multiplicity.  Outputs keep the uniform
 * [batch][frames(samples)] shapes; rows past a segment's valid frames are zero, ids past its own physical frames are -1.
 * Needs a tensor-core precision mode. */
FA_API int fa_front_half_ragged(fa_ctx* ctx, const float* audio_host, int batch, int64_t samples, const int64_t* ilens,
                         const int64_t* phys, float* enc_host, float* adaptor_host, int32_t* ids_host);
FA_API int fa_front_half_ragged_dev(fa_ctx* ctx, const float* audio_dev, int batch, int64_t samples, const int64_t* ilens_host,
                             const int64_t* phys_host, float* enc_dev, float* adaptor_dev, int32_t* ids_dev);

/* ---- embedding handoff ----------------------------------------------------------------
 * fa_front_half, except that of each segment's adaptor_output only the rows the LLM reads — [0, target_len), what
 * nano_onnx.py:131-133 slices out — leave the device, written straight to embd_rows[b] (host or device memory,
 * target_len x 1024 fp32, dense): e.g. into llama_batch.embd behind the prefix prompt's rows, instead of a
 * [T][1024] array that the caller slices, concatenates with the prompt and memmoves (core/decoder.py:199,
 * llama.py:536-547).  rows_out[b] = target_len (may be NULL); enc_host may be NULL. */
FA_API int fa_front_half_embd(fa_ctx* ctx, const float* audio_host, int batch, int64_t samples, const int64_t* ilens,
                       float* enc_host, float* const* embd_rows, int64_t* rows_out, int32_t* ids_host);

/* ---- greedy collapse on device: the integer part of decode_ctc (nano_ctc.py:70-99) ----------------
 * ids [batch][frames] -> tokens/starts [batch][frames] (first counts[b] entries valid), blank = vocab-1 */
FA_API int fa_ctc_collapse_dev(fa_ctx* ctx, const int32_t* ids_dev, int batch, int frames, int32_t* tokens_dev,
                        int32_t* starts_dev, int32_t* counts_dev);

/* ---- debug taps and kernel-level test hooks (used by tests/, not by the product path) -------------- */
FA_API int fa_debug_enable_taps(fa_ctx* ctx, int on);
FA_API int fa_debug_read_tap(fa_ctx* ctx, const char* name, float* out_host, int64_t capacity, int64_t* rows, int64_t* cols);
FA_API int fa_test_linear(fa_ctx* ctx, const float* a, const float* w, const float* bias, const float* resid, int m, int n,
                   int k, int relu, int precision, float* out, float* out_planes_sum);
FA_API int fa_test_vocab_argmax(fa_ctx* ctx, const float* a, const float* w, const float* bias, int m, int n, int k,
                         int precision, int32_t* ids);
FA_API int fa_test_attention(fa_ctx* ctx, const float* qkv, int batch, int frames, int heads, int dk, const int32_t* kv_len,
                      int precision, float* out);
FA_API int fa_test_layernorm(fa_ctx* ctx, const float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                      float* out, float* out_planes_sum);
FA_API int fa_test_fsmn(fa_ctx* ctx, const float* v, const float* w, const int32_t* t_valid, int batch, int frames,
                 const float* resid, float* out);
FA_API int fa_test_front_end(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens,
                      float* logmel /*[B][T_mel][80]*/, float* x0 /*[B][T][560]*/);

#ifdef __cplusplus
}
#endif
#endif /* FUNASR_B200_H */
