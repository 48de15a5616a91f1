// Context object behind the C ABI: weights, workspace and the two graph executions
// (encoder+adaptor, CTC head) of the reference's exported sessions.
#pragma once
#include "kernels.h"

#include <map>
#include <tuple>
#include <memory>
#include <string>
#include <vector>

namespace fa {

enum Precision { kFp32 = 0, kBf16x3 = 1, kBf16 = 2, kFp8 = 3 };

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n;
        if (n) FA_CUDA(cudaMalloc(&p, n));
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Linear {                 // one nn.Linear: y = x W^T + b
    const float* w = nullptr;   // [n][k] fp32
    const float* b = nullptr;   // [n]
    int n = 0, k = 0;
    __nv_bfloat16* planes = nullptr;   // [2][n][k] bf16 hi|lo (tensor-core modes)
    TcOperand op;
    uint8_t* w8 = nullptr;             // [n][k] e4m3(w / wscale[row]) (fp8 mode)
    float* wscale = nullptr;           // [n] per-output-channel scale
    TcOperand op8;
};

struct Act {                    // an activation matrix as the GEMMs consume it
    const float* f32 = nullptr; // fp32 mode
    Planes pl;                  // tensor-core modes
    uint8_t* f8 = nullptr;      // fp8 mode: e4m3 bytes (null: this operand stays bf16 even in fp8 mode)
    int ld = 0;
};

struct SanmLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *fsmn_w;
    Linear qkv, out, w1, w2;
    int d_in;
};
struct MhaBlock {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    Linear qkv, out, w1, w2;    // qkv = linear_q|linear_k|linear_v stacked at load time
};
struct Projector {
    Linear lin1, lin2;
    std::vector<MhaBlock> blocks;
    int d = 0, heads = 0;
};

void prof_begin();
std::string prof_end();

class Context {
public:
    Context(int device, int max_batch, int64_t max_samples, int precision);
    ~Context();

    void load_tensor(const std::string& name, const float* host, const std::vector<int64_t>& shape);
    void finalize();

    // device-pointer graph executions; asynchronous on stream()
    void encode_dev(const float* d_audio, int batch, int64_t s_phys, const int64_t* h_ilens, float* d_enc,
                    float* d_adaptor);
    void ctc_dev(const float* d_enc, int batch, int frames, int32_t* d_ids);
    // both graphs back to back on device pointers (d_enc / d_adaptor may be null: the context's own buffers are used)
    void front_half_dev(const float* d_audio, int batch, int64_t s_phys, const int64_t* h_ilens, float* d_enc, float* d_adaptor,
                        int32_t* d_ids, const int64_t* h_phys = nullptr);
    void collapse_dev(const int32_t* d_ids, int batch, int frames, int32_t* d_tokens, int32_t* d_starts,
                      int32_t* d_counts);
    // host-pointer executions (copies inside); synchronous
    void encode_host(const float* audio, int batch, int64_t s_phys, const int64_t* ilens, float* enc, float* adaptor);
    void ctc_host(const float* enc, int batch, int frames, int32_t* ids);
    // embd_rows != nullptr: segment b's adaptor rows [0, target_len) go to embd_rows[b] (host or device memory) and
    // rows_out[b] = target_len; `adaptor` is then ignored
    // phys != nullptr (ragged batch): segment b is computed as the reference computes it at ITS OWN physical length phys[b]
    // (ilens[b] <= phys[b] <= s_phys, s_phys being the row stride of `audio`); needs a tensor-core precision mode
    void front_half_host(const float* audio, int batch, int64_t s_phys, const int64_t* ilens, float* enc,
                         float* adaptor, int32_t* ids, float* const* embd_rows = nullptr, int64_t* rows_out = nullptr,
                         const int64_t* phys = nullptr);

    void sync() { set_device(); FA_CUDA(cudaStreamSynchronize(stream_)); }
    void set_stream(cudaStream_t s);
    cudaStream_t stream() const { return stream_; }
    int max_batch() const { return max_batch_; }
    int64_t max_samples() const { return max_samples_; }
    int precision() const { return prec_; }
    int device() const { return device_; }
    int vocab() const { return vocab_; }
    void enable_taps(bool on) { taps_on_ = on; }
    bool read_tap(const std::string& name, std::vector<float>& out, std::vector<int64_t>& shape);
    std::vector<std::string> tap_names() const;

    // kernel-level test hooks
    void test_linear(const float* a, const float* w, const float* bias, const float* resid, int m, int n, int k,
                     int relu, int precision, float* out, float* out_planes_sum);
    void test_vocab_argmax(const float* a, const float* w, const float* bias, int m, int n, int k, int precision,
                           int32_t* ids);
    void test_attention(const float* qkv, int batch, int frames, int heads, int dk, const int32_t* kv_len,
                        int precision, float* out);

private:
    void set_device() const { FA_CUDA(cudaSetDevice(device_)); }
    const float* T(const std::string& name, std::vector<int64_t> shape) const;
    Linear make_linear(const std::string& prefix, int n, int k);
    Linear make_linear_from(const float* w, const float* b, int n, int k);
    void build_planes(Linear& l);
    void linear(const Act& a, const Linear& w, int m, const Epilogue& ep);
    void attention(const float* qkv, int ld, int d_model, int batch, int frames, int heads, const int* kv_len,
                   float* ctx_f32, Planes ctx_pl, int ldo, const Packing* pk = nullptr);
    void sanm_layer(const SanmLayer& L, bool first, int batch, int frames);
    void projector(const Projector& P, const Act& in, int batch, int frames, const int* kv_len, const Packing* pk);
    // the CTC head on `in` (uniform rows, or the head's packed rows when pk is given) -> ids of every physical frame
    void ctc_graph(const Act& in, int batch, int frames, int32_t* d_ids, const Packing* pk);
    // the CTC head right after encoder_graph in the same call: packed batches keep their packing (SURVEY §7 hard part 2)
    void ctc_after_encoder(const float* d_enc, int batch, int frames, int32_t* d_ids);
    // LayerNorm into the form the next projection reads at this precision (fp32 / bf16 planes / e4m3)
    void layernorm_to(const float* x, int rows, int d, const float* g, const float* b, float eps, const Act& dst);
    void tap(const char* name, const float* d, int64_t rows, int64_t cols);
    void ensure_room(int batch, int64_t s_phys) const;
    // encode_dev in three parts, so that the host variants can overlap copies with the front end
    void stage_lengths(int batch, int64_t s_phys, const int64_t* h_ilens, bool allow_packed, const int64_t* h_phys = nullptr);
    void front_end(const float* d_audio, int b0, int nb, int64_t s_phys);
    void encoder_graph(int batch, int64_t s_phys, float* d_enc, float* d_adaptor, bool record_events = true);
    // Small batches are launch-bound (some 620 launches per 60 s segment against a few milliseconds of GPU work):
    // the host variants replay them as CUDA graphs, captured once per (call kind, batch, samples).
    bool use_graph(int batch) const;
    template <class F> void run_graphed(int kind, int batch, int64_t size, F&& body);
    void upload_and_front_end(const float* audio_host, int nb, int64_t s_phys);
    void download_async(void* host, const void* dev, size_t bytes, cudaEvent_t after);
    Planes qkv_planes() const;
    Epilogue qkv_epilogue(int d_model, int heads, bool need_v_f32) const;
    Act h_act(int ld) const;
    Act ctx_act(int ld) const;
    Act ffn_act(int ld) const;
    // a14 on the tensor cores: argmax over the vocabulary of x[M][k] * W^T + b without the logits reaching HBM
    void vocab_argmax(const float* x, Planes x_pl, const Linear& lo, float w_norm_max, int m, VocabCand cand,
                      float* amax_val, int32_t* amax_idx, int32_t* d_ids);
    VocabCand cand_workspace() const;

    int device_, max_batch_, prec_, vocab_ = 0;
    int64_t max_samples_;
    int t_mel_max_, t_max_;
    int64_t m_max_;
    cudaStream_t stream_ = nullptr;
    cudaStream_t copy_stream_ = nullptr;      // host<->device copies of the host variants, overlapped with compute
    cudaEvent_t ev_up_[8] = {}, ev_enc_ = nullptr, ev_ad_ = nullptr;
    bool capturable_ = true;                  // stream_ can be captured into a CUDA graph (not the legacy default stream)
    bool own_stream_ = true, finalized_ = false, taps_on_ = false, simt_attention_ = true;

    struct HostTensor { std::vector<int64_t> shape; std::unique_ptr<DevBuf> buf; };
    std::map<std::string, HostTensor> tensors_;
    std::vector<std::unique_ptr<DevBuf>> derived_;      // fused / transposed / plane copies of weights

    std::vector<SanmLayer> enc_layers_;
    const float *after_g_, *after_b_, *tp_g_, *tp_b_;
    Projector adaptor_, ctc_;
    Linear ctc_lo_;
    const float *dft_t_ = nullptr, *melfb_t_ = nullptr, *pos_enc_ = nullptr;
    const int* mel_range_ = nullptr;          // [80][2] support of each mel filter
    int pos_rows_ = 0;

    // workspace
    DevBuf audio_, partials_, logmel_, x0_, x_, h32_, hpl_, qkv_, qkvpl_, ctx32_, ctxpl_, ffn32_, ffnpl_, encpl_, enc_,
        adaptor_out_, ids_, amax_val_, amax_idx_, logits_, lens_, tokens_, cand_meta_, cand_list_, yplanes_, power_, h8_, ffn8_, enc8_;
    float vocab_wnorm_ = 0.f;                 // max_c |w_c|_2 of ctc_lo (bound of the one-product vocabulary pass)
    bool vocab_rescore_ = false;              // bf16x3: one-product pass + exact rescoring of the candidates
    int* d_nvalid_ = nullptr;
    int* d_tvalid_ = nullptr;
    int* d_tlen_ = nullptr;
    int* h_lens_ = nullptr;      // pinned staging for the length vectors and the packing tables
    // Padding-free execution of the encoder and the adaptor (kernels.h Packing): chosen per call by stage_lengths when
    // the batch holds padded frames and is not replayed as a CUDA graph (a graph's launch geometry is fixed)
    bool packed_ = false, allow_packed_env_ = true;
    Packing pk_;                 // encoder + adaptor rows
    Packing pk_ctc_;             // CTC head rows: the valid frames of a segment plus ONE row for all its padded frames
    const int* d_len_ctc_ = nullptr;          // [batch] rows of each segment in the head's packing
    const int* d_tphys_ = nullptr;            // [batch] LFR frames of each segment's own physical length (ragged batches), else null
    bool fused_ctc_next_ = false;             // the CTC head follows in this call: encoder_graph prepares its packed input
    Act ctc_packed_input() const;
    bool ctc_packed_ready_ = false;           // encoder_graph left the head's packed input in encpl_ / enc8_
    int max_tiles_ = 0, tab_off_ = 0;         // attention tile tables inside the staged block (kernels.h Packing::tile_tab)
    int len_ints_ = 0;           // ints per staging slot: 3 length vectors + seg_off, order, tile_off (+ the head's tables)
    const Packing* packing() const { return packed_ ? &pk_ : nullptr; }
    int enc_rows(int batch, int frames) const { return packed_ ? pk_.total_rows : batch * frames; }
    static constexpr int kLenSlots = 4;       // ring of staging slots: a slot is reused once its upload has completed
    cudaEvent_t len_ev_[kLenSlots] = {};
    int len_next_ = 0;
    int logits_rows_ = 0;
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
    std::map<std::tuple<int, int, int64_t>, GraphEntry> graphs_;
    bool fbank_tc_ = true;                    // DFT of the front end on the tensor cores (FUNASR_B200_FBANK=simt: CUDA cores)
    void* dft_planes_ = nullptr;              // fp16 hi/lo planes of the DFT table (tensor-core front end)
    int graph_max_batch_ = 4;                 // FUNASR_B200_GRAPH_MAX_BATCH (0 turns graph replay off)
    std::map<std::string, std::pair<std::vector<int64_t>, std::unique_ptr<DevBuf>>> taps_;
};

}  // namespace fa
