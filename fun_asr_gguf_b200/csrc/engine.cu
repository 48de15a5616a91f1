// Graph execution: the reference's two exported sessions as sequences of the kernels in this
// directory.  Layer order and every quirk follow fun_asr_gguf/model_definition.py:
//   encoder  :205-214 (70 SAN-M layers, after_norm / tp_norm with mask sweeps), layer body :100-116
//            (layer 0 returns after attention+FSMN: no residual, no FFN — SURVEY F9)
//   adaptor  :179-185 + :154-163, length control :317-321 (rows >= target_len zeroed — F10)
//   CTC head :335-337 (5 blocks with mask=None over every physical frame — F7; argmax -> int32)
// Batches are independent rows (the reference graph is batch-1 only — F8).
#include "engine.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

namespace fa {

thread_local int64_t g_launches = 0;
thread_local bool g_prof_on = false;
bool g_pdl = [] { const char* s = getenv("FUNASR_B200_PDL"); return !(s && s[0] == '0'); }();

namespace {
struct ProfRec { std::string name; cudaEvent_t a, b; double flops, bytes; };
thread_local std::vector<ProfRec> g_prof;
thread_local double g_next_flops = 0.0, g_next_bytes = 0.0;
thread_local std::string g_next_tag;
}  // namespace

void prof_note_work(double flops, double bytes) { g_next_flops = flops; g_next_bytes = bytes; }
// Shape class of the next launch ("n1536_k512", "gated", ...): appended to the kernel's name in the profile so that
// bench.py can report each class of one kernel against its own roofline.
void prof_note_tag(const char* tag) { if (g_prof_on) g_next_tag = tag; }

void prof_before(const char* kernel, cudaStream_t st) {
    ProfRec r;
    r.name = kernel;
    if (!r.name.empty() && r.name[0] == '(') r.name.erase(0, 1);      // FA_LAUNCH((k<a, b>), ...) stringifies with its parentheses
    if (!r.name.empty() && r.name.back() == ')') r.name.pop_back();
    const size_t lt = r.name.find('<');
    if (lt != std::string::npos && r.name.find("k_gemm_tc") == std::string::npos) r.name = r.name.substr(0, lt);
    if (!g_next_tag.empty()) { r.name += "/" + g_next_tag; g_next_tag.clear(); }
    r.flops = g_next_flops; r.bytes = g_next_bytes;
    g_next_flops = g_next_bytes = 0.0;
    FA_CUDA(cudaEventCreate(&r.a));
    FA_CUDA(cudaEventCreate(&r.b));
    FA_CUDA(cudaEventRecord(r.a, st));
    g_prof.push_back(r);
}

void prof_after(cudaStream_t st) { FA_CUDA(cudaEventRecord(g_prof.back().b, st)); }

void prof_begin() {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
}

// Stops profiling and returns {"kernel": {"launches": n, "ms": total, "flops": sum, "bytes": sum}, ...} as JSON.
std::string prof_end() {
    g_prof_on = false;
    std::map<std::string, std::array<double, 4>> agg;
    for (auto& r : g_prof) {
        FA_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        FA_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        auto& a = agg[r.name];
        a[0] += 1; a[1] += ms; a[2] += r.flops; a[3] += r.bytes;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %.0f, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
                 first ? "" : ", ", kv.first.c_str(), kv.second[0], kv.second[1], kv.second[2], kv.second[3]);
        out += buf;
        first = false;
    }
    return out + "}";
}

namespace {
int lfr_frames_of(int64_t samples) { return (int)((samples / kHop + 1 + kLfrN - 1) / kLfrN); }
int target_len_of(int64_t n_valid) {
    const int t = lfr_frames_of(n_valid);
    const int o1 = 1 + (t - 3 + 2) / 2;                  // floor division on non-negative operands only when t >= 1
    return (1 + (o1 - 3 + 2) / 2 - 1) / 2 + 1;
}
}  // namespace

// ------------------------------------------------------------------------------------ construction

Context::Context(int device, int max_batch, int64_t max_samples, int precision)
    : device_(device), max_batch_(max_batch), prec_(precision), max_samples_(max_samples) {
    FA_REQUIRE(max_batch >= 1 && max_samples >= 1, "max_batch and max_samples must be positive");
    FA_REQUIRE(precision >= kFp32 && precision <= kFp8, "unknown precision mode");
    int count = 0;
    FA_CUDA(cudaGetDeviceCount(&count));
    FA_REQUIRE(device >= 0 && device < count, "no such CUDA device");
    set_device();
    cudaDeviceProp prop;
    FA_CUDA(cudaGetDeviceProperties(&prop, device));
    FA_REQUIRE(prop.major == 10, "this library is built for sm_100a (B200) only; found compute capability " +
                                     std::to_string(prop.major) + "." + std::to_string(prop.minor));
    FA_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    FA_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (auto& ev : ev_up_) FA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FA_CUDA(cudaEventCreateWithFlags(&ev_enc_, cudaEventDisableTiming));
    FA_CUDA(cudaEventCreateWithFlags(&ev_ad_, cudaEventDisableTiming));
    frontend_init_device();
    attention_init_device();
    tc_init_device();
    attention_tc_init_device();
    const char* att = getenv("FUNASR_B200_ATTENTION");     // debugging aid: "simt" keeps fp32 attention in the tensor-core modes
    simt_attention_ = precision == kFp32 || (att && std::string(att) == "simt");
    if (const char* g = getenv("FUNASR_B200_GRAPH_MAX_BATCH")) graph_max_batch_ = std::max(0, atoi(g));
    {
        const char* g = getenv("FUNASR_B200_FBANK");          // comparison aid
        fbank_tc_ = precision != kFp32 && !(g && std::string(g) == "simt");
        if (fbank_tc_) fbank_tc_init_device();
    }
    t_mel_max_ = (int)(max_samples / kHop + 1);
    t_max_ = lfr_frames_of(max_samples);
    m_max_ = (int64_t)max_batch * t_max_;
    for (auto& e : len_ev_) FA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

Context::~Context() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (auto& e : len_ev_) if (e) cudaEventDestroy(e);
    for (auto& kv : graphs_) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (h_lens_) cudaFreeHost(h_lens_);
    for (auto& ev : ev_up_) if (ev) cudaEventDestroy(ev);
    if (ev_enc_) cudaEventDestroy(ev_enc_);
    if (ev_ad_) cudaEventDestroy(ev_ad_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (own_stream_ && stream_) cudaStreamDestroy(stream_);
}

void Context::set_stream(cudaStream_t s) {
    set_device();
    FA_CUDA(cudaStreamSynchronize(stream_));
    if (own_stream_ && stream_) cudaStreamDestroy(stream_);
    stream_ = s;
    own_stream_ = false;
    // the legacy default stream (and the per-thread one) cannot be captured: calls on it launch eagerly
    capturable_ = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
}

void Context::load_tensor(const std::string& name, const float* host, const std::vector<int64_t>& shape) {
    FA_REQUIRE(!finalized_, "context already finalized");
    set_device();
    size_t n = 1;
    for (auto d : shape) { FA_REQUIRE(d > 0, "tensor dims must be positive"); n *= (size_t)d; }
    HostTensor t;
    t.shape = shape;
    t.buf.reset(new DevBuf);
    t.buf->alloc(n * sizeof(float));
    FA_CUDA(cudaMemcpy(t.buf->p, host, n * sizeof(float), cudaMemcpyHostToDevice));
    tensors_[name] = std::move(t);
}

const float* Context::T(const std::string& name, std::vector<int64_t> shape) const {
    auto it = tensors_.find(name);
    if (it == tensors_.end()) throw Error("missing tensor: " + name);
    if (it->second.shape != shape) {
        std::string got, want;
        for (auto d : it->second.shape) got += std::to_string(d) + ",";
        for (auto d : shape) want += std::to_string(d) + ",";
        throw Error("tensor " + name + " has shape (" + got + ") but the path needs (" + want + ")");
    }
    return it->second.buf->as<float>();
}

void Context::build_planes(Linear& l) {
    if (prec_ == kFp32) return;
    derived_.emplace_back(new DevBuf);
    DevBuf& b = *derived_.back();
    const int64_t n = (int64_t)l.n * l.k;
    b.alloc(2 * n * sizeof(__nv_bfloat16));
    l.planes = b.as<__nv_bfloat16>();
    launch_split_planes(l.w, n, Planes{l.planes, l.planes + n}, stream_);
    l.op = tc_make_weight(l.planes, l.n, l.k, n, 2);
    if (prec_ == kFp8 && l.k % 16 == 0) {
        derived_.emplace_back(new DevBuf);
        DevBuf& q = *derived_.back();
        const size_t scale_off = ((size_t)n + 255) / 256 * 256;
        q.alloc(scale_off + (size_t)l.n * sizeof(float));
        l.w8 = q.as<uint8_t>();
        l.wscale = reinterpret_cast<float*>(q.as<uint8_t>() + scale_off);
        launch_quant_rows_e4m3(l.w, l.n, l.k, l.w8, l.wscale, stream_);
        l.op8 = tc_make_operand_f8(l.w8, l.n, l.k, l.k, 64);
    }
}

Linear Context::make_linear_from(const float* w, const float* b, int n, int k) {
    Linear l;
    l.w = w; l.b = b; l.n = n; l.k = k;
    build_planes(l);
    return l;
}

Linear Context::make_linear(const std::string& prefix, int n, int k) {
    return make_linear_from(T(prefix + ".weight", {n, k}), T(prefix + ".bias", {n}), n, k);
}

void Context::finalize() {
    FA_REQUIRE(!finalized_, "context already finalized");
    set_device();
    // ---- encoder
    auto sanm = [&](const std::string& p, int d_in) {
        SanmLayer L;
        L.d_in = d_in;
        L.ln1_g = T(p + ".norm1.weight", {d_in});
        L.ln1_b = T(p + ".norm1.bias", {d_in});
        L.ln2_g = T(p + ".norm2.weight", {kDenc});
        L.ln2_b = T(p + ".norm2.bias", {kDenc});
        L.fsmn_w = T(p + ".self_attn.fsmn_block.weight", {kDenc, 1, kFsmnK});
        L.qkv = make_linear(p + ".self_attn.linear_q_k_v", 3 * kDenc, d_in);
        L.out = make_linear(p + ".self_attn.linear_out", kDenc, kDenc);
        if (d_in == kDenc) {      // layer 0's FFN / norm2 exist in the checkpoint but are never executed (F9)
            L.w1 = make_linear(p + ".feed_forward.w_1", kDffn, kDenc);
            L.w2 = make_linear(p + ".feed_forward.w_2", kDenc, kDffn);
        }
        return L;
    };
    enc_layers_.clear();
    enc_layers_.push_back(sanm("audio_encoder.encoders0.0", kDin));
    for (int i = 0; i < 49; ++i) enc_layers_.push_back(sanm("audio_encoder.encoders." + std::to_string(i), kDenc));
    for (int i = 0; i < 20; ++i) enc_layers_.push_back(sanm("audio_encoder.tp_encoders." + std::to_string(i), kDenc));
    after_g_ = T("audio_encoder.after_norm.weight", {kDenc});
    after_b_ = T("audio_encoder.after_norm.bias", {kDenc});
    tp_g_ = T("audio_encoder.tp_norm.weight", {kDenc});
    tp_b_ = T("audio_encoder.tp_norm.bias", {kDenc});

    // ---- adaptor / CTC decoder (CorrectTransformerAdaptor): q|k|v stacked into one [3d][d] projection
    auto proj = [&](const std::string& p, int d, int n_blocks, int heads) {
        Projector P;
        P.d = d; P.heads = heads;
        P.lin1 = make_linear(p + ".linear1", kDffn, kDenc);
        P.lin2 = make_linear(p + ".linear2", d, kDffn);
        for (int i = 0; i < n_blocks; ++i) {
            const std::string b = p + ".blocks." + std::to_string(i);
            MhaBlock B;
            B.ln1_g = T(b + ".norm1.weight", {d}); B.ln1_b = T(b + ".norm1.bias", {d});
            B.ln2_g = T(b + ".norm2.weight", {d}); B.ln2_b = T(b + ".norm2.bias", {d});
            derived_.emplace_back(new DevBuf);
            DevBuf& wq = *derived_.back();
            wq.alloc(((size_t)3 * d * d + 3 * d) * sizeof(float));
            float* w = wq.as<float>();
            float* bias = w + (size_t)3 * d * d;
            const char* names[3] = {"linear_q", "linear_k", "linear_v"};
            for (int j = 0; j < 3; ++j) {
                const std::string n = b + ".self_attn." + names[j];
                FA_CUDA(cudaMemcpyAsync(w + (size_t)j * d * d, T(n + ".weight", {d, d}), (size_t)d * d * sizeof(float),
                                        cudaMemcpyDeviceToDevice, stream_));
                FA_CUDA(cudaMemcpyAsync(bias + j * d, T(n + ".bias", {d}), d * sizeof(float), cudaMemcpyDeviceToDevice,
                                        stream_));
            }
            B.qkv = make_linear_from(w, bias, 3 * d, d);
            B.out = make_linear(b + ".self_attn.linear_out", d, d);
            B.w1 = make_linear(b + ".feed_forward.w_1", d / 4, d);
            B.w2 = make_linear(b + ".feed_forward.w_2", d, d / 4);
            P.blocks.push_back(B);
        }
        return P;
    };
    adaptor_ = proj("audio_adaptor", kDllm, 2, 8);
    ctc_ = proj("ctc_decoder", kDenc, 5, 8);
    {
        auto it = tensors_.find("ctc_proj.ctc_lo.weight");
        if (it == tensors_.end() || it->second.shape.size() != 2 || it->second.shape[1] != kDenc)
            throw Error("missing or malformed tensor: ctc_proj.ctc_lo.weight");
        vocab_ = (int)it->second.shape[0];
        ctc_lo_ = make_linear("ctc_proj.ctc_lo", vocab_, kDenc);
        // FUNASR_B200_VOCAB=full keeps the three-product projection with the fused running argmax (comparison aid)
        const char* vm = getenv("FUNASR_B200_VOCAB");
        vocab_rescore_ = prec_ != kFp32 && !(vm && std::string(vm) == "full");
        if (vocab_rescore_) {
            DevBuf nm;
            nm.alloc(sizeof(float));
            launch_row_norm_max(ctc_lo_.w, vocab_, kDenc, nm.as<float>(), stream_);
            FA_CUDA(cudaMemcpyAsync(&vocab_wnorm_, nm.p, sizeof(float), cudaMemcpyDeviceToHost, stream_));
            FA_CUDA(cudaStreamSynchronize(stream_));
        }
    }

    // ---- front-end constants: DFT kernels transposed to [n][cos|sin] and the mel matrix to [bin][mel]
    {
        const float* c = T("const.dft_cos", {kBins, kNfft});
        const float* s = T("const.dft_sin", {kBins, kNfft});
        const float* mf = T("const.mel_fbank", {kMels, kBins});
        std::vector<float> hc((size_t)kBins * kNfft), hs(hc.size()), hm((size_t)kMels * kBins);
        FA_CUDA(cudaMemcpy(hc.data(), c, hc.size() * 4, cudaMemcpyDeviceToHost));
        FA_CUDA(cudaMemcpy(hs.data(), s, hs.size() * 4, cudaMemcpyDeviceToHost));
        FA_CUDA(cudaMemcpy(hm.data(), mf, hm.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<float> dt((size_t)kNfft * kDftLd, 0.f), mt((size_t)kBins * kMels);
        for (int k = 0; k < kBins; ++k)
            for (int n = 0; n < kNfft; ++n) {
                dt[(size_t)n * kDftLd + k] = hc[(size_t)k * kNfft + n];
                dt[(size_t)n * kDftLd + kBins + k] = hs[(size_t)k * kNfft + n];
            }
        for (int j = 0; j < kMels; ++j)
            for (int k = 0; k < kBins; ++k) mt[(size_t)k * kMels + j] = hm[(size_t)j * kBins + k];
        derived_.emplace_back(new DevBuf);
        derived_.back()->alloc(dt.size() * 4);
        FA_CUDA(cudaMemcpy(derived_.back()->p, dt.data(), dt.size() * 4, cudaMemcpyHostToDevice));
        dft_t_ = derived_.back()->as<float>();
        derived_.emplace_back(new DevBuf);
        derived_.back()->alloc(mt.size() * 4);
        FA_CUDA(cudaMemcpy(derived_.back()->p, mt.data(), mt.size() * 4, cudaMemcpyHostToDevice));
        melfb_t_ = derived_.back()->as<float>();
        // support of each mel filter: [first, last+1) non-zero bin.  A zero weight contributes exactly 0 to the
        // reference's dense 80 x 201 product, so the kernel only walks the support (HTK triangles are ~5 bins wide)
        std::vector<int> range(2 * kMels);
        for (int j = 0; j < kMels; ++j) {
            int lo = kBins, hi = 0;
            for (int k = 0; k < kBins; ++k)
                if (hm[(size_t)j * kBins + k] != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
            if (hi <= lo) { lo = 0; hi = 0; }
            range[2 * j] = lo; range[2 * j + 1] = hi;
        }
        derived_.emplace_back(new DevBuf);
        derived_.back()->alloc(range.size() * sizeof(int));
        FA_CUDA(cudaMemcpy(derived_.back()->p, range.data(), range.size() * sizeof(int), cudaMemcpyHostToDevice));
        mel_range_ = derived_.back()->as<int>();
        if (fbank_tc_) {
            derived_.emplace_back(new DevBuf);
            derived_.back()->alloc(fbank_tc_table_bytes());
            dft_planes_ = derived_.back()->p;
            launch_fbank_table_planes(dft_t_, dft_planes_, stream_);
        }
        auto it = tensors_.find("const.pos_enc");
        if (it == tensors_.end() || it->second.shape.size() != 2 || it->second.shape[1] != kDin)
            throw Error("missing or malformed tensor: const.pos_enc");
        pos_rows_ = (int)it->second.shape[0];
        FA_REQUIRE(pos_rows_ >= t_max_, "const.pos_enc has fewer rows than the longest segment needs");
        pos_enc_ = it->second.buf->as<float>();
    }

    // ---- workspace, sized for max_batch segments of max_samples
    const size_t M = (size_t)m_max_;
    const bool tc = prec_ != kFp32;
    audio_.alloc((size_t)max_batch_ * max_samples_ * 4);
    partials_.alloc((size_t)max_batch_ * kMeanParts * 8);
    logmel_.alloc((size_t)max_batch_ * t_mel_max_ * kMels * 4);
    if (fbank_tc_) {
        yplanes_.alloc((size_t)2 * max_batch_ * fbank_tc_padded_samples(max_samples_) * 2);
        power_.alloc((size_t)max_batch_ * t_mel_max_ * fbank_tc_power_ld() * 4);
    }
    x0_.alloc(M * kDin * 4);
    x_.alloc(M * kDllm * 4);
    qkv_.alloc(M * 3 * kDllm * 4);
    enc_.alloc(M * kDenc * 4);
    adaptor_out_.alloc(M * kDllm * 4);
    ids_.alloc(M * 4);
    tokens_.alloc(M * 4 * 2 + (size_t)max_batch_ * 4);
    if (tc) {
        hpl_.alloc(2 * M * kDllm * 2);
        qkvpl_.alloc(2 * M * 3 * kDllm * 2);
        ctxpl_.alloc(2 * M * kDllm * 2);
        ffnpl_.alloc(2 * M * kDffn * 2);
        encpl_.alloc(2 * M * kDenc * 2);
        if (prec_ == kFp8) {
            h8_.alloc(M * kDllm);
            ffn8_.alloc(M * kDffn);
            enc8_.alloc(M * kDenc);
        }
        const size_t tiles = (size_t)tc_argmax_tiles(vocab_);
        amax_val_.alloc(M * tiles * 4);
        amax_idx_.alloc(M * tiles * 4);
        if (vocab_rescore_) {
            cand_meta_.alloc((M * 3 + 1) * 4);
            cand_list_.alloc(M * kVocabCandCap * sizeof(int2));
        }
    } else {
        h32_.alloc(M * kDllm * 4);
        ctx32_.alloc(M * kDllm * 4);
        ffn32_.alloc(M * kDffn * 4);
        logits_rows_ = (int)std::min<size_t>(M, 2048);
        logits_.alloc((size_t)logits_rows_ * vocab_ * 4);
    }
    // the staged block (layout at stage_lengths): nine B-sized vectors + two (B+1)-sized prefix arrays' extra entries,
    // then, 16-byte aligned, the two attention tile tables (encoder/adaptor, CTC head)
    max_tiles_ = max_batch_ * (cdiv(t_max_, 128) + 1);
    tab_off_ = (8 * max_batch_ + 2 + 3) / 4 * 4;
    len_ints_ = tab_off_ + 2 * 4 * max_tiles_;
    lens_.alloc((size_t)len_ints_ * sizeof(int));
    d_nvalid_ = lens_.as<int>();
    d_tvalid_ = d_nvalid_ + max_batch_;
    d_tlen_ = d_tvalid_ + max_batch_;
    FA_CUDA(cudaMallocHost(&h_lens_, (size_t)kLenSlots * len_ints_ * sizeof(int)));
    if (const char* pe = getenv("FUNASR_B200_PACKED")) allow_packed_env_ = !(pe[0] == '0');     // comparison aid
    FA_CUDA(cudaStreamSynchronize(stream_));
    finalized_ = true;
}

// ------------------------------------------------------------------------------------ helpers

Act Context::h_act(int ld) const {
    Act a; a.ld = ld; a.f32 = h32_.as<float>();
    if (hpl_.p) a.pl = Planes{hpl_.as<__nv_bfloat16>(), hpl_.as<__nv_bfloat16>() + m_max_ * kDllm};
    a.f8 = h8_.as<uint8_t>();
    return a;
}
Act Context::ctx_act(int ld) const {
    Act a; a.ld = ld; a.f32 = ctx32_.as<float>();
    if (ctxpl_.p) a.pl = Planes{ctxpl_.as<__nv_bfloat16>(), ctxpl_.as<__nv_bfloat16>() + m_max_ * kDllm};
    return a;
}
Act Context::ffn_act(int ld) const {
    Act a; a.ld = ld; a.f32 = ffn32_.as<float>();
    if (ffnpl_.p) a.pl = Planes{ffnpl_.as<__nv_bfloat16>(), ffnpl_.as<__nv_bfloat16>() + m_max_ * kDffn};
    a.f8 = ffn8_.as<uint8_t>();
    return a;
}

static Epilogue into(const Act& dst, int prec) {
    Epilogue e;
    if (prec == kFp32) { e.out_f32 = const_cast<float*>(dst.f32); e.ldc = dst.ld; }
    else if (prec == kFp8 && dst.f8) { e.out_f8 = dst.f8; e.ld8 = dst.ld; }
    else { e.out_pl = dst.pl; e.ldp = dst.ld; }
    return e;
}

void Context::layernorm_to(const float* x, int rows, int d, const float* g, const float* b, float eps, const Act& dst) {
    if (prec_ == kFp32) launch_layernorm(x, rows, d, g, b, eps, nullptr, 0, const_cast<float*>(dst.f32), Planes{}, stream_);
    else if (prec_ == kFp8 && dst.f8) launch_layernorm(x, rows, d, g, b, eps, nullptr, 0, nullptr, Planes{}, stream_, nullptr, dst.f8);
    else launch_layernorm(x, rows, d, g, b, eps, nullptr, 0, nullptr, dst.pl, stream_);
}

void Context::linear(const Act& a, const Linear& w, int m, const Epilogue& ep_in) {
    Epilogue ep = ep_in;
    ep.bias = w.b;
    if (prec_ == kFp32) {
        launch_gemm_simt(a.f32, a.ld, w.w, m, w.n, w.k, ep, stream_);
    } else if (prec_ == kFp8 && a.f8 && w.w8) {
        // e4m3 x e4m3 (activations as they are, weights per output channel): acc * wscale[col] + bias
        const TcOperand opa = tc_make_operand_f8(a.f8, m, w.k, a.ld, kTcBlockM);
        ep.f8 = true; ep.col_scale = w.wscale;
        launch_gemm_tc(opa, w.op8, m, w.n, w.k, 1, ep, stream_);
    } else {
        const int64_t plane_stride = a.pl.lo - a.pl.hi;
        const TcOperand opa = tc_make_operand(a.pl.hi, m, w.k, a.ld, plane_stride, 2, kTcBlockM);
        launch_gemm_tc(opa, w.op, m, w.n, w.k, prec_ == kBf16x3 ? 2 : 1, ep, stream_);
    }
}

void Context::attention(const float* qkv, int ld, int d_model, int batch, int frames, int heads, const int* kv_len,
                        float* ctx_f32, Planes ctx_pl, int ldo, const Packing* pk) {
    if (simt_attention_) {
        FA_REQUIRE(pk == nullptr, "the CUDA-core attention runs on the uniform layout only");
        launch_attention_simt(qkv, qkv + d_model, qkv + 2 * d_model, ld, batch, frames, heads, d_model / heads, kv_len,
                              ctx_f32, ctx_pl, ldo, stream_);
    } else {
        const Planes pl = qkv_planes();
        launch_attention_tc(pl, pl.lo - pl.hi, ld, d_model, batch, frames, heads, d_model / heads, kv_len, ctx_f32, ctx_pl,
                            ldo, stream_, pk);
    }
}

Planes Context::qkv_planes() const {
    return Planes{qkvpl_.as<__nv_bfloat16>(), qkvpl_.as<__nv_bfloat16>() + m_max_ * 3 * kDllm};
}

// Epilogue of a fused q|k|v projection: what the attention kernel in use wants to read.
Epilogue Context::qkv_epilogue(int d_model, int heads, bool need_v_f32) const {
    Epilogue e;
    if (simt_attention_) {
        e.out_f32 = qkv_.as<float>(); e.ldc = 3 * d_model;
    } else {
        e.out_pl = qkv_planes(); e.ldp = 3 * d_model;
        e.pl_col_scale = (float)(1.4426950408889634 / std::sqrt((double)(d_model / heads)));   // d_k^-0.5 * log2(e)
        e.pl_col_scale_end = d_model;
        if (need_v_f32) { e.out_f32 = qkv_.as<float>(); e.ldc = 3 * d_model; e.f32_col_begin = 2 * d_model; }
    }
    return e;
}

void Context::tap(const char* name, const float* d, int64_t rows, int64_t cols) {
    if (!taps_on_) return;
    auto& slot = taps_[name];
    slot.first = {rows, cols};
    slot.second.reset(new DevBuf);
    slot.second->alloc((size_t)rows * cols * 4);
    FA_CUDA(cudaMemcpyAsync(slot.second->p, d, (size_t)rows * cols * 4, cudaMemcpyDeviceToDevice, stream_));
}

bool Context::read_tap(const std::string& name, std::vector<float>& out, std::vector<int64_t>& shape) {
    auto it = taps_.find(name);
    if (it == taps_.end()) return false;
    sync();
    shape = it->second.first;
    out.resize((size_t)shape[0] * shape[1]);
    FA_CUDA(cudaMemcpy(out.data(), it->second.second->p, out.size() * 4, cudaMemcpyDeviceToHost));
    return true;
}

std::vector<std::string> Context::tap_names() const {
    std::vector<std::string> v;
    for (auto& kv : taps_) v.push_back(kv.first);
    return v;
}

void Context::ensure_room(int batch, int64_t s_phys) const {
    FA_REQUIRE(finalized_, "context not finalized (load every tensor, then fa_ctx_finalize)");
    FA_REQUIRE(batch >= 1 && batch <= max_batch_, "batch exceeds the context's max_batch");
    FA_REQUIRE(s_phys >= 1 && s_phys <= max_samples_, "segment longer than the context's max_samples");
}

// ------------------------------------------------------------------------------------ layers

void Context::sanm_layer(const SanmLayer& L, bool first, int batch, int frames) {
    const int M = enc_rows(batch, frames);
    const bool f32 = prec_ == kFp32;
    float* x = x_.as<float>();
    const float* xin = first ? x0_.as<float>() : x;
    const Act h = h_act(L.d_in);
    layernorm_to(xin, M, L.d_in, L.ln1_g, L.ln1_b, 1e-5f, h);
    linear(h, L.qkv, M, qkv_epilogue(kDenc, 4, true));      // fp32 v feeds the FSMN branch
    const float* qkv = qkv_.as<float>();
    // x <- (x) + fsmn(v*m): the memory branch plus, except in layer 0, the block's residual
    launch_fsmn(qkv + 2 * kDenc, 3 * kDenc, L.fsmn_w, d_tvalid_, batch, frames, first ? nullptr : x, x, stream_, packing());
    const Act c = ctx_act(kDenc);
    attention(qkv, 3 * kDenc, kDenc, batch, frames, 4, d_tvalid_, f32 ? const_cast<float*>(c.f32) : nullptr,
              f32 ? Planes{} : c.pl, kDenc, packing());
    Epilogue eo;
    eo.resid = x; eo.ldr = kDenc; eo.out_f32 = x; eo.ldc = kDenc;
    linear(c, L.out, M, eo);
    if (first) return;
    const Act h2 = h_act(kDenc);
    layernorm_to(x, M, kDenc, L.ln2_g, L.ln2_b, 1e-5f, h2);
    const Act f = ffn_act(kDffn);
    Epilogue e1 = into(f, prec_);
    e1.relu = true;
    linear(h2, L.w1, M, e1);
    Epilogue e2;
    e2.resid = x; e2.ldr = kDenc; e2.out_f32 = x; e2.ldc = kDenc;
    linear(f, L.w2, M, e2);
}

void Context::projector(const Projector& P, const Act& in, int batch, int frames, const int* kv_len, const Packing* pk) {
    const int M = pk ? pk->total_rows : batch * frames, d = P.d;
    const bool f32 = prec_ == kFp32;
    float* x = x_.as<float>();
    const Act f = ffn_act(kDffn);
    Epilogue e1 = into(f, prec_);
    e1.relu = true;
    linear(in, P.lin1, M, e1);
    Epilogue e2;
    e2.out_f32 = x; e2.ldc = d;
    linear(f, P.lin2, M, e2);
    for (const MhaBlock& B : P.blocks) {
        const Act h = h_act(d);
        layernorm_to(x, M, d, B.ln1_g, B.ln1_b, 1e-12f, h);
        linear(h, B.qkv, M, qkv_epilogue(d, P.heads, false));
        const Act c = ctx_act(d);
        attention(qkv_.as<float>(), 3 * d, d, batch, frames, P.heads, kv_len, f32 ? const_cast<float*>(c.f32) : nullptr,
                  f32 ? Planes{} : c.pl, d, pk);
        Epilogue eo;
        eo.resid = x; eo.ldr = d; eo.out_f32 = x; eo.ldc = d;
        linear(c, B.out, M, eo);
        layernorm_to(x, M, d, B.ln2_g, B.ln2_b, 1e-12f, h);
        const Act ff = ffn_act(d / 4);
        Epilogue ea = into(ff, prec_);
        ea.relu = true;
        linear(h, B.w1, M, ea);
        Epilogue eb;
        eb.resid = x; eb.ldr = d; eb.out_f32 = x; eb.ldc = d;
        linear(ff, B.w2, M, eb);
    }
}

// ------------------------------------------------------------------------------------ graphs

// Per-call lengths and packing tables, staged in one pinned block and uploaded with one copy.  Layout of the block
// (ints; B = max_batch_): n_valid[B] | t_valid[B] | target_len[B] | seg_off[B+1] | ctc seg_off[B+1] | ctc len[B] |
// ctc key bias[B] (float) | t_phys[B] | pad to 16 bytes | tile table[max_tiles_] (int4) | ctc tile table[max_tiles_] (int4)
void Context::stage_lengths(int batch, int64_t s_phys, const int64_t* h_ilens, bool allow_packed, const int64_t* h_phys) {
    const int slot = len_next_;
    len_next_ = (len_next_ + 1) % kLenSlots;
    FA_CUDA(cudaEventSynchronize(len_ev_[slot]));
    const int B = max_batch_;
    int* hl = h_lens_ + (size_t)slot * len_ints_;
    int* h_tv = hl + B;
    int* h_off = hl + 3 * B;
    int* h_off_c = h_off + B + 1;
    int* h_len_c = h_off_c + B + 1;
    float* h_bias_c = reinterpret_cast<float*>(h_len_c + B);
    int* h_tphys = h_len_c + 2 * B;
    int* h_tab = hl + tab_off_;
    int* h_tab_c = h_tab + 4 * max_tiles_;
    const int* d0 = d_nvalid_;
    const int frames = lfr_frames_of(s_phys);
    int total = 0, longest = 0;
    double sq = 0.0;
    for (int b = 0; b < batch; ++b) {
        const int64_t nv = h_ilens[b];
        FA_REQUIRE(nv >= 1 && nv <= s_phys, "ilens must satisfy 1 <= ilens[b] <= samples");
        FA_REQUIRE(!h_phys || (h_phys[b] >= nv && h_phys[b] <= s_phys), "phys must satisfy ilens[b] <= phys[b] <= samples");
        h_tphys[b] = h_phys ? lfr_frames_of(h_phys[b]) : frames;
        hl[b] = (int)nv;
        h_tv[b] = lfr_frames_of(nv);
        hl[2 * B + b] = target_len_of(nv);
        h_off[b] = total;
        total += h_tv[b];
        longest = std::max(longest, h_tv[b]);
        sq += (double)h_tv[b] * h_tv[b];
    }
    h_off[batch] = total;
    // Packed execution only pays (and only differs) when the batch holds padded frames
    packed_ = allow_packed && allow_packed_env_ && prec_ != kFp32 && !simt_attention_ && !taps_on_ && total < batch * frames;
    if (h_phys) {
        // a ragged batch only exists in the packed layout (every segment keeps its own physical length)
        FA_REQUIRE(allow_packed && prec_ != kFp32 && !simt_attention_ && !taps_on_,
                   "per-segment physical lengths need the padding-free path (a tensor-core precision mode, no debug taps)");
        packed_ = true;
    }
    d_tphys_ = nullptr;
    ctc_packed_ready_ = false;
    if (packed_) {
        // attention items are dealt round-robin over the CTAs: longest segments first, so that every CTA's share
        // mixes long and short items (the cost of an item is proportional to its segment's length)
        std::vector<int> order(batch);
        for (int b = 0; b < batch; ++b) order[b] = b;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return h_tv[x] > h_tv[y]; });
        auto fill_table = [&](int* tab, const int* off, const int* len) {       // int4 per query tile: {row0, len, tile, segment}
            int tiles = 0;
            for (int b : order)
                for (int qt = 0; qt < cdiv(len[b], 128); ++qt, ++tiles) {
                    tab[4 * tiles] = off[b]; tab[4 * tiles + 1] = len[b]; tab[4 * tiles + 2] = qt; tab[4 * tiles + 3] = b;
                }
            return tiles;
        };
        pk_ = Packing{};
        pk_.seg_off = d0 + (h_off - hl);
        pk_.tile_tab = reinterpret_cast<const int4*>(d0 + tab_off_);
        pk_.total_rows = total; pk_.total_tiles = fill_table(h_tab, h_off, h_tv); pk_.max_len = longest; pk_.sum_len_sq = sq;
        // the CTC head's rows: valid frames, then ONE row standing for all of the segment's zero-padded frames (counted
        // at the segment's own physical length in a ragged batch), whose key weighs n_pad in every softmax
        int total_c = 0;
        double sq_c = 0.0;
        for (int b = 0; b < batch; ++b) {
            const int n_pad = h_tphys[b] - h_tv[b];
            h_len_c[b] = h_tv[b] + (n_pad > 0 ? 1 : 0);
            h_bias_c[b] = n_pad > 1 ? std::log2((float)n_pad) : 0.f;
            h_off_c[b] = total_c;
            total_c += h_len_c[b];
            sq_c += (double)h_len_c[b] * h_len_c[b];
        }
        h_off_c[batch] = total_c;
        pk_ctc_ = Packing{};
        pk_ctc_.seg_off = d0 + (h_off_c - hl);
        pk_ctc_.tile_tab = pk_.tile_tab + max_tiles_;
        pk_ctc_.last_key_bias = reinterpret_cast<const float*>(d0 + (h_len_c - hl) + B);
        pk_ctc_.total_rows = total_c; pk_ctc_.total_tiles = fill_table(h_tab_c, h_off_c, h_len_c); pk_ctc_.max_len = longest + 1;
        pk_ctc_.sum_len_sq = sq_c;
        d_len_ctc_ = d0 + (h_len_c - hl);
        if (h_phys) d_tphys_ = d0 + (h_tphys - hl);
    }
    FA_CUDA(cudaMemcpyAsync(lens_.p, hl, (size_t)len_ints_ * sizeof(int), cudaMemcpyHostToDevice, stream_));
    FA_CUDA(cudaEventRecord(len_ev_[slot], stream_));
}

// a1-a2 for segments b0 .. b0+nb-1 of the batch (segments are independent): waveform -> log-mel
void Context::front_end(const float* d_audio, int b0, int nb, int64_t s_phys) {
    const int t_mel = (int)(s_phys / kHop + 1);
    double* parts = partials_.as<double>() + (size_t)b0 * kMeanParts;
    launch_segment_sums(d_audio, nb, s_phys, d_nvalid_ + b0, parts, stream_);
    if (fbank_tc_) {
        launch_fbank_tc(d_audio, nb, s_phys, d_nvalid_ + b0, parts, dft_planes_, melfb_t_, mel_range_, yplanes_.p,
                        power_.as<float>(), logmel_.as<float>() + (size_t)b0 * t_mel * kMels, t_mel, stream_);
    } else {
        launch_fbank(d_audio, nb, s_phys, d_nvalid_ + b0, parts, dft_t_, melfb_t_, mel_range_,
                     logmel_.as<float>() + (size_t)b0 * t_mel * kMels, t_mel, stream_);
    }
}

void Context::encode_dev(const float* d_audio, int batch, int64_t s_phys, const int64_t* h_ilens, float* d_enc,
                         float* d_adaptor) {
    ensure_room(batch, s_phys);
    set_device();
    stage_lengths(batch, s_phys, h_ilens, true);
    front_end(d_audio, 0, batch, s_phys);
    encoder_graph(batch, s_phys, d_enc, d_adaptor);
}

// a3 onwards: LFR + position -> 70 SAN-M layers -> enc ; adaptor -> adaptor_output
void Context::encoder_graph(int batch, int64_t s_phys, float* d_enc, float* d_adaptor, bool record_events) {
    const int t_mel = (int)(s_phys / kHop + 1), frames = lfr_frames_of(s_phys);
    const int M = enc_rows(batch, frames);                  // packed: valid frames only
    const Packing* pk = packing();
    const int* seg_off = pk ? pk->seg_off : nullptr;
    tap("logmel", logmel_.as<float>(), (int64_t)batch * t_mel, kMels);
    float* lfr_raw = taps_on_ ? adaptor_out_.as<float>() : nullptr;     // borrowed scratch for the tap
    launch_lfr_embed(logmel_.as<float>(), batch, t_mel, frames, d_nvalid_, pos_enc_, x0_.as<float>(), lfr_raw, stream_, seg_off);
    if (lfr_raw) tap("lfr", lfr_raw, M, kDin);

    // the mask sweeps of model_definition.py:210,213 zero the padded frames; packed rows have none
    const int* sweep = pk ? nullptr : d_tvalid_;
    float* x = x_.as<float>();
    for (int i = 0; i < kEncLayers; ++i) {
        sanm_layer(enc_layers_[i], i == 0, batch, frames);
        if (i == 0) tap("layer0", x, M, kDenc);
        if (i == 1) tap("layer1", x, M, kDenc);
        if (i == 49) {
            launch_layernorm(x, M, kDenc, after_g_, after_b_, 1e-5f, sweep, frames, x, Planes{}, stream_);
            tap("layer49", x, M, kDenc);
        }
    }
    const bool f32 = prec_ == kFp32;
    Planes encpl;
    if (!f32) encpl = Planes{encpl_.as<__nv_bfloat16>(), encpl_.as<__nv_bfloat16>() + m_max_ * kDenc};
    const bool f8 = prec_ == kFp8;
    Act in; in.f32 = d_enc; in.pl = encpl; in.f8 = enc8_.as<uint8_t>(); in.ld = kDenc;
    if (pk) {
        // enc_output leaves in the reference's physical [batch][frames] shape, padded frames zero (the CTC head runs on
        // it as it is: unmasked, every physical frame — F7); the adaptor goes on with the packed rows
        launch_layernorm(x, batch * frames, kDenc, tp_g_, tp_b_, 1e-5f, d_tvalid_, frames, d_enc, Planes{}, stream_, seg_off);
        in = h_act(kDenc);
        layernorm_to(x, M, kDenc, tp_g_, tp_b_, 1e-5f, in);
    } else {
        launch_layernorm(x, M, kDenc, tp_g_, tp_b_, 1e-5f, d_tvalid_, frames, d_enc, f8 ? Planes{} : encpl, stream_, nullptr,
                         f8 ? in.f8 : nullptr);
    }
    if (record_events) FA_CUDA(cudaEventRecord(ev_enc_, stream_));       // enc_output is final: a download may start

    if (pk && fused_ctc_next_) {
        // the CTC head's input, in its own packing, before the adaptor reuses the buffer `in` lives in
        const Act c = ctc_packed_input();
        if (f8) launch_repack_rows(in.f8, c.f8, kDenc, pk->seg_off, pk_ctc_.seg_off, d_tvalid_, batch, pk->max_len, stream_);
        else {
            launch_repack_rows(in.pl.hi, c.pl.hi, kDenc * 2, pk->seg_off, pk_ctc_.seg_off, d_tvalid_, batch, pk->max_len, stream_);
            launch_repack_rows(in.pl.lo, c.pl.lo, kDenc * 2, pk->seg_off, pk_ctc_.seg_off, d_tvalid_, batch, pk->max_len, stream_);
        }
        ctc_packed_ready_ = true;
    }
    projector(adaptor_, in, batch, frames, d_tvalid_, pk);
    launch_row_keep(x, d_adaptor, batch, frames, kDllm, d_tlen_, stream_, seg_off);
    if (record_events) FA_CUDA(cudaEventRecord(ev_ad_, stream_));
}

// ------------------------------------------------------------------------------------ graph replay
bool Context::use_graph(int batch) const { return capturable_ && batch <= graph_max_batch_ && !g_prof_on && !taps_on_; }

// Runs `body` (kernel launches on stream_ only, no host synchronisation) as a CUDA graph: captured and instantiated
// the first time a (kind, batch, size) is seen, replayed afterwards.  Every pointer a launch carries is a workspace
// or weight address of this context and the per-call lengths are read from device memory, so a replay is exact.
template <class F>
void Context::run_graphed(int kind, int batch, int64_t size, F&& body) {
    const auto key = std::make_tuple(kind, batch, size);
    auto it = graphs_.find(key);
    if (it == graphs_.end()) {
        if (graphs_.size() >= 32) {                      // odd lengths every call: do not hoard executables
            for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second.exec);
            graphs_.clear();
        }
        const int64_t before = g_launches;
        FA_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        cudaGraph_t graph = nullptr;
        try {
            body();
        } catch (...) {
            cudaStreamEndCapture(stream_, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        FA_CUDA(cudaStreamEndCapture(stream_, &graph));
        GraphEntry ge;
        ge.launches = g_launches - before;
        g_launches = before;                             // nothing has run yet; the launch below counts them
        const cudaError_t e = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        FA_CUDA(e);
        it = graphs_.emplace(key, ge).first;
    }
    FA_CUDA(cudaGraphLaunch(it->second.exec, stream_));
    g_launches += it->second.launches;
}


Act Context::ctc_packed_input() const {
    Act a; a.ld = kDenc;
    a.pl = Planes{encpl_.as<__nv_bfloat16>(), encpl_.as<__nv_bfloat16>() + m_max_ * kDenc};
    a.f8 = enc8_.as<uint8_t>();
    return a;
}

void Context::ctc_dev(const float* d_enc, int batch, int frames, int32_t* d_ids) {
    FA_REQUIRE(finalized_, "context not finalized");
    FA_REQUIRE(batch >= 1 && batch <= max_batch_ && frames >= 1 && frames <= t_max_, "CTC input exceeds the context's capacity");
    set_device();
    const int M = batch * frames;
    Act in; in.f32 = d_enc; in.ld = kDenc;
    if (prec_ == kFp8) {
        in.f8 = enc8_.as<uint8_t>();
        launch_to_e4m3(d_enc, (int64_t)M * kDenc, in.f8, stream_);
    } else if (prec_ != kFp32) {
        in.pl = Planes{encpl_.as<__nv_bfloat16>(), encpl_.as<__nv_bfloat16>() + m_max_ * kDenc};
        launch_split_planes(d_enc, (int64_t)M * kDenc, in.pl, stream_);
    }
    ctc_graph(in, batch, frames, d_ids, nullptr);
}

// The head in the same call as the encoder: a packed batch keeps the head's packing (one row per segment for all its
// zero-padded frames, whose key counts n_pad times); otherwise the physical rows of d_enc.
void Context::ctc_after_encoder(const float* d_enc, int batch, int frames, int32_t* d_ids) {
    if (packed_ && ctc_packed_ready_) {
        ctc_graph(ctc_packed_input(), batch, frames, d_ids, &pk_ctc_);
        return;
    }
    ctc_dev(d_enc, batch, frames, d_ids);
}

void Context::ctc_graph(const Act& in, int batch, int frames, int32_t* d_ids, const Packing* pk) {
    const int M = pk ? pk->total_rows : batch * frames;
    const bool f32 = prec_ == kFp32;
    projector(ctc_, in, batch, frames, pk ? d_len_ctc_ : nullptr, pk);
    float* x = x_.as<float>();
    tap("ctc_h", x, M, kDenc);
    int32_t* ids_rows = pk ? tokens_.as<int32_t>() : d_ids;           // packed rows -> scratch, then unpacked
    if (f32) {
        for (int r0 = 0; r0 < M; r0 += logits_rows_) {
            const int rows = std::min(logits_rows_, M - r0);
            Epilogue e;
            e.bias = ctc_lo_.b; e.out_f32 = logits_.as<float>(); e.ldc = vocab_;
            launch_gemm_simt(x + (size_t)r0 * kDenc, kDenc, ctc_lo_.w, rows, vocab_, kDenc, e, stream_);
            launch_argmax_rows(logits_.as<float>(), rows, vocab_, vocab_, ids_rows + r0, stream_);
        }
    } else {
        vocab_argmax(x, h_act(kDenc).pl, ctc_lo_, vocab_wnorm_, M, vocab_rescore_ ? cand_workspace() : VocabCand{},
                     amax_val_.as<float>(), amax_idx_.as<int32_t>(), ids_rows);
    }
    if (pk) launch_unpack_ids(ids_rows, d_ids, batch, frames, pk->seg_off, d_tvalid_, stream_, d_tphys_);
}

void Context::front_half_dev(const float* d_audio, int batch, int64_t s_phys, const int64_t* h_ilens, float* d_enc,
                             float* d_adaptor, int32_t* d_ids, const int64_t* h_phys) {
    ensure_room(batch, s_phys);
    set_device();
    if (!d_enc) d_enc = enc_.as<float>();
    if (!d_adaptor) d_adaptor = adaptor_out_.as<float>();
    stage_lengths(batch, s_phys, h_ilens, true, h_phys);
    front_end(d_audio, 0, batch, s_phys);
    fused_ctc_next_ = true;
    try { encoder_graph(batch, s_phys, d_enc, d_adaptor); } catch (...) { fused_ctc_next_ = false; throw; }
    fused_ctc_next_ = false;
    ctc_after_encoder(d_enc, batch, lfr_frames_of(s_phys), d_ids);
}

VocabCand Context::cand_workspace() const {
    VocabCand c;
    c.run_max = cand_meta_.as<int32_t>();
    c.count = c.run_max + m_max_;
    c.bound2 = reinterpret_cast<const float*>(c.count + m_max_);
    c.overflowed = c.count + 2 * m_max_;
    c.list = cand_list_.as<int2>();
    c.cap = kVocabCandCap;
    return c;
}

// The logits never reach HBM.  With candidate lists (bf16x3 mode, batches of at least half a wave of 256-row blocks):
// one bf16 product per element finds every column that can hold the row maximum and those few are rescored exactly
// (kernels.h); rows with more than kVocabCandCap such columns take the gated second-chance pass.  Without: the
// projection runs at the context's precision and the epilogue keeps a running (max, first index) per 128 columns.
void Context::vocab_argmax(const float* x, Planes x_pl, const Linear& lo, float w_norm_max, int m, VocabCand cand,
                           float* amax_val, int32_t* amax_idx, int32_t* d_ids) {
    const int64_t plane_stride = x_pl.lo - x_pl.hi;
    const TcOperand opa = tc_make_operand(x_pl.hi, m, lo.k, lo.k, plane_stride, 2, kTcBlockM);
    Epilogue e;
    e.bias = lo.b;
    // a short batch would spread every row over many concurrent pairs, each starting its running maximum from
    // nothing: long lists for flat logits and no time to win back, so it keeps the three-product projection
    const bool use_cand = cand.list && 2 * cdiv(m, 2 * kTcBlockM) >= tc_num_pairs();
    // either way the decision is the fp32 argmax over rescored columns (kernels.h), so a segment's ids do not depend
    // on how many segments share the call; the one-product speed mode keeps the plain combine
    const bool rescore_slots = cand.list != nullptr;
    if (use_cand) {
        launch_vocab_prepare(x, m, lo.k, w_norm_max, x_pl, cand, stream_);
        e.cand = cand;
        launch_gemm_tc(opa, lo.op, m, lo.n, lo.k, 1, e, stream_);
        launch_vocab_rescore(x, lo.w, lo.b, m, lo.k, lo.n, cand, d_ids, stream_);
        Epilogue e2;
        e2.bias = lo.b; e2.amax_val = amax_val; e2.amax_idx = amax_idx; e2.gate = cand.overflowed;
        launch_gemm_tc(opa, lo.op, m, lo.n, lo.k, 2, e2, stream_);
        launch_vocab_rescore_slots(x, lo.w, lo.b, m, lo.k, lo.n, w_norm_max, amax_val, tc_argmax_tiles(lo.n), d_ids, stream_,
                                   cand.count, cand.cap);
        return;
    }
    launch_split_planes(x, (int64_t)m * lo.k, x_pl, stream_);
    e.amax_val = amax_val; e.amax_idx = amax_idx;
    launch_gemm_tc(opa, lo.op, m, lo.n, lo.k, prec_ == kBf16x3 ? 2 : 1, e, stream_);
    if (rescore_slots && prec_ == kBf16x3)
        launch_vocab_rescore_slots(x, lo.w, lo.b, m, lo.k, lo.n, w_norm_max, amax_val, tc_argmax_tiles(lo.n), d_ids, stream_);
    else
        launch_argmax_combine(amax_val, amax_idx, m, tc_argmax_tiles(lo.n), d_ids, stream_);
}

void Context::collapse_dev(const int32_t* d_ids, int batch, int frames, int32_t* d_tokens, int32_t* d_starts,
                           int32_t* d_counts) {
    set_device();
    launch_ctc_collapse(d_ids, batch, frames, vocab_ - 1, d_tokens, d_starts, d_counts, stream_);
}

// Host variants.  The audio goes up in groups of a few segments on the copy stream and the front end of a
// group starts as soon as its samples have landed; enc_output goes down while the adaptor runs and
// adaptor_output while the CTC head runs.  Pinned host buffers make all of it asynchronous; pageable
// buffers still work (the copies then serialise on the host).
void Context::upload_and_front_end(const float* audio_host, int nb, int64_t s_phys) {
    const int groups = std::min<int>(8, nb);
    for (int g = 0; g < groups; ++g) {
        const int b0 = (int)((int64_t)nb * g / groups), b1 = (int)((int64_t)nb * (g + 1) / groups);
        FA_CUDA(cudaMemcpyAsync(audio_.as<float>() + (size_t)b0 * s_phys, audio_host + (size_t)b0 * s_phys,
                                (size_t)(b1 - b0) * s_phys * 4, cudaMemcpyHostToDevice, copy_stream_));
        FA_CUDA(cudaEventRecord(ev_up_[g], copy_stream_));
        FA_CUDA(cudaStreamWaitEvent(stream_, ev_up_[g], 0));
        front_end(audio_.as<float>() + (size_t)b0 * s_phys, b0, b1 - b0, s_phys);
    }
}

void Context::download_async(void* host, const void* dev, size_t bytes, cudaEvent_t after) {
    FA_CUDA(cudaStreamWaitEvent(copy_stream_, after, 0));
    FA_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, copy_stream_));
}

void Context::encode_host(const float* audio, int batch, int64_t s_phys, const int64_t* ilens, float* enc,
                          float* adaptor) {
    front_half_host(audio, batch, s_phys, ilens, enc, adaptor, nullptr);
}

void Context::ctc_host(const float* enc, int batch, int frames, int32_t* ids) {
    // checked before the first copy: an oversized `frames` must not write past the workspace
    FA_REQUIRE(finalized_, "context not finalized");
    FA_REQUIRE(batch >= 1 && frames >= 1 && frames <= t_max_, "CTC input exceeds the context's capacity");
    set_device();
    for (int b0 = 0; b0 < batch; b0 += max_batch_) {
        const int nb = std::min(max_batch_, batch - b0);
        FA_CUDA(cudaMemcpyAsync(enc_.p, enc + (size_t)b0 * frames * kDenc, (size_t)nb * frames * kDenc * 4, cudaMemcpyHostToDevice, stream_));
        if (use_graph(nb)) run_graphed(2, nb, frames, [&] { ctc_dev(enc_.as<float>(), nb, frames, ids_.as<int32_t>()); });
        else ctc_dev(enc_.as<float>(), nb, frames, ids_.as<int32_t>());
        FA_CUDA(cudaMemcpyAsync(ids + (size_t)b0 * frames, ids_.p, (size_t)nb * frames * 4, cudaMemcpyDeviceToHost, stream_));
        FA_CUDA(cudaStreamSynchronize(stream_));
    }
}

// ids == nullptr: encoder session only
void Context::front_half_host(const float* audio, int batch, int64_t s_phys, const int64_t* ilens, float* enc,
                              float* adaptor, int32_t* ids, float* const* embd_rows, int64_t* rows_out, const int64_t* phys) {
    FA_REQUIRE(!phys || ids, "a ragged batch runs both graphs in one call");
    ensure_room(1, s_phys);
    set_device();
    const int frames = lfr_frames_of(s_phys);
    if (embd_rows) adaptor = nullptr;
    // embedding handoff: only the rows the LLM reads leave the device, straight to where the caller wants them
    auto hand_off = [&](int b0, int nb, cudaStream_t st) {
        for (int i = 0; embd_rows && i < nb; ++i) {
            const int tl = target_len_of(ilens[b0 + i]);
            if (rows_out) rows_out[b0 + i] = tl;
            FA_REQUIRE(embd_rows[b0 + i] != nullptr, "embd_rows holds a null destination");
            FA_CUDA(cudaMemcpyAsync(embd_rows[b0 + i], adaptor_out_.as<float>() + (size_t)i * frames * kDllm, (size_t)tl * kDllm * 4,
                                    cudaMemcpyDefault, st));
        }
    };
    for (int b0 = 0; b0 < batch; b0 += max_batch_) {
        const int nb = std::min(max_batch_, batch - b0);
        stage_lengths(nb, s_phys, ilens + b0, phys || !use_graph(nb), phys ? phys + b0 : nullptr);
        if (!phys && use_graph(nb)) {
            // launch-bound regime: one upload, one graph (front end, encoder, adaptor and, if asked for, the CTC head),
            // downloads behind it on the same stream
            FA_CUDA(cudaMemcpyAsync(audio_.p, audio + (size_t)b0 * s_phys, (size_t)nb * s_phys * 4, cudaMemcpyHostToDevice, stream_));
            run_graphed(ids ? 1 : 0, nb, s_phys, [&] {
                front_end(audio_.as<float>(), 0, nb, s_phys);
                encoder_graph(nb, s_phys, enc_.as<float>(), adaptor_out_.as<float>(), false);
                if (ids) ctc_dev(enc_.as<float>(), nb, frames, ids_.as<int32_t>());
            });
            if (enc) FA_CUDA(cudaMemcpyAsync(enc + (size_t)b0 * frames * kDenc, enc_.p, (size_t)nb * frames * kDenc * 4, cudaMemcpyDeviceToHost, stream_));
            if (adaptor) FA_CUDA(cudaMemcpyAsync(adaptor + (size_t)b0 * frames * kDllm, adaptor_out_.p, (size_t)nb * frames * kDllm * 4, cudaMemcpyDeviceToHost, stream_));
            hand_off(b0, nb, stream_);
            if (ids) FA_CUDA(cudaMemcpyAsync(ids + (size_t)b0 * frames, ids_.p, (size_t)nb * frames * 4, cudaMemcpyDeviceToHost, stream_));
            FA_CUDA(cudaStreamSynchronize(stream_));
            continue;
        }
        upload_and_front_end(audio + (size_t)b0 * s_phys, nb, s_phys);
        fused_ctc_next_ = ids != nullptr;
        try { encoder_graph(nb, s_phys, enc_.as<float>(), adaptor_out_.as<float>()); } catch (...) { fused_ctc_next_ = false; throw; }
        fused_ctc_next_ = false;
        if (enc) download_async(enc + (size_t)b0 * frames * kDenc, enc_.p, (size_t)nb * frames * kDenc * 4, ev_enc_);
        if (adaptor) download_async(adaptor + (size_t)b0 * frames * kDllm, adaptor_out_.p, (size_t)nb * frames * kDllm * 4, ev_ad_);
        if (embd_rows) { FA_CUDA(cudaStreamWaitEvent(copy_stream_, ev_ad_, 0)); hand_off(b0, nb, copy_stream_); }
        if (ids) {
            ctc_after_encoder(enc_.as<float>(), nb, frames, ids_.as<int32_t>());
            FA_CUDA(cudaMemcpyAsync(ids + (size_t)b0 * frames, ids_.p, (size_t)nb * frames * 4, cudaMemcpyDeviceToHost, stream_));
        }
        FA_CUDA(cudaStreamSynchronize(stream_));
        FA_CUDA(cudaStreamSynchronize(copy_stream_));
    }
}

// ------------------------------------------------------------------------------------ kernel-level test hooks

void Context::test_linear(const float* a, const float* w, const float* bias, const float* resid, int m, int n, int k,
                          int relu, int precision, float* out, float* out_planes_sum) {
    set_device();
    DevBuf da, dw, db, dr, dout, dpl_a, dpl_w, dpl_o;
    da.alloc((size_t)m * k * 4); dw.alloc((size_t)n * k * 4); db.alloc((size_t)n * 4); dout.alloc((size_t)m * n * 4);
    FA_CUDA(cudaMemcpy(da.p, a, da.bytes, cudaMemcpyHostToDevice));
    FA_CUDA(cudaMemcpy(dw.p, w, dw.bytes, cudaMemcpyHostToDevice));
    FA_CUDA(cudaMemcpy(db.p, bias, db.bytes, cudaMemcpyHostToDevice));
    if (resid) { dr.alloc((size_t)m * n * 4); FA_CUDA(cudaMemcpy(dr.p, resid, dr.bytes, cudaMemcpyHostToDevice)); }
    Epilogue e;
    e.bias = db.as<float>(); e.resid = dr.as<float>(); e.ldr = n; e.relu = relu != 0; e.out_f32 = dout.as<float>(); e.ldc = n;
    if (resid && getenv("FUNASR_B200_TEST_INPLACE")) {      // the engine's form: the residual is the output buffer (x += ...)
        FA_CUDA(cudaMemcpy(dout.p, resid, dout.bytes, cudaMemcpyHostToDevice));
        e.resid = dout.as<float>();
    }
    const int ldp = precision == kFp8 ? (n + 15) / 16 * 16 : (n + 7) / 8 * 8;
    if (out_planes_sum && getenv("FUNASR_B200_TEST_NO_F32")) e.out_f32 = nullptr;     // timing aid: the engine's planes-only epilogue
    if (out_planes_sum && precision != kFp8) {
        dpl_o.alloc((size_t)2 * m * ldp * 2);
        FA_CUDA(cudaMemsetAsync(dpl_o.p, 0, dpl_o.bytes, stream_));
        e.out_pl = Planes{dpl_o.as<__nv_bfloat16>(), dpl_o.as<__nv_bfloat16>() + (size_t)m * ldp};
        e.ldp = ldp;
    }
    DevBuf d8a, d8w, d8o;
    if (precision == kFp32) {
        launch_gemm_simt(da.as<float>(), k, dw.as<float>(), m, n, k, e, stream_);
    } else if (precision == kFp8) {
        // e4m3 operands: activations converted as they are, weights per output channel; with out_planes_sum the e4m3
        // output form is exercised too (handed back decoded)
        FA_REQUIRE(k % 16 == 0, "fp8 test_linear needs K % 16 == 0");
        const size_t scale_off = ((size_t)n * k + 255) / 256 * 256;
        d8a.alloc((size_t)m * k); d8w.alloc(scale_off + (size_t)n * 4);
        launch_to_e4m3(da.as<float>(), (int64_t)m * k, d8a.as<uint8_t>(), stream_);
        float* sc = reinterpret_cast<float*>(d8w.as<uint8_t>() + scale_off);
        launch_quant_rows_e4m3(dw.as<float>(), n, k, d8w.as<uint8_t>(), sc, stream_);
        e.out_pl = Planes{}; e.ldp = 0;
        e.f8 = true; e.col_scale = sc;
        const TcOperand o8a = tc_make_operand_f8(d8a.as<uint8_t>(), m, k, k, kTcBlockM), o8w = tc_make_operand_f8(d8w.as<uint8_t>(), n, k, k, 64);
        launch_gemm_tc(o8a, o8w, m, n, k, 1, e, stream_);
        if (out_planes_sum) {                       // the e4m3 output form is a launch of its own (it shares the staging tile)
            d8o.alloc((size_t)m * ldp);
            FA_CUDA(cudaMemsetAsync(d8o.p, 0, d8o.bytes, stream_));
            Epilogue e8 = e;
            e8.out_f32 = nullptr; e8.resid = nullptr; e8.out_f8 = d8o.as<uint8_t>(); e8.ld8 = ldp;
            launch_gemm_tc(o8a, o8w, m, n, k, 1, e8, stream_);
        }
    } else {
        dpl_a.alloc((size_t)2 * m * k * 2); dpl_w.alloc((size_t)2 * n * k * 2);
        Planes pa{dpl_a.as<__nv_bfloat16>(), dpl_a.as<__nv_bfloat16>() + (size_t)m * k};
        Planes pw{dpl_w.as<__nv_bfloat16>(), dpl_w.as<__nv_bfloat16>() + (size_t)n * k};
        launch_split_planes(da.as<float>(), (int64_t)m * k, pa, stream_);
        launch_split_planes(dw.as<float>(), (int64_t)n * k, pw, stream_);
        const TcOperand oa = tc_make_operand(pa.hi, m, k, k, (int64_t)m * k, 2, kTcBlockM);
        const TcOperand ow = tc_make_weight(pw.hi, n, k, (int64_t)n * k, 2);
        launch_gemm_tc(oa, ow, m, n, k, precision == kBf16x3 ? 2 : 1, e, stream_);
    }
    FA_CUDA(cudaStreamSynchronize(stream_));
    FA_CUDA(cudaMemcpy(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
    if (out_planes_sum && precision == kFp8) {
        std::vector<uint8_t> h8((size_t)m * ldp);
        FA_CUDA(cudaMemcpy(h8.data(), d8o.p, d8o.bytes, cudaMemcpyDeviceToHost));
        auto decode = [](uint8_t b) {                    // e4m3fn: 1-4-3, bias 7, no infinities
            const int e = (b >> 3) & 15, mant = b & 7;
            const float v = e == 0 ? std::ldexp((float)mant, -9) : std::ldexp(1.0f + mant / 8.0f, e - 7);
            return (b & 0x80) ? -v : v;
        };
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) out_planes_sum[(size_t)i * n + j] = decode(h8[(size_t)i * ldp + j]);
        return;
    }
    if (out_planes_sum) {
        std::vector<__nv_bfloat16> hp((size_t)2 * m * ldp);
        FA_CUDA(cudaMemcpy(hp.data(), dpl_o.p, dpl_o.bytes, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j)
                out_planes_sum[(size_t)i * n + j] = __bfloat162float(hp[(size_t)i * ldp + j]) +
                                                    __bfloat162float(hp[(size_t)m * ldp + (size_t)i * ldp + j]);
    }
}

void Context::test_vocab_argmax(const float* a, const float* w, const float* bias, int m, int n, int k, int precision,
                                int32_t* ids) {
    set_device();
    DevBuf da, dw, db, dids;
    da.alloc((size_t)m * k * 4); dw.alloc((size_t)n * k * 4); db.alloc((size_t)n * 4); dids.alloc((size_t)m * 4);
    FA_CUDA(cudaMemcpy(da.p, a, da.bytes, cudaMemcpyHostToDevice));
    FA_CUDA(cudaMemcpy(dw.p, w, dw.bytes, cudaMemcpyHostToDevice));
    FA_CUDA(cudaMemcpy(db.p, bias, db.bytes, cudaMemcpyHostToDevice));
    if (precision == kFp32) {
        DevBuf dl;
        dl.alloc((size_t)m * n * 4);
        Epilogue e;
        e.bias = db.as<float>(); e.out_f32 = dl.as<float>(); e.ldc = n;
        launch_gemm_simt(da.as<float>(), k, dw.as<float>(), m, n, k, e, stream_);
        launch_argmax_rows(dl.as<float>(), m, n, n, dids.as<int32_t>(), stream_);
        FA_CUDA(cudaStreamSynchronize(stream_));
    } else {
        DevBuf dpl_a, dpl_w, dv, di, dmeta, dlist, dnorm;
        const int tiles = tc_argmax_tiles(n);
        dpl_a.alloc((size_t)2 * m * k * 2); dpl_w.alloc((size_t)2 * n * k * 2);
        dv.alloc((size_t)m * tiles * 4); di.alloc((size_t)m * tiles * 4);
        Planes pa{dpl_a.as<__nv_bfloat16>(), dpl_a.as<__nv_bfloat16>() + (size_t)m * k};
        Planes pw{dpl_w.as<__nv_bfloat16>(), dpl_w.as<__nv_bfloat16>() + (size_t)n * k};
        launch_split_planes(dw.as<float>(), (int64_t)n * k, pw, stream_);
        Linear lo;
        lo.w = dw.as<float>(); lo.b = db.as<float>(); lo.n = n; lo.k = k;
        lo.op = tc_make_weight(pw.hi, n, k, (int64_t)n * k, 2);
        VocabCand cand;
        float wnorm = 0.f;
        const char* vm = getenv("FUNASR_B200_VOCAB");
        if (precision == kBf16x3 && !(vm && std::string(vm) == "full")) {
            dmeta.alloc(((size_t)m * 3 + 1) * 4); dlist.alloc((size_t)m * kVocabCandCap * sizeof(int2)); dnorm.alloc(4);
            cand.run_max = dmeta.as<int32_t>(); cand.count = cand.run_max + m;
            cand.bound2 = reinterpret_cast<const float*>(cand.count + m);
            cand.overflowed = cand.count + 2 * m;
            cand.list = dlist.as<int2>(); cand.cap = kVocabCandCap;
            launch_row_norm_max(lo.w, n, k, dnorm.as<float>(), stream_);
            FA_CUDA(cudaMemcpyAsync(&wnorm, dnorm.p, 4, cudaMemcpyDeviceToHost, stream_));
            FA_CUDA(cudaStreamSynchronize(stream_));
        }
        const int saved = prec_;
        prec_ = precision;
        try {
            vocab_argmax(da.as<float>(), pa, lo, wnorm, m, cand, dv.as<float>(), di.as<int32_t>(), dids.as<int32_t>());
        } catch (...) { prec_ = saved; throw; }
        prec_ = saved;
        FA_CUDA(cudaStreamSynchronize(stream_));
        if (cand.list && 2 * cdiv(m, 2 * kTcBlockM) >= tc_num_pairs() && getenv("FUNASR_B200_VOCAB_STATS")) {       // tuning aid: how long the lists get
            std::vector<int32_t> meta((size_t)3 * m);
            std::vector<int2> lst((size_t)m * kVocabCandCap);
            FA_CUDA(cudaMemcpy(meta.data(), dmeta.p, meta.size() * 4, cudaMemcpyDeviceToHost));
            FA_CUDA(cudaMemcpy(lst.data(), dlist.p, lst.size() * sizeof(int2), cudaMemcpyDeviceToHost));
            double sum = 0, surv = 0;
            int mx = 0, over = 0;
            for (int r = 0; r < m; ++r) {
                const int c = meta[(size_t)m + r];
                sum += c; mx = std::max(mx, c); over += c > kVocabCandCap;
                const int32_t o = meta[r];
                const int32_t bits = o >= 0 ? o : o ^ 0x7fffffff;
                float rm, b2;
                memcpy(&rm, &bits, 4); memcpy(&b2, &meta[(size_t)2 * m + r], 4);
                for (int i = 0; i < std::min(c, kVocabCandCap); ++i) {
                    float v; memcpy(&v, &lst[(size_t)r * kVocabCandCap + i].y, 4);
                    surv += v >= rm - b2;
                }
            }
            fprintf(stderr, "vocab candidates: rows %d, listed mean %.1f max %d, overflowed rows %d, rescored mean %.2f\n", m,
                    sum / m, mx, over, surv / m);
        }
    }
    FA_CUDA(cudaMemcpy(ids, dids.p, dids.bytes, cudaMemcpyDeviceToHost));
}

void Context::test_attention(const float* qkv, int batch, int frames, int heads, int dk, const int32_t* kv_len,
                             int precision, float* out) {
    set_device();
    const int d = heads * dk;
    const size_t M = (size_t)batch * frames;
    DevBuf dq, dout, dl, dpl;
    dq.alloc(M * 3 * d * 4); dout.alloc(M * d * 4);
    if (kv_len) { dl.alloc((size_t)batch * 4); FA_CUDA(cudaMemcpy(dl.p, kv_len, dl.bytes, cudaMemcpyHostToDevice)); }
    if (precision == kFp32) {
        FA_CUDA(cudaMemcpy(dq.p, qkv, dq.bytes, cudaMemcpyHostToDevice));
        launch_attention_simt(dq.as<float>(), dq.as<float>() + d, dq.as<float>() + 2 * d, 3 * d, batch, frames, heads, dk,
                              dl.as<int>(), dout.as<float>(), Planes{}, d, stream_);
    } else {
        // the producing GEMM would have folded d_k^-0.5 * log2(e) into the q planes; do the same here
        std::vector<float> scaled(qkv, qkv + M * 3 * d);
        const float s = (float)(1.4426950408889634 / std::sqrt((double)dk));
        for (size_t r = 0; r < M; ++r)
            for (int c = 0; c < d; ++c) scaled[r * 3 * d + c] *= s;
        FA_CUDA(cudaMemcpy(dq.p, scaled.data(), dq.bytes, cudaMemcpyHostToDevice));
        dpl.alloc(M * 3 * d * 2 * 2);
        Planes pl{dpl.as<__nv_bfloat16>(), dpl.as<__nv_bfloat16>() + M * 3 * d};
        launch_split_planes(dq.as<float>(), (int64_t)(M * 3 * d), pl, stream_);
        // FUNASR_B200_TEST_ATTN_OUT=planes: write the bf16 hi/lo planes the out-projection reads (what every engine launch
        // does) instead of fp32, and hand back hi + lo
        const char* om = getenv("FUNASR_B200_TEST_ATTN_OUT");
        if (om && !strcmp(om, "planes")) {
            DevBuf dop;
            dop.alloc(M * d * 2 * 2);
            Planes op{dop.as<__nv_bfloat16>(), dop.as<__nv_bfloat16>() + M * d};
            launch_attention_tc(pl, (int64_t)(M * 3 * d), 3 * d, d, batch, frames, heads, dk, dl.as<int>(), nullptr, op, d, stream_);
            FA_CUDA(cudaStreamSynchronize(stream_));
            std::vector<uint16_t> h(M * d * 2);
            FA_CUDA(cudaMemcpy(h.data(), dop.p, dop.bytes, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < M * d; ++i) {
                const uint32_t hb = (uint32_t)h[i] << 16, lb = (uint32_t)h[M * d + i] << 16;
                float hf, lf;
                memcpy(&hf, &hb, 4); memcpy(&lf, &lb, 4);
                out[i] = hf + lf;
            }
            return;
        }
        launch_attention_tc(pl, (int64_t)(M * 3 * d), 3 * d, d, batch, frames, heads, dk, dl.as<int>(), dout.as<float>(),
                            Planes{}, d, stream_);
    }
    FA_CUDA(cudaStreamSynchronize(stream_));
    FA_CUDA(cudaMemcpy(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
}

}  // namespace fa
