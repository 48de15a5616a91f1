// extern "C" boundary (include/funasr_b200.h).  Exceptions stop here: every entry point returns an
// integer code and leaves the message in a thread-local string.
#include "../../include/funasr_b200.h"
#include "engine.h"

#include <cstring>

struct fa_ctx {
    fa::Context impl;
    fa_ctx(int device, int max_batch, int64_t max_samples, int precision)
        : impl(device, max_batch, max_samples, precision) {}
};

namespace {
thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const fa::Error& e) {
        g_error = e.what();
        return 1;
    } catch (const std::exception& e) {
        g_error = std::string("unexpected: ") + e.what();
        return 2;
    } catch (...) {
        g_error = "unknown failure";
        return 3;
    }
}
#define NEED(p) FA_REQUIRE((p) != nullptr, "null pointer: " #p)
}  // namespace

extern "C" {

int fa_abi_version(void) { return FA_ABI_VERSION; }
const char* fa_last_error(void) { return g_error.c_str(); }

int fa_device_count(int* count) {
    return guarded([&] { NEED(count); FA_CUDA(cudaGetDeviceCount(count)); });
}

int64_t fa_frames_for_samples(int64_t samples) { return (samples / fa::kHop + 1 + fa::kLfrN - 1) / fa::kLfrN; }

int64_t fa_adaptor_rows_for_samples(int64_t n_valid) {
    const int64_t t = fa_frames_for_samples(n_valid);
    const int64_t o1 = 1 + (t - 3 + 2) / 2;
    return (1 + (o1 - 3 + 2) / 2 - 1) / 2 + 1;
}

int fa_ctx_create(int device, int max_batch, int64_t max_samples, int precision, fa_ctx** out) {
    return guarded([&] { NEED(out); *out = new fa_ctx(device, max_batch, max_samples, precision); });
}

int fa_ctx_destroy(fa_ctx* ctx) {
    return guarded([&] { delete ctx; });
}

int fa_ctx_load_tensor(fa_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim) {
    return guarded([&] {
        NEED(ctx); NEED(name); NEED(data); NEED(shape);
        FA_REQUIRE(ndim >= 1 && ndim <= 4, "tensor rank must be 1..4");
        ctx->impl.load_tensor(name, data, std::vector<int64_t>(shape, shape + ndim));
    });
}

int fa_ctx_finalize(fa_ctx* ctx) { return guarded([&] { NEED(ctx); ctx->impl.finalize(); }); }

int fa_ctx_set_stream(fa_ctx* ctx, void* cuda_stream) {
    return guarded([&] { NEED(ctx); ctx->impl.set_stream(static_cast<cudaStream_t>(cuda_stream)); });
}

int fa_ctx_sync(fa_ctx* ctx) { return guarded([&] { NEED(ctx); ctx->impl.sync(); }); }

int fa_ctx_vocab(fa_ctx* ctx, int* vocab) { return guarded([&] { NEED(ctx); NEED(vocab); *vocab = ctx->impl.vocab(); }); }

int64_t fa_launch_count(void) { return fa::g_launches; }

int fa_prof_begin(void) { return guarded([&] { fa::prof_begin(); }); }

int fa_prof_end(char* json, int64_t capacity) {
    return guarded([&] {
        NEED(json);
        const std::string s = fa::prof_end();
        FA_REQUIRE((int64_t)s.size() + 1 <= capacity, "profile buffer too small");
        std::memcpy(json, s.c_str(), s.size() + 1);
    });
}

int fa_encode(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* enc,
              float* adaptor) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(enc); NEED(adaptor);
        ctx->impl.encode_host(audio, batch, samples, ilens, enc, adaptor);
    });
}

int fa_encode_dev(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* enc,
                  float* adaptor) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(enc); NEED(adaptor);
        ctx->impl.encode_dev(audio, batch, samples, ilens, enc, adaptor);
    });
}

int fa_ctc(fa_ctx* ctx, const float* enc, int batch, int frames, int32_t* ids) {
    return guarded([&] { NEED(ctx); NEED(enc); NEED(ids); ctx->impl.ctc_host(enc, batch, frames, ids); });
}

int fa_ctc_dev(fa_ctx* ctx, const float* enc, int batch, int frames, int32_t* ids) {
    return guarded([&] { NEED(ctx); NEED(enc); NEED(ids); ctx->impl.ctc_dev(enc, batch, frames, ids); });
}

int fa_front_half(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* enc,
                  float* adaptor, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(ids);
        ctx->impl.front_half_host(audio, batch, samples, ilens, enc, adaptor, ids);
    });
}

int fa_front_half_ragged(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, const int64_t* phys,
                         float* enc, float* adaptor, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(phys); NEED(ids);
        ctx->impl.front_half_host(audio, batch, samples, ilens, enc, adaptor, ids, nullptr, nullptr, phys);
    });
}

int fa_front_half_ragged_dev(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, const int64_t* phys,
                             float* enc, float* adaptor, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(phys); NEED(ids);
        ctx->impl.front_half_dev(audio, batch, samples, ilens, enc, adaptor, ids, phys);
    });
}

int fa_front_half_dev(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* enc,
                      float* adaptor, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(ids);
        ctx->impl.front_half_dev(audio, batch, samples, ilens, enc, adaptor, ids);
    });
}

int fa_front_half_embd(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* enc,
                       float* const* embd_rows, int64_t* rows_out, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens); NEED(embd_rows); NEED(ids);
        ctx->impl.front_half_host(audio, batch, samples, ilens, enc, nullptr, ids, embd_rows, rows_out);
    });
}

int fa_ctc_collapse_dev(fa_ctx* ctx, const int32_t* ids, int batch, int frames, int32_t* tokens, int32_t* starts,
                        int32_t* counts) {
    return guarded([&] {
        NEED(ctx); NEED(ids); NEED(tokens); NEED(starts); NEED(counts);
        ctx->impl.collapse_dev(ids, batch, frames, tokens, starts, counts);
    });
}

int fa_debug_enable_taps(fa_ctx* ctx, int on) { return guarded([&] { NEED(ctx); ctx->impl.enable_taps(on != 0); }); }

int fa_debug_read_tap(fa_ctx* ctx, const char* name, float* out, int64_t capacity, int64_t* rows, int64_t* cols) {
    return guarded([&] {
        NEED(ctx); NEED(name); NEED(rows); NEED(cols);
        std::vector<float> v;
        std::vector<int64_t> shape;
        if (!ctx->impl.read_tap(name, v, shape)) throw fa::Error(std::string("no such tap: ") + name);
        *rows = shape[0]; *cols = shape[1];
        if (out) {
            FA_REQUIRE((int64_t)v.size() <= capacity, "tap buffer too small");
            std::memcpy(out, v.data(), v.size() * sizeof(float));
        }
    });
}

int fa_test_linear(fa_ctx* ctx, const float* a, const float* w, const float* bias, const float* resid, int m, int n,
                   int k, int relu, int precision, float* out, float* out_planes_sum) {
    return guarded([&] {
        NEED(ctx); NEED(a); NEED(w); NEED(bias); NEED(out);
        ctx->impl.test_linear(a, w, bias, resid, m, n, k, relu, precision, out, out_planes_sum);
    });
}

int fa_test_vocab_argmax(fa_ctx* ctx, const float* a, const float* w, const float* bias, int m, int n, int k,
                         int precision, int32_t* ids) {
    return guarded([&] {
        NEED(ctx); NEED(a); NEED(w); NEED(bias); NEED(ids);
        ctx->impl.test_vocab_argmax(a, w, bias, m, n, k, precision, ids);
    });
}

int fa_test_attention(fa_ctx* ctx, const float* qkv, int batch, int frames, int heads, int dk, const int32_t* kv_len,
                      int precision, float* out) {
    return guarded([&] {
        NEED(ctx); NEED(qkv); NEED(out);
        ctx->impl.test_attention(qkv, batch, frames, heads, dk, kv_len, precision, out);
    });
}

int fa_test_layernorm(fa_ctx* ctx, const float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                      float* out, float* out_planes_sum) {
    return guarded([&] {
        NEED(ctx); NEED(x); NEED(gamma); NEED(beta); NEED(out);
        FA_CUDA(cudaSetDevice(ctx->impl.device()));
        fa::DevBuf dx, dg, db, dy, dp;
        const size_t n = (size_t)rows * d;
        dx.alloc(n * 4); dg.alloc((size_t)d * 4); db.alloc((size_t)d * 4); dy.alloc(n * 4); dp.alloc(2 * n * 2);
        FA_CUDA(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
        FA_CUDA(cudaMemcpy(dg.p, gamma, (size_t)d * 4, cudaMemcpyHostToDevice));
        FA_CUDA(cudaMemcpy(db.p, beta, (size_t)d * 4, cudaMemcpyHostToDevice));
        fa::Planes pl{dp.as<__nv_bfloat16>(), dp.as<__nv_bfloat16>() + n};
        fa::launch_layernorm(dx.as<float>(), rows, d, dg.as<float>(), db.as<float>(), eps, nullptr, rows, dy.as<float>(), pl,
                             ctx->impl.stream());
        ctx->impl.sync();
        FA_CUDA(cudaMemcpy(out, dy.p, n * 4, cudaMemcpyDeviceToHost));
        if (out_planes_sum) {
            std::vector<__nv_bfloat16> h(2 * n);
            FA_CUDA(cudaMemcpy(h.data(), dp.p, 2 * n * 2, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) out_planes_sum[i] = __bfloat162float(h[i]) + __bfloat162float(h[n + i]);
        }
    });
}

int fa_test_fsmn(fa_ctx* ctx, const float* v, const float* w, const int32_t* t_valid, int batch, int frames,
                 const float* resid, float* out) {
    return guarded([&] {
        NEED(ctx); NEED(v); NEED(w); NEED(t_valid); NEED(out);
        FA_CUDA(cudaSetDevice(ctx->impl.device()));
        fa::DevBuf dv, dw, dt, dr, dout;
        const size_t n = (size_t)batch * frames * fa::kDenc;
        dv.alloc(n * 4); dw.alloc((size_t)fa::kDenc * fa::kFsmnK * 4); dt.alloc((size_t)batch * 4); dout.alloc(n * 4);
        FA_CUDA(cudaMemcpy(dv.p, v, n * 4, cudaMemcpyHostToDevice));
        FA_CUDA(cudaMemcpy(dw.p, w, dw.bytes, cudaMemcpyHostToDevice));
        FA_CUDA(cudaMemcpy(dt.p, t_valid, dt.bytes, cudaMemcpyHostToDevice));
        if (resid) { dr.alloc(n * 4); FA_CUDA(cudaMemcpy(dr.p, resid, n * 4, cudaMemcpyHostToDevice)); }
        fa::launch_fsmn(dv.as<float>(), fa::kDenc, dw.as<float>(), dt.as<int>(), batch, frames, dr.as<float>(),
                        dout.as<float>(), ctx->impl.stream());
        ctx->impl.sync();
        FA_CUDA(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    });
}

int fa_test_front_end(fa_ctx* ctx, const float* audio, int batch, int64_t samples, const int64_t* ilens, float* logmel,
                      float* x0) {
    return guarded([&] {
        NEED(ctx); NEED(audio); NEED(ilens);
        fa::Context& c = ctx->impl;
        const int64_t frames = fa_frames_for_samples(samples);
        std::vector<float> enc((size_t)batch * frames * fa::kDenc), ad((size_t)batch * frames * fa::kDllm);
        c.enable_taps(true);
        c.encode_host(audio, batch, samples, ilens, enc.data(), ad.data());
        c.enable_taps(false);
        std::vector<float> v;
        std::vector<int64_t> shape;
        if (logmel) {
            FA_REQUIRE(c.read_tap("logmel", v, shape), "logmel tap missing");
            std::memcpy(logmel, v.data(), v.size() * 4);
        }
        if (x0) {
            FA_REQUIRE(c.read_tap("lfr", v, shape), "lfr tap missing");
            std::memcpy(x0, v.data(), v.size() * 4);
        }
    });
}

}  // extern "C"
