// Flash-style attention in fp32 on the CUDA cores (model_definition.py:68-78,
// 80-90 for the SAN-M encoder, :132-145 for the adaptor / CTC blocks).
//
// scores = (q * d_k^-0.5) k^T, additive key mask (m-1)*10000, softmax, times v.  In fp32 the
// additive -10000 makes a masked key's weight exactly zero, so masked keys are simply not visited:
// kv_len[b] keys are attended (the valid frames for encoder/adaptor; every physical frame for the
// unmasked CTC head).  Online softmax over 64-key tiles; scores never leave the SM.
// This is the precision-exact mode and the arbiter for attn_tc.cu.
#include "kernels.h"

namespace fa {

namespace {

constexpr int kQT = 64, kKT = 64, kLd = 68;

template <int DK>
struct AttnSmem {
    float qt[DK][kLd];      // q tile, transposed, pre-scaled
    float kt[DK][kLd];      // k tile, transposed
    float vs[kKT][DK];
    float ps[kQT][kLd];
};

template <int DK>
__global__ void __launch_bounds__(256)
k_attention_simt(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
                 int frames, const int* __restrict__ kv_len, float* __restrict__ ctx,
                 __nv_bfloat16* __restrict__ ctx_hi, __nv_bfloat16* __restrict__ ctx_lo, int ldo, float scale) {
    grid_dependency_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnSmem<DK>& s = *reinterpret_cast<AttnSmem<DK>*>(smem_raw);
    constexpr int DV = DK / 4;            // float4 per head row
    constexpr int NO = DK / 64;           // float4 output groups per thread
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * kQT, h = blockIdx.y, b = blockIdx.z;
    const int klen = kv_len ? kv_len[b] : frames;
    const int64_t base = (int64_t)b * frames;
    const int hoff = h * DK;

    for (int i = tid; i < kQT * DV; i += 256) {
        const int r = i / DV, d4 = i - r * DV;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < frames) x = *reinterpret_cast<const float4*>(q + (base + q0 + r) * ld + hoff + d4 * 4);
        s.qt[d4 * 4 + 0][r] = __fmul_rn(x.x, scale);
        s.qt[d4 * 4 + 1][r] = __fmul_rn(x.y, scale);
        s.qt[d4 * 4 + 2][r] = __fmul_rn(x.z, scale);
        s.qt[d4 * 4 + 3][r] = __fmul_rn(x.w, scale);
    }

    float o[4][NO * 4];
    float mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY;
        lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NO * 4; ++j) o[i][j] = 0.f;
    }

    for (int kv0 = 0; kv0 < klen; kv0 += kKT) {
        __syncthreads();      // previous tile fully consumed (and q tile visible on the first pass)
        for (int i = tid; i < kKT * DV; i += 256) {
            const int c = i / DV, d4 = i - c * DV;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (kv0 + c < klen) {
                kk = *reinterpret_cast<const float4*>(k + (base + kv0 + c) * ld + hoff + d4 * 4);
                vv = *reinterpret_cast<const float4*>(v + (base + kv0 + c) * ld + hoff + d4 * 4);
            }
            s.kt[d4 * 4 + 0][c] = kk.x;
            s.kt[d4 * 4 + 1][c] = kk.y;
            s.kt[d4 * 4 + 2][c] = kk.z;
            s.kt[d4 * 4 + 3][c] = kk.w;
            *reinterpret_cast<float4*>(&s.vs[c][d4 * 4]) = vv;
        }
        __syncthreads();

        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < DK; ++d) {
            const float4 qa = *reinterpret_cast<const float4*>(&s.qt[d][ty * 4]);
            const float4 kb = *reinterpret_cast<const float4*>(&s.kt[d][tx * 4]);
            const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kvv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(qv[i], kvv[j], sc[i][j]);
        }

#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float tmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (kv0 + tx * 4 + j >= klen) sc[i][j] = -INFINITY;
                tmax = fmaxf(tmax, sc[i][j]);
            }
            for (int off = 8; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
            const float mnew = fmaxf(mrow[i], tmax);
            const float alpha = expf(mrow[i] - mnew);
            float psum = 0.f;
            float4 pv;
            float* pp = &pv.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(sc[i][j] - mnew);
                pp[j] = p;
                psum += p;
            }
            for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
            lrow[i] = lrow[i] * alpha + psum;
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < NO * 4; ++j) o[i][j] *= alpha;
            *reinterpret_cast<float4*>(&s.ps[ty * 4 + i][tx * 4]) = pv;
        }
        __syncthreads();

#pragma unroll 2
        for (int c = 0; c < kKT; c += 4) {
            float4 p4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p4[i] = *reinterpret_cast<const float4*>(&s.ps[ty * 4 + i][c]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                for (int g = 0; g < NO; ++g) {
                    const float4 vv = *reinterpret_cast<const float4*>(&s.vs[c + cc][g * 64 + tx * 4]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p = (&p4[i].x)[cc];
                        o[i][g * 4 + 0] = fmaf(p, vv.x, o[i][g * 4 + 0]);
                        o[i][g * 4 + 1] = fmaf(p, vv.y, o[i][g * 4 + 1]);
                        o[i][g * 4 + 2] = fmaf(p, vv.z, o[i][g * 4 + 2]);
                        o[i][g * 4 + 3] = fmaf(p, vv.w, o[i][g * 4 + 3]);
                    }
                }
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= frames) continue;
        const float inv = 1.0f / lrow[i];
#pragma unroll
        for (int g = 0; g < NO; ++g) {
            float4 r4 = make_float4(o[i][g * 4 + 0] * inv, o[i][g * 4 + 1] * inv, o[i][g * 4 + 2] * inv,
                                    o[i][g * 4 + 3] * inv);
            const int64_t off = (base + r) * ldo + hoff + g * 64 + tx * 4;
            if (ctx) *reinterpret_cast<float4*>(ctx + off) = r4;
            if (ctx_hi) {
                __nv_bfloat16 hh[4], ll[4];
                split_bf16(r4.x, hh[0], ll[0]);
                split_bf16(r4.y, hh[1], ll[1]);
                split_bf16(r4.z, hh[2], ll[2]);
                split_bf16(r4.w, hh[3], ll[3]);
                *reinterpret_cast<uint2*>(ctx_hi + off) = *reinterpret_cast<uint2*>(hh);
                if (ctx_lo) *reinterpret_cast<uint2*>(ctx_lo + off) = *reinterpret_cast<uint2*>(ll);
            }
        }
    }
}

}  // namespace

void attention_init_device() {
    FA_CUDA(cudaFuncSetAttribute(k_attention_simt<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(AttnSmem<128>)));
    FA_CUDA(cudaFuncSetAttribute(k_attention_simt<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(AttnSmem<64>)));
}

void launch_attention_simt(const float* q, const float* k, const float* v, int ld, int batch, int frames, int heads,
                           int dk, const int* kv_len, float* ctx_f32, Planes ctx_pl, int ldo, cudaStream_t st) {
    const dim3 grid(cdiv(frames, kQT), heads, batch);
    prof_note_work(4.0 * batch * heads * (double)frames * frames * dk, 0.0);   // upper bound: all keys attended
    const float scale = (float)(1.0 / sqrt((double)dk));    // fp32(d_k ** -0.5), as q_h * (d_k ** -0.5) does
    if (dk == 128) {
        FA_LAUNCH(k_attention_simt<128>, grid, 256, sizeof(AttnSmem<128>), st, q, k, v, ld, frames, kv_len, ctx_f32,
                  ctx_pl.hi, ctx_pl.lo, ldo, scale);
    } else if (dk == 64) {
        FA_LAUNCH(k_attention_simt<64>, grid, 256, sizeof(AttnSmem<64>), st, q, k, v, ld, frames, kv_len, ctx_f32,
                  ctx_pl.hi, ctx_pl.lo, ldo, scale);
    } else {
        throw Error("attention head width must be 64 or 128");
    }
}

}  // namespace fa
