// Bandwidth-bound row kernels: LayerNorm, FSMN memory block, length control, argmax, greedy collapse
// (SURVEY §8a rows a5, a7, a12, a14 tail, a15).  All are one pass over their input with 128-bit
// accesses; each row/time-strip is owned by one warp or thread so no atomics are needed.
#include "kernels.h"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace fa {

namespace {

// ------------------------------------------------------------------------------------ LayerNorm
// model_definition.py:42-44 (eps 1e-5, encoder) and :151 (nn.LayerNorm eps 1e-12, adaptor / CTC blocks).
// One warp per row, the row held in registers (d <= 1024).  Output as fp32 and/or bf16 hi/lo planes;
// rows at or past t_valid[b] are written as zeros when a length vector is given (the "sweeping"
// multiplies at model_definition.py:210,213).
constexpr int kLnMaxVec = 8;     // float4 per lane -> d <= 1024

template <int NV, int kLnRows = 2>   // NV: float4 per lane actually needed (d <= 128 * NV); kLnRows: rows per warp
__global__ void __launch_bounds__(256)
k_layernorm(const float* x, int rows, int d, const float* __restrict__ gamma,
            const float* __restrict__ beta, float eps, const int* __restrict__ t_valid, int frames,
            float* y /* may alias x: a warp reads its rows before it writes them */,
            __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo,
            const int* __restrict__ seg_off /* unpack: x packed, outputs physical */, uint8_t* __restrict__ y_f8) {
    grid_dependency_wait();
    const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kLnRows, lane = threadIdx.x & 31;
    if (row0 >= rows) return;
    const int nvec = d >> 2;
    float4 v[kLnRows][NV];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
        bool have = row0 + r < rows;
        int64_t src = row0 + r;
        if (seg_off && have) {
            const int b = (row0 + r) / frames, t = (row0 + r) - b * frames;
            have = t < t_valid[b];
            src = (int64_t)seg_off[b] + t;
        }
        const float4* xr = reinterpret_cast<const float4*>(x + src * d);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            v[r][i] = (have && idx < nvec) ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
        const int row = row0 + r;
        if (row >= rows) break;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + 32 * i < nvec) {
                const float a = v[r][i].x - mean, b = v[r][i].y - mean, c = v[r][i].z - mean, e = v[r][i].w - mean;
                q += (a * a + b * b) + (c * c + e * e);
            }
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.0f / sqrtf(q / (float)d + eps);
        bool live = true;
        if (t_valid) {
            const int b = row / frames, t = row - b * frames;
            live = t < t_valid[b];
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            if (idx >= nvec) continue;
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float4 g = g4[idx], bb = b4[idx];
                o4.x = (v[r][i].x - mean) * rstd * g.x + bb.x;
                o4.y = (v[r][i].y - mean) * rstd * g.y + bb.y;
                o4.z = (v[r][i].z - mean) * rstd * g.z + bb.z;
                o4.w = (v[r][i].w - mean) * rstd * g.w + bb.w;
            }
            const int64_t o = (int64_t)row * d + idx * 4;
            if (y) *reinterpret_cast<float4*>(y + o) = o4;
            if (y_hi) {
                uint2 h, l;
                split_bf16x2(o4.x, o4.y, h.x, l.x);
                split_bf16x2(o4.z, o4.w, h.y, l.y);
                *reinterpret_cast<uint2*>(y_hi + o) = h;
                if (y_lo) *reinterpret_cast<uint2*>(y_lo + o) = l;
            }
            if (y_f8) *reinterpret_cast<uint32_t*>(y_f8 + o) = f32x4_to_e4m3(o4.x, o4.y, o4.z, o4.w);
        }
    }
}

// ------------------------------------------------------------------------------------ FSMN
// MultiHeadedAttentionSANM.forward_fsmn (model_definition.py:60-66): v*m, zero-pad 5|5, depthwise
// 11-tap correlation over time (weight (512,1,11), no bias), plus the masked input.  The residual
// add of EncoderLayerSANM.forward (:110) is folded in (resid may alias out; layer 0 passes null, F9).
// Thread = 4 adjacent channels (128-bit accesses), strip of kFsmnT consecutive frames; the strip's
// kFsmnT+10 input rows and kFsmnT residual rows are all requested before any arithmetic, so each
// thread keeps ~26 independent 16-byte loads in flight (the kernel is pure HBM streaming).
constexpr int kFsmnT = 8;

__global__ void __launch_bounds__(kDenc / 4)
k_fsmn(const float* __restrict__ v, int ldv, const float* __restrict__ w, const int* __restrict__ t_valid,
       int frames, const float* resid, float* out, const int* __restrict__ seg_off) {
    grid_dependency_wait();
    const int c = threadIdx.x * 4, b = blockIdx.y, t0 = blockIdx.x * kFsmnT;
    const int tv = t_valid[b];
    // packed layout: the segment is rows seg_off[b] .. + tv - 1 and has no padded frames
    const int64_t base = seg_off ? seg_off[b] : (int64_t)b * frames;
    if (seg_off) { frames = tv; if (t0 >= tv) return; }
    const float* vb = v + base * ldv + c;
    float4 win[kFsmnT + kFsmnK - 1];
#pragma unroll
    for (int i = 0; i < kFsmnT + kFsmnK - 1; ++i) {
        const int t = t0 + i - 5;
        win[i] = (t >= 0 && t < tv) ? *reinterpret_cast<const float4*>(vb + (int64_t)t * ldv) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 rs[kFsmnT];
    if (resid) {
#pragma unroll
        for (int i = 0; i < kFsmnT; ++i) {
            const int t = t0 + i;
            rs[i] = t < frames ? *reinterpret_cast<const float4*>(resid + (base + t) * kDenc + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float wk[4][kFsmnK];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kFsmnK; ++j) wk[i][j] = w[(c + i) * kFsmnK + j];
#pragma unroll
    for (int i = 0; i < kFsmnT; ++i) {
        const int t = t0 + i;
        if (t >= frames) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kFsmnK; ++j) {
            acc.x = fmaf(wk[0][j], win[i + j].x, acc.x);
            acc.y = fmaf(wk[1][j], win[i + j].y, acc.y);
            acc.z = fmaf(wk[2][j], win[i + j].z, acc.z);
            acc.w = fmaf(wk[3][j], win[i + j].w, acc.w);
        }
        float4 r = make_float4(__fadd_rn(acc.x, win[i + 5].x), __fadd_rn(acc.y, win[i + 5].y),
                               __fadd_rn(acc.z, win[i + 5].z), __fadd_rn(acc.w, win[i + 5].w));
        if (resid) {
            r.x = __fadd_rn(rs[i].x, r.x); r.y = __fadd_rn(rs[i].y, r.y);
            r.z = __fadd_rn(rs[i].z, r.z); r.w = __fadd_rn(rs[i].w, r.w);
        }
        *reinterpret_cast<float4*>(out + (base + t) * kDenc + c) = r;
    }
}

// The same block as a persistent, double-buffered streaming kernel (FUNASR_B200_FSMN=stream; measured equal to the strip
// kernel on B200, 4.07 vs 4.08 ms per step, so the block is bound by the memory system and not by latency or halo re-reads).  One CTA per SM walks strips of
// kFsT frames of one segment; one warp asks the copy engine for the strip's kFsT + 10 input rows and kFsT residual rows
// (2 KB bulk copies, mbarrier transaction count) into the other buffer while the CTA computes the current strip out of
// shared memory, so ~84 KB per SM are in flight at all times without holding a register per byte (the strip kernel
// above keeps 26 loads per thread in flight at 150+ registers and 3 CTAs per SM, and re-reads 10 halo rows per 8).
// Rows outside [0, t_valid) are never copied: the reader substitutes zeros.  Arithmetic and its order are those of
// k_fsmn, so the two agree bit for bit.
constexpr int kFsT = 16, kFsRows = kFsT + kFsmnK - 1, kFsThreads = 256;
constexpr int kFsBufBytes = (kFsRows + kFsT) * kDenc * 4;                 // 86 016
constexpr int kFsSmem = 2 * kFsBufBytes + 64;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kFsThreads, 1)
k_fsmn_stream(const float* __restrict__ v, int ldv, const float* __restrict__ w, const int* __restrict__ t_valid, int batch,
              int frames, const float* resid, float* out) {
    extern __shared__ __align__(128) unsigned char fs_smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(fs_smem);
    const uint32_t bar0 = base + 2 * kFsBufBytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = (tid & 127) * 4, fh = tid >> 7;                          // 4 channels; frames fh*8 .. fh*8+7 of the strip
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    grid_dependency_wait();
    float wk[4][kFsmnK];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kFsmnK; ++j) wk[i][j] = w[(c + i) * kFsmnK + j];
    const int n_strips = (frames + kFsT - 1) / kFsT, total = batch * n_strips;

    // warp 0: request strip s into buffer `buf` (lane i: window row i, then residual row i)
    auto request = [&](int s, int buf) {
        const int b = s / n_strips, t0 = (s - b * n_strips) * kFsT, tv = t_valid[b];
        const uint32_t dst = base + buf * kFsBufBytes, bar = bar0 + 8 * buf;
        const int lo = t0 - 5 < 0 ? 0 : t0 - 5, hi = t0 + kFsT + 5 < tv ? t0 + kFsT + 5 : tv;     // copied input rows [lo, hi)
        const int n_in = hi > lo ? hi - lo : 0;
        const int n_res = resid ? (t0 + kFsT < frames ? kFsT : frames - t0) : 0;
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(n_in + n_res) * kDenc * 4);
        __syncwarp();
        const int t = t0 - 5 + lane;
        if (lane < kFsRows && t >= lo && t < hi)
            bulk_load_1d(dst + lane * kDenc * 4, v + ((int64_t)b * frames + t) * ldv, kDenc * 4, bar);
        if (lane < n_res)
            bulk_load_1d(dst + (kFsRows + lane) * kDenc * 4, resid + ((int64_t)b * frames + t0 + lane) * kDenc, kDenc * 4, bar);
    };

    if (warp == 0 && (int)blockIdx.x < total) request(blockIdx.x, 0);
    int it = 0;
    for (int s = blockIdx.x; s < total; s += gridDim.x, ++it) {
        const int buf = it & 1;
        if (warp == 0 && s + (int)gridDim.x < total) request(s + gridDim.x, buf ^ 1);   // that buffer was read in the previous iteration
        mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);                         // bounded: traps instead of hanging
        const int b = s / n_strips, t0 = (s - b * n_strips) * kFsT, tv = t_valid[b];
        const float* wbuf = reinterpret_cast<const float*>(fs_smem + buf * kFsBufBytes);
        const float* rbuf = wbuf + kFsRows * kDenc;
        float4 win[8 + kFsmnK - 1];
#pragma unroll
        for (int i = 0; i < 8 + kFsmnK - 1; ++i) {
            const int r = fh * 8 + i, t = t0 + r - 5;
            win[i] = (t >= 0 && t < tv) ? *reinterpret_cast<const float4*>(wbuf + r * kDenc + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + fh * 8 + i;
            if (t < frames) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kFsmnK; ++j) {
                    acc.x = fmaf(wk[0][j], win[i + j].x, acc.x);
                    acc.y = fmaf(wk[1][j], win[i + j].y, acc.y);
                    acc.z = fmaf(wk[2][j], win[i + j].z, acc.z);
                    acc.w = fmaf(wk[3][j], win[i + j].w, acc.w);
                }
                float4 r = make_float4(__fadd_rn(acc.x, win[i + 5].x), __fadd_rn(acc.y, win[i + 5].y),
                                       __fadd_rn(acc.z, win[i + 5].z), __fadd_rn(acc.w, win[i + 5].w));
                if (resid) {
                    const float4 rs = *reinterpret_cast<const float4*>(rbuf + (fh * 8 + i) * kDenc + c);
                    r.x = __fadd_rn(rs.x, r.x); r.y = __fadd_rn(rs.y, r.y);
                    r.z = __fadd_rn(rs.z, r.z); r.w = __fadd_rn(rs.w, r.w);
                }
                *reinterpret_cast<float4*>(out + ((int64_t)b * frames + t) * kDenc + c) = r;
            }
        }
        __syncthreads();                                                  // everyone is done reading this buffer
    }
}

__global__ void __launch_bounds__(256)
k_row_keep(const float4* __restrict__ in, float4* __restrict__ out, int frames, int d4, const int* __restrict__ keep,
           int64_t total4, const int* __restrict__ seg_off) {
    grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int64_t row = i / d4;
    const int b = (int)(row / frames), t = (int)(row - (int64_t)b * frames);
    const int64_t src = seg_off ? ((int64_t)seg_off[b] + t) * d4 + (i - row * d4) : i;
    out[i] = t < keep[b] ? in[src] : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(128)
k_repack_rows(const uint4* __restrict__ src, uint4* __restrict__ dst, int row16, const int* __restrict__ src_off,
              const int* __restrict__ dst_off, const int* __restrict__ tv) {
    grid_dependency_wait();
    const int t = blockIdx.x, b = blockIdx.y;
    const int n = tv[b], rows = dst_off[b + 1] - dst_off[b];      // rows = n, or n + 1 with the pad row
    if (t >= rows) return;
    uint4* d = dst + ((int64_t)dst_off[b] + t) * row16;
    const uint4* s = src + ((int64_t)src_off[b] + t) * row16;
    for (int i = threadIdx.x; i < row16; i += blockDim.x) d[i] = t < n ? s[i] : make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256)
k_unpack_ids(const int32_t* __restrict__ packed, int32_t* __restrict__ ids, int frames, const int* __restrict__ off,
             const int* __restrict__ tv, const int* __restrict__ tphys) {
    grid_dependency_wait();
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= frames) return;
    const int n = tv[b];
    // frames past a segment's own physical length (ragged batches) do not exist in the reference's run of it: -1
    ids[(int64_t)b * frames + t] = (tphys && t >= tphys[b]) ? -1 : packed[off[b] + (t < n ? t : n)];
}

__global__ void __launch_bounds__(256)
k_split_planes(const float4* __restrict__ x, int64_t n4, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    uint2 h, l;
    split_bf16x2(v.x, v.y, h.x, l.x);
    split_bf16x2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + i * 4) = h;
    if (lo) *reinterpret_cast<uint2*>(lo + i * 4) = l;
}

__global__ void __launch_bounds__(256)
k_to_e4m3(const float4* __restrict__ x, int64_t n4, uint32_t* __restrict__ out) {
    grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    out[i] = f32x4_to_e4m3(v.x, v.y, v.z, v.w);
}

// one warp per weight row: scale = max|w| / 448 (1 for an all-zero row), w8 = e4m3(w / scale)
__global__ void __launch_bounds__(256)
k_quant_rows_e4m3(const float* __restrict__ w, int rows, int k, uint8_t* __restrict__ w8, float* __restrict__ scale) {
    grid_dependency_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* wr = w + (int64_t)row * k;
    float mx = 0.f;
    for (int i = lane * 4; i < k; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(wr + i);
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float sc = mx > 0.f ? mx / 448.0f : 1.0f;
    for (int i = lane * 4; i < k; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(wr + i);
        *reinterpret_cast<uint32_t*>(w8 + (int64_t)row * k + i) = f32x4_to_e4m3(v.x / sc, v.y / sc, v.z / sc, v.w / sc);
    }
    if (lane == 0) scale[row] = sc;
}

// ------------------------------------------------------------------------------------ argmax
// torch.argmax(..., dim=-1) returns the first index on ties (model_definition.py:337).
__device__ __forceinline__ void amax_merge(float& bv, int& bi, float v, int i) {
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}

__global__ void __launch_bounds__(256)
k_argmax_rows(const float* __restrict__ logits, int n, int ld, int32_t* __restrict__ ids) {
    grid_dependency_wait();
    const int row = blockIdx.x;
    const float* p = logits + (int64_t)row * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) amax_merge(bv, bi, p[i], i);
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        amax_merge(bv, bi, ov, oi);
    }
    __shared__ float sv[8];
    __shared__ int si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) amax_merge(bv, bi, sv[i], si[i]);
        ids[row] = bi;
    }
}

__global__ void __launch_bounds__(256)
k_argmax_combine(const float* __restrict__ pmax, const int32_t* __restrict__ pidx, int rows, int tiles,
                 int32_t* __restrict__ ids, const int32_t* __restrict__ only_if_over, int over) {
    grid_dependency_wait();
    // one warp per row: lanes stride over the tiles, then merge (value desc, index asc)
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (only_if_over && only_if_over[row] <= over) return;      // second-chance pass of the candidate path: overflowed rows only
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lane; t < tiles; t += 32) amax_merge(bv, bi, pmax[(int64_t)row * tiles + t], pidx[(int64_t)row * tiles + t]);
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        amax_merge(bv, bi, ov, oi);
    }
    if (lane == 0) ids[row] = bi;
}

// ------------------------------------------------------------------------------------ vocabulary candidates
// order-preserving int encoding of a float (for atomicMax on signed ints)
__device__ __forceinline__ int32_t f2ord(float f) {
    const int32_t b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int32_t o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
k_row_norm_max(const float* __restrict__ w, int rows, int d, float* out) {
    grid_dependency_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) { const float v = w[(int64_t)row * d + i]; s = fmaf(v, v, s); }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(s) * 1.000001f));   // non-negative floats order like ints
}

__global__ void __launch_bounds__(256)
k_vocab_prepare(const float* __restrict__ x, int rows, int d, float w_norm_max, __nv_bfloat16* __restrict__ x_hi,
                __nv_bfloat16* __restrict__ x_lo, int32_t* __restrict__ run_max, int32_t* __restrict__ count,
                float* __restrict__ bound2, int32_t* __restrict__ overflowed) {
    grid_dependency_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int i = lane * 4; i < d; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)row * d + i);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        uint2 h, l;
        split_bf16x2(v.x, v.y, h.x, l.x);
        split_bf16x2(v.z, v.w, h.y, l.y);
        *reinterpret_cast<uint2*>(x_hi + (int64_t)row * d + i) = h;
        *reinterpret_cast<uint2*>(x_lo + (int64_t)row * d + i) = l;      // only the second-chance pass reads it
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        // |L - L~| <= 2^-8 (1 + 2^-10) sum |x_i||w_i| + accumulation rounding (K * 2^-22 of the same sum at worst)
        // <= 2^-8 * 1.05 * |x| |w|; the list test uses twice that
        bound2[row] = 2.0f * 0.00390625f * 1.05f * sqrtf(s) * w_norm_max;
        run_max[row] = f2ord(-INFINITY);
        count[row] = 0;
        if (row == 0) *overflowed = 0;
    }
}

// fp32 logit of one column, by one warp: fixed lane / shuffle order, so every path that rescores a column gets the same bits
__device__ __forceinline__ float vocab_score(const float* __restrict__ xr, const float* __restrict__ w,
                                             const float* __restrict__ bias, int col, int d, int lane) {
    const float* wr = w + (int64_t)col * d;
    float s = 0.f;
    for (int i = lane * 4; i < d; i += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xr + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(wr + i));
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return __fadd_rn(s, bias[col]);
}

// One warp per row.  The list holds every column that was within bound2 of the running maximum when its tile was
// finished; only those within bound2 of the FINAL approximate maximum can hold the true maximum, and only they are
// rescored: an fp32 dot product in a fixed lane / shuffle order, so the result does not depend on the
// (non-deterministic) order of the list; ties go to the lowest index like torch.argmax.  A row whose list
// overflowed is left to the second-chance pass (it is counted in *overflowed, which gates that pass).
__global__ void __launch_bounds__(256)
k_vocab_rescore(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int rows, int d,
                int n, const int32_t* __restrict__ run_max, const int32_t* __restrict__ count, const float* __restrict__ bound2,
                const int2* __restrict__ list, int cap, int32_t* __restrict__ ids, int32_t* __restrict__ overflowed) {
    grid_dependency_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (int64_t)row * d;
    const int cnt = count[row];
    if (cnt > cap) {
        if (lane == 0) atomicAdd(overflowed, 1);
        return;
    }
    const float thr = ord2f(run_max[row]) - bound2[row];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        int2 e = make_int2(0, 0);
        bool keep = false;
        if (c0 + lane < cnt) {
            e = list[(int64_t)row * cap + c0 + lane];
            keep = __int_as_float(e.y) >= thr;
        }
        unsigned live = __ballot_sync(0xffffffffu, keep);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const int col = __shfl_sync(0xffffffffu, e.x, src);
            amax_merge(bv, bi, vocab_score(xr, w, bias, col, d, lane), col);
        }
    }
    if (lane == 0) ids[row] = bi;
}

// The three-product projection's per-128-column partial maxima (small batches, and the second-chance pass of the
// candidate path) decided the same way: every 128-column slot whose partial maximum is within the three-product
// error bound of the row's best may hold the true fp32 maximum, so ALL columns of those slots (one or two, as a rule)
// are rescored with vocab_score and the first maximal index wins.  Both vocabulary paths therefore return the argmax
// of the same fp32 scores: the ids of a segment do not depend on how many segments share the call.
// |L - L3| <= 3 * 2^-17 (1 + eps) sum |x_i||w_i| + accumulation rounding  <=  2^-15 |x| |w|; the test uses twice that.
__global__ void __launch_bounds__(256)
k_vocab_rescore_slots(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int rows, int d,
                      int n, float w_norm_max, const float* __restrict__ pmax, int slots, int32_t* __restrict__ ids,
                      const int32_t* __restrict__ only_if_over, int over) {
    grid_dependency_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (only_if_over && only_if_over[row] <= over) return;
    const float* xr = x + (int64_t)row * d;
    float s2 = 0.f, mx = -INFINITY;
    for (int i = lane * 4; i < d; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        s2 = fmaf(v.x, v.x, s2); s2 = fmaf(v.y, v.y, s2); s2 = fmaf(v.z, v.z, s2); s2 = fmaf(v.w, v.w, s2);
    }
    for (int t = lane; t < slots; t += 32) mx = fmaxf(mx, pmax[(int64_t)row * slots + t]);
    for (int o = 16; o > 0; o >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const float thr = mx - 2.0f * 0x1p-15f * sqrtf(s2) * w_norm_max;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int t0 = 0; t0 < slots; t0 += 32) {
        const int t = t0 + lane;
        unsigned live = __ballot_sync(0xffffffffu, t < slots && pmax[(int64_t)row * slots + t] >= thr);
        while (live) {
            const int slot = t0 + __ffs(live) - 1;
            live &= live - 1;
            const int c1 = min(n, (slot + 1) * 128);
            for (int col = slot * 128; col < c1; ++col) amax_merge(bv, bi, vocab_score(xr, w, bias, col, d, lane), col);
        }
    }
    if (lane == 0) ids[row] = bi;
}

// ------------------------------------------------------------------------------------ greedy collapse
// decode_ctc (nano_ctc.py:70-99): merge runs, drop blanks, keep the first frame of each run.
// One CTA per segment; block-wide exclusive scan of the "run starts here and is not blank" flags.
__global__ void __launch_bounds__(1024)
k_ctc_collapse(const int32_t* __restrict__ ids, int frames, int blank, int32_t* __restrict__ tokens,
               int32_t* __restrict__ starts, int32_t* __restrict__ counts) {
    grid_dependency_wait();
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* p = ids + (int64_t)b * frames;
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < frames; t0 += 1024) {
        const int t = t0 + tid;
        int flag = 0, tok = 0;
        if (t < frames) {
            tok = p[t];
            flag = (t == 0 || p[t - 1] != tok) && tok != blank;
        }
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        const int in_warp = __popc(m & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        if (flag) {
            tokens[(int64_t)b * frames + off + in_warp] = tok;
            starts[(int64_t)b * frames + off + in_warp] = t;
        }
        __syncthreads();
        if (tid == 0) {
            int s = base_s;
            for (int w = 0; w < 32; ++w) s += warp_tot[w];
            base_s = s;
        }
        __syncthreads();
    }
    if (tid == 0) counts[b] = base_s;
}

}  // namespace

void launch_layernorm(const float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                      const int* t_valid, int frames, float* y_f32, Planes y_pl, cudaStream_t st, const int* seg_off,
                      uint8_t* y_f8) {
    FA_REQUIRE(d % 4 == 0 && d <= 128 * kLnMaxVec, "layernorm width must be a multiple of 4 and <= 1024");
    FA_REQUIRE(!seg_off || (t_valid && y_f32 != x), "layernorm unpack needs the length vector and an output that is not its input");
    // algorithmic bytes: the row in, and whichever outputs are written (fp32 row, bf16 hi plane, bf16 lo plane)
    prof_note_work(0.0, (double)rows * d * (4.0 + (y_f32 ? 4.0 : 0.0) + (y_pl.hi ? 2.0 : 0.0) + (y_pl.lo ? 2.0 : 0.0) + (y_f8 ? 1.0 : 0.0)));
    // d <= 512 (141 of the 155 launches of a step): one row per warp — 40 registers, 48 resident warps per SM; two rows
    // per warp (62 registers) measured 6 % slower on the same box, four rows 18 % slower
    if (d <= 512) {
        FA_LAUNCH((k_layernorm<4, 1>), cdiv(rows, 8), 256, 0, st, x, rows, d, gamma, beta, eps, t_valid, frames, y_f32, y_pl.hi, y_pl.lo, seg_off, y_f8);
        return;
    }
    const int grid = cdiv(rows, 8 * 2);
    if (d <= 640) {
        FA_LAUNCH(k_layernorm<5>, grid, 256, 0, st, x, rows, d, gamma, beta, eps, t_valid, frames, y_f32, y_pl.hi, y_pl.lo, seg_off, y_f8);
    } else {
        FA_LAUNCH(k_layernorm<8>, grid, 256, 0, st, x, rows, d, gamma, beta, eps, t_valid, frames, y_f32, y_pl.hi, y_pl.lo, seg_off, y_f8);
    }
}

void launch_fsmn(const float* v, int ldv, const float* w, const int* t_valid, int batch, int frames,
                 const float* resid, float* out, cudaStream_t st, const Packing* pk) {
    const int* seg_off = pk ? pk->seg_off : nullptr;
    if (pk) frames = pk->max_len;
    FA_REQUIRE(ldv % 4 == 0, "fsmn input stride must be a multiple of 4");
    FA_REQUIRE(resid == nullptr || resid != v, "fsmn: the residual may alias the output, not the input");
    prof_note_work(0.0, (pk ? (double)pk->total_rows : (double)batch * frames) * kDenc * 4.0 * (resid ? 3.0 : 2.0));    // v in, residual in, x out
    const char* fe = getenv("FUNASR_B200_FSMN");                          // comparison aid, read at every launch
    const bool strips = seg_off || !(fe && !strcmp(fe, "stream"));                     // the streaming kernel measured equal (4.07 vs 4.08 ms per step): strips stay the default
    static int sms = 0;
    if (!strips && sms == 0) {
        int dev = 0;
        FA_CUDA(cudaGetDevice(&dev));
        FA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FA_CUDA(cudaFuncSetAttribute(k_fsmn_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, kFsSmem));
    }
    if (strips) {
        FA_LAUNCH(k_fsmn, dim3(cdiv(frames, kFsmnT), batch), kDenc / 4, 0, st, v, ldv, w, t_valid, frames, resid, out, seg_off);
        return;
    }
    const int total = batch * cdiv(frames, kFsT);
    FA_LAUNCH(k_fsmn_stream, total < sms ? total : sms, kFsThreads, kFsSmem, st, v, ldv, w, t_valid, batch, frames, resid, out);
}

void launch_row_keep(const float* in, float* out, int batch, int frames, int d, const int* keep, cudaStream_t st,
                     const int* seg_off) {
    const int64_t total4 = (int64_t)batch * frames * d / 4;
    FA_LAUNCH(k_row_keep, cdiv(total4, 256), 256, 0, st, reinterpret_cast<const float4*>(in),
              reinterpret_cast<float4*>(out), frames, d / 4, keep, total4, seg_off);
}

void launch_repack_rows(const void* src, void* dst, int row_bytes, const int* src_off, const int* dst_off, const int* tv,
                        int batch, int max_len, cudaStream_t st) {
    FA_REQUIRE(row_bytes % 16 == 0, "repack_rows needs rows of whole 16-byte words");
    FA_LAUNCH(k_repack_rows, dim3(max_len + 1, batch), 128, 0, st, reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
              row_bytes / 16, src_off, dst_off, tv);
}

void launch_unpack_ids(const int32_t* packed, int32_t* ids, int batch, int frames, const int* off, const int* tv, cudaStream_t st,
                       const int* tphys) {
    FA_LAUNCH(k_unpack_ids, dim3(cdiv(frames, 256), batch), 256, 0, st, packed, ids, frames, off, tv, tphys);
}

void launch_split_planes(const float* x, int64_t n, Planes out, cudaStream_t st) {
    FA_REQUIRE(n % 4 == 0, "split_planes length must be a multiple of 4");
    FA_LAUNCH(k_split_planes, cdiv(n / 4, 256), 256, 0, st, reinterpret_cast<const float4*>(x), n / 4, out.hi, out.lo);
}

void launch_to_e4m3(const float* x, int64_t n, uint8_t* out, cudaStream_t st) {
    FA_REQUIRE(n % 4 == 0, "to_e4m3 length must be a multiple of 4");
    FA_LAUNCH(k_to_e4m3, cdiv(n / 4, 256), 256, 0, st, reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<uint32_t*>(out));
}

void launch_quant_rows_e4m3(const float* w, int rows, int k, uint8_t* w8, float* scale, cudaStream_t st) {
    FA_REQUIRE(k % 4 == 0, "quant_rows needs K % 4 == 0");
    FA_LAUNCH(k_quant_rows_e4m3, cdiv(rows, 8), 256, 0, st, w, rows, k, w8, scale);
}

void launch_argmax_rows(const float* logits, int rows, int n, int ld, int32_t* ids, cudaStream_t st) {
    FA_LAUNCH(k_argmax_rows, rows, 256, 0, st, logits, n, ld, ids);
}

void launch_argmax_combine(const float* pmax, const int32_t* pidx, int rows, int tiles, int32_t* ids, cudaStream_t st,
                           const int32_t* only_if_over, int over) {
    FA_LAUNCH(k_argmax_combine, cdiv(rows, 8), 256, 0, st, pmax, pidx, rows, tiles, ids, only_if_over, over);
}

void launch_row_norm_max(const float* w, int rows, int d, float* out, cudaStream_t st) {
    FA_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    FA_LAUNCH(k_row_norm_max, cdiv(rows, 8), 256, 0, st, w, rows, d, out);
}

void launch_vocab_prepare(const float* x, int rows, int d, float w_norm_max, Planes x_pl, VocabCand c, cudaStream_t st) {
    FA_REQUIRE(d % 4 == 0, "vocab_prepare needs a width that is a multiple of 4");
    FA_LAUNCH(k_vocab_prepare, cdiv(rows, 8), 256, 0, st, x, rows, d, w_norm_max, x_pl.hi, x_pl.lo, c.run_max, c.count,
              const_cast<float*>(c.bound2), c.overflowed);
}

void launch_vocab_rescore(const float* x, const float* w, const float* bias, int rows, int d, int n, VocabCand c, int32_t* ids,
                          cudaStream_t st) {
    FA_REQUIRE(d % 4 == 0, "vocab_rescore needs a width that is a multiple of 4");
    FA_LAUNCH(k_vocab_rescore, cdiv(rows, 8), 256, 0, st, x, w, bias, rows, d, n, c.run_max, c.count, c.bound2, c.list, c.cap, ids,
              c.overflowed);
}

void launch_vocab_rescore_slots(const float* x, const float* w, const float* bias, int rows, int d, int n, float w_norm_max,
                                const float* pmax, int slots, int32_t* ids, cudaStream_t st, const int32_t* only_if_over,
                                int over) {
    FA_REQUIRE(d % 4 == 0, "vocab_rescore_slots needs a width that is a multiple of 4");
    FA_LAUNCH(k_vocab_rescore_slots, cdiv(rows, 8), 256, 0, st, x, w, bias, rows, d, n, w_norm_max, pmax, slots, ids, only_if_over,
              over);
}

void launch_ctc_collapse(const int32_t* ids, int batch, int frames, int blank, int32_t* tokens, int32_t* starts,
                         int32_t* counts, cudaStream_t st) {
    FA_LAUNCH(k_ctc_collapse, batch, 1024, 0, st, ids, frames, blank, tokens, starts, counts);
}

}  // namespace fa
