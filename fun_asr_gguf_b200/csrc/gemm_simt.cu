// fp32 GEMM on the CUDA cores:  C[M][N] = A[M][K] * W[N][K]^T (+bias, ReLU, residual).
//
// This is the precision-exact execution mode (FA_PREC_FP32) of every nn.Linear on the path
// (model_definition.py:30-40, 80-90, 132-145, 179-185, 216-221) and the on-device arbiter the
// tcgen05 kernels are unit-tested against.  It is not the throughput path: gemm_tc.cu is.
// 128x128x16 tiles, 256 threads, 8x8 outputs per thread, operands transposed into shared memory
// so the inner loop is four 128-bit LDS per 64 FMA.
#include "kernels.h"

namespace fa {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;

__global__ void __launch_bounds__(256)
k_gemm_simt(const float* __restrict__ a, int lda, const float* __restrict__ w, int m, int n, int k,
            const float* __restrict__ bias, const float* resid, int ldr, int relu, float* out, int ldc,
            __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int ldp) {
    grid_dependency_wait();
    __shared__ __align__(16) float As[BK][LDS_];
    __shared__ __align__(16) float Bs[BK][LDS_];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    for (int k0 = 0; k0 < k; k0 += BK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (m0 + r < m) va = *reinterpret_cast<const float4*>(a + (int64_t)(m0 + r) * lda + k0 + lk);
            if (n0 + r < n) vb = *reinterpret_cast<const float4*>(w + (int64_t)(n0 + r) * k + k0 + lk);
            As[lk + 0][r] = va.x; As[lk + 1][r] = va.y; As[lk + 2][r] = va.z; As[lk + 3][r] = va.w;
            Bs[lk + 0][r] = vb.x; Bs[lk + 1][r] = vb.y; Bs[lk + 2][r] = vb.z; Bs[lk + 3][r] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (col >= n) continue;
            float v = acc[i][j];
            if (bias) v = __fadd_rn(v, bias[col]);
            if (relu) v = fmaxf(v, 0.f);
            if (resid) v = __fadd_rn(resid[(int64_t)row * ldr + col], v);
            if (out) out[(int64_t)row * ldc + col] = v;
            if (out_hi) {
                __nv_bfloat16 h, l;
                split_bf16(v, h, l);
                out_hi[(int64_t)row * ldp + col] = h;
                if (out_lo) out_lo[(int64_t)row * ldp + col] = l;
            }
        }
    }
}

}  // namespace

void launch_gemm_simt(const float* a, int lda, const float* w, int m, int n, int k, const Epilogue& ep,
                      cudaStream_t st) {
    FA_REQUIRE(k % BK == 0 && lda % 4 == 0, "simt gemm needs K % 16 == 0 and lda % 4 == 0");
    FA_REQUIRE(ep.amax_val == nullptr, "simt gemm has no fused argmax; materialise a logits chunk instead");
    FA_REQUIRE(ep.pl_col_scale_end == 0 && ep.f32_col_begin == 0, "simt gemm does not implement the column-range epilogue options");
    prof_note_work(2.0 * m * (double)n * k, 0.0);
    FA_LAUNCH(k_gemm_simt, dim3(cdiv(n, BN), cdiv(m, BM)), 256, 0, st, a, lda, w, m, n, k, ep.bias, ep.resid, ep.ldr,
              ep.relu ? 1 : 0, ep.out_f32, ep.ldc, ep.out_pl.hi, ep.out_pl.lo, ep.ldp);
}

}  // namespace fa
