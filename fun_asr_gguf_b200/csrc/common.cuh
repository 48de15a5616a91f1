// Shared host/device helpers for the Fun-ASR-Nano front-half library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace fa {

// Geometry of the path. Reference: fun_asr_gguf/model_definition.py:190-200, 223-229 and
// 01-Export-Encoder-Adaptor-CTC.py:41-45.
constexpr int kHop = 160, kNfft = 400, kBins = 201, kMels = 80, kLfrM = 7, kLfrN = 6;
constexpr int kDin = 560, kDenc = 512, kDffn = 2048, kDllm = 1024, kFsmnK = 11;
constexpr int kEncLayers = 70;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        throw Error(buf);
    }
}
#define FA_CUDA(x) ::fa::cuda_check((x), #x, __FILE__, __LINE__)
#define FA_REQUIRE(cond, msg)                                                        \
    do {                                                                             \
        if (!(cond)) throw ::fa::Error(std::string("requirement failed: ") + (msg)); \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Launch accounting: every kernel this library launches goes through FA_LAUNCH so that
// bench.py can report gpu_launches from a counter instead of an estimate.
extern thread_local int64_t g_launches;
// Optional per-launch CUDA-event timing (bench.py's roofline leg): off by default, and when off the
// two hooks are a single predictable branch.
extern thread_local bool g_prof_on;
void prof_before(const char* kernel, cudaStream_t st);
void prof_after(cudaStream_t st);
void prof_note_work(double flops, double bytes);   // algorithmic work of the next launch
void prof_note_tag(const char* tag);               // shape class of the next launch (profile key suffix)
// Every launch carries the programmatic-stream-serialization attribute: the next kernel's CTAs may be placed
// and run their prologue (barrier init, TMEM allocation, descriptor prefetch) while the previous kernel drains;
// each kernel calls grid_dependency_wait() before it touches global memory.  FUNASR_B200_PDL=0 turns it off.
extern bool g_pdl;
template <class... KArgs, class... Args>
inline void launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = g_pdl ? 1 : 0;
    FA_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
#define FA_LAUNCH(kernel, grid, block, smem, stream, ...)                                        \
    do {                                                                                         \
        if (::fa::g_prof_on) ::fa::prof_before(#kernel, (stream));                               \
        ::fa::launch_ex(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); \
        ++::fa::g_launches;                                                                      \
        if (::fa::g_prof_on) ::fa::prof_after((stream));                                         \
    } while (0)

#ifdef __CUDACC__
// Wait until the kernels this launch depends on have completed and their writes are visible (no-op without PDL).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// bf16 "planes": a fp32 value v is carried as hi = bf16(v), lo = bf16(v - hi).  hi+lo keeps
// 16 mantissa bits, and the three products hi*hi + hi*lo + lo*hi on the tensor cores
// (fp32 accumulate) reproduce an fp32 GEMM to ~2^-16 relative per product.
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Packed form: two values -> one 32-bit word of hi halves and one of lo halves (element 0 in the low
// 16 bits).  cvt.rn.bf16x2.f32 runs on the FMA/ALU pipes; the scalar cvt above is a quarter-rate XU op.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

#ifdef __CUDACC__
// Four floats -> four e4m3 bytes (element 0 in the low byte): round to nearest even, saturating at +-448.
__device__ __forceinline__ uint32_t f32x4_to_e4m3(float a, float b, float c, float d) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
#endif

}  // namespace fa
