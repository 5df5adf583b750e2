// Shared helpers: status/error plumbing for the C ABI, dtype traits, small device utilities.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/styler_b200.h"

namespace sb {

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Debug timeline (styler_debug_trace): when enabled, every leaf C-ABI entry records a CUDA event pair on its stream around
// the kernels it enqueues -- also when it is called from inside a composite entry (decoder, FFT block, predictor, PostNet).
struct TraceScope {
  TraceScope(const char* name, void* stream, int a = 0, int b = 0, int c = 0, int d = 0);
  ~TraceScope();
  int idx;
};

// Per-DEVICE one-shot state.  A process may drive several GPUs from several threads (nn.DataParallel, train.py:33 of the
// reference): function attributes such as the >48 KB dynamic-smem opt-in are per device, so "already done" flags are kept
// per device ordinal; a race at worst repeats an idempotent cudaFuncSetAttribute.
inline int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : 0;
}
struct DeviceFlags {
  std::atomic<unsigned long long> bits{0};
  bool test(int dev) const { return dev >= 0 && dev < 64 && ((bits.load(std::memory_order_acquire) >> dev) & 1ull) != 0; }
  void set(int dev) { if (dev >= 0 && dev < 64) bits.fetch_or(1ull << dev, std::memory_order_release); }
};
int num_sms();   // SM count of the current device (cached per device)

// A/B switches: read once from the environment (STYLER_<NAME>), overridable at run time through styler_set_tuning()
// (tests and tools/prof_kernels.py flip them inside one process).
enum Tuning { TUNE_TC_2CTA = 0, TUNE_TC_PERSIST, TUNE_CONV_WIN, TUNE_TC_BN, TUNE_TC_SMEM_KB, TUNE_PDL, TUNE_ATTN_PERSIST, TUNE_TC_WIDE, TUNE_LSTM_MULTI, TUNE_ATTN_POLY, TUNE_LSTM_MMA, TUNE_STFT_OCC, TUNE_COUNT };
int tuning(Tuning t);

#define SB_OPT_IN_SMEM(flags, kern, bytes)                                                                     \
  do {                                                                                                         \
    const int dev__ = ::sb::current_device();                                                                  \
    if (!(flags).test(dev__)) {                                                                                \
      SB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes))); \
      (flags).set(dev__);                                                                                      \
    }                                                                                                          \
  } while (0)

#define SB_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::sb::set_error(__VA_ARGS__);    \
      return -1;                       \
    }                                  \
  } while (0)

#define SB_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::sb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return static_cast<int>(e__);                                                        \
    }                                                                                      \
  } while (0)

#define SB_LAUNCH_OK()                                                                     \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      ::sb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return static_cast<int>(e__);                                                        \
    }                                                                                      \
    ::sb::count_launch();                                                                  \
  } while (0)

// ---- dtype traits ---------------------------------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float> {
  static constexpr int code = STYLER_F32;
  __device__ static __forceinline__ float ld(const float* p) { return *p; }
  __device__ static __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct DT<__nv_bfloat16> {
  static constexpr int code = STYLER_BF16;
  __device__ static __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct DT<__half> {     // fp16 storage: 11 significand bits (tf32-class accuracy) at the bf16 tensor-pipe rate
  static constexpr int code = STYLER_F16;
  __device__ static __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  __device__ static __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};
inline bool dtype_ok(int dtype) { return dtype == STYLER_F32 || dtype == STYLER_BF16 || dtype == STYLER_F16; }
// two floats -> one 32-bit word of two T elements (round to nearest even)
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// bf16 storage tolerates MUFU.TANH (2^-11); fp16 storage (tf32-class accuracy mode) uses ex2/rcp based forms (~1e-6 absolute)
template <typename T> constexpr bool kBf16Math = DT<T>::code == STYLER_BF16;
__device__ __forceinline__ float sigmoid_ex2(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_ex2(float x) { return fmaf(2.f, sigmoid_ex2(2.f * x), -1.f); }

// 8 consecutive elements <-> 8 floats (16-byte access for bf16, 2x16 for fp32); pointers must be 16B aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope = 0.f) {
  if (act == STYLER_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == STYLER_ACT_TANH) return tanhf(v);
  if (act == STYLER_ACT_LRELU) return fmaxf(v, v * slope);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch with programmatic stream serialization (PDL): the kernel must call pdl_grid_dependency_wait() before it reads
// anything a previous kernel wrote.  Opt-in with STYLER_PDL=1 (it measured slower on the full forward: the early CTAs of
// the next kernel hold smem/TMEM slots while they wait).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// Launch as thread-block clusters of two consecutive CTAs along x (CTA pairs for tcgen05 cta_group::2).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster2(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// dispatch on the ABI dtype code
#define SB_DISPATCH_DTYPE(code, T, ...)                          \
  do {                                                           \
    if ((code) == STYLER_F32) { using T = float; __VA_ARGS__; }  \
    else if ((code) == STYLER_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else if ((code) == STYLER_F16) { using T = __half; __VA_ARGS__; } \
    else { ::sb::set_error("unsupported dtype code %d", (int)(code)); return -1; } \
  } while (0)

// entry points implemented per translation unit
int conv1d_simt(const styler_conv1d_args& a, cudaStream_t s);
int conv1d_tc(const styler_conv1d_args& a, cudaStream_t s);
bool conv1d_tc_supported(const styler_conv1d_args& a, const char** why);
bool conv1d_win_supported(const styler_conv1d_args& a);   // conv_win.cu: few channels, many taps (HiFi-GAN resblocks)
int conv1d_win(const styler_conv1d_args& a, cudaStream_t s);
void set_phase_buffer(long long* buf, int cap);
int attention_simt(const void* qk, int64_t qk_bs, int qk_ld, const void* vt, int64_t vt_bs, int vt_ld,
                   const int64_t* lens, void* ctx, int64_t ctx_bs, int ctx_ld, int B, int T, int H, int dtype,
                   cudaStream_t s);
int attention_tc(const void* qk, int64_t qk_bs, int qk_ld, const void* vt, int64_t vt_bs, int vt_ld,
                 const int64_t* lens, void* ctx, int64_t ctx_bs, int ctx_ld, int B, int T, int H, int dtype,
                 cudaStream_t s);

}  // namespace sb
