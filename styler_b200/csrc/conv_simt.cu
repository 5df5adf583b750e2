// fp32 CUDA-core Conv1d / Linear with the same fused-epilogue contract as conv_tc.cu.
// Used (a) as the exact-fp32 parity mode, (b) for shapes the tensor-core kernel does not take (N=2/4, Cin=4,
// one-row batches) and (c) as the on-device cross-check of the tcgen05 kernel.  Classic 64x64x16 smem tiling,
// 4x4 outputs per thread; LayerNorm / row-dot run as a second, warp-per-row kernel over the stored rows.
#include "common.cuh"

namespace sb {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv1d_simt_kernel(styler_conv1d_args a, int tiles_per_utt, bool fuse_tail) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int b = mt / tiles_per_utt, t0 = (mt % tiles_per_utt) * TM, n0 = nt * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> n, ty -> m
  const T* xb = static_cast<const T*>(a.x) + b * a.x_bstride;
  const T* w = static_cast<const T*>(a.w);
  const int dil = a.dilation > 1 ? a.dilation : 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < a.KS; ++tap) {
    for (int c0 = 0; c0 < a.Cin; c0 += TK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int kk = idx & 15, m = idx >> 4;
        const int c = c0 + kk;
        const int t = t0 + m + tap * dil - a.pad;
        float va = 0.f, vb = 0.f;
        if (c < a.Cin && t >= 0 && t < a.T) va = DT<T>::ld(xb + static_cast<long long>(t) * a.x_ld + c);
        const int n = n0 + m;
        if (c < a.Cin && n < a.N) vb = DT<T>::ld(w + (static_cast<long long>(tap) * a.N + n) * a.Cin + c);
        As[kk][m] = va;
        Bs[kk][m] = vb;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.T) continue;
    const bool masked = fuse_tail && a.lens != nullptr && t >= static_cast<int>(a.lens[b]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias != nullptr) v += a.bias[n];
      v = apply_act(v, a.act, a.act_slope);
      if (a.residual != nullptr) {
        const long long off = b * a.r_bstride + static_cast<long long>(t) * a.r_ld + n;
        float rv = a.residual_is_f32 ? static_cast<const float*>(a.residual)[off] : DT<T>::ld(static_cast<const T*>(a.residual) + off);
        if (a.residual_inv_lrelu != 0 && rv < 0.f) rv = rv / a.act_slope;
        v += rv;
      }
      if (fuse_tail) {
        v = apply_act(v, a.act2, a.act_slope);
        if (masked) v = 0.f;
      }
      if (a.vt != nullptr && n >= a.vt_col0) {
        DT<T>::st(static_cast<T*>(a.vt) + b * a.vt_bstride + static_cast<long long>(n - a.vt_col0) * a.vt_ld + t, v);
      } else {
        if (a.out != nullptr)
          DT<T>::st(static_cast<T*>(a.out) + b * a.o_bstride + static_cast<long long>(t) * a.o_ld + n, v);
        if (a.out_f32 != nullptr) a.out_f32[b * a.of_bstride + static_cast<long long>(t) * a.of_ld + n] = v;
    if (a.out2_f32 != nullptr) a.out2_f32[b * a.of_bstride + static_cast<long long>(t) * a.of_ld + n] = v;
        if (a.out2_f32 != nullptr) a.out2_f32[b * a.of_bstride + static_cast<long long>(t) * a.of_ld + n] = v;
      }
    }
  }
}

// N <= 4 outputs per time step (HiFi-GAN conv_post: 32 -> 1, k = 7; the 64x64 tile above would waste 63 of 64 columns):
// one thread per output time step, the [256 + (KS-1)*dil][Cin] input window and the weights staged in smem as fp32,
// HBM-bound by construction (every input row is read once per CTA).
constexpr int kSmallT = 256;

template <typename T>
__global__ void __launch_bounds__(kSmallT) conv1d_smalln_kernel(styler_conv1d_args a, int tiles_per_utt, int win_rows) {
  extern __shared__ float sm_small[];
  float* xs = sm_small;                                   // [win_rows][Cin + 1]
  const int ldx = a.Cin + 1;
  float* ws = xs + static_cast<size_t>(win_rows) * ldx;   // [KS][N][Cin]
  const int dil = a.dilation > 1 ? a.dilation : 1;
  const int b = blockIdx.x / tiles_per_utt, t0 = (blockIdx.x % tiles_per_utt) * kSmallT;
  const T* xb = static_cast<const T*>(a.x) + b * a.x_bstride;
  for (int i = threadIdx.x; i < win_rows * a.Cin; i += kSmallT) {
    const int r = i / a.Cin, c = i % a.Cin;
    const int t = t0 + r - a.pad;
    xs[r * ldx + c] = (t >= 0 && t < a.T) ? DT<T>::ld(xb + static_cast<long long>(t) * a.x_ld + c) : 0.f;
  }
  for (int i = threadIdx.x; i < a.KS * a.N * a.Cin; i += kSmallT) ws[i] = DT<T>::ld(static_cast<const T*>(a.w) + i);
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= a.T) return;
  const bool masked = a.lens != nullptr && t >= static_cast<int>(a.lens[b]);
  for (int n = 0; n < a.N; ++n) {
    float acc = 0.f;
    for (int tap = 0; tap < a.KS; ++tap) {
      const float* xr = xs + (threadIdx.x + tap * dil) * ldx;
      const float* wr = ws + (tap * a.N + n) * a.Cin;
      for (int c = 0; c < a.Cin; ++c) acc = fmaf(xr[c], wr[c], acc);
    }
    float v = acc + (a.bias != nullptr ? a.bias[n] : 0.f);
    v = apply_act(v, a.act, a.act_slope);
    if (a.residual != nullptr) {
      const long long off = b * a.r_bstride + static_cast<long long>(t) * a.r_ld + n;
      float rv = a.residual_is_f32 ? static_cast<const float*>(a.residual)[off] : DT<T>::ld(static_cast<const T*>(a.residual) + off);
      if (a.residual_inv_lrelu != 0 && rv < 0.f) rv = rv / a.act_slope;
      v += rv;
    }
    v = apply_act(v, a.act2, a.act_slope);
    if (masked) v = 0.f;
    if (a.out != nullptr) DT<T>::st(static_cast<T*>(a.out) + b * a.o_bstride + static_cast<long long>(t) * a.o_ld + n, v);
    if (a.out_f32 != nullptr) a.out_f32[b * a.of_bstride + static_cast<long long>(t) * a.of_ld + n] = v;
    if (a.out2_f32 != nullptr) a.out2_f32[b * a.of_bstride + static_cast<long long>(t) * a.of_ld + n] = v;
  }
}

// Warp per row: LayerNorm (two-pass, fp32), act2, row-dot, padding mask -- in place on the stored rows.
template <typename T>
__global__ void row_tail_kernel(styler_conv1d_args a) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.B * a.T) return;
  const int b = row / a.T, t = row % a.T;
  T* orow = a.out != nullptr ? static_cast<T*>(a.out) + b * a.o_bstride + static_cast<long long>(t) * a.o_ld : nullptr;
  float* frow = a.out_f32 != nullptr ? a.out_f32 + b * a.of_bstride + static_cast<long long>(t) * a.of_ld : nullptr;
  const int N = a.N;
  auto rd = [&](int n) { return frow != nullptr ? frow[n] : DT<T>::ld(orow + n); };
  float mean = 0.f, rstd = 1.f;
  if (a.ln_gamma != nullptr) {
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s += rd(n);
    mean = warp_sum(s) / N;
    float s2 = 0.f;
    for (int n = lane; n < N; n += 32) { const float d = rd(n) - mean; s2 += d * d; }
    rstd = rsqrtf(warp_sum(s2) / N + a.ln_eps);
  }
  const bool masked = a.lens != nullptr && t >= static_cast<int>(a.lens[b]);
  float dot = 0.f;
  for (int n = lane; n < N; n += 32) {
    float v = rd(n);
    if (a.ln_gamma != nullptr) v = (v - mean) * rstd * a.ln_gamma[n] + a.ln_beta[n];
    v = apply_act(v, a.act2, a.act_slope);
    if (a.dot_w != nullptr) dot += v * a.dot_w[n];
    if (masked) v = 0.f;
    if (orow != nullptr) DT<T>::st(orow + n, v);
    if (frow != nullptr) frow[n] = v;
  }
  if (a.dot_out != nullptr) {
    dot = warp_sum(dot);
    if (lane == 0) a.dot_out[static_cast<long long>(b) * a.T + t] = masked ? 0.f : dot + a.dot_b;
  }
}

template <typename T>
int launch(const styler_conv1d_args& a, cudaStream_t s) {
  const bool need_tail = a.ln_gamma != nullptr || a.dot_w != nullptr;
  SB_REQUIRE(!need_tail || a.out != nullptr || a.out_f32 != nullptr,
             "conv1d_simt: LayerNorm/dot epilogue needs an output buffer to stage rows");
  SB_REQUIRE(!need_tail || a.vt == nullptr, "conv1d_simt: LayerNorm/dot with vt is unsupported");
  if (!need_tail && a.vt == nullptr && a.N <= 4 && a.Cin <= 128) {
    const int dil = a.dilation > 1 ? a.dilation : 1;
    const int win_rows = kSmallT + (a.KS - 1) * dil;
    const size_t smem = (static_cast<size_t>(win_rows) * (a.Cin + 1) + static_cast<size_t>(a.KS) * a.N * a.Cin) * sizeof(float);
    if (smem <= 200 * 1024) {
      auto kern = conv1d_smalln_kernel<T>;
      if (smem > 48 * 1024) SB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      const int tpu = ceil_div(a.T, kSmallT);
      kern<<<a.B * tpu, kSmallT, smem, s>>>(a, tpu, win_rows);
      SB_LAUNCH_OK();
      return 0;
    }
  }
  const int tiles_per_utt = ceil_div(a.T, TM);
  dim3 grid(a.B * tiles_per_utt, ceil_div(a.N, TN));
  conv1d_simt_kernel<T><<<grid, 256, 0, s>>>(a, tiles_per_utt, !need_tail);
  SB_LAUNCH_OK();
  if (need_tail) {
    const int rows = a.B * a.T;
    row_tail_kernel<T><<<ceil_div(rows, 8), 256, 0, s>>>(a);
    SB_LAUNCH_OK();
  }
  return 0;
}

}  // namespace

int conv1d_simt(const styler_conv1d_args& a, cudaStream_t s) {
  SB_DISPATCH_DTYPE(a.dtype, T, return launch<T>(a, s));
  return 0;
}

}  // namespace sb

extern "C" int styler_conv1d_fwd(const styler_conv1d_args* a, void* stream) {
  sb::TraceScope trace__("conv1d", stream, a ? a->B : 0, a ? a->T : 0, a ? a->Cin : 0, a ? a->N * 100 + a->KS : 0);
  using namespace sb;
  SB_REQUIRE(a != nullptr && a->x != nullptr && a->w != nullptr, "conv1d: null x/w");
  SB_REQUIRE(a->B > 0 && a->T > 0 && a->Cin > 0 && a->N > 0 && a->KS > 0, "conv1d: bad shape B=%d T=%d Cin=%d N=%d KS=%d",
             a->B, a->T, a->Cin, a->N, a->KS);
  SB_REQUIRE(a->out != nullptr || a->out_f32 != nullptr || a->dot_out != nullptr || a->vt != nullptr, "conv1d: no output");
  SB_REQUIRE(sb::dtype_ok(a->dtype), "conv1d: bad dtype %d", a->dtype);
  SB_REQUIRE(a->dilation >= 0, "conv1d: bad dilation %d", a->dilation);
  SB_REQUIRE((a->act != STYLER_ACT_LRELU && a->act2 != STYLER_ACT_LRELU && a->residual_inv_lrelu == 0) ||
                 (a->act_slope > 0.f && a->act_slope < 1.f),
             "conv1d: leaky ReLU needs 0 < act_slope < 1 (got %g)", static_cast<double>(a->act_slope));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int impl = a->impl;
  if (impl == STYLER_IMPL_AUTO) impl = conv1d_tc_supported(*a, nullptr) && a->B * a->T >= 64 ? STYLER_IMPL_TC : STYLER_IMPL_SIMT;
  if (impl == STYLER_IMPL_TC) return conv1d_tc(*a, s);
  SB_REQUIRE(impl == STYLER_IMPL_SIMT, "conv1d: bad impl %d", a->impl);
  SB_REQUIRE(a->gn_partial == nullptr, "conv1d: gn_partial (fused GroupNorm statistics) exists on the tensor-core path only");
  return conv1d_simt(*a, s);
}

// Debug/tuning hook (tools/phase_timing.py): per-CTA clock64() phase stamps of the tcgen05 conv kernel; NULL disables.
extern "C" int styler_debug_set_phase_buffer(int64_t* buf, int32_t capacity_ctas) {
  sb::set_phase_buffer(reinterpret_cast<long long*>(buf), capacity_ctas);
  return 0;
}
