// One layer of a bidirectional LSTM over the padded, un-packed [B, L] grid (modules.py:179-182 of the reference
// runs nn.LSTM on padded tensors, so the reverse direction starts inside the padding -- reproduced here).
// The input projection x@W_ih^T + b_ih + b_hh for all steps and both directions is a tensor-core GEMM done by
// styler_conv1d_fwd (gx, fp32 [B][L][8H]); this kernel is the latency-bound recurrence:
//   one CTA per (direction, NB utterances), 4H threads; thread j keeps row j of W_hh in registers, h lives in
//   shared memory (broadcast reads), c in a register; gx for the next step is prefetched during the current one.
// Gate order i, f, g, o (PyTorch).  State and gates are fp32 regardless of the activation dtype.
#include "common.cuh"

namespace sb {
namespace {

template <typename T, int H, int NB>
__global__ void __launch_bounds__(4 * H) bilstm_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                       T* __restrict__ out, long long o_bs, int o_ld, int B, int L) {
  constexpr int G = 4 * H;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * NB;
  const int j = threadIdx.x;
  __shared__ __align__(16) float hs[NB][H];
  __shared__ float gs[NB][G];
  float w[H];
  const float* wrow = whh + (static_cast<long long>(dir) * G + j) * H;
#pragma unroll
  for (int k = 0; k < H; ++k) w[k] = wrow[k];
  for (int i = j; i < NB * H; i += G) (&hs[0][0])[i] = 0.f;
  float c = 0.f;
  const int my_nb = j / H, my_k = j % H;  // valid when j < NB*H
  float gcur[NB], gnext[NB];
  auto gx_at = [&](int nb, int step) -> float {
    const int b = b0 + nb;
    if (b >= B || step >= L) return 0.f;
    const int t = dir ? L - 1 - step : step;
    return gx[(static_cast<long long>(b) * L + t) * (2 * G) + dir * G + j];
  };
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) gcur[nb] = gx_at(nb, 0);
  __syncthreads();
  for (int step = 0; step < L; ++step) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) gnext[nb] = gx_at(nb, step + 1);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      // four independent accumulation chains (the H-long dependent FMA chain was the per-step critical path)
      float acc0 = gcur[nb], acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      const float4* h4 = reinterpret_cast<const float4*>(&hs[nb][0]);
#pragma unroll
      for (int k = 0; k < H / 4; ++k) {
        const float4 hv = h4[k];
        acc0 = fmaf(w[4 * k], hv.x, acc0);
        acc1 = fmaf(w[4 * k + 1], hv.y, acc1);
        acc2 = fmaf(w[4 * k + 2], hv.z, acc2);
        acc3 = fmaf(w[4 * k + 3], hv.w, acc3);
      }
      gs[nb][j] = (acc0 + acc1) + (acc2 + acc3);
    }
    __syncthreads();
    if (j < NB * H && b0 + my_nb < B) {
      const float gi = gs[my_nb][my_k], gf = gs[my_nb][H + my_k], gg = gs[my_nb][2 * H + my_k], go = gs[my_nb][3 * H + my_k];
      float i_, f_, o_, g_, h;
      if constexpr (sizeof(T) == 2) {
        // bf16 activations: MUFU.TANH based gates (rel. error 2^-11, far below the bf16 rounding of h); the fp32 modes
        // keep the accurate expf/tanhf forms
        auto tanh_fast = [](float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
        i_ = fmaf(0.5f, tanh_fast(0.5f * gi), 0.5f);
        f_ = fmaf(0.5f, tanh_fast(0.5f * gf), 0.5f);
        o_ = fmaf(0.5f, tanh_fast(0.5f * go), 0.5f);
        g_ = tanh_fast(gg);
        c = f_ * c + i_ * g_;
        h = o_ * tanh_fast(c);
      } else {
        i_ = 1.f / (1.f + expf(-gi)); f_ = 1.f / (1.f + expf(-gf)); o_ = 1.f / (1.f + expf(-go));
        g_ = tanhf(gg);
        c = f_ * c + i_ * g_;
        h = o_ * tanhf(c);
      }
      hs[my_nb][my_k] = h;
      const int t = dir ? L - 1 - step : step;
      DT<T>::st(out + (b0 + my_nb) * o_bs + static_cast<long long>(t) * o_ld + dir * H + my_k, h);
    }
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) gcur[nb] = gnext[nb];
  }
}

template <typename T, int H>
int launch(const float* gx, const float* whh, void* out, int64_t o_bs, int o_ld, int B, int L, cudaStream_t s) {
  constexpr int NB = 2;
  dim3 grid(ceil_div(B, NB), 2);
  bilstm_kernel<T, H, NB><<<grid, 4 * H, 0, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, B, L);
  SB_LAUNCH_OK();
  return 0;
}

}  // namespace
}  // namespace sb

extern "C" int styler_bilstm_layer_fwd(const float* gx, const float* whh, void* out, int64_t o_bstride, int32_t o_ld,
                                       int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream) {
  using namespace sb;
  SB_REQUIRE(gx && whh && out, "bilstm: null pointer");
  SB_REQUIRE(B > 0 && L > 0, "bilstm: bad shape");
  SB_REQUIRE(H == 64 || H == 80, "bilstm: hidden size %d not instantiated (64, 80)", H);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB_DISPATCH_DTYPE(dtype, T, {
    if (H == 64) return launch<T, 64>(gx, whh, out, o_bstride, o_ld, B, L, s);
    return launch<T, 80>(gx, whh, out, o_bstride, o_ld, B, L, s);
  });
  return 0;
}
