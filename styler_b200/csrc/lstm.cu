// One layer of a bidirectional LSTM over the padded, un-packed [B, L] grid (modules.py:179-182 of the reference
// runs nn.LSTM on padded tensors, so the reverse direction starts inside the padding -- reproduced here).
// The input projection x@W_ih^T + b_ih + b_hh for all steps and both directions is a tensor-core GEMM done by
// styler_conv1d_fwd (gx, fp32 [B][L][8H]); this kernel is the latency-bound recurrence (128 dependent steps).
//
// Round-2 form (the round-1 kernel spent 1650 clk per step: 2 utterances per CTA = 160 dependent-issue FFMA per thread
// on the rt=2 FMA pipe, plus two block barriers and a smem gate exchange per step; ncu: profiles/ncu_bilstm_r2.md):
//   * one CTA per (utterance, direction) -> B*2 CTAs (128 at the bench shape: one wave on 148 SMs), 4H threads;
//   * thread t owns gate g = t & 3 of hidden unit k = t >> 2: the four gates of a unit live in one warp QUAD, so they are
//     exchanged with warp shuffles -- no shared-memory gate buffer and no barrier for it;
//   * row g*H + k of W_hh sits in registers as H/2 float2 pairs and the dot product with h (broadcast float4 reads from
//     shared memory) runs on packed FFMA2 (two MACs per issue slot; the FMA pipe issues one warp instruction per 2 clk);
//   * each thread applies the activation of ITS gate (the quad's four MUFU ops run in parallel), the activated values are
//     shuffled, every thread of the quad keeps c redundantly, the g == 0 thread publishes h;
//   * h is double-buffered in shared memory -> ONE block barrier per step;
//   * gx is prefetched four steps ahead (it comes from L2 / HBM, 600-1000 clk away; a step takes ~400).
// Gate order i, f, g, o (PyTorch).  State and gates are fp32 regardless of the activation dtype.
#include "common.cuh"

namespace sb {
namespace {

__device__ __forceinline__ float tanh_fast(float x) {   // MUFU.TANH, rel. error 2^-11
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, int H>
__global__ void __launch_bounds__(4 * H) bilstm_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                       T* __restrict__ out, long long o_bs, int o_ld, int L) {
  constexpr int G = 4 * H;
  const int b = blockIdx.x, dir = blockIdx.y;
  const int t = threadIdx.x, gate = t & 3, k = t >> 2;
  const int row = gate * H + k;                             // PyTorch row of this (gate, unit)
  __shared__ __align__(16) float hs[2][H];
  float2 w2[H / 2];
  {
    const float2* wrow = reinterpret_cast<const float2*>(whh + (static_cast<long long>(dir) * G + row) * H);
#pragma unroll
    for (int i = 0; i < H / 2; ++i) w2[i] = wrow[i];
  }
  if (t < H) hs[0][t] = 0.f;
  float c = 0.f;
  const float* gxb = gx + static_cast<long long>(b) * L * (2 * G) + dir * G + row;
  auto gx_at = [&](int step) -> float {
    if (step >= L) return 0.f;
    const int tt = dir ? L - 1 - step : step;
    return __ldg(gxb + static_cast<long long>(tt) * (2 * G));
  };
  // gx is read from L2 / HBM (600-1000 clk away) while a step takes ~400 clk: keep kPF steps of it in flight
  constexpr int kPF = 4;
  float gq[kPF];
#pragma unroll
  for (int u = 0; u < kPF; ++u) gq[u] = gx_at(u);
  T* ob = out + b * o_bs + dir * H + k;
  __syncthreads();
  const int qbase = (threadIdx.x & 31) & ~3;                // first lane of this quad
  for (int step0 = 0; step0 < L; step0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int step = step0 + u;
      if (step >= L) break;                                 // block-uniform
      const float gcur = gq[u];
      gq[u] = gx_at(step + kPF);
      const float4* h4 = reinterpret_cast<const float4*>(hs[step & 1]);
      float2 a0 = make_float2(gcur, 0.f), a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f), a3 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < H / 8; ++i) {                     // 8 hidden units per iteration: two float4 = four float2 pairs
        const float4 ha = h4[2 * i], hb = h4[2 * i + 1];
        a0 = __ffma2_rn(w2[4 * i], make_float2(ha.x, ha.y), a0);
        a1 = __ffma2_rn(w2[4 * i + 1], make_float2(ha.z, ha.w), a1);
        a2 = __ffma2_rn(w2[4 * i + 2], make_float2(hb.x, hb.y), a2);
        a3 = __ffma2_rn(w2[4 * i + 3], make_float2(hb.z, hb.w), a3);
      }
      const float2 s = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
      const float pre = s.x + s.y;
      // activation of this thread's own gate: sigmoid for i, f, o (gate 0, 1, 3), tanh for g (gate 2)
      float act;
      if constexpr (sizeof(T) == 2) {
        // bf16 activations: MUFU.TANH based gates (far below the bf16 rounding of h); sigmoid(x) = 0.5 tanh(x/2) + 0.5
        const float th = tanh_fast(gate == 2 ? pre : 0.5f * pre);
        act = gate == 2 ? th : fmaf(0.5f, th, 0.5f);
      } else {
        act = gate == 2 ? tanhf(pre) : 1.f / (1.f + expf(-pre));
      }
      const float i_ = __shfl_sync(0xffffffffu, act, qbase);
      const float f_ = __shfl_sync(0xffffffffu, act, qbase + 1);
      const float g_ = __shfl_sync(0xffffffffu, act, qbase + 2);
      const float o_ = __shfl_sync(0xffffffffu, act, qbase + 3);
      c = fmaf(f_, c, i_ * g_);
      float h;
      if constexpr (sizeof(T) == 2) h = o_ * tanh_fast(c);
      else h = o_ * tanhf(c);
      if (gate == 0) {
        hs[(step + 1) & 1][k] = h;
        const int tt = dir ? L - 1 - step : step;
        DT<T>::st(ob + static_cast<long long>(tt) * o_ld, h);
      }
      __syncthreads();                                      // h(step) visible; everyone is done reading hs[step & 1]
    }
  }
}

template <typename T, int H>
int launch(const float* gx, const float* whh, void* out, int64_t o_bs, int o_ld, int B, int L, cudaStream_t s) {
  dim3 grid(B, 2);
  bilstm_kernel<T, H><<<grid, 4 * H, 0, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, L);
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
  SB_REQUIRE((reinterpret_cast<uintptr_t>(whh) & 7) == 0, "bilstm: W_hh must be 8-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB_DISPATCH_DTYPE(dtype, T, {
    if (H == 64) return launch<T, 64>(gx, whh, out, o_bstride, o_ld, B, L, s);
    return launch<T, 80>(gx, whh, out, o_bstride, o_ld, B, L, s);
  });
  return 0;
}
