// One layer of a bidirectional LSTM over the padded, un-packed [B, L] grid (modules.py:179-182 of the reference
// runs nn.LSTM on padded tensors, so the reverse direction starts inside the padding -- reproduced here).
// The input projection x@W_ih^T + b_ih + b_hh for all steps and both directions is a tensor-core GEMM done by
// styler_conv1d_fwd (gx, fp32 [B][L][2][H][4]); this kernel is the latency-bound recurrence (128 dependent steps).
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
//   * gx is streamed into shared memory by bulk async copies (cp.async.bulk + mbarrier), three 8-step chunks ahead; it is
//     stored by the projection GEMM in QUAD order [B][L][2 dirs][H units][4 gates] (the rows of W_ih are permuted at pack
//     time), so element t of a row belongs to thread t: contiguous copies, conflict-free reads.
// Gate order i, f, g, o (PyTorch).  State and gates are fp32 regardless of the activation dtype.
#include "common.cuh"
#include "ptx.cuh"

namespace sb {
namespace {

__device__ __forceinline__ float tanh_fast(float x) {   // MUFU.TANH, rel. error 2^-11
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kChunk = 8;      // steps of gx per bulk copy group
constexpr int kBufs = 4;       // chunks in flight / in use: gx is fetched 3 chunks = 24 steps (~10 k clk) ahead of its use

template <typename T, int H>
__global__ void __launch_bounds__(4 * H) bilstm_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                       T* __restrict__ out, long long o_bs, int o_ld, int L) {
  constexpr int G = 4 * H;
  const int b = blockIdx.x, dir = blockIdx.y;
  const int t = threadIdx.x, gate = t & 3, k = t >> 2;
  const int row = gate * H + k;                             // PyTorch row of this (gate, unit) in W_hh
  __shared__ __align__(16) float hs[2][H];
  __shared__ __align__(128) float gbuf[kBufs][kChunk][G];   // gx chunks, element t of a row belongs to thread t (quad order)
  __shared__ __align__(8) uint64_t gbar[kBufs];
  float2 w2[H / 2];
  {
    const float2* wrow = reinterpret_cast<const float2*>(whh + (static_cast<long long>(dir) * G + row) * H);
#pragma unroll
    for (int i = 0; i < H / 2; ++i) w2[i] = wrow[i];
  }
  if (t < H) hs[0][t] = 0.f;
  if (t == 0) {
    for (int i = 0; i < kBufs; ++i) mbar_init(&gbar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // gx comes from HBM / L2 (ncu on the register-prefetch version: 34 % of all stall samples on the first use of the loaded
  // value): one thread streams it into shared memory with bulk async copies (one 4H-float row per step), kBufs - 1 chunks ahead
  const float* gxb = gx + static_cast<long long>(b) * L * (2 * G) + dir * G;
  const int n_chunks = (L + kChunk - 1) / kChunk;
  auto fetch = [&](int c) {                                 // thread 0 only
    const int rows = min(kChunk, L - c * kChunk);
    uint64_t* bar = &gbar[c % kBufs];
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(rows * G * sizeof(float)));
    for (int u = 0; u < rows; ++u) {
      const int step = c * kChunk + u;
      const int tt = dir ? L - 1 - step : step;
      bulk_load_1d(&gbuf[c % kBufs][u][0], gxb + static_cast<long long>(tt) * (2 * G), G * sizeof(float), bar);
    }
  };
  if (t == 0)
    for (int c = 0; c < kBufs - 1 && c < n_chunks; ++c) fetch(c);
  float c_state = 0.f;
  T* ob = out + b * o_bs + dir * H + k;
  const int qbase = (threadIdx.x & 31) & ~3;                // first lane of this quad
  for (int c = 0; c < n_chunks; ++c) {
    // every thread passed the barrier that ended chunk c - 1, so its buffer (= that of chunk c + kBufs - 1) is free again
    if (t == 0 && c + kBufs - 1 < n_chunks) fetch(c + kBufs - 1);
    mbar_wait(&gbar[c % kBufs], (c / kBufs) & 1);
    const int rows = min(kChunk, L - c * kChunk);
    for (int u = 0; u < rows; ++u) {
      const int step = c * kChunk + u;
      const float gcur = gbuf[c % kBufs][u][t];
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
      c_state = fmaf(f_, c_state, i_ * g_);
      float h;
      if constexpr (sizeof(T) == 2) h = o_ * tanh_fast(c_state);
      else h = o_ * tanhf(c_state);
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
  SB_REQUIRE((reinterpret_cast<uintptr_t>(gx) & 15) == 0, "bilstm: gx must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB_DISPATCH_DTYPE(dtype, T, {
    if (H == 64) return launch<T, 64>(gx, whh, out, o_bstride, o_ld, B, L, s);
    return launch<T, 80>(gx, whh, out, o_bstride, o_ld, B, L, s);
  });
  return 0;
}
