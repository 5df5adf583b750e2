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
      if constexpr (kBf16Math<T>) {
        // bf16 activations: MUFU.TANH based gates (far below the bf16 rounding of h); sigmoid(x) = 0.5 tanh(x/2) + 0.5
        const float th = tanh_fast(gate == 2 ? pre : 0.5f * pre);
        act = gate == 2 ? th : fmaf(0.5f, th, 0.5f);
      } else if constexpr (sizeof(T) == 2) {                  // fp16 activations: ex2/rcp forms (~1e-6)
        act = gate == 2 ? tanh_ex2(pre) : sigmoid_ex2(pre);
      } else {
        act = gate == 2 ? tanhf(pre) : 1.f / (1.f + expf(-pre));
      }
      const float i_ = __shfl_sync(0xffffffffu, act, qbase);
      const float f_ = __shfl_sync(0xffffffffu, act, qbase + 1);
      const float g_ = __shfl_sync(0xffffffffu, act, qbase + 2);
      const float o_ = __shfl_sync(0xffffffffu, act, qbase + 3);
      c_state = fmaf(f_, c_state, i_ * g_);
      float h;
      if constexpr (kBf16Math<T>) h = o_ * tanh_fast(c_state);
      else if constexpr (sizeof(T) == 2) h = o_ * tanh_ex2(c_state);
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

// Multi-utterance, register-blocked form.  The recurrence is bound by shared-memory reads of h, not by the FMA pipe: a
// broadcast LDS.128 costs ~2 LSU wavefronts (ncu: 472 wavefronts per CTA-step above) and feeds only two FFMA2 when a thread
// owns ONE row of W_hh.  Here
//   * a thread owns TWO rows (gates i,f or g,o of one hidden unit: 2H threads, the rows as H float2 pairs in registers) and
//   * a CTA carries U = 4 utterances of one direction,
// so every 16 bytes of h read from shared memory feed four FFMA2, the barrier / shuffle / MUFU latency of a step is paid once
// per four utterances, and a layer is B/4 x 2 CTAs: the BiLSTMs of the four style-factor branches (four streams) are all
// resident at once (128 CTAs at B = 64) instead of 512 CTAs queueing for SM slots at one or two per SM (measured 0.15-0.28 ms
// per layer inside the forward against 0.07 ms alone).  The arithmetic order of every dot product is that of the
// one-utterance kernel, so the two forms agree bitwise.
constexpr int kChunkM = 4;     // steps of gx per bulk copy group (U rows of 4H floats per step)
constexpr int kBufsM = 4;

template <typename T>
__device__ __forceinline__ float lstm_act(float pre, bool is_tanh) {
  if constexpr (kBf16Math<T>) {
    const float th = tanh_fast(is_tanh ? pre : 0.5f * pre);
    return is_tanh ? th : fmaf(0.5f, th, 0.5f);
  } else if constexpr (sizeof(T) == 2) {
    return is_tanh ? tanh_ex2(pre) : sigmoid_ex2(pre);
  } else {
    return is_tanh ? tanhf(pre) : 1.f / (1.f + expf(-pre));
  }
}

template <typename T, int H, int U>
__global__ void __launch_bounds__(2 * H, 1) bilstm_multi_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                                T* __restrict__ out, long long o_bs, int o_ld, int B, int L) {
  constexpr int G = 4 * H;
  const int b0 = blockIdx.x * U, dir = blockIdx.y;
  const int nu = min(U, B - b0);                            // utterances of this CTA that exist
  const int t = threadIdx.x, half = t & 1, k = t >> 1;      // half 0: gates i, f of unit k; half 1: gates g, o
  extern __shared__ __align__(128) uint8_t lstm_smem[];
  float (*gbuf)[kChunkM][U][G] = reinterpret_cast<float (*)[kChunkM][U][G]>(lstm_smem);          // [kBufsM]
  float (*hs)[U][H] = reinterpret_cast<float (*)[U][H]>(lstm_smem + sizeof(float) * kBufsM * kChunkM * U * G);   // [2]
  uint64_t* gbar = reinterpret_cast<uint64_t*>(lstm_smem + sizeof(float) * (kBufsM * kChunkM * U * G + 2 * U * H));
  float2 wa[H / 2], wb[H / 2];                              // rows (2 half) * H + k and (2 half + 1) * H + k of W_hh
  {
    const float2* ra = reinterpret_cast<const float2*>(whh + (static_cast<long long>(dir) * G + (2 * half) * H + k) * H);
    const float2* rb = reinterpret_cast<const float2*>(whh + (static_cast<long long>(dir) * G + (2 * half + 1) * H + k) * H);
#pragma unroll
    for (int i = 0; i < H / 2; ++i) { wa[i] = ra[i]; wb[i] = rb[i]; }
  }
  for (int i = t; i < 2 * U * H; i += 2 * H) (&hs[0][0][0])[i] = 0.f;
  for (int i = t; i < kBufsM * kChunkM * U * G; i += 2 * H) (&gbuf[0][0][0][0])[i] = 0.f;   // rows of absent utterances stay zero
  if (t == 0) {
    for (int i = 0; i < kBufsM; ++i) mbar_init(&gbar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();                                 // the zero fill (generic proxy) is ordered before the bulk copies
  __syncthreads();
  const float* gxb = gx + static_cast<long long>(b0) * L * (2 * G) + dir * G;
  const int n_chunks = (L + kChunkM - 1) / kChunkM;
  auto fetch = [&](int c) {                                 // thread 0 only
    const int rows = min(kChunkM, L - c * kChunkM);
    uint64_t* bar = &gbar[c % kBufsM];
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(rows * nu * G * sizeof(float)));
    for (int u = 0; u < rows; ++u) {
      const int step = c * kChunkM + u;
      const int tt = dir ? L - 1 - step : step;
      for (int j = 0; j < nu; ++j)
        bulk_load_1d(&gbuf[c % kBufsM][u][j][0], gxb + (static_cast<long long>(j) * L + tt) * (2 * G), G * sizeof(float), bar);
    }
  };
  if (t == 0)
    for (int c = 0; c < kBufsM - 1 && c < n_chunks; ++c) fetch(c);
  float c_state[U];
#pragma unroll
  for (int j = 0; j < U; ++j) c_state[j] = 0.f;
  T* ob = out + b0 * o_bs + dir * H + k;
  for (int c = 0; c < n_chunks; ++c) {
    if (t == 0 && c + kBufsM - 1 < n_chunks) fetch(c + kBufsM - 1);
    mbar_wait(&gbar[c % kBufsM], (c / kBufsM) & 1);
    const int rows = min(kChunkM, L - c * kChunkM);
    for (int u = 0; u < rows; ++u) {
      const int step = c * kChunkM + u;
      // accumulators: [utterance][row a|b][4 chains], seeded with gx exactly as in the one-utterance kernel (chain 0 = (gx, 0))
      float2 acc[U][2][4];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const float2 g2 = *reinterpret_cast<const float2*>(&gbuf[c % kBufsM][u][j][4 * k + 2 * half]);   // quad order: [unit][gate]
        acc[j][0][0] = make_float2(g2.x, 0.f);
        acc[j][1][0] = make_float2(g2.y, 0.f);
#pragma unroll
        for (int q = 1; q < 4; ++q) acc[j][0][q] = acc[j][1][q] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < H / 8; ++i) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float4* h4 = reinterpret_cast<const float4*>(hs[step & 1][j]);
          const float4 ha = h4[2 * i], hb = h4[2 * i + 1];
          const float2 h0 = make_float2(ha.x, ha.y), h1 = make_float2(ha.z, ha.w), h2 = make_float2(hb.x, hb.y), h3 = make_float2(hb.z, hb.w);
          acc[j][0][0] = __ffma2_rn(wa[4 * i], h0, acc[j][0][0]);
          acc[j][0][1] = __ffma2_rn(wa[4 * i + 1], h1, acc[j][0][1]);
          acc[j][0][2] = __ffma2_rn(wa[4 * i + 2], h2, acc[j][0][2]);
          acc[j][0][3] = __ffma2_rn(wa[4 * i + 3], h3, acc[j][0][3]);
          acc[j][1][0] = __ffma2_rn(wb[4 * i], h0, acc[j][1][0]);
          acc[j][1][1] = __ffma2_rn(wb[4 * i + 1], h1, acc[j][1][1]);
          acc[j][1][2] = __ffma2_rn(wb[4 * i + 2], h2, acc[j][1][2]);
          acc[j][1][3] = __ffma2_rn(wb[4 * i + 3], h3, acc[j][1][3]);
        }
      }
      const int tt = dir ? L - 1 - step : step;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const float2 sa = __fadd2_rn(__fadd2_rn(acc[j][0][0], acc[j][0][1]), __fadd2_rn(acc[j][0][2], acc[j][0][3]));
        const float2 sb = __fadd2_rn(__fadd2_rn(acc[j][1][0], acc[j][1][1]), __fadd2_rn(acc[j][1][2], acc[j][1][3]));
        // half 0: (i, f) both sigmoid; half 1: (g tanh, o sigmoid)
        const float va = lstm_act<T>(sa.x + sa.y, half == 1);
        const float vb = lstm_act<T>(sb.x + sb.y, false);
        const float pa = __shfl_xor_sync(0xffffffffu, va, 1);
        const float pb = __shfl_xor_sync(0xffffffffu, vb, 1);
        const float i_ = half == 0 ? va : pa, f_ = half == 0 ? vb : pb, g_ = half == 0 ? pa : va, o_ = half == 0 ? pb : vb;
        c_state[j] = fmaf(f_, c_state[j], i_ * g_);
        float h;
        if constexpr (kBf16Math<T>) h = o_ * tanh_fast(c_state[j]);
        else if constexpr (sizeof(T) == 2) h = o_ * tanh_ex2(c_state[j]);
        else h = o_ * tanhf(c_state[j]);
        if (half == 0) {
          hs[(step + 1) & 1][j][k] = h;
          if (j < nu) DT<T>::st(ob + j * o_bs + static_cast<long long>(tt) * o_ld, h);
        }
      }
      __syncthreads();
    }
  }
}

// Tensor-core form for 16-bit activations (bf16 / fp16 modes): the recurrent product of a step is
//     gates^T [4H x 16 utterances] = W_hh [4H x H] . h^T [H x 16]
// on warp-level mma.sync (m16n8k16, fp32 accumulate, the accumulator seeded with gx).  The CUDA-core forms above are bound by
// the FMA pipe (4H x H MACs per utterance and step at 64 register-operand MACs per clock per SM: >= 400 clk per utterance-step
// before any latency), so one layer of the four style-factor branches keeps most of the machine busy for ~0.2 ms with
// latency-bound CTAs that fill the register file (the decoder kernels of a pipelined neighbour batch cannot co-reside).
// Here a CTA carries SIXTEEN utterances of one direction and a layer is B/16 x 2 CTAs (8 at B = 64; 32 for the four branches):
//   * warp w owns hidden units 8w .. 8w+7; the rows of W_hh are permuted so that the two m16 tiles of a warp hold gates (i,f)
//     and (g,o) of those units in the row halves of the fragment: the thread that receives D[row = lane/4 (+8)][col = 2 (lane%4)
//     (+1)] of both tiles then owns ALL FOUR gates of one unit for two utterances per n-tile -- the cell update needs no
//     shuffles and no shared-memory gate exchange;
//   * W_hh lives in registers as A fragments (2 x H/16 x 4 regs), h(t-1) in shared memory as [utterance][unit] 16-bit rows
//     (B fragments are plain 32-bit LDS, conflict-free with an 8-element row pad), double-buffered -> ONE barrier per step;
//   * gx is read straight from global memory as one float4 (the quad order i,f,g,o of a unit) per (unit, utterance), two steps
//     ahead of its use.
// h enters the product rounded to the activation type (it is stored in that type anyway); c and the gates stay fp32.
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};


template <typename T, int H, int NT>      // NT: n8 tiles (8 utterances each) per CTA
__global__ void __launch_bounds__(4 * H, 1) bilstm_mma_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                              T* __restrict__ out, long long o_bs, int o_ld, int B, int L) {
  static_assert(sizeof(T) == 2 && H % 16 == 0, "16-bit activations, H a multiple of the k16 tile");
  constexpr int G = 4 * H, KT = H / 16, LD = H + 8, kMmaUtt = 8 * NT;         // row pitch of the h buffer in elements (pad: conflict-free LDS)
  const int b0 = blockIdx.x * kMmaUtt, dir = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = lane >> 2, cq = lane & 3;
  const int u = 8 * warp + r0;                              // this thread's hidden unit
  __shared__ __align__(16) T hsm[2][kMmaUtt][LD];
  // A fragments: tile m, k-tile kt: a0 = (gate 2m, k..k+1), a1 = (gate 2m+1, k..k+1), a2 / a3 = the same rows at k + 8
  uint32_t wf[2][KT][4];
  {
    const float* wd = whh + static_cast<long long>(dir) * G * H;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const float* ra = wd + (static_cast<long long>(2 * m) * H + u) * H + kt * 16 + 2 * cq;
        const float* rb = wd + (static_cast<long long>(2 * m + 1) * H + u) * H + kt * 16 + 2 * cq;
        wf[m][kt][0] = pack2<T>(ra[0], ra[1]);
        wf[m][kt][1] = pack2<T>(rb[0], rb[1]);
        wf[m][kt][2] = pack2<T>(ra[8], ra[9]);
        wf[m][kt][3] = pack2<T>(rb[8], rb[9]);
      }
  }
  for (int i = threadIdx.x; i < 2 * kMmaUtt * LD; i += 4 * H) DT<T>::st(&hsm[0][0][0] + i, 0.f);
  // this thread's utterances: n = nt * 8 + 2 cq + e  (nt, e in {0, 1})
  const float* gp[NT][2];
  T* op[NT][2];
  bool ok[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = nt * 8 + 2 * cq + e;
      ok[nt][e] = b0 + n < B;
      const int bb = ok[nt][e] ? b0 + n : b0;
      gp[nt][e] = gx + static_cast<long long>(bb) * L * (2 * G) + dir * G + 4 * u;
      op[nt][e] = out + bb * o_bs + dir * H + u;
    }
  auto load_gx = [&](int step, float4 (&g)[NT][2]) {
    const int tt = dir ? L - 1 - step : step;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        g[nt][e] = (ok[nt][e] && step < L) ? __ldg(reinterpret_cast<const float4*>(gp[nt][e] + static_cast<long long>(tt) * (2 * G)))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 gA[NT][2], gB[NT][2];                                // gx of step t (gA) and t + 1 (gB); t + 2 is fetched into gC below
  load_gx(0, gA);
  load_gx(1, gB);
  float c_state[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) c_state[nt][0] = c_state[nt][1] = 0.f;
  __syncthreads();
  for (int step = 0; step < L; ++step) {
    float4 gC[NT][2];
    load_gx(step + 2, gC);
    const int cur = step & 1;
    // accumulators seeded with gx: d[m][nt] = {gate 2m (n0), gate 2m (n0+1), gate 2m+1 (n0), gate 2m+1 (n0+1)}
    float d[2][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      d[0][nt][0] = gA[nt][0].x; d[0][nt][1] = gA[nt][1].x; d[0][nt][2] = gA[nt][0].y; d[0][nt][3] = gA[nt][1].y;
      d[1][nt][0] = gA[nt][0].z; d[1][nt][1] = gA[nt][1].z; d[1][nt][2] = gA[nt][0].w; d[1][nt][3] = gA[nt][1].w;
    }
    // two partial accumulators per tile (even / odd k-tiles): the HMMA dependency chain is the latency of a step (ncu: 29 % of
    // the stall samples sit on the first consumer of the chain), so it is cut from KT to ceil(KT / 2) links
    float d2[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[m][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const T* hrow = &hsm[cur][nt * 8 + r0][kt * 16 + 2 * cq];
        const uint32_t bf0 = *reinterpret_cast<const uint32_t*>(hrow);
        const uint32_t bf1 = *reinterpret_cast<const uint32_t*>(hrow + 8);
        if (kt & 1) {
          MmaOp<T>::run(d2[0][nt], wf[0][kt], bf0, bf1);
          MmaOp<T>::run(d2[1][nt], wf[1][kt], bf0, bf1);
        } else {
          MmaOp<T>::run(d[0][nt], wf[0][kt], bf0, bf1);
          MmaOp<T>::run(d[1][nt], wf[1][kt], bf0, bf1);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[m][nt][i] += d2[m][nt][i];
    const int tt = dir ? L - 1 - step : step;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float i_ = lstm_act<T>(d[0][nt][e], false), f_ = lstm_act<T>(d[0][nt][2 + e], false);
        const float g_ = lstm_act<T>(d[1][nt][e], true), o_ = lstm_act<T>(d[1][nt][2 + e], false);
        c_state[nt][e] = fmaf(f_, c_state[nt][e], i_ * g_);
        float h;
        if constexpr (kBf16Math<T>) h = o_ * tanh_fast(c_state[nt][e]);
        else h = o_ * tanh_ex2(c_state[nt][e]);
        T hv;
        DT<T>::st(&hv, h);
        hsm[cur ^ 1][nt * 8 + 2 * cq + e][u] = hv;
        if (ok[nt][e]) op[nt][e][static_cast<long long>(tt) * o_ld] = hv;
      }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) { gA[nt][e] = gB[nt][e]; gB[nt][e] = gC[nt][e]; }
    __syncthreads();                                        // h(step) complete; everyone is done reading hsm[cur]
  }
}

template <typename T, int H>
int launch(const float* gx, const float* whh, void* out, int64_t o_bs, int o_ld, int B, int L, cudaStream_t s) {
  // LSTM_MMA (default 1): tensor-core recurrence for 16-bit activations, sixteen utterances per CTA
  if constexpr (sizeof(T) == 2) {
    if (tuning(TUNE_LSTM_MMA) != 0 && B >= 8) {
      // legacy mma.sync issues slowly on sm_100 (measured: 16 utterances per CTA = 200 mma per step = 2200 clk per step), so a
      // CTA takes 8 utterances while the four concurrent branch layers still fit one CTA per SM, 16 beyond
      if (ceil_div(B, 8) * 2 * 4 <= num_sms()) {
        bilstm_mma_kernel<T, H, 1><<<dim3(ceil_div(B, 8), 2), 4 * H, 0, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, B, L);
      } else {
        bilstm_mma_kernel<T, H, 2><<<dim3(ceil_div(B, 16), 2), 4 * H, 0, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, B, L);
      }
      SB_LAUNCH_OK();
      return 0;
    }
  }
  // LSTM_MULTI (default 0): four utterances per CTA, two rows of W_hh per thread (FMA-pipe bound: measured 0.205 ms against
  // 0.074 ms for the one-utterance form at B = 64, H = 80; kept as an A/B switch)
  constexpr int U = 4;
  if (tuning(TUNE_LSTM_MULTI) != 0 && B >= 4 * U) {
    constexpr size_t smem = sizeof(float) * (kBufsM * kChunkM * U * 4 * H + 2 * U * H) + kBufsM * sizeof(uint64_t);
    static DeviceFlags attr_set;
    SB_OPT_IN_SMEM(attr_set, (bilstm_multi_kernel<T, H, U>), smem);
    dim3 grid(ceil_div(B, U), 2);
    bilstm_multi_kernel<T, H, U><<<grid, 2 * H, smem, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, B, L);
    SB_LAUNCH_OK();
    return 0;
  }
  dim3 grid(B, 2);
  bilstm_kernel<T, H><<<grid, 4 * H, 0, s>>>(gx, whh, static_cast<T*>(out), o_bs, o_ld, L);
  SB_LAUNCH_OK();
  return 0;
}

}  // namespace
}  // namespace sb

extern "C" int styler_bilstm_layer_fwd(const float* gx, const float* whh, void* out, int64_t o_bstride, int32_t o_ld,
                                       int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream) {
  sb::TraceScope trace__("bilstm", stream, B, L, H, 0);
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
