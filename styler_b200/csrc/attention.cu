// Multi-head scaled-dot-product self-attention with key-padding mask (transformer/Modules.py:14-25,
// SubLayers.py:44-56 of the reference), never materialising the [T,T] attention matrix.
//
//  * attention_tc  : flash-style on tcgen05.  One CTA per (utterance, head, 128-query tile).
//        warp 0  TMA producer: Q tile once, then K tile [128 keys x 64] and V^T tile [64 x 128 keys] per step
//                (double-buffered for bf16), all 128B-swizzled K-major;
//        warp 1  MMA issuer: S = Q.K^T (128x128 fp32 in TMEM cols 0..127), then O_j = P.V (128x64, cols 128..191);
//        warps 2-5 softmax: thread r owns query row r; two passes over S in TMEM (row max, then exp / row sum),
//                P written to swizzled smem as the next MMA's A operand, running max/sum + O accumulated in
//                registers (online softmax), key tiles beyond lens[b] are skipped entirely.
//    1/temperature is folded into W_q at pack time (exact: 1/8 is a power of two).
//  * attention_simt: fp32 warp-per-query reference implementation (exact-fp32 mode and on-device cross-check).
//
// Bounding roofline: tensor pipe (4*T*64 FLOP per query row per head) with an SFU (exp) co-bound.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <stdlib.h>

namespace sb {
namespace {

// ----------------------------------------------------------------------------------------------- SIMT
template <typename T>
__global__ void __launch_bounds__(128) attention_simt_kernel(const T* qk, long long qk_bs, int qk_ld, const T* vt,
                                                             long long vt_bs, int vt_ld, const int64_t* lens, T* ctx,
                                                             long long ctx_bs, int ctx_ld, int Tlen, int H) {
  extern __shared__ float psm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int tq = blockIdx.x * 4 + warp;
  if (tq >= Tlen) return;
  float* p = psm + static_cast<size_t>(warp) * Tlen;
  int len = lens != nullptr ? static_cast<int>(lens[b]) : Tlen;
  len = len < Tlen ? len : Tlen;
  const int D = H * 64;
  float q[64];
  const T* qrow = qk + b * qk_bs + static_cast<long long>(tq) * qk_ld + h * 64;
#pragma unroll
  for (int d = 0; d < 64; ++d) q[d] = DT<T>::ld(qrow + d);
  float mx = -INFINITY;
  for (int k = lane; k < len; k += 32) {
    const T* krow = qk + b * qk_bs + static_cast<long long>(k) * qk_ld + D + h * 64;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s = fmaf(q[d], DT<T>::ld(krow + d), s);
    p[k] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < len; k += 32) {
    const float e = expf(p[k] - mx);
    p[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  T* orow = ctx + b * ctx_bs + static_cast<long long>(tq) * ctx_ld + h * 64;
  for (int d = lane; d < 64; d += 32) {
    float acc = 0.f;
    if (vt != nullptr) {
      const T* vrow = vt + b * vt_bs + static_cast<long long>(h * 64 + d) * vt_ld;
      for (int k = 0; k < len; ++k) acc = fmaf(p[k], DT<T>::ld(vrow + k), acc);
    } else {   // V stored row-major next to Q and K: qk[b][k][2*D + h*64 + d]
      const T* vcol = qk + b * qk_bs + 2 * D + h * 64 + d;
      for (int k = 0; k < len; ++k) acc = fmaf(p[k], DT<T>::ld(vcol + static_cast<long long>(k) * qk_ld), acc);
    }
    DT<T>::st(orow + d, len > 0 ? acc / sum : 0.f);
  }
}

// ----------------------------------------------------------------------------------------------- tcgen05
constexpr int kQ = 128;    // queries per CTA
constexpr int kKV = 128;   // keys per step
constexpr int kThreadsTc = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two per TMEM lane quarter)

__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2, flush-to-zero; inputs are <= 0 here
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic fit of 2^f (max relative error
// 7.7e-5, well inside the bf16 rounding of P), exponent patched in with one shift-add.  Used for a quarter of the scores
// of the bf16 kernel, where the 16 MUFU lanes per SM - not the tensor pipe - set the softmax rate.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float tt = x + 12582912.f;                   // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (tt - 12582912.f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(tt) << 23));
}

template <typename T> struct AttnCfg {
  static constexpr int es = sizeof(T);
  static constexpr int bke = 128 / es;              // elements per 128-byte slice
  static constexpr int qk_slices = 64 / bke;        // slices covering the 64-wide head dim  (bf16 1, fp32 2)
  static constexpr int pv_slices = kKV / bke;       // slices covering 128 keys              (bf16 2, fp32 4)
  static constexpr int q_bytes = qk_slices * kQ * 128;
  static constexpr int k_bytes = qk_slices * kKV * 128;
  static constexpr int v_bytes = pv_slices * 64 * 128;
  static constexpr int p_bytes = pv_slices * kQ * 128;
  static constexpr int kv_stages = es == 2 ? 2 : 1;
  static constexpr int umma_k = 32 / es;
  static constexpr size_t smem = q_bytes + kv_stages * (k_bytes + v_bytes) + p_bytes + 512 + 256;  // bf16: 2 CTAs/SM
};

template <typename T>
__global__ void __launch_bounds__(kThreadsTc) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK,
                                                                  const __grid_constant__ CUtensorMap tmVT,
                                                                  const int64_t* __restrict__ lens, T* __restrict__ ctx,
                                                                  long long ctx_bs, int ctx_ld, int Tlen, int H,
                                                                  int q_tiles, int v_mn) {
  using C = AttnCfg<T>;
  constexpr bool kTf32 = C::es == 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + C::q_bytes;                     // per stage: K then V^T
  uint8_t* sP = sKV + C::kv_stages * (C::k_bytes + C::v_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::p_bytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                       // [2]
  uint64_t* kv_empty = bars + 3;                      // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  if (threadIdx.x == 0 && static_cast<size_t>(reinterpret_cast<uint8_t*>(tmem_slot + 1) - smem_raw) > C::smem) {
    printf("styler_b200: attention smem carve-up overflows the allocation (base misaligned)\n");
    __trap();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int h = (blockIdx.x / q_tiles) % H;
  const int b = blockIdx.x / (q_tiles * H);
  const int t0 = qt * kQ;
  int len = lens != nullptr ? static_cast<int>(lens[b]) : Tlen;
  len = len < Tlen ? len : Tlen;
  const int nkt = (len + kKV - 1) / kKV;
  const int D = H * 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
  pdl_launch_dependents();
  pdl_grid_dependency_wait();     // the QKV projection (previous kernel) is complete and visible from here on

  if (warp == 0) {
    if (lane == 0 && nkt > 0) {
      mbar_arrive_expect_tx(q_full, C::q_bytes);
      for (int sl = 0; sl < C::qk_slices; ++sl)
        tma_load_3d(sQ + sl * kQ * 128, &tmQK, q_full, h * 64 + sl * C::bke, t0, b);
      for (int j = 0; j < nkt; ++j) {
        const int s = j % C::kv_stages;
        const uint32_t ph = (j / C::kv_stages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], C::k_bytes + C::v_bytes);
        uint8_t* sK = sKV + s * (C::k_bytes + C::v_bytes);
        uint8_t* sV = sK + C::k_bytes;
        for (int sl = 0; sl < C::qk_slices; ++sl)
          tma_load_3d(sK + sl * kKV * 128, &tmQK, &kv_full[s], D + h * 64 + sl * C::bke, j * kKV, b);
        if (v_mn) {   // V row-major in the qkv tensor: [128 keys x 128-byte span of d] boxes, consumed as an MN-major B operand
          for (int sl = 0; sl < C::qk_slices; ++sl)
            tma_load_3d(sV + sl * kKV * 128, &tmQK, &kv_full[s], 2 * D + h * 64 + sl * C::bke, j * kKV, b);
        } else {
          for (int sl = 0; sl < C::pv_slices; ++sl)
            tma_load_3d(sV + sl * 64 * 128, &tmVT, &kv_full[s], j * kKV + sl * C::bke, h * 64, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkt > 0) {
      const uint32_t fmt = kTf32 ? UMMA_FMT_TF32 : UMMA_FMT_BF16;
      const uint32_t idesc_s = umma_idesc(fmt, kQ, kKV);
      const uint32_t idesc_o = umma_idesc(fmt, kQ, 64, v_mn ? 1u : 0u);
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkt; ++j) {
        const int s = j % C::kv_stages;
        mbar_wait(&kv_full[s], (j / C::kv_stages) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ);
        const uint32_t k_addr = smem_u32(sKV + s * (C::k_bytes + C::v_bytes));
        const uint32_t v_addr = k_addr + C::k_bytes;
        const uint32_t p_addr = smem_u32(sP);
#pragma unroll
        for (int kk = 0; kk < 64 / C::umma_k; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          umma_ss<kTf32>(tmem_S, umma_desc_k_sw128(q_addr + sl * kQ * 128 + off),
                         umma_desc_k_sw128(k_addr + sl * kKV * 128 + off), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kKV / C::umma_k; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          const uint64_t bdesc = v_mn ? umma_desc_mn_sw128(v_addr + kk * C::umma_k * 128, kKV * 128, 1024)
                                      : umma_desc_k_sw128(v_addr + sl * 64 * 128 + off);
          umma_ss<kTf32>(tmem_O, umma_desc_k_sw128(p_addr + sl * kQ * 128 + off), bdesc, idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
    }
    __syncwarp();
  } else {
    // 8 softmax warps: warp w may only touch TMEM lanes 32*(w%4)..+31, so each row (lane) is served by two threads:
    // both scan all 128 score columns for the row max (cheap, keeps the running max identical in both), then each
    // exponentiates / packs its own 64 key columns and accumulates its own 32 of the 64 output columns.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;                    // column half handled by this thread
    const int r = q * 32 + lane;
    const int t = t0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < nkt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * kKV;
      const bool full_tile = kbase + kKV <= len;           // only the last key tile needs the padding mask
      float mx = m_run;
      for (int c = 0; c < kKV; c += 32) {
        uint32_t ra[16], rb[16];
        tmem_ld16(tmem_S + lane_off + c, ra);
        tmem_ld16(tmem_S + lane_off + c + 16, rb);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (kbase + c + i < len) mx = fmaxf(mx, __uint_as_float(ra[i]));
            if (kbase + c + 16 + i < len) mx = fmaxf(mx, __uint_as_float(rb[i]));
          }
        }
      }
      const float alpha = m_run == -INFINITY ? 0.f : fast_exp2((m_run - mx) * kLog2e);
      l_run *= alpha;
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] *= alpha;
      const float mxl = mx * kLog2e;
      for (int cc = 0; cc < 64; cc += 16) {
        const int c = hf * 64 + cc;
        uint32_t raw[16];
        tmem_ld16(tmem_S + lane_off + c, raw);
        tmem_ld_wait();
        float pv[16];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            pv[i] = fast_exp2(fmaf(__uint_as_float(raw[i]), kLog2e, -mxl));
            l_run += pv[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e = kbase + c + i < len ? fast_exp2(fmaf(__uint_as_float(raw[i]), kLog2e, -mxl)) : 0.f;
            pv[i] = e;
            l_run += e;
          }
        }
        // P[r][c..c+15] -> K-major 128B-swizzled smem (A operand of the PV MMA)
        const int byte0 = c * C::es;                       // byte offset of key c within the row
        uint8_t* slice = p_row + (byte0 / 128) * (kQ * 128);
        const int ch0 = (byte0 % 128) / 16;
        if constexpr (C::es == 2) {
          float a0[8], a1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a0[i] = pv[i]; a1[i] = pv[8 + i]; }
          store8(reinterpret_cast<__nv_bfloat16*>(slice + ((ch0 ^ sw) * 16)), a0);
          store8(reinterpret_cast<__nv_bfloat16*>(slice + (((ch0 + 1) ^ sw) * 16)), a1);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(slice + (((ch0 + g) ^ sw) * 16)) =
                make_float4(pv[4 * g], pv[4 * g + 1], pv[4 * g + 2], pv[4 * g + 3]);
        }
      }
      m_run = mx;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t ra[16], rb[16];
        tmem_ld16(tmem_O + lane_off + hf * 32, ra);
        tmem_ld16(tmem_O + lane_off + hf * 32 + 16, rb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i] += __uint_as_float(ra[i]); o[16 + i] += __uint_as_float(rb[i]); }
      }
    }
    // combine the two partial row sums through smem (the Q tile is dead once the last S MMA has completed)
    float* s_l = reinterpret_cast<float*>(sQ);
    s_l[hf * kQ + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = s_l[r] + s_l[kQ + r];
    if (t < Tlen) {
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      T* orow = ctx + b * ctx_bs + static_cast<long long>(t) * ctx_ld + h * 64 + hf * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[c + i] * inv;
        store8(orow + c, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}


// ----------------------------------------------------------------------------------------------- tcgen05, v2
// Same tiling and operand layouts as attention_tc_kernel, restructured so that the per-tile dependency chain is
//   S MMA -> softmax -> (PV MMA + next S MMA back to back)
// instead of S MMA -> softmax -> PV MMA -> O read-back -> next S MMA:
//   * the 128 keys of a tile are split into two halves of 64, each with its OWN running max / row sum and its OWN
//     accumulator in TMEM (O_a = cols 128..191, O_b = cols 192..255).  Thread (row r, half hf) therefore never has to
//     exchange a row max with the thread holding the other 64 columns (they sit in different warps because of the
//     TMEM lane-quarter rule); the two halves are merged once, after the last key tile.
//   * O stays in TMEM for the whole key loop (tcgen05.mma accumulates across tiles); the running max is only raised -
//     and O / l rescaled, by a TMEM ld-mul-st of the thread's own 64 columns - when the tile max exceeds the stale
//     max by more than 2^8 (warp-uniform vote, so tcgen05.ld/st stay warp-collective).  exp2 arguments are then
//     bounded by 8, P <= 256 (exact range for bf16 / tf32 operands), and the result is mathematically unchanged.
//   * the MMA thread issues PV(j) and S(j+1) back to back; one tcgen05.commit (s_full) covers both, so observing
//     S(j+1) also tells the softmax threads that PV(j) has drained: sP may be overwritten and O may be rescaled.
template <typename T, int POLY, bool SPIN>
__global__ void __launch_bounds__(kThreadsTc, 2) attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQK,
                                                                      const __grid_constant__ CUtensorMap tmVT,
                                                                      const int64_t* __restrict__ lens, T* __restrict__ ctx,
                                                                      long long ctx_bs, int ctx_ld, int Tlen, int H,
                                                                      int q_tiles, int v_mn) {
  using C = AttnCfg<T>;
  constexpr bool kTf32 = C::es == 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + C::q_bytes;                     // per stage: K then V
  uint8_t* sP = sKV + C::kv_stages * (C::k_bytes + C::v_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::p_bytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                       // [2]
  uint64_t* kv_empty = bars + 3;                      // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  if (threadIdx.x == 0 && static_cast<size_t>(reinterpret_cast<uint8_t*>(tmem_slot + 1) - smem_raw) > C::smem) {
    printf("styler_b200: attention smem carve-up overflows the allocation (base misaligned)\n");
    __trap();
  }

  auto bwait = [](uint64_t* bar, uint32_t parity) {
    if constexpr (SPIN) mbar_wait_spin(bar, parity);
    else mbar_wait(bar, parity);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int h = (blockIdx.x / q_tiles) % H;
  const int b = blockIdx.x / (q_tiles * H);
  const int t0 = qt * kQ;
  int len = lens != nullptr ? static_cast<int>(lens[b]) : Tlen;
  len = len < Tlen ? len : Tlen;
  const int nkt = (len + kKV - 1) / kKV;
  const int D = H * 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;    // O_a = tmem_O, O_b = tmem_O + 64
  pdl_launch_dependents();
  pdl_grid_dependency_wait();     // the QKV projection (previous kernel) is complete and visible from here on

  if (warp == 0) {
    if (lane == 0 && nkt > 0) {
      mbar_arrive_expect_tx(q_full, C::q_bytes);
      for (int sl = 0; sl < C::qk_slices; ++sl)
        tma_load_3d(sQ + sl * kQ * 128, &tmQK, q_full, h * 64 + sl * C::bke, t0, b);
      for (int j = 0; j < nkt; ++j) {
        const int s = j % C::kv_stages;
        const uint32_t ph = (j / C::kv_stages) & 1;
        bwait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], C::k_bytes + C::v_bytes);
        uint8_t* sK = sKV + s * (C::k_bytes + C::v_bytes);
        uint8_t* sV = sK + C::k_bytes;
        for (int sl = 0; sl < C::qk_slices; ++sl)
          tma_load_3d(sK + sl * kKV * 128, &tmQK, &kv_full[s], D + h * 64 + sl * C::bke, j * kKV, b);
        if (v_mn) {   // V row-major in the qkv tensor: [128 keys x 128-byte span of d] boxes, consumed as an MN-major B operand
          for (int sl = 0; sl < C::qk_slices; ++sl)
            tma_load_3d(sV + sl * kKV * 128, &tmQK, &kv_full[s], 2 * D + h * 64 + sl * C::bke, j * kKV, b);
        } else {
          for (int sl = 0; sl < C::pv_slices; ++sl)
            tma_load_3d(sV + sl * 64 * 128, &tmVT, &kv_full[s], j * kKV + sl * C::bke, h * 64, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkt > 0) {
      const uint32_t fmt = kTf32 ? UMMA_FMT_TF32 : UMMA_FMT_BF16;
      const uint32_t idesc_s = umma_idesc(fmt, kQ, kKV);
      const uint32_t idesc_o = umma_idesc(fmt, kQ, 64, v_mn ? 1u : 0u);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      constexpr int kPvSteps = kKV / C::umma_k;          // K steps of the PV product (bf16 8, tf32 16)
      auto issue_s = [&](int j) {
        const int s = j % C::kv_stages;
        bwait(&kv_full[s], (j / C::kv_stages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * (C::k_bytes + C::v_bytes));
#pragma unroll
        for (int kk = 0; kk < 64 / C::umma_k; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          umma_ss<kTf32>(tmem_S, umma_desc_k_sw128(q_addr + sl * kQ * 128 + off),
                         umma_desc_k_sw128(k_addr + sl * kKV * 128 + off), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      bwait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkt; ++j) {
        const int s = j % C::kv_stages;
        const uint32_t v_addr = smem_u32(sKV + s * (C::k_bytes + C::v_bytes)) + C::k_bytes;
        bwait(p_ready, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kPvSteps; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          const int half = kk / (kPvSteps / 2);           // keys 0..63 -> O_a, 64..127 -> O_b
          const uint64_t bdesc = v_mn ? umma_desc_mn_sw128(v_addr + kk * C::umma_k * 128, kKV * 128, 1024)
                                      : umma_desc_k_sw128(v_addr + sl * 64 * 128 + off);
          umma_ss<kTf32>(tmem_O + half * 64, umma_desc_k_sw128(p_addr + sl * kQ * 128 + off), bdesc, idesc_o,
                         (j != 0 || (kk % (kPvSteps / 2)) != 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        if (j + 1 < nkt) issue_s(j + 1);                 // its commit also covers PV(j)
        else umma_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                            // TMEM lane quarter this warp may touch
    const int hf = (warp - 2) >> 2;                    // key half (64 of the tile's 128 score columns) owned by this thread
    const int r = q * 32 + lane;
    const int t = t0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_S + lane_off + hf * 64;
    const uint32_t tO = tmem_O + lane_off + hf * 64;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescale = 8.0f;                   // raise the running max only when it is stale by > 2^8
    float m_l = -INFINITY;                             // running (stale) max, log2 domain
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;      // row sum of this key half (4 chains)
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < nkt; ++j) {
      bwait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * kKV + hf * 64;
      const bool full_tile = kbase + 64 <= len;        // only the last key tile needs the padding mask
      uint32_t sv[64];
      {
        uint32_t (*sv16)[16] = reinterpret_cast<uint32_t (*)[16]>(sv);
        tmem_ld16(tS, sv16[0]);
        tmem_ld16(tS + 16, sv16[1]);
        tmem_ld16(tS + 32, sv16[2]);
        tmem_ld16(tS + 48, sv16[3]);
        tmem_ld_wait();
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (kbase + i < len) mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
      }
      const float mx_l = fmaxf(mx0, mx1) * kLog2e;
      const bool raise = mx_l > m_l + kRescale;        // false when this half has no valid key yet (-inf > -inf)
      if (__any_sync(0xffffffffu, raise)) {
        const float m_new = raise ? mx_l : m_l;
        const float alpha = m_l == -INFINITY ? 0.f : fast_exp2(m_l - m_new);   // 1 for the lanes that keep their max
        l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
        if (j > 0) {                                   // PV(j-1) has drained (s_full covers it): rescale O in place
#pragma unroll 1
          for (int c = 0; c < 64; c += 16) {           // rare path: keep its register footprint small
            uint32_t ra[16];
            tmem_ld16(tO + c, ra);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ra[i] = __float_as_uint(__uint_as_float(ra[i]) * alpha);
            tmem_st16(tO + c, ra);
          }
          tmem_st_wait();
        }
        m_l = m_new;
      }
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        float pv[16];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = fmaf(__uint_as_float(sv[cc + i]), kLog2e, -m_l);
            pv[i] = (POLY != 0 && (i & 3) == 3) ? poly_exp2(x) : fast_exp2(x);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pv[i] = kbase + cc + i < len ? fast_exp2(fmaf(__uint_as_float(sv[cc + i]), kLog2e, -m_l)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) { l0 += pv[i]; l1 += pv[i + 1]; l2 += pv[i + 2]; l3 += pv[i + 3]; }
        // P[r][c..c+15] -> K-major 128B-swizzled smem (A operand of the PV MMA)
        const int byte0 = (hf * 64 + cc) * C::es;          // byte offset of key c within the row
        uint8_t* slice = p_row + (byte0 / 128) * (kQ * 128);
        const int ch0 = (byte0 % 128) / 16;
        if constexpr (C::es == 2) {
          float a0[8], a1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a0[i] = pv[i]; a1[i] = pv[8 + i]; }
          store8(reinterpret_cast<__nv_bfloat16*>(slice + ((ch0 ^ sw) * 16)), a0);
          store8(reinterpret_cast<__nv_bfloat16*>(slice + (((ch0 + 1) ^ sw) * 16)), a1);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(slice + (((ch0 + g) ^ sw) * 16)) =
                make_float4(pv[4 * g], pv[4 * g + 1], pv[4 * g + 2], pv[4 * g + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // merge the two key halves: out = (O_a w_a + O_b w_b) / (l_a w_a + l_b w_b), w_x = 2^(m_x - max(m_a, m_b)).
    // The exchange goes through smem (the Q tile is dead once the last S MMA has completed).
    float2* s_ml = reinterpret_cast<float2*>(sQ);
    if (nkt > 0) {
      bwait(o_full, 0);
      tc_fence_after();
    }
    s_ml[hf * kQ + r] = make_float2(m_l, (l0 + l1) + (l2 + l3));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 ma = s_ml[r], mb = s_ml[kQ + r];
    const float m_tot = fmaxf(ma.x, mb.x);
    const float wa = ma.x == -INFINITY ? 0.f : fast_exp2(ma.x - m_tot);
    const float wb = mb.x == -INFINITY ? 0.f : fast_exp2(mb.x - m_tot);
    const float l_tot = ma.y * wa + mb.y * wb;
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const float fa = wa * inv, fb = wb * inv;
    T* orow = ctx + b * ctx_bs + static_cast<long long>(t) * ctx_ld + h * 64 + hf * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 16) {
      float o[16];
      if (nkt > 0) {
        uint32_t ra[16], rb[16];
        tmem_ld16(tmem_O + lane_off + hf * 32 + c, ra);
        tmem_ld16(tmem_O + lane_off + 64 + hf * 32 + c, rb);
        tmem_ld_wait();
        // a weight of exactly 0 must win over whatever the (zero-probability) accumulator holds
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o[i] = (wa != 0.f ? __uint_as_float(ra[i]) * fa : 0.f) + (wb != 0.f ? __uint_as_float(rb[i]) * fb : 0.f);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.f;
      }
      if (t < Tlen) {
        float v0[8], v1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v0[i] = o[i]; v1[i] = o[8 + i]; }
        store8(orow + c, v0);
        store8(orow + c + 8, v1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

bool attn_poly() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STYLER_ATTN_POLY"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
bool attn_spin() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STYLER_ATTN_SPIN"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
bool attn_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STYLER_ATTN_V1"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

template <typename T>
int launch_tc(const void* qk, int64_t qk_bs, int qk_ld, const void* vt, int64_t vt_bs, int vt_ld,
              const int64_t* lens, void* ctx, int64_t ctx_bs, int ctx_ld, int B, int Tlen, int H, cudaStream_t s) {
  using C = AttnCfg<T>;
  const int D = H * 64;
  const int v_mn = vt == nullptr ? 1 : 0;
  CUtensorMap tmQK, tmVT;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>((v_mn ? 3 : 2) * D), static_cast<uint64_t>(Tlen), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(qk_ld) * C::es, static_cast<uint64_t>(qk_bs) * C::es};
    const uint32_t box[3] = {static_cast<uint32_t>(C::bke), 128, 1};
    int rc = make_tmap(&tmQK, qk, C::es == 2 ? 1 : 0, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  if (v_mn) {
    tmVT = tmQK;
  } else {
    const uint64_t dims[3] = {static_cast<uint64_t>(Tlen), static_cast<uint64_t>(D), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(vt_ld) * C::es, static_cast<uint64_t>(vt_bs) * C::es};
    const uint32_t box[3] = {static_cast<uint32_t>(C::bke), 64, 1};
    int rc = make_tmap(&tmVT, vt, C::es == 2 ? 1 : 0, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  auto kern = attn_v1() ? attention_tc_kernel<T>
              : (sizeof(T) == 2 && attn_poly()) ? (attn_spin() ? attention_tc2_kernel<T, sizeof(T) == 2 ? 1 : 0, true>
                                                               : attention_tc2_kernel<T, sizeof(T) == 2 ? 1 : 0, false>)
                                                : (attn_spin() ? attention_tc2_kernel<T, 0, true> : attention_tc2_kernel<T, 0, false>);
  static decltype(kern) attr_set_for = nullptr;
  if (attr_set_for != kern) {
    SB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::smem)));
    attr_set_for = kern;
  }
  const int q_tiles = ceil_div(Tlen, kQ);
  SB_CUDA_OK(launch_pdl(kern, dim3(B * H * q_tiles), dim3(kThreadsTc), C::smem, s, tmQK, tmVT, lens, static_cast<T*>(ctx),
                        static_cast<long long>(ctx_bs), ctx_ld, Tlen, H, q_tiles, v_mn));
  SB_LAUNCH_OK();
  return 0;
}

}  // namespace

int attention_simt(const void* qk, int64_t qk_bs, int qk_ld, const void* vt, int64_t vt_bs, int vt_ld,
                   const int64_t* lens, void* ctx, int64_t ctx_bs, int ctx_ld, int B, int T, int H, int dtype,
                   cudaStream_t s) {
  const size_t smem = static_cast<size_t>(4) * T * sizeof(float);
  SB_REQUIRE(smem <= 200 * 1024, "attention_simt: T=%d too long", T);
  dim3 grid(ceil_div(T, 4), B * H);
  SB_DISPATCH_DTYPE(dtype, TT, {
    auto kern = attention_simt_kernel<TT>;
    if (smem > 48 * 1024)
      SB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, 128, smem, s>>>(static_cast<const TT*>(qk), qk_bs, qk_ld, static_cast<const TT*>(vt), vt_bs, vt_ld,
                                 lens, static_cast<TT*>(ctx), ctx_bs, ctx_ld, T, H);
  });
  SB_LAUNCH_OK();
  return 0;
}

int attention_tc(const void* qk, int64_t qk_bs, int qk_ld, const void* vt, int64_t vt_bs, int vt_ld,
                 const int64_t* lens, void* ctx, int64_t ctx_bs, int ctx_ld, int B, int T, int H, int dtype,
                 cudaStream_t s) {
  const int es = dtype == STYLER_BF16 ? 2 : 4;
  SB_REQUIRE(vt != nullptr || dtype == STYLER_BF16,
             "attention_tc: row-major V (vt == NULL) is only validated for bf16; pass V^T for fp32/tf32");
  SB_REQUIRE((static_cast<int64_t>(ctx_ld) * es) % 16 == 0 && (ctx_bs * es) % 16 == 0 &&
                 (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
             "attention_tc: ctx must be 16-byte aligned/strided");
  if (dtype == STYLER_BF16)
    return launch_tc<__nv_bfloat16>(qk, qk_bs, qk_ld, vt, vt_bs, vt_ld, lens, ctx, ctx_bs, ctx_ld, B, T, H, s);
  return launch_tc<float>(qk, qk_bs, qk_ld, vt, vt_bs, vt_ld, lens, ctx, ctx_bs, ctx_ld, B, T, H, s);
}

}  // namespace sb

extern "C" int styler_attention_fwd(const void* qk, int64_t qk_bstride, int32_t qk_ld, const void* vt,
                                    int64_t vt_bstride, int32_t vt_ld, const int64_t* lens, void* ctx,
                                    int64_t ctx_bstride, int32_t ctx_ld, int32_t B, int32_t T, int32_t H,
                                    int32_t dtype, int32_t impl, void* stream) {
  using namespace sb;
  SB_REQUIRE(qk != nullptr && ctx != nullptr, "attention: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && H > 0, "attention: bad shape B=%d T=%d H=%d", B, T, H);
  SB_REQUIRE(dtype == STYLER_F32 || dtype == STYLER_BF16, "attention: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == STYLER_IMPL_AUTO) impl = STYLER_IMPL_TC;
  if (impl == STYLER_IMPL_TC)
    return attention_tc(qk, qk_bstride, qk_ld, vt, vt_bstride, vt_ld, lens, ctx, ctx_bstride, ctx_ld, B, T, H, dtype, s);
  SB_REQUIRE(impl == STYLER_IMPL_SIMT, "attention: bad impl %d", impl);
  return attention_simt(qk, qk_bstride, qk_ld, vt, vt_bstride, vt_ld, lens, ctx, ctx_bstride, ctx_ld, B, T, H, dtype, s);
}
