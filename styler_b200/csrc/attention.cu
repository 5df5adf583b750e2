// Multi-head scaled-dot-product self-attention with key-padding mask (transformer/Modules.py:14-25,
// SubLayers.py:44-56 of the reference), never materialising the [T,T] attention matrix.
//
//  * attention_tc  : flash-style on tcgen05.  One CTA per (utterance, head, 128-query tile), two CTAs per SM.
//        warp 0  TMA producer: Q tile once, then per key tile K [128 keys x 64] and V (row-major from the fused QKV
//                buffer, consumed as an MN-major B operand; or V^T for fp32/tf32 operands), double-buffered for bf16;
//        warp 1  MMA issuer: S = Q.K^T (128x128 fp32, TMEM cols 0..127); O_a += P[:, :64].V[:64], O_b += P[:, 64:].V[64:]
//                (cols 128..191 / 192..255) accumulated in TMEM over the whole key loop;
//        warps 2-9 softmax: thread (row r, key half hf) holds its 64 scores in registers, keeps its OWN running max and
//                row sum (the two halves of a row live in different warps - TMEM lane-quarter rule - and are merged once
//                after the last key tile), raises the max lazily (only when stale by > 2^8; O and l are then rescaled
//                in place in TMEM), writes P to swizzled smem as the next MMA's A operand; key tiles beyond lens[b] are
//                skipped entirely.
//    1/temperature is folded into W_q at pack time (exact: 1/8 is a power of two).
//  * attention_simt: fp32 warp-per-query reference implementation (exact-fp32 mode and on-device cross-check).
//
// Bounding roofline: tensor pipe (4*T*64 FLOP per query row per head) with an SFU (exp) co-bound.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

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
// 2^x for a PAIR of scores on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x) taken from the low mantissa bits of
// x + 1.5 * 2^23, 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (max. relative error 7.5e-5: below the rounding of P to
// bf16 / fp16), the exponent n added into the result's exponent field with one integer multiply-add.  The softmax warps are
// bound by the XU pipe (MUFU.EX2: 4 lanes per clock per SM sub-partition, 64 exponentials per thread per key tile); sending
// kPolyOf8 of every 8 score pairs through this form moves that share of the work to pipes that were half idle.
__device__ __forceinline__ void exp2_poly_pair(float2 x, uint32_t& o0, uint32_t& o1) {
  constexpr float kMagic = 12582912.f;                 // 1.5 * 2^23
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 xf = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 rn = __fadd2_rn(xf, make_float2(-kMagic, -kMagic));
  const float2 f = __ffma2_rn(rn, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(make_float2(0.05517027899622917f, 0.05517027899622917f), f, make_float2(0.2426076978445053f, 0.2426076978445053f));
  p = __ffma2_rn(p, f, make_float2(0.693260908126831f, 0.693260908126831f));
  p = __ffma2_rn(p, f, make_float2(0.9999282956123352f, 0.9999282956123352f));
  o0 = __float_as_uint(p.x) + (__float_as_uint(xf.x) << 23);
  o1 = __float_as_uint(p.y) + (__float_as_uint(xf.y) << 23);
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
  static constexpr size_t smem = q_bytes + kv_stages * (k_bytes + v_bytes) + p_bytes + 512 + 256;  // bf16: 2 CTAs/SM (113.25 KB each)
};

// Pipeline (per key tile j), arranged so that the tensor pipe is off the softmax critical path.  ncu on the first
// version of this kernel (S -> softmax -> PV -> O read-back per tile) showed the softmax warps 28 % of their time in the
// wait for the next S tile and 36 % in the MUFU-bound exponentials, the tensor pipe 19 % busy; TMEM reads are not a
// limit (tools/tmem_bench.cu: 885 B/clk/SM).
//   * a softmax warp releases the S columns as soon as the scores are in its registers (s_taken); the MMA thread issues
//     the NEXT tile's S right then, so it runs under the exponentials of this tile;
//   * PV(j) is issued when P(j) lands; its completion (pv_done) is only waited for by the next tile just before it overwrites
//     sP / rescales O - after its own TMEM load, row max and all 64 exponentials;
//   * O never leaves TMEM inside the loop.
// With a single K/V stage (fp32/tf32 operands) K(j+1) cannot be resident before PV(j) has drained, so that
// instantiation issues the next S after PV(j) instead.  Measured (B=64, T=1024, 4 heads, bf16): 0.176 -> 0.145 ms.
//
// Round 2: PERSISTENT work loop.  The round-1 kernel ran one (utterance, head, query tile) per CTA: 8 key tiles of work
// behind a cold prologue (TMEM alloc, barrier init, Q/K/V first touch ~2-5 k clk) and a drain (merge, exit barrier): the
// source-level ncu samples put 14 % of the softmax warps' time on the FIRST s_full wait of a CTA and 8 % on the exit
// (profiles/ncu_attn_r2.md).  Now a CTA walks work items item = blockIdx.x + i * gridDim.x with the K/V ring, the S / P /
// PV hand-shakes and the TMEM accumulators running across item boundaries as one flat sequence of key tiles: the Q tile
// of the next item is fetched as soon as the last S MMA of this item has been issued, and the first S of the next item
// runs under the exponentials of this item's last tile.  (grid = number of items gives back the one-item form.)
struct AttnItem {
  int b, h, t0, len, nkt;
};

template <typename T, int kPolyOf8>
__global__ void __launch_bounds__(kThreadsTc, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK,
                                                                      const __grid_constant__ CUtensorMap tmVT,
                                                                      const int64_t* __restrict__ lens, T* __restrict__ ctx,
                                                                      long long ctx_bs, int ctx_ld, int Tlen, int H,
                                                                      int q_tiles, int v_mn, int total_items) {
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
  uint64_t* s_full = bars + 5;                        // MMA -> softmax: S complete in TMEM
  uint64_t* s_taken = bars + 6;                       // softmax -> MMA: S is in registers, its TMEM columns may be overwritten
  uint64_t* p_ready = bars + 7;                       // softmax -> MMA: P is in smem (and O rescaled if needed)
  uint64_t* pv_done = bars + 8;                       // MMA -> softmax: PV drained: sP may be rewritten, O may be rescaled / read
  uint64_t* q_empty = bars + 9;                       // MMA -> producer: the last S of an item has read sQ
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  if (threadIdx.x == 0 && static_cast<size_t>(reinterpret_cast<uint8_t*>(tmem_slot + 1) - smem_raw) > C::smem) {
    printf("styler_b200: attention smem carve-up overflows the allocation (base misaligned)\n");
    __trap();
  }

  auto bwait = [](uint64_t* bar, uint32_t parity) { mbar_wait_spin(bar, parity); };
  auto bwait_long = [](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };   // producer / MMA threads: parked in hardware
  constexpr bool kEarlyS = C::kv_stages >= 2;         // K(j+1) can be resident while V(j) is still needed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * 64;
  auto item_at = [&](int item) {                      // every role decodes the same item sequence
    AttnItem it;
    const int qt = item % q_tiles;
    it.h = (item / q_tiles) % H;
    it.b = item / (q_tiles * H);
    it.t0 = qt * kQ;
    int len = lens != nullptr ? static_cast<int>(lens[it.b]) : Tlen;
    it.len = len < Tlen ? len : Tlen;
    it.nkt = (it.len + kKV - 1) / kKV;
    return it;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_taken, 8);                            // one elected arrival per softmax warp
    mbar_init(p_ready, 8);
    mbar_init(pv_done, 1);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int qi = 0, g = 0;                               // items with key tiles seen so far; key tiles loaded so far (ring position)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const AttnItem it = item_at(item);
        if (it.nkt == 0) continue;
        if (qi > 0) bwait_long(q_empty, (qi - 1) & 1);  // the previous item's last S MMA has consumed sQ
        mbar_arrive_expect_tx(q_full, C::q_bytes);
        for (int sl = 0; sl < C::qk_slices; ++sl)
          tma_load_3d(sQ + sl * kQ * 128, &tmQK, q_full, it.h * 64 + sl * C::bke, it.t0, it.b);
        ++qi;
        for (int j = 0; j < it.nkt; ++j, ++g) {
          const int s = g % C::kv_stages;
          const uint32_t ph = (g / C::kv_stages) & 1;
          bwait_long(&kv_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], C::k_bytes + C::v_bytes);
          uint8_t* sK = sKV + s * (C::k_bytes + C::v_bytes);
          uint8_t* sV = sK + C::k_bytes;
          for (int sl = 0; sl < C::qk_slices; ++sl)
            tma_load_3d(sK + sl * kKV * 128, &tmQK, &kv_full[s], D + it.h * 64 + sl * C::bke, j * kKV, it.b);
          if (v_mn) {   // V row-major in the qkv tensor: [128 keys x 128-byte span of d] boxes, consumed as an MN-major B operand
            for (int sl = 0; sl < C::qk_slices; ++sl)
              tma_load_3d(sV + sl * kKV * 128, &tmQK, &kv_full[s], 2 * D + it.h * 64 + sl * C::bke, j * kKV, it.b);
          } else {
            for (int sl = 0; sl < C::pv_slices; ++sl)
              tma_load_3d(sV + sl * 64 * 128, &tmVT, &kv_full[s], j * kKV + sl * C::bke, it.h * 64, it.b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one flat sequence of key tiles
    if (lane == 0) {
      const uint32_t fmt = umma_fmt_of<T>();
      const uint32_t idesc_s = umma_idesc(fmt, kQ, kKV);
      const uint32_t idesc_o = umma_idesc(fmt, kQ, 64, v_mn ? 1u : 0u);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      constexpr int kPvSteps = kKV / C::umma_k;          // K steps of the PV product (bf16 8, tf32 16)
      // cursor over (item, key tile); `g` = key tiles issued so far (ring position), `qi` = items started
      int item_c = blockIdx.x, nkt_c = 0, j_c = 0;       // current tile (whose PV is next)
      auto next_item_with_tiles = [&](int item, int& nkt) {
        for (; item < total_items; item += gridDim.x) {
          nkt = item_at(item).nkt;
          if (nkt > 0) return item;
        }
        nkt = 0;
        return total_items;
      };
      item_c = next_item_with_tiles(item_c, nkt_c);
      int gs = 0, qi = 0;                                // S tiles issued, items whose Q has been waited for
      auto issue_s = [&](bool first_of_item, bool last_of_item) {
        if (first_of_item) { bwait_long(q_full, qi & 1); ++qi; }
        const int s = gs % C::kv_stages;
        bwait_long(&kv_full[s], (gs / C::kv_stages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * (C::k_bytes + C::v_bytes));
#pragma unroll
        for (int kk = 0; kk < 64 / C::umma_k; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          umma_ss<kTf32>(tmem_S, umma_desc_k_sw128(q_addr + sl * kQ * 128 + off),
                         umma_desc_k_sw128(k_addr + sl * kKV * 128 + off), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (last_of_item) umma_commit(q_empty);          // sQ may be refilled once this MMA has completed
        ++gs;
      };
      if (item_c < total_items) issue_s(true, nkt_c == 1);
      int g = 0;                                         // PV tiles issued
      while (item_c < total_items) {
        // the tile after (item_c, j_c) in the flat sequence
        int item_n = item_c, nkt_n = nkt_c, j_n = j_c + 1;
        if (j_n == nkt_c) { j_n = 0; item_n = next_item_with_tiles(item_c + static_cast<int>(gridDim.x), nkt_n); }
        const bool has_next = item_n < total_items;
        const int s = g % C::kv_stages;
        const uint32_t v_addr = smem_u32(sKV + s * (C::k_bytes + C::v_bytes)) + C::k_bytes;
        if (kEarlyS && has_next) {                       // the next S runs under the exponentials of this tile
          bwait_long(s_taken, g & 1);
          tc_fence_after();
          issue_s(j_n == 0, j_n == nkt_n - 1);
        }
        bwait_long(p_ready, g & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kPvSteps; ++kk) {
          const int sl = (kk * 32) / 128, off = (kk * 32) % 128;
          const int half = kk / (kPvSteps / 2);           // keys 0..63 -> O_a, 64..127 -> O_b
          const uint64_t bdesc = v_mn ? umma_desc_mn_sw128(v_addr + kk * C::umma_k * 128, kKV * 128, 1024)
                                      : umma_desc_k_sw128(v_addr + sl * 64 * 128 + off);
          umma_ss<kTf32>(tmem_O + half * 64, umma_desc_k_sw128(p_addr + sl * kQ * 128 + off), bdesc, idesc_o,
                         (j_c != 0 || (kk % (kPvSteps / 2)) != 0) ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[s]);
        ++g;
        if (!kEarlyS && has_next) issue_s(j_n == 0, j_n == nkt_n - 1);   // single K/V stage: K(next) can only land after PV
        item_c = item_n; nkt_c = nkt_n; j_c = j_n;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int q = warp & 3;                            // TMEM lane quarter this warp may touch
    const int hf = (warp - 2) >> 2;                    // key half (64 of the tile's 128 score columns) owned by this thread
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_S + lane_off + hf * 64;
    const uint32_t tO = tmem_O + lane_off + hf * 64;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescale = 8.0f;                   // raise the running max only when it is stale by > 2^8
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    float2* s_ml = reinterpret_cast<float2*>(sP);      // half-merge exchange; sP is dead between an item's last PV and the next P
    int g = 0;                                         // key tiles processed so far (barrier phases run across items)
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    const AttnItem it = item_at(item);
    const int nkt = it.nkt, len = it.len;
    const int t = it.t0 + r;
    float m_l = -INFINITY;                             // running (stale) max, log2 domain
    float2 l01 = make_float2(0.f, 0.f), l23 = make_float2(0.f, 0.f);   // row sum of this key half (two packed chains)
    for (int j = 0; j < nkt; ++j, ++g) {
      bwait(s_full, g & 1);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_taken);               // S is in registers: the MMA thread may start the next S
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
      const bool any_raise = __any_sync(0xffffffffu, raise);
      float alpha = 1.f;
      if (any_raise) {
        const float m_new = raise ? mx_l : m_l;
        alpha = m_l == -INFINITY ? 0.f : fast_exp2(m_l - m_new);   // 1 for the lanes that keep their max
        const float2 a2 = make_float2(alpha, alpha);
        l01 = __fmul2_rn(l01, a2); l23 = __fmul2_rn(l23, a2);
        m_l = m_new;
      }
      // exponentials in place (sv[i] <- bits of p_i): everything up to here overlaps the previous PV and the next S on the
      // tensor pipe.  x = s * log2(e) - m on packed FFMA2 (two scores per issue slot), then MUFU.EX2.
      {
        const float2 k2 = make_float2(kLog2e, kLog2e), nm2 = make_float2(-m_l, -m_l);
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), k2, nm2);
            if (((i >> 1) & 7) < kPolyOf8) {           // compile-time split between the FMA-pipe form and MUFU.EX2
              exp2_poly_pair(x, sv[i], sv[i + 1]);
            } else {
              sv[i] = __float_as_uint(fast_exp2(x.x));
              sv[i + 1] = __float_as_uint(fast_exp2(x.y));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), k2, nm2);
            sv[i] = __float_as_uint(kbase + i < len ? fast_exp2(x.x) : 0.f);
            sv[i + 1] = __float_as_uint(kbase + i + 1 < len ? fast_exp2(x.y) : 0.f);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        l01 = __fadd2_rn(l01, make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        l23 = __fadd2_rn(l23, make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
      }
      if (j > 0) {                                     // PV(j-1) must have drained before sP is rewritten / O rescaled
        bwait(pv_done, (g - 1) & 1);                   // (j == 0: the previous item's last PV was waited for by its merge)
        tc_fence_after();
        if (any_raise) {
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
      }
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        // P[r][c..c+15] -> K-major 128B-swizzled smem (A operand of the PV MMA)
        const int byte0 = (hf * 64 + cc) * C::es;          // byte offset of key c within the row
        uint8_t* slice = p_row + (byte0 / 128) * (kQ * 128);
        const int ch0 = (byte0 % 128) / 16;
        if constexpr (C::es == 2) {
          float a0[8], a1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a0[i] = __uint_as_float(sv[cc + i]); a1[i] = __uint_as_float(sv[cc + 8 + i]); }
          store8(reinterpret_cast<T*>(slice + ((ch0 ^ sw) * 16)), a0);
          store8(reinterpret_cast<T*>(slice + (((ch0 + 1) ^ sw) * 16)), a1);
        } else {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            *reinterpret_cast<uint4*>(slice + (((ch0 + gq) ^ sw) * 16)) =
                make_uint4(sv[cc + 4 * gq], sv[cc + 4 * gq + 1], sv[cc + 4 * gq + 2], sv[cc + 4 * gq + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // merge the two key halves: out = (O_a w_a + O_b w_b) / (l_a w_a + l_b w_b), w_x = 2^(m_x - max(m_a, m_b)).
    // The exchange goes through sP (dead once the item's last PV has completed; the next item's first P is only written
    // after the second barrier below).
    if (nkt > 0) {
      bwait(pv_done, (g - 1) & 1);
      tc_fence_after();
    }
    s_ml[hf * kQ + r] = make_float2(m_l, (l01.x + l01.y) + (l23.x + l23.y));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 ma = s_ml[r], mb = s_ml[kQ + r];
    asm volatile("bar.sync 1, 256;" ::: "memory");    // both halves have read the exchange: sP may take the next item's P
    const float m_tot = fmaxf(ma.x, mb.x);
    const float wa = ma.x == -INFINITY ? 0.f : fast_exp2(ma.x - m_tot);
    const float wb = mb.x == -INFINITY ? 0.f : fast_exp2(mb.x - m_tot);
    const float l_tot = ma.y * wa + mb.y * wb;
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const float fa = wa * inv, fb = wb * inv;
    T* orow = ctx + it.b * ctx_bs + static_cast<long long>(t) * ctx_ld + it.h * 64 + hf * 32;
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
    tc_fence_before();                                 // this item's O reads are ordered before the p_ready arrive that lets the
    }   // item loop                                   // next item's first PV overwrite the accumulators
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
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
  // ATTN_POLY (0..8, default 3): share (in eighths) of the exponentials computed on the FMA pipe, 16-bit operands only
  const int poly = C::es == 2 ? tuning(TUNE_ATTN_POLY) : 0;
  auto kern = poly == 0 ? attention_tc_kernel<T, 0> : (poly <= 2 ? attention_tc_kernel<T, 2> : (poly == 3 ? attention_tc_kernel<T, 3> : attention_tc_kernel<T, 4>));
  static DeviceFlags attr_set[4];
  SB_OPT_IN_SMEM(attr_set[poly == 0 ? 0 : (poly <= 2 ? 1 : (poly == 3 ? 2 : 3))], kern, C::smem);
  const int q_tiles = ceil_div(Tlen, kQ);
  const long long total = static_cast<long long>(B) * H * q_tiles;
  SB_REQUIRE(total < (1LL << 30), "attention_tc: too many work items");
  // persistent: one resident CTA per slot (two per SM for bf16 operands, one for fp32) walks the items; ATTN_PERSIST=0 -> one item per CTA
  const int slots = (C::es == 2 ? 2 : 1) * num_sms();
  const int grid = (tuning(TUNE_ATTN_PERSIST) != 0 && total > slots) ? slots : static_cast<int>(total);
  SB_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kThreadsTc), C::smem, s, tmQK, tmVT, lens, static_cast<T*>(ctx),
                        static_cast<long long>(ctx_bs), ctx_ld, Tlen, H, q_tiles, v_mn, static_cast<int>(total)));
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
  const int es = dtype != STYLER_F32 ? 2 : 4;
  SB_REQUIRE(vt != nullptr || dtype != STYLER_F32,
             "attention_tc: row-major V (vt == NULL) is only validated for 16-bit operands; pass V^T for fp32/tf32");
  SB_REQUIRE((static_cast<int64_t>(ctx_ld) * es) % 16 == 0 && (ctx_bs * es) % 16 == 0 &&
                 (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
             "attention_tc: ctx must be 16-byte aligned/strided");
  if (dtype == STYLER_BF16)
    return launch_tc<__nv_bfloat16>(qk, qk_bs, qk_ld, vt, vt_bs, vt_ld, lens, ctx, ctx_bs, ctx_ld, B, T, H, s);
  if (dtype == STYLER_F16)
    return launch_tc<__half>(qk, qk_bs, qk_ld, vt, vt_bs, vt_ld, lens, ctx, ctx_bs, ctx_ld, B, T, H, s);
  return launch_tc<float>(qk, qk_bs, qk_ld, vt, vt_bs, vt_ld, lens, ctx, ctx_bs, ctx_ld, B, T, H, s);
}

}  // namespace sb

extern "C" int styler_attention_fwd(const void* qk, int64_t qk_bstride, int32_t qk_ld, const void* vt,
                                    int64_t vt_bstride, int32_t vt_ld, const int64_t* lens, void* ctx,
                                    int64_t ctx_bstride, int32_t ctx_ld, int32_t B, int32_t T, int32_t H,
                                    int32_t dtype, int32_t impl, void* stream) {
  sb::TraceScope trace__("attention", stream, B, T, H, 0);
  using namespace sb;
  SB_REQUIRE(qk != nullptr && ctx != nullptr, "attention: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && H > 0, "attention: bad shape B=%d T=%d H=%d", B, T, H);
  SB_REQUIRE(sb::dtype_ok(dtype), "attention: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == STYLER_IMPL_AUTO) impl = STYLER_IMPL_TC;
  if (impl == STYLER_IMPL_TC)
    return attention_tc(qk, qk_bstride, qk_ld, vt, vt_bstride, vt_ld, lens, ctx, ctx_bstride, ctx_ld, B, T, H, dtype, s);
  SB_REQUIRE(impl == STYLER_IMPL_SIMT, "attention: bad impl %d", impl);
  return attention_simt(qk, qk_bstride, qk_ld, vt, vt_bstride, vt_ld, lens, ctx, ctx_bstride, ctx_ld, B, T, H, dtype, s);
}
