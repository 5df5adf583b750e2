// Implicit-GEMM Conv1d / Linear on the 5th-gen tensor cores (tcgen05) for sm_100a.
//
//   y[b,t,n] = epilogue( sum_{tap,c} x[b, t+tap-pad, c] * w[tap][n][c] )
//
// One CTA computes a 128(t) x BN(n) output tile of ONE utterance.  Roles (warp-specialised, 192 threads):
//   warp 0   : TMA producer.  Per k-block (one tap x one 128-byte slice of channels) it loads
//              A = x[b, t0+tap-pad .. +128, c0..]   through a 3-D tensor map {C, T, B}  (rows outside [0,T) and
//                  channels >= Cin are zero-filled by TMA: that IS the Conv1d zero padding, and a tile can
//                  never bleed into the neighbouring utterance);
//              B = w[tap][n0 .. n0+BN][c0..]       through a 3-D tensor map {C, N, KS};
//              both land in 128B-swizzled K-major smem, signalled by an mbarrier (complete_tx).
//   warp 1   : allocates TMEM, then one elected lane issues tcgen05.mma (kind::f16 for bf16 operands,
//              kind::tf32 for fp32 operands), fp32 accumulator 128 lanes x BN columns in TMEM;
//              tcgen05.commit releases smem stages back to the producer and finally signals the epilogue.
//   warps 2-5: epilogue.  Thread r owns output row t0+r (TMEM lane r).  tcgen05.ld 16 columns at a time ->
//              bias, activation, residual, LayerNorm over the full N-wide row (two passes over TMEM, the
//              pre-norm values parked back in TMEM with tcgen05.st), padding mask, optional 256->1 row dot,
//              stores as bf16/fp32 (+ optional fp32 copy, + optional transposed store for V^T).
//
// Bounding roofline: tensor pipe (dense contraction).  Algorithmic FLOPs per launch = 2*B*T*N*KS*Cin.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <stdlib.h>

namespace sb {

namespace {

constexpr int kBM = 128;
constexpr int kAStageBytes = kBM * 128;
constexpr int kThreads = 192;

// Optional per-CTA phase timestamps (tools/phase_timing.py): 8 x int64 per CTA written with clock64():
// [0] kernel entry  [1] after TMEM alloc + setup sync  [2] MMA thread: first stage full  [3] MMA thread: all issued
// [4] epilogue: tmem_full observed  [5] epilogue: LN pass 1 done  [6] epilogue done  [7] before exit
long long* g_phase_buf = nullptr;
int g_phase_cap = 0;

struct EpiParams {
  long long* dbg;
  const float* bias;
  int act, act2;
  const void* residual; long long r_bstride; int r_ld; int res_f32;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  const int64_t* lens;
  const float* dot_w; float dot_b; float* dot_out;
  void* out; long long o_bstride; int o_ld;
  float* out_f32; long long of_bstride; int of_ld;
  void* vt; int vt_col0; long long vt_bstride; int vt_ld;
};

template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&v)[16]) {
  float a[8], b[8];
  load8(p, a);
  load8(p + 8, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = a[i]; v[8 + i] = b[i]; }
}
template <typename T>
__device__ __forceinline__ void store16(T* p, const float (&v)[16]) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = v[i]; b[i] = v[8 + i]; }
  store8(p, a);
  store8(p + 8, b);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const EpiParams ep, int Tlen, int n_tiles,
                                                             int tiles_per_utt, int KS, int pad, int kb_per_tap,
                                                             int BN, int stages) {
  constexpr bool kTf32 = sizeof(T) == 4;
  constexpr int kBKE = 128 / sizeof(T);  // elements per 128-byte k-slice

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_stage_bytes = BN * 128;
  const int stage_bytes = kAStageBytes + b_stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_par = reinterpret_cast<float*>(smem + stages * stage_bytes + 256);   // [4][256]: bias, gamma, beta, dot_w

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = ep.dbg != nullptr ? ep.dbg + static_cast<long long>(blockIdx.x) * 8 : nullptr;
  if (dbg != nullptr && threadIdx.x == 0) dbg[0] = clock64();
  const int nt = blockIdx.x % n_tiles, mt = blockIdx.x / n_tiles;
  const int b = mt / tiles_per_utt, t0 = (mt % tiles_per_utt) * kBM;
  const int n0 = nt * BN;
  const int num_kb = KS * kb_per_tap;
  const uint32_t tmem_cols = tmem_cols_pow2(BN);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg != nullptr && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        const int tap = kb / kb_per_tap, kc = (kb % kb_per_tap) * kBKE;
        uint8_t* sa = smem + s * stage_bytes;
        tma_load_3d(sa, &tmA, &full[s], kc, t0 + tap - pad, b);
        tma_load_3d(sa + kAStageBytes, &tmB, &full[s], kc, n0, tap);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(kTf32 ? UMMA_FMT_TF32 : UMMA_FMT_BF16, kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full[s], ph);
        if (dbg != nullptr && kb == 0) dbg[2] = clock64();
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + kAStageBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (32 bytes of K) per 128-byte slice
          umma_ss<kTf32>(tmem_base, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
      if (dbg != nullptr) dbg[3] = clock64();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int t = t0 + r;
    const bool row_ok = t < Tlen;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool masked = ep.lens != nullptr && row_ok && t >= static_cast<int>(ep.lens[b]);
    const bool has_ln = ep.ln_gamma != nullptr;
    const bool has_res = ep.residual != nullptr;
    const bool to_vt = ep.vt != nullptr && n0 >= ep.vt_col0;
    const T* res_row = nullptr;
    const float* res_row_f = nullptr;
    if (has_res) {
      const long long off = b * ep.r_bstride + static_cast<long long>(t) * ep.r_ld + n0;
      if (ep.res_f32) res_row_f = static_cast<const float*>(ep.residual) + off;
      else res_row = static_cast<const T*>(ep.residual) + off;
    }
    T* out_row = ep.out != nullptr
                     ? static_cast<T*>(ep.out) + b * ep.o_bstride + static_cast<long long>(t) * ep.o_ld + n0
                     : nullptr;
    float* of_row = ep.out_f32 != nullptr
                        ? ep.out_f32 + b * ep.of_bstride + static_cast<long long>(t) * ep.of_ld + n0
                        : nullptr;
    T* vt_base = to_vt ? static_cast<T*>(ep.vt) + b * ep.vt_bstride +
                             static_cast<long long>(n0 - ep.vt_col0) * ep.vt_ld + t
                       : nullptr;

    // Per-column parameters of this N tile staged in smem once per CTA (while the mainloop runs): reading them with
    // 16 dependent global loads per chunk was the dominant epilogue cost.
    {
      const int te = threadIdx.x - 64;
      for (int i = te; i < BN; i += 128) {
        s_par[i] = ep.bias != nullptr ? ep.bias[n0 + i] : 0.f;
        s_par[256 + i] = has_ln ? ep.ln_gamma[n0 + i] : 1.f;
        s_par[512 + i] = has_ln ? ep.ln_beta[n0 + i] : 0.f;
        s_par[768 + i] = ep.dot_w != nullptr ? ep.dot_w[n0 + i] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const float* s_bias = s_par;
    const float* s_gamma = s_par + 256;
    const float* s_beta = s_par + 512;
    const float* s_dot = s_par + 768;

    auto load_res = [&](int c, float (&rr)[16]) {       // residual columns c..c+15 of this thread's row
      if (res_row_f != nullptr) load16(res_row_f + c, rr); else load16(res_row + c, rr);
    };
    auto act_inplace = [&](float (&v)[16], int act) {
      if (act == STYLER_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (act == STYLER_ACT_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
      }
    };

    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (dbg != nullptr && threadIdx.x == 64) dbg[4] = clock64();

    float mean = 0.f, rstd = 1.f;
    if (has_ln) {
      // pass 1: v = act(acc + bias) + residual, parked back in TMEM; shifted sums for mean/variance.
      // Two 16-column chunks per iteration so TMEM and residual loads of both are in flight together.
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      for (int c = 0; c < BN; c += 32) {
        const bool two = c + 16 < BN;
        uint32_t ra[16], rb[16];
        float xa[16], xb[16];
        tmem_ld16(taddr + c, ra);
        if (two) tmem_ld16(taddr + c + 16, rb);
        if (has_res && row_ok) { load_res(c, xa); if (two) load_res(c + 16, xb); }
        tmem_ld_wait();
        {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]) + s_bias[c + i];
          act_inplace(v, ep.act);
          if (has_res && row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += xa[i];
          }
          if (c == 0) shift = v[0];
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - shift; s1 += d; s2 += d * d; ra[i] = __float_as_uint(v[i]); }
          tmem_st16(taddr + c, ra);
        }
        if (two) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rb[i]) + s_bias[c + 16 + i];
          act_inplace(v, ep.act);
          if (has_res && row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += xb[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - shift; s1 += d; s2 += d * d; rb[i] = __float_as_uint(v[i]); }
          tmem_st16(taddr + c + 16, rb);
        }
      }
      tmem_st_wait();
      const float inv_n = 1.0f / static_cast<float>(BN);
      const float dm = s1 * inv_n;
      mean = shift + dm;
      const float var = fmaxf(s2 * inv_n - dm * dm, 0.f);
      rstd = rsqrtf(var + ep.ln_eps);
    }
    if (dbg != nullptr && threadIdx.x == 64) dbg[5] = clock64();

    float dot = 0.f;
    const bool has_dot = ep.dot_w != nullptr;
    auto finish_chunk = [&](int c, float (&v)[16]) {   // v = final values of columns c..c+15
      if (has_dot) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dot = fmaf(v[i], s_dot[c + i], dot);
      }
      if (masked) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (row_ok) {
        if (to_vt) {
#pragma unroll
          for (int i = 0; i < 16; ++i) DT<T>::st(vt_base + static_cast<long long>(c + i) * ep.vt_ld, v[i]);
        } else {
          if (out_row != nullptr) store16(out_row + c, v);
          if (of_row != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(of_row + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
    };
    for (int c = 0; c < BN; c += 32) {
      const bool two = c + 16 < BN;
      uint32_t ra[16], rb[16];
      float xa[16], xb[16];
      tmem_ld16(taddr + c, ra);
      if (two) tmem_ld16(taddr + c + 16, rb);
      const bool need_res = has_res && !has_ln && row_ok;
      if (need_res) { load_res(c, xa); if (two) load_res(c + 16, xb); }
      tmem_ld_wait();
      float v[16];
      if (has_ln) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (__uint_as_float(ra[i]) - mean) * rstd * s_gamma[c + i] + s_beta[c + i];
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]) + s_bias[c + i];
        act_inplace(v, ep.act);
        if (need_res) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += xa[i];
        }
      }
      act_inplace(v, ep.act2);
      finish_chunk(c, v);
      if (two) {
        if (has_ln) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = (__uint_as_float(rb[i]) - mean) * rstd * s_gamma[c + 16 + i] + s_beta[c + 16 + i];
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rb[i]) + s_bias[c + 16 + i];
          act_inplace(v, ep.act);
          if (need_res) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += xb[i];
          }
        }
        act_inplace(v, ep.act2);
        finish_chunk(c + 16, v);
      }
    }
    if (ep.dot_out != nullptr && row_ok)
      ep.dot_out[static_cast<long long>(b) * Tlen + t] = masked ? 0.f : dot + ep.dot_b;
    if (dbg != nullptr && threadIdx.x == 64) dbg[6] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  if (dbg != nullptr && threadIdx.x == 0) dbg[7] = clock64();
}

int smem_budget_bytes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STYLER_TC_SMEM_KB");
    v = (e != nullptr ? atoi(e) : 110) * 1024;
    if (v < 64 * 1024) v = 64 * 1024;
    if (v > 220 * 1024) v = 220 * 1024;
  }
  return v;
}

int pick_bn(const styler_conv1d_args& a, int m_tiles) {
  if (a.ln_gamma != nullptr || a.dot_w != nullptr) return (a.N <= 256 && a.N % 16 == 0) ? a.N : 0;
  static int forced = -1;   // tuning override: STYLER_TC_BN
  if (forced < 0) { const char* e = getenv("STYLER_TC_BN"); forced = e != nullptr ? atoi(e) : 0; }
  if (forced > 0 && forced % 16 == 0 && forced <= 256 && a.N % forced == 0 && (a.vt == nullptr || a.vt_col0 % forced == 0))
    return forced;
  // largest tile that still gives >= 2 waves of CTAs; otherwise the smallest tile >= 64 (more CTAs);
  // otherwise the largest tile available.
  int largest = 0, smallest64 = 0;
  for (int bn = 256; bn >= 16; bn -= 16) {
    if (a.N % bn != 0) continue;
    if (a.vt != nullptr && a.vt_col0 % bn != 0) continue;
    if (largest == 0) largest = bn;
    if (static_cast<long long>(m_tiles) * (a.N / bn) >= 296) return bn;
    if (bn >= 64) smallest64 = bn;
  }
  return smallest64 != 0 ? smallest64 : largest;
}

template <typename T>
int launch(const styler_conv1d_args& a, cudaStream_t stream) {
  constexpr int es = sizeof(T);
  constexpr int bke = 128 / es;
  const int tiles_per_utt = ceil_div(a.T, kBM);
  const int m_tiles = a.B * tiles_per_utt;
  const int BN = pick_bn(a, m_tiles);
  SB_REQUIRE(BN > 0, "conv1d_tc: no valid N tile for N=%d", a.N);
  const int n_tiles = a.N / BN;
  const int kb_per_tap = ceil_div(a.Cin, bke);
  const int num_kb = a.KS * kb_per_tap;
  const int stage_bytes = kAStageBytes + BN * 128;
  int stages = smem_budget_bytes() / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > num_kb) stages = num_kb;
  if (stages < 2) stages = num_kb >= 2 ? 2 : 1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024 /*align*/ + 256 /*barriers*/ + 4096 /*params*/;
  SB_REQUIRE(smem <= 227 * 1024, "conv1d_tc: smem %zu too large", smem);

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.Cin), static_cast<uint64_t>(a.T), static_cast<uint64_t>(a.B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.x_ld) * es,
                                 static_cast<uint64_t>(a.B > 1 ? a.x_bstride : static_cast<int64_t>(a.x_ld) * a.T) * es};
    const uint32_t box[3] = {static_cast<uint32_t>(bke), static_cast<uint32_t>(kBM), 1};
    int rc = make_tmap(&tmA, a.x, es == 2 ? 1 : 0, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.Cin), static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.KS)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.Cin) * es, static_cast<uint64_t>(a.Cin) * a.N * es};
    const uint32_t box[3] = {static_cast<uint32_t>(bke), static_cast<uint32_t>(BN), 1};
    int rc = make_tmap(&tmB, a.w, es == 2 ? 1 : 0, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  EpiParams ep;
  ep.dbg = (g_phase_buf != nullptr && m_tiles * n_tiles <= g_phase_cap) ? g_phase_buf : nullptr;
  ep.bias = a.bias; ep.act = a.act; ep.act2 = a.act2;
  ep.residual = a.residual; ep.r_bstride = a.r_bstride; ep.r_ld = a.r_ld; ep.res_f32 = a.residual_is_f32;
  ep.ln_gamma = a.ln_gamma; ep.ln_beta = a.ln_beta; ep.ln_eps = a.ln_eps;
  ep.lens = a.lens;
  ep.dot_w = a.dot_w; ep.dot_b = a.dot_b; ep.dot_out = a.dot_out;
  ep.out = a.out; ep.o_bstride = a.o_bstride; ep.o_ld = a.o_ld;
  ep.out_f32 = a.out_f32; ep.of_bstride = a.of_bstride; ep.of_ld = a.of_ld;
  ep.vt = a.vt; ep.vt_col0 = a.vt_col0; ep.vt_bstride = a.vt_bstride; ep.vt_ld = a.vt_ld;

  auto kern = conv1d_tc_kernel<T>;
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  kern<<<m_tiles * n_tiles, kThreads, smem, stream>>>(tmA, tmB, ep, a.T, n_tiles, tiles_per_utt, a.KS, a.pad,
                                                      kb_per_tap, BN, stages);
  SB_LAUNCH_OK();
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool conv1d_tc_supported(const styler_conv1d_args& a, const char** why) {
  const int es = a.dtype == STYLER_BF16 ? 2 : 4;
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (a.N % 16 != 0) return fail("N not a multiple of 16");
  if ((a.Cin * es) % 16 != 0) return fail("Cin row not a multiple of 16 bytes");
  if (!aligned16(a.x) || (static_cast<int64_t>(a.x_ld) * es) % 16 != 0 || (a.x_bstride * es) % 16 != 0)
    return fail("x not 16-byte aligned/strided");
  if (!aligned16(a.w)) return fail("w not 16-byte aligned");
  if ((a.ln_gamma != nullptr || a.dot_w != nullptr) && a.N > 256) return fail("LayerNorm/dot epilogue needs N <= 256");
  if (a.out != nullptr && (!aligned16(a.out) || (static_cast<int64_t>(a.o_ld) * es) % 16 != 0 || (a.o_bstride * es) % 16 != 0))
    return fail("out not 16-byte aligned/strided");
  if (a.out_f32 != nullptr && (!aligned16(a.out_f32) || (a.of_ld % 4) != 0 || (a.of_bstride % 4) != 0))
    return fail("out_f32 not 16-byte aligned/strided");
  const int res_es = a.residual_is_f32 ? 4 : es;
  if (a.residual != nullptr && (!aligned16(a.residual) || (static_cast<int64_t>(a.r_ld) * res_es) % 16 != 0 || (a.r_bstride * res_es) % 16 != 0))
    return fail("residual not 16-byte aligned/strided");
  if (a.vt != nullptr && a.vt_col0 % 16 != 0) return fail("vt_col0 not a multiple of 16");
  if (a.T < 1 || a.B < 1) return fail("empty problem");
  return true;
}

void set_phase_buffer(long long* buf, int cap) { g_phase_buf = buf; g_phase_cap = cap; }

int conv1d_tc(const styler_conv1d_args& a, cudaStream_t s) {
  const char* why = nullptr;
  SB_REQUIRE(conv1d_tc_supported(a, &why), "conv1d_tc: unsupported arguments: %s", why ? why : "?");
  if (a.dtype == STYLER_BF16) return launch<__nv_bfloat16>(a, s);
  return launch<float>(a, s);
}

}  // namespace sb
