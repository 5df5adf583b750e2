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
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter, half the columns each)

// Optional per-CTA phase timestamps (tools/phase_timing.py): 8 x int64 per CTA written with clock64():
// [0] kernel entry  [1] after TMEM alloc + setup sync  [2] MMA thread: first stage full  [3] MMA thread: all issued
// [4] epilogue: tmem_full observed  [5] epilogue: LN pass 1 done  [6] epilogue done  [7] before exit
long long* g_phase_buf = nullptr;
int g_phase_cap = 0;

struct EpiParams {
  long long* dbg;
  int stage_out, stage_res;   // 1: output rows / residual rows go through the freed pipeline smem with TMA (coalesced)
  const float* bias;
  int act, act2;
  float slope, inv_slope;     // leaky ReLU negative-side slope and its reciprocal
  int res_inv;                // residual is stored as lrelu(r): undo before the add
  const void* residual; long long r_bstride; int r_ld; int res_f32;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  const int64_t* lens;
  const float* dot_w; float dot_b; float* dot_out;
  void* out; long long o_bstride; int o_ld;
  float* out_f32; long long of_bstride; int of_ld;
  float* out2_f32;            // optional second fp32 destination (same strides): peer-mapped gather slice
  int contig_f32;             // non-staged epilogue, one N tile, full fp32 rows (of_ld == N): the tile's fp32 output (and fp32
                              // residual) is one CONTIGUOUS block in global memory -> staged through plain smem, moved with
                              // coalesced float4 accesses by all epilogue threads (mel_linear, last PostNet conv: N = 80)
  float* gn_partial;          // optional GroupNorm partial sums [B][tiles_per_utt][N/16][2] (sum, sum of squares per 16 channels)
  int n_total;                // N of the whole problem (gn_partial indexing)
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

// Packed fp32 arithmetic (FADD2 / FFMA2, sm_100: two lanes of work per issue slot; the FMA pipe accepts one warp instruction
// every 2 clk, and these epilogues are bound by instruction issue, not by memory -- profiles/prof_kernels_r1e.txt).
__device__ __forceinline__ void add16(float (&v)[16], const float (&b)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 r = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b[i], b[i + 1]));
    v[i] = r.x; v[i + 1] = r.y;
  }
}
__device__ __forceinline__ void add16_bits(float (&v)[16], const uint32_t (&a)[16], const float (&b)[16]) {   // v = a + b
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 r = __fadd2_rn(make_float2(__uint_as_float(a[i]), __uint_as_float(a[i + 1])), make_float2(b[i], b[i + 1]));
    v[i] = r.x; v[i + 1] = r.y;
  }
}
// shifted sums for LayerNorm: s1 += (v - shift), s2 += (v - shift)^2, two packed chains
__device__ __forceinline__ void stats16(const float (&v)[16], float shift, float2 (&s1)[2], float2 (&s2)[2]) {
  const float2 ns = make_float2(-shift, -shift);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 d = __fadd2_rn(make_float2(v[i], v[i + 1]), ns);
    s1[(i >> 1) & 1] = __fadd2_rn(s1[(i >> 1) & 1], d);
    s2[(i >> 1) & 1] = __ffma2_rn(d, d, s2[(i >> 1) & 1]);
  }
}
// v = ((x * rstd + nmr) * gamma + beta)
__device__ __forceinline__ void norm16(float (&v)[16], const uint32_t (&x)[16], float rstd, float nmr, const float (&g)[16],
                                       const float (&b)[16]) {
  const float2 r2 = make_float2(rstd, rstd), n2 = make_float2(nmr, nmr);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 t = __ffma2_rn(make_float2(__uint_as_float(x[i]), __uint_as_float(x[i + 1])), r2, n2);
    const float2 r = __ffma2_rn(t, make_float2(g[i], g[i + 1]), make_float2(b[i], b[i + 1]));
    v[i] = r.x; v[i + 1] = r.y;
  }
}

// ACT (activation before residual/LN) and LN are compile-time so each instance carries one tight epilogue loop: the
// generic runtime-switched body was ~1850 SASS instructions per 32 columns (tanhf inlined 64x) and ran at ~7 clk/instr.
// FAST: output staged through smem + TMA store, residual (if any) staged through smem, no V^T / fp32 copy / row dot:
// the common case, compiled without the alternative paths (instruction issue, not memory, bounds this epilogue).
// PERSIST (small-K GEMMs, FAST epilogue only): the CTA loops over output tiles (tile = blockIdx.x + i * gridDim.x) with
// TWO accumulators in TMEM, a dedicated output/residual staging buffer next to the operand ring, and the operand ring
// running on across tile boundaries.  The loads and MMAs of tile i+1 then overlap the epilogue and the TMA store of tile
// i inside one CTA: with K <= 1024 a tile's mainloop (2-8 k cycles of MMA) is shorter than its TMA round trip plus its
// epilogue (~6 k cycles), which the one-tile-per-CTA form pays serially (phase stamps: profiles/phase_timing_r1c.txt).
// CG2 (bf16, FAST, non-persistent, BN = 256): CTA pairs.  A cluster of two CTAs owns two vertically adjacent
// 128-row tiles of the same N tile; each CTA stages its own A rows and 128 of the 256 weight rows, the leader issues
// tcgen05.mma.cta_group::2 (M = 256) and every CTA runs the epilogue of its own 128 accumulator rows (ptx.cuh).  The big
// convolutions are bound by the operand bytes each SM has to pull from L2 (FFN k9: 48 KB per 128x256x64 k-block,
// 15.5 TB/s chip-wide at 0.455 ms); the pair needs 32 KB per SM for the same MMA work.
template <typename T, int ACT, bool LN, bool FAST, bool PERSIST, bool CG2 = false>
__global__ void __launch_bounds__(kThreads, 2) conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmOut,
                                                             const __grid_constant__ CUtensorMap tmRes,
                                                             const EpiParams ep, int Tlen, int n_tiles,
                                                             int tiles_per_utt, int KS, int pad, int kb_per_tap,
                                                             int BN, int stages, int total_tiles, int dil) {
  static_assert(!PERSIST || FAST, "the persistent tile loop only exists for the staged (FAST) epilogue");
  static_assert(!CG2 || (FAST && !PERSIST && sizeof(T) == 2), "CTA pairs: bf16, staged epilogue, one tile per CTA");
  constexpr bool kTf32 = sizeof(T) == 4;
  constexpr int kBKE = 128 / sizeof(T);  // elements per 128-byte k-slice

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_stage_bytes = (CG2 ? BN / 2 : BN) * 128;           // CTA pair: this CTA holds half of the weight rows
  const int stage_bytes = kAStageBytes + b_stage_bytes;
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;        // 0 = leader (issues the MMAs of the pair)
  const int n_box = (BN * static_cast<int>(sizeof(T))) / 128;     // 16 KB [128 rows x 128 B] boxes per output tile
  // epilogue staging (output rows / residual rows): its own region when persistent, the idle operand ring otherwise
  uint8_t* stg = PERSIST ? smem + stages * stage_bytes : smem;
  uint8_t* tail = smem + stages * stage_bytes + (PERSIST ? n_box * kAStageBytes : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;        // [2] MMA -> epilogue, one per accumulator
  uint64_t* tmem_empty = tmem_full + 2;        // [2] epilogue -> MMA (persistent only)
  uint64_t* res_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);
  const int pstr = BN > 256 ? 512 : 256;                                        // per-column parameter vectors: [4][pstr]
  float* s_par = reinterpret_cast<float*>(tail + 256);                          // bias, gamma, beta, dot_w
  float* s_x = s_par + 4 * pstr;                                                // [256 threads][4]: LN / dot exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = ep.dbg != nullptr ? ep.dbg + static_cast<long long>(blockIdx.x) * 8 : nullptr;
  if (dbg != nullptr && threadIdx.x == 0) dbg[0] = clock64();
  const int num_kb = KS * kb_per_tap;
  const uint32_t acc_cols = tmem_cols_pow2(BN);                      // columns of one accumulator
  const uint32_t tmem_cols = PERSIST ? 2 * acc_cols : acc_cols;
  const int tile_step = PERSIST ? static_cast<int>(gridDim.x) : total_tiles;   // non-persistent: exactly one tile per CTA
  const int tile_first = CG2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // CG2: total_tiles counts PAIRS
  auto m_tile_of = [&](int tile) { return CG2 ? 2 * (tile / n_tiles) + static_cast<int>(cta_rank) : tile / n_tiles; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 1); }
    mbar_init(res_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { if (CG2) tmem_alloc_2cta(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols); }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();     // the peer's barriers are initialised before any remote arrive / complete_tx can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();        // the next kernel may start its prologue in SM slots this grid no longer needs
  pdl_grid_dependency_wait();     // everything above overlapped the previous kernel's tail; its outputs are visible now
  if (dbg != nullptr && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0;                                   // k-block counter over all tiles of this CTA: the ring never drains
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int nt = tile % n_tiles, mt = m_tile_of(tile);
        const int b = mt / tiles_per_utt, t0 = (mt % tiles_per_utt) * kBM, n0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % stages;
          const uint32_t ph = (g / stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          const int tap = kb / kb_per_tap, kc = (kb % kb_per_tap) * kBKE;
          uint8_t* sa = smem + s * stage_bytes;
          if constexpr (CG2) {
            // both CTAs of the pair credit the LEADER's full barrier; only the leader arms it (with the bytes of both)
            if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * stage_bytes);
            const uint32_t bar = leader_cta_addr(smem_u32(&full[s]));
            tma_load_3d_2cta(sa, &tmA, bar, kc, t0 + tap * dil - pad, b);
            tma_load_3d_2cta(sa + kAStageBytes, &tmB, bar, kc, n0 + static_cast<int>(cta_rank) * (BN / 2), tap);
          } else {
            mbar_arrive_expect_tx(&full[s], stage_bytes);
            tma_load_3d(sa, &tmA, &full[s], kc, t0 + tap * dil - pad, b);
            tma_load_3d(sa + kAStageBytes, &tmB, &full[s], kc, n0, tap);
            if (BN > 256)        // wide tile: a TMA box has at most 256 rows, the weight tile arrives as two boxes of BN/2 rows
              tma_load_3d(sa + kAStageBytes + (BN / 2) * 128, &tmB, &full[s], kc, n0 + BN / 2, tap);
          }
        }
        if (!PERSIST && (FAST ? ep.residual != nullptr : ep.stage_res != 0)) {
          // all MMAs done -> the pipeline stages are free: stage the residual tile there
          mbar_wait(&tmem_full[0], 0);
          mbar_arrive_expect_tx(res_full, n_box * kAStageBytes);
          for (int bx = 0; bx < n_box; ++bx)
            tma_load_3d(stg + bx * kAStageBytes, &tmRes, res_full, n0 + bx * kBKE, t0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {
      // WIDE tile (256 < BN <= 512, e.g. the 320-channel audio-encoder convs): one A stage feeds TWO MMAs of BN/2 columns each,
      // so the activations are pulled from L2 once per 128 x BN outputs (these convs are bound by operand bytes, not by the pipe)
      const int mma_n = BN > 256 ? BN / 2 : BN;
      const uint32_t idesc = umma_idesc(umma_fmt_of<T>(), CG2 ? 2 * kBM : kBM, mma_n);
      int g = 0, it = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++it) {
        const int ab = PERSIST ? (it & 1) : 0;
        const uint32_t acc = tmem_base + ab * acc_cols;
        if (PERSIST) {                             // the epilogue has drained this accumulator (tile it-2)
          mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % stages;
          const uint32_t ph = (g / stages) & 1;
          mbar_wait(&full[s], ph);
          if (dbg != nullptr && kb == 0) dbg[2] = clock64();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + kAStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (32 bytes of K) per 128-byte slice
            if constexpr (CG2)
              umma_ss_2cta_f16(acc, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                               (kb | k) != 0 ? 1u : 0u);
            else {
              umma_ss<kTf32>(acc, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                             (kb | k) != 0 ? 1u : 0u);
              if (mma_n != BN)     // second half of the wide tile: weight rows mma_n.., accumulator columns mma_n..
                umma_ss<kTf32>(acc + mma_n, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + mma_n * 128 + k * 32),
                               idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (CG2) umma_commit_2cta(&empty[s], 3); else umma_commit(&empty[s]);   // pair: frees the stage in BOTH CTAs
        }
        if constexpr (CG2) umma_commit_2cta(&tmem_full[ab], 3); else umma_commit(&tmem_full[ab]);
        if (dbg != nullptr) dbg[3] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps; 128 TMEM lanes x 2 column halves)
    // A warp may only read TMEM lanes 32*(warp%4)..+31, so every output row is served by two threads, each owning
    // half of the tile's columns: one warp per scheduler was latency-bound (~7 clk/instr), two halve the critical path.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const bool split = (BN % 32) == 0;
    const int c_begin = split ? hf * (BN / 2) : (hf == 0 ? 0 : BN);
    const int c_end = split ? c_begin + BN / 2 : BN;
    int it = 0;
    if (PERSIST && ep.residual != nullptr && threadIdx.x == 64 && static_cast<int>(blockIdx.x) < total_tiles) {   // (never CG2)
      // residual rows of this CTA's first tile -> staging buffer (later tiles: issued when the previous store has drained)
      const int nt = blockIdx.x % n_tiles, mt = blockIdx.x / n_tiles;
      mbar_arrive_expect_tx(res_full, n_box * kAStageBytes);
      for (int bx = 0; bx < n_box; ++bx)
        tma_load_3d(stg + bx * kAStageBytes, &tmRes, res_full, nt * BN + bx * kBKE, (mt % tiles_per_utt) * kBM, mt / tiles_per_utt);
    }
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++it) {
    const int nt = tile % n_tiles, mt = m_tile_of(tile);
    const int b = mt / tiles_per_utt, t0 = (mt % tiles_per_utt) * kBM;
    const int n0 = nt * BN;
    const int ab = PERSIST ? (it & 1) : 0;
    const bool tile_to_vt = !FAST && ep.vt != nullptr && n0 >= ep.vt_col0;
    const bool st_out = FAST || (ep.stage_out != 0 && !tile_to_vt);            // TMA-store this tile's output from smem
    const bool st_res = FAST ? ep.residual != nullptr : ep.stage_res != 0;      // this tile's residual arrives in smem by TMA
    // contiguous fp32 tile I/O (see EpiParams::contig_f32): s_of = [128][cld] output staging, s_rf = [128][cld] residual
    const bool contig = !FAST && ep.contig_f32 != 0;
    // row pitch of the staging (floats): padded against bank conflicts -- unless a second (peer) destination exists: then the
    // staged tile must be one contiguous block, because it leaves as ONE bulk async copy (TMA engine -> NVLink)
    const int cld = ep.out2_f32 != nullptr ? BN : BN + 4;
    float* s_of = reinterpret_cast<float*>(smem);
    float* s_rf = s_of + kBM * cld;
    const int rows_ok = min(kBM, Tlen - t0);                           // rows of this tile that exist
    const int t = t0 + r;
    const bool row_ok = t < Tlen;
    const uint32_t taddr = tmem_base + ab * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
    const bool masked = ep.lens != nullptr && row_ok && t >= static_cast<int>(ep.lens[b]);
    constexpr bool has_ln = LN;
    const bool has_res = ep.residual != nullptr;
    const bool to_vt = tile_to_vt;
    const T* res_row = nullptr;
    const float* res_row_f = nullptr;
    if (has_res && !FAST) {
      const long long off = b * ep.r_bstride + static_cast<long long>(t) * ep.r_ld + n0;
      if (ep.res_f32) res_row_f = static_cast<const float*>(ep.residual) + off;
      else res_row = static_cast<const T*>(ep.residual) + off;
    }
    T* out_row = !FAST && ep.out != nullptr
                     ? static_cast<T*>(ep.out) + b * ep.o_bstride + static_cast<long long>(t) * ep.o_ld + n0
                     : nullptr;
    float* of_row = !FAST && ep.out_f32 != nullptr
                        ? ep.out_f32 + b * ep.of_bstride + static_cast<long long>(t) * ep.of_ld + n0
                        : nullptr;
    float* of2_row = !FAST && ep.out2_f32 != nullptr
                         ? ep.out2_f32 + b * ep.of_bstride + static_cast<long long>(t) * ep.of_ld + n0
                         : nullptr;
    T* vt_base = to_vt ? static_cast<T*>(ep.vt) + b * ep.vt_bstride +
                             static_cast<long long>(n0 - ep.vt_col0) * ep.vt_ld + t
                       : nullptr;

    // Per-column parameters of this N tile staged in smem once per CTA (while the mainloop runs): reading them with
    // 16 dependent global loads per chunk was the dominant epilogue cost.
    if (it == 0 || n_tiles > 1) {   // (persistent: every reader of the previous tile's values has passed the pre-store barrier)
      const int te = threadIdx.x - 64;
      for (int i = te; i < BN; i += 256) {
        s_par[i] = ep.bias != nullptr ? ep.bias[n0 + i] : 0.f;
        s_par[pstr + i] = has_ln ? ep.ln_gamma[n0 + i] : 1.f;
        s_par[2 * pstr + i] = has_ln ? ep.ln_beta[n0 + i] : 0.f;
        s_par[3 * pstr + i] = ep.dot_w != nullptr ? ep.dot_w[n0 + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const float* s_bias = s_par;
    const float* s_gamma = s_par + pstr;
    const float* s_beta = s_par + 2 * pstr;
    const float* s_dot = s_par + 3 * pstr;
    auto ld16s = [](const float* p, float (&o)[16]) {   // 16 consecutive smem floats as 4 x LDS.128 (p is 64-byte aligned)
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const float4 f = reinterpret_cast<const float4*>(p)[g4];
        o[4 * g4] = f.x; o[4 * g4 + 1] = f.y; o[4 * g4 + 2] = f.z; o[4 * g4 + 3] = f.w;
      }
    };

    // smem staging address of columns c..c+15 of this thread's row: box (c*es/128), 128-byte row r, 16-byte chunks XOR (r&7)
    uint8_t* stage_row = stg + r * 128;
    const int sw = r & 7;
    auto stage_ptr = [&](int c, int chunk) -> uint8_t* {     // c is a multiple of 16 and every loop over c is unrolled by the
      const int byte0 = c * static_cast<int>(sizeof(T));     // compiler only partially: keep the arithmetic to shifts and one XOR
      return stage_row + (byte0 >> 7) * kAStageBytes + (((((byte0 & 127) >> 4) + chunk) ^ sw) << 4);
    };
    auto load_res = [&](int c, float (&rr)[16]) {       // residual columns c..c+15 of this thread's row
      if (FAST || st_res) {
        if constexpr (sizeof(T) == 2) {
          float a[8], bq[8];
          load8(reinterpret_cast<const T*>(stage_ptr(c, 0)), a);
          load8(reinterpret_cast<const T*>(stage_ptr(c, 1)), bq);
#pragma unroll
          for (int i = 0; i < 8; ++i) { rr[i] = a[i]; rr[8 + i] = bq[i]; }
        } else {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 f = *reinterpret_cast<const float4*>(stage_ptr(c, g4));
            rr[4 * g4] = f.x; rr[4 * g4 + 1] = f.y; rr[4 * g4 + 2] = f.z; rr[4 * g4 + 3] = f.w;
          }
        }
      } else if (contig) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          const float4 f = *reinterpret_cast<const float4*>(s_rf + r * cld + c + 4 * g4);
          rr[4 * g4] = f.x; rr[4 * g4 + 1] = f.y; rr[4 * g4 + 2] = f.z; rr[4 * g4 + 3] = f.w;
        }
      } else if (res_row_f != nullptr) {
        load16(res_row_f + c, rr);
      } else {
        load16(res_row + c, rr);
      }
      if (ep.res_inv != 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) rr[i] = rr[i] < 0.f ? rr[i] * ep.inv_slope : rr[i];
      }
    };
    auto store_out = [&](int c, const float (&v)[16]) {   // output columns c..c+15 (dtype T)
      if (FAST || st_out) {
        if constexpr (sizeof(T) == 2) {
          float a[8], bq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] = v[i]; bq[i] = v[8 + i]; }
          store8(reinterpret_cast<T*>(stage_ptr(c, 0)), a);
          store8(reinterpret_cast<T*>(stage_ptr(c, 1)), bq);
        } else {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            *reinterpret_cast<float4*>(stage_ptr(c, g4)) = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
        }
      } else if (out_row != nullptr && row_ok) {
        store16(out_row + c, v);
      }
    };
    auto act1 = [&](float (&v)[16]) {                  // compile-time activation (before residual / LayerNorm)
      if constexpr (ACT == STYLER_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if constexpr (ACT == STYLER_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * ep.slope);
      } else if constexpr (ACT == STYLER_ACT_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if constexpr (kBf16Math<T>) {                // bf16 storage: MUFU.TANH (rel. error 2^-11 < bf16 rounding)
            float y;
            asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
            v[i] = y;
          } else if constexpr (sizeof(T) == 2) {       // fp16 storage
            v[i] = tanh_ex2(v[i]);
          } else {
            v[i] = tanhf(v[i]);
          }
        }
      }
    };
    const bool relu2 = ep.act2 == STYLER_ACT_RELU || ep.act2 == STYLER_ACT_LRELU;   // final activation: none | relu | lrelu
    const float slope2 = ep.act2 == STYLER_ACT_LRELU ? ep.slope : 0.f;
    auto act2f = [&](float (&v)[16]) {
      if (relu2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * slope2);
      }
    };

    mbar_wait(&tmem_full[ab], PERSIST ? ((it >> 1) & 1) : 0);
    tc_fence_after();
    if (st_res) mbar_wait(res_full, PERSIST ? (it & 1) : 0);
    else if (PERSIST && it > 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // thread 64 has seen the previous store drain
    if (dbg != nullptr && threadIdx.x == 64) dbg[4] = clock64();
    if (contig && has_res) {
      const float* src = static_cast<const float*>(ep.residual) + b * ep.r_bstride + static_cast<long long>(t0) * BN;
      const int n4 = rows_ok * (BN / 4);
      for (int i = threadIdx.x - 64; i < n4; i += 256) {
        const int rr_ = i / (BN / 4), c4 = i - rr_ * (BN / 4);
        *reinterpret_cast<float4*>(s_rf + rr_ * cld + 4 * c4) = __ldg(reinterpret_cast<const float4*>(src) + i);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }

    float mean = 0.f, rstd = 1.f, nmr = 0.f;   // nmr = -mean * rstd: (v - mean) * rstd = fma(v, rstd, nmr)
    if (has_ln) {
      // pass 1: v = act(acc + bias) + residual, parked back in TMEM; shifted sums for mean/variance.
      // Two 16-column chunks per iteration so TMEM and residual loads of both are in flight together.
      float shift = 0.f;
      float2 s1p[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, s2p[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      for (int c = c_begin; c < c_end; c += 32) {
        const bool two = c + 16 < c_end;
        uint32_t ra[16], rb[16];
        float xa[16], xb[16];
        tmem_ld16(taddr + c, ra);
        if (two) tmem_ld16(taddr + c + 16, rb);
        const bool do_res = has_res && (row_ok || st_res);
        if (do_res) { load_res(c, xa); if (two) load_res(c + 16, xb); }
        tmem_ld_wait();
        {
          float v[16], pb[16];
          ld16s(s_bias + c, pb);
          add16_bits(v, ra, pb);
          act1(v);
          if (do_res) add16(v, xa);
          if (c == c_begin) shift = v[0];
          stats16(v, shift, s1p, s2p);
#pragma unroll
          for (int i = 0; i < 16; ++i) ra[i] = __float_as_uint(v[i]);
          tmem_st16(taddr + c, ra);
        }
        if (two) {
          float v[16], pb[16];
          ld16s(s_bias + c + 16, pb);
          add16_bits(v, rb, pb);
          act1(v);
          if (do_res) add16(v, xb);
          stats16(v, shift, s1p, s2p);
#pragma unroll
          for (int i = 0; i < 16; ++i) rb[i] = __float_as_uint(v[i]);
          tmem_st16(taddr + c + 16, rb);
        }
      }
      tmem_st_wait();
      // combine the two half-row statistics (shifted sums are merged exactly; n_h = columns owned by half h)
      float* mine = s_x + (hf * 128 + r) * 4;
      mine[0] = shift; mine[1] = (s1p[0].x + s1p[0].y) + (s1p[1].x + s1p[1].y); mine[2] = (s2p[0].x + s2p[0].y) + (s2p[1].x + s2p[1].y);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* h0 = s_x + r * 4;
      const float* h1 = s_x + (128 + r) * 4;
      const float n0f = static_cast<float>(split ? BN / 2 : BN), n1f = static_cast<float>(split ? BN / 2 : 0);
      const float inv_n = 1.0f / static_cast<float>(BN);
      mean = (n0f * h0[0] + h0[1] + n1f * h1[0] + h1[1]) * inv_n;
      const float d0 = mean - h0[0], d1 = mean - h1[0];
      const float var = fmaxf((h0[2] - 2.f * d0 * h0[1] + n0f * d0 * d0 + h1[2] - 2.f * d1 * h1[1] + n1f * d1 * d1) * inv_n, 0.f);
      rstd = rsqrtf(var + ep.ln_eps);
      nmr = -mean * rstd;
    }
    if (dbg != nullptr && threadIdx.x == 64 && has_ln) dbg[5] = clock64();

    float dot = 0.f;
    const bool has_dot = !FAST && ep.dot_w != nullptr;
    auto finish_chunk = [&](int c, float (&v)[16]) {   // v = final values of columns c..c+15
      if (has_dot) {
        float pd[16];
        ld16s(s_dot + c, pd);
#pragma unroll
        for (int i = 0; i < 16; ++i) dot = fmaf(v[i], pd[i], dot);
      }
      if (masked) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (to_vt) {
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) DT<T>::st(vt_base + static_cast<long long>(c + i) * ep.vt_ld, v[i]);
        }
      } else if (contig) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(s_of + r * cld + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
        store_out(c, v);
        if (row_ok) {
          if (of_row != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(of_row + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
          if (of2_row != nullptr) {   // e.g. rank 0's receive buffer over NVLink: the gather is fused into this epilogue
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(of2_row + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
    };
    // GroupNorm statistics of the tile (modules.py:113 normalises 16-channel groups over ALL time steps of the padded grid):
    // a 16-column chunk is exactly one group, a warp holds 32 of the tile's rows -> shuffle-reduce (sum, sum of squares) per
    // chunk, the four lane-quarter warps meet in smem, one pair per (tile, group) goes to HBM.  The stand-alone statistics
    // pass over the stored tensor (12 launches, ~15 us each per forward) disappears; fixed reduction order = deterministic.
    float* s_gn = s_x;                                        // [4 quarters][BN/16][2]; s_x is otherwise used by LN / dot only
    auto gn_accum = [&](int c, const float (&v)[16]) {
      float sg = 0.f, qg = 0.f;
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { sg += v[i]; qg = fmaf(v[i], v[i], qg); }
      }
      sg = warp_sum(sg);
      qg = warp_sum(qg);
      if (lane == 0) { s_gn[(q * (BN / 16) + (c >> 4)) * 2] = sg; s_gn[(q * (BN / 16) + (c >> 4)) * 2 + 1] = qg; }
    };
    const bool do_gn = !LN && ep.gn_partial != nullptr;
    long long t_wait = 0;
    for (int c = c_begin; c < c_end; c += 32) {
      const bool two = c + 16 < c_end;
      uint32_t ra[16], rb[16];
      float xa[16], xb[16];
      long long tq0 = 0;
      if (dbg != nullptr) tq0 = clock64();
      tmem_ld16(taddr + c, ra);
      if (two) tmem_ld16(taddr + c + 16, rb);
      const bool need_res = has_res && !has_ln && (row_ok || st_res);
      if (need_res) { load_res(c, xa); if (two) load_res(c + 16, xb); }
      tmem_ld_wait();
      if (dbg != nullptr) t_wait += clock64() - tq0;
      float v[16], pa[16], pb[16];
      if (has_ln) {
        ld16s(s_gamma + c, pa);
        ld16s(s_beta + c, pb);
        norm16(v, ra, rstd, nmr, pa, pb);
      } else {
        ld16s(s_bias + c, pb);
        add16_bits(v, ra, pb);
        act1(v);
        if (need_res) add16(v, xa);
      }
      act2f(v);
      if (do_gn) gn_accum(c, v);
      finish_chunk(c, v);
      if (two) {
        if (has_ln) {
          ld16s(s_gamma + c + 16, pa);
          ld16s(s_beta + c + 16, pb);
          norm16(v, rb, rstd, nmr, pa, pb);
        } else {
          ld16s(s_bias + c + 16, pb);
          add16_bits(v, rb, pb);
          act1(v);
          if (need_res) add16(v, xb);
        }
        act2f(v);
        if (do_gn) gn_accum(c + 16, v);
        finish_chunk(c + 16, v);
      }
    }
    if (do_gn) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int te = threadIdx.x - 64;
      if (te < BN / 16) {
        float sg = 0.f, qg = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) { sg += s_gn[(qq * (BN / 16) + te) * 2]; qg += s_gn[(qq * (BN / 16) + te) * 2 + 1]; }
        float* dst = ep.gn_partial + ((static_cast<long long>(b) * tiles_per_utt + t0 / kBM) * (ep.n_total / 16) + n0 / 16 + te) * 2;
        dst[0] = sg; dst[1] = qg;
      }
      if (PERSIST) asm volatile("bar.sync 1, 256;" ::: "memory");   // s_gn is reused by this CTA's next tile
    }
    if (!FAST && ep.dot_out != nullptr) {               // row dot: sum the two column halves
      if (has_ln) asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone has consumed the LN exchange slots
      s_x[(hf * 128 + r) * 4 + 3] = dot;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hf == 0 && row_ok)
        ep.dot_out[static_cast<long long>(b) * Tlen + t] = masked ? 0.f : s_x[r * 4 + 3] + s_x[(128 + r) * 4 + 3] + ep.dot_b;
    }
    if (dbg != nullptr && threadIdx.x == 64 && !has_ln) dbg[5] = dbg[4] + t_wait;   // non-LN tiles: slot 5 = TMEM ld+wait time
    long long t_pre_store = 0;
    if (dbg != nullptr) t_pre_store = clock64();
    if (contig) {   // the tile's rows are one contiguous block of the fp32 output (and of the dtype output): coalesced copies
      if (ep.out2_f32 != nullptr) fence_proxy_async_smem();     // the staged tile will be read by the async proxy (bulk copy)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const long long base = b * ep.of_bstride + static_cast<long long>(t0) * BN;
      float4* dst = reinterpret_cast<float4*>(ep.out_f32 + base);
      // Second destination = this rank's slice of rank 0's receive region (peer memory over NVLink): the whole tile leaves as
      // one cp.async.bulk (smem -> global) issued by one thread; the copy engine streams it while the other threads write the
      // local outputs -- remote stores from the LSU path stalled these epilogues (+0.18 ms per step at N = 2).
      if (ep.out2_f32 != nullptr && threadIdx.x == 64 && rows_ok > 0) {
        bulk_store_1d(ep.out2_f32 + base, s_of, static_cast<uint32_t>(rows_ok) * BN * sizeof(float));
        tma_store_commit();
      }
      float4* dst2 = nullptr;
      T* dsto = ep.out != nullptr ? static_cast<T*>(ep.out) + b * ep.o_bstride + static_cast<long long>(t0) * BN : nullptr;
      const int n4 = rows_ok * (BN / 4);
      for (int i = threadIdx.x - 64; i < n4; i += 256) {
        const int rr_ = i / (BN / 4), c4 = i - rr_ * (BN / 4);
        const float4 f = *reinterpret_cast<const float4*>(s_of + rr_ * cld + 4 * c4);
        dst[i] = f;
        if (dst2 != nullptr) dst2[i] = f;                      // e.g. rank 0's receive buffer over NVLink
        if (dsto != nullptr) {
          if constexpr (sizeof(T) == 2) {
            uint2 u;
            u.x = pack2<T>(f.x, f.y); u.y = pack2<T>(f.z, f.w);
            reinterpret_cast<uint2*>(dsto)[i] = u;
          } else {
            reinterpret_cast<float4*>(dsto)[i] = f;
          }
        }
      }
      if (ep.out2_f32 != nullptr && threadIdx.x == 64) tma_store_wait_all();   // the peer write has completed before this CTA can exit
    }
    if (st_out) {   // smem tile -> global with TMA (rows >= T are clipped by the tensor map)
      fence_proxy_async_smem();
      if (PERSIST) tc_fence_before();              // this tile's TMEM reads are ordered before the hand-back below
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        if (PERSIST) mbar_arrive(&tmem_empty[ab]); // the MMA thread may overwrite this accumulator (tile it+2)
        for (int bx = 0; bx < n_box; ++bx) tma_store_3d(&tmOut, stg + bx * kAStageBytes, n0 + bx * kBKE, t0, b);
        tma_store_commit();
        tma_store_wait_read();
        const int nx = tile + tile_step;
        if (PERSIST && st_res && nx < total_tiles) {   // staging is free again: fetch the next tile's residual rows
          const int nnt = nx % n_tiles, nmt = nx / n_tiles;
          mbar_arrive_expect_tx(res_full, n_box * kAStageBytes);
          for (int bx = 0; bx < n_box; ++bx)
            tma_load_3d(stg + bx * kAStageBytes, &tmRes, res_full, nnt * BN + bx * kBKE, (nmt % tiles_per_utt) * kBM,
                        nmt / tiles_per_utt);
        }
      }
    }
    if (dbg != nullptr && threadIdx.x == 64) { dbg[6] = clock64(); if (!has_ln) dbg[3] = t_pre_store; }
    }   // tile loop
  }

  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();     // neither CTA may exit (or free TMEM) while the pair's MMAs / multicast arrivals can still touch it
  if (warp == 1) { if (CG2) tmem_dealloc_2cta(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols); }
  if (dbg != nullptr && threadIdx.x == 0) dbg[7] = clock64();
}

// Operand-ring budget of a non-persistent CTA.  TWO CTAs must fit an SM: 228 KB per SM minus 1 KB the driver reserves per
// CTA = 113 KB (115,712 B) each, INCLUDING the ~9.25 KB of barriers / parameter vectors / exchange behind the ring.  (Round 1
// sized the ring alone at 110 KB: a 4 x 26 KB ring for N = 80 came to 115,968 B -- 256 B over -- and the last PostNet conv and
// mel_linear silently ran one CTA per SM: 0.215 ms at B=128.)
constexpr int kFixedSmemBytes = 1024 /*align*/ + 256 /*barriers*/ + 4096 /*params*/ + 4096 /*exchange*/;
int smem_budget_bytes() {
  const int total = tuning(TUNE_TC_SMEM_KB) * 1024;
  return (total < 115712 ? total : 115712) - kFixedSmemBytes;
}
// persistent form: 0 = one tile per CTA everywhere, 1 (default) = where two CTAs per SM still fit (two accumulators of <= 128
// TMEM columns), 2 = also the one-CTA-per-SM form (N = 256 LayerNorm rows; measured slower: out-proj 0.037 -> 0.045 ms, its
// epilogue is issue-bound and loses the second CTA's warps)
int persist_mode() { return tuning(TUNE_TC_PERSIST); }
// CTA pairs (cta_group::2, M = 256) for the big bf16 convolutions: 0 = never, 1 (default) = where it pays, 2 = wherever legal
int cg2_mode() { return tuning(TUNE_TC_2CTA); }

int pick_bn(const styler_conv1d_args& a, int m_tiles) {
  if (a.ln_gamma != nullptr || a.dot_w != nullptr) return (a.N <= 256 && a.N % 16 == 0) ? a.N : 0;
  const int es0 = a.dtype != STYLER_F32 ? 2 : 4;
  // full-width tile for 256 < N <= 512 that is not a multiple of 256 (the 320-channel convs): see the MMA issuer
  if (tuning(TUNE_TC_WIDE) != 0 && es0 == 2 && a.N > 256 && a.N <= 512 && a.N % 256 != 0 && a.N % 32 == 0 && (a.N * es0) % 128 == 0 &&
      a.out != nullptr && a.vt == nullptr && a.out_f32 == nullptr && m_tiles >= num_sms())
    return a.N;
  const int forced = tuning(TUNE_TC_BN);   // tuning override: STYLER_TC_BN
  if (forced > 0 && forced % 16 == 0 && forced <= 256 && a.N % forced == 0 && (a.vt == nullptr || a.vt_col0 % forced == 0))
    return forced;
  // Prefer tiles whose rows are whole 128-byte boxes (coalesced TMA epilogue); among those the largest tile that still
  // gives >= 2 waves of CTAs, otherwise the smallest tile >= 64 (more CTAs), otherwise the largest tile available.
  const int es = a.dtype != STYLER_F32 ? 2 : 4;
  for (int pass = 0; pass < 2; ++pass) {
    int largest = 0, smallest64 = 0;
    for (int bn = 256; bn >= 16; bn -= 16) {
      if (a.N % bn != 0) continue;
      if (a.vt != nullptr && a.vt_col0 % bn != 0) continue;
      if (pass == 0) {   // coalesced-epilogue candidates: whole boxes and a tile that fits the pipeline smem
        const int sb = kAStageBytes + bn * 128;
        int st = smem_budget_bytes() / sb;
        st = st > 8 ? 8 : (st < 2 ? 2 : st);
        if ((bn * es) % 128 != 0 || a.out == nullptr || static_cast<long long>(bn) * es * 128 > static_cast<long long>(st) * sb) continue;
      }
      if (largest == 0) largest = bn;
      if (static_cast<long long>(m_tiles) * (a.N / bn) >= 296) return bn;
      if (bn >= 64) smallest64 = bn;
    }
    if (smallest64 != 0) return smallest64;
    if (largest != 0) return largest;
  }
  return 0;
}

template <typename T>
int launch(const styler_conv1d_args& a, cudaStream_t stream) {
  constexpr int es = sizeof(T);
  constexpr int bke = 128 / es;
  const int tiles_per_utt = ceil_div(a.T, kBM);
  const int m_tiles = a.B * tiles_per_utt;
  const int kb_per_tap = ceil_div(a.Cin, bke);
  const int num_kb = a.KS * kb_per_tap;
  const bool has_ln = a.ln_gamma != nullptr;
  // Persistent form (see the kernel comment): bf16, staged epilogue only, short mainloops, more tiles than CTA slots.
  const bool fast_like = a.out != nullptr && a.vt == nullptr && a.out_f32 == nullptr && a.dot_w == nullptr &&
                         (a.residual == nullptr || (!a.residual_is_f32 && a.r_ld != 0));
  bool persist = persist_mode() != 0 && es == 2 && fast_like && num_kb <= 16;
  int BN = 0, ctas_per_sm = 2;
  if (persist) {
    if (has_ln) BN = (a.N <= 256 && a.N % 64 == 0) ? a.N : 0;
    else BN = a.N % 128 == 0 ? 128 : ((a.N <= 128 && a.N % 64 == 0) ? a.N : 0);
    if (BN == 0) persist = false;
    else ctas_per_sm = 2 * tmem_cols_pow2(BN) <= 256 ? 2 : 1;     // two accumulators per CTA, 512 TMEM columns per SM
    if (ctas_per_sm == 1 && persist_mode() != 2) persist = false;
    if (persist && static_cast<long long>(m_tiles) * (a.N / BN) <= static_cast<long long>(ctas_per_sm) * num_sms()) persist = false;
  }
  if (!persist) BN = pick_bn(a, m_tiles);
  SB_REQUIRE(BN > 0, "conv1d_tc: no valid N tile for N=%d", a.N);
  const int n_tiles = a.N / BN;
  const int total_tiles = m_tiles * n_tiles;
  // CTA pairs (see the kernel comment): long mainloops on a 256-wide N tile, staged epilogue without residual, an even number
  // of M tiles, and enough tiles that pairing costs no occupancy
  // (LayerNorm rows, N = 256, included: their B tile is 2/3 of the operand bytes and their 48 KB stages only fit twice)
  const bool cg2 = cg2_mode() != 0 && es == 2 && !persist && fast_like && BN == 256 && (m_tiles % 2) == 0 &&
                   (cg2_mode() == 2 || (num_kb >= 8 && total_tiles >= 2 * num_sms()));   // (K = 256 out-proj measured slower paired)
  const int stage_bytes = kAStageBytes + (cg2 ? BN / 2 : BN) * 128;
  const int staging_bytes = persist ? BN * es * kBM : 0;
  const bool wide = BN > 256;
  const int fixed_bytes = kFixedSmemBytes + (wide ? 4096 : 0) /*wide tile: 512-column parameter vectors*/;
  int stages = ((persist ? (ctas_per_sm == 2 ? 113 : 226) * 1024 - staging_bytes - fixed_bytes
                         : (wide ? 200 * 1024 : smem_budget_bytes()))) / stage_bytes;
  if (stages > (persist ? 4 : 8)) stages = persist ? 4 : 8;
  if (!persist && stages > num_kb) stages = num_kb;
  if (stages < 2) stages = (persist || num_kb >= 2) ? 2 : 1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + staging_bytes + fixed_bytes;
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
    const uint32_t box[3] = {static_cast<uint32_t>(bke), static_cast<uint32_t>((cg2 || BN > 256) ? BN / 2 : BN), 1};
    int rc = make_tmap(&tmB, a.w, es == 2 ? 1 : 0, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  // Coalesced epilogue I/O through the (by then idle) pipeline smem: needs whole 128-byte boxes and room for the tile.
  const bool boxes_ok = (BN * es) % 128 == 0 && (persist || static_cast<size_t>(BN) * es * 128 <= static_cast<size_t>(stages) * stage_bytes);
  const bool stage_out = a.out != nullptr && boxes_ok;
  const bool stage_res = a.residual != nullptr && !a.residual_is_f32 && a.r_ld != 0 && boxes_ok;
  CUtensorMap tmOut = tmA, tmRes = tmA;
  if (stage_out) {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.vt != nullptr ? a.vt_col0 : a.N), static_cast<uint64_t>(a.T),
                              static_cast<uint64_t>(a.B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.o_ld) * es,
                                 static_cast<uint64_t>(a.B > 1 ? a.o_bstride : static_cast<int64_t>(a.o_ld) * a.T) * es};
    const uint32_t box[3] = {static_cast<uint32_t>(bke), static_cast<uint32_t>(kBM), 1};
    int rc = make_tmap(&tmOut, a.out, es == 2 ? 1 : 2, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  if (stage_res) {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.T), static_cast<uint64_t>(a.B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.r_ld) * es,
                                 static_cast<uint64_t>(a.B > 1 ? a.r_bstride : static_cast<int64_t>(a.r_ld) * a.T) * es};
    const uint32_t box[3] = {static_cast<uint32_t>(bke), static_cast<uint32_t>(kBM), 1};
    int rc = make_tmap(&tmRes, a.residual, es == 2 ? 1 : 2, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  EpiParams ep;
  ep.stage_out = stage_out ? 1 : 0;
  ep.stage_res = stage_res ? 1 : 0;
  ep.dbg = (!persist && g_phase_buf != nullptr && total_tiles <= g_phase_cap) ? g_phase_buf : nullptr;
  ep.bias = a.bias; ep.act = a.act; ep.act2 = a.act2;
  ep.slope = a.act_slope; ep.inv_slope = a.act_slope > 0.f ? 1.0f / a.act_slope : 1.0f; ep.res_inv = a.residual_inv_lrelu;
  ep.residual = a.residual; ep.r_bstride = a.r_bstride; ep.r_ld = a.r_ld; ep.res_f32 = a.residual_is_f32;
  ep.ln_gamma = a.ln_gamma; ep.ln_beta = a.ln_beta; ep.ln_eps = a.ln_eps;
  ep.lens = a.lens;
  ep.dot_w = a.dot_w; ep.dot_b = a.dot_b; ep.dot_out = a.dot_out;
  ep.out = a.out; ep.o_bstride = a.o_bstride; ep.o_ld = a.o_ld;
  ep.out_f32 = a.out_f32; ep.of_bstride = a.of_bstride; ep.of_ld = a.of_ld; ep.out2_f32 = a.out2_f32;
  ep.gn_partial = a.gn_partial; ep.n_total = a.N;
  // contiguous-tile staging for narrow fp32 outputs (see EpiParams): non-staged epilogue, one N tile, full rows everywhere
  ep.contig_f32 = (!stage_out && a.out_f32 != nullptr && n_tiles == 1 && a.of_ld == a.N && a.N % 4 == 0 && a.vt == nullptr &&
                   a.dot_w == nullptr && !has_ln && (a.out == nullptr || a.o_ld == a.N) &&
                   (a.residual == nullptr || (a.residual_is_f32 && a.r_ld == a.N)) &&
                   static_cast<size_t>(2) * kBM * (a.N + 4) * sizeof(float) <= static_cast<size_t>(stages) * stage_bytes &&
                   (a.out == nullptr || (reinterpret_cast<uintptr_t>(a.out) % 8 == 0 && (a.o_bstride * es) % 8 == 0)))
                      ? 1 : 0;
  ep.vt = a.vt; ep.vt_col0 = a.vt_col0; ep.vt_bstride = a.vt_bstride; ep.vt_ld = a.vt_ld;

  using KernFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, EpiParams, int, int, int, int, int, int, int, int, int, int);
#define SB_K(A, L, F, P) conv1d_tc_kernel<T, A, L, F, P>
#define SB_KROW(A, L) {{SB_K(A, L, false, false), nullptr}, {SB_K(A, L, true, false), SB_K(A, L, true, true)}}
  static const KernFn table[4][2][2][2] = {{SB_KROW(STYLER_ACT_NONE, false), SB_KROW(STYLER_ACT_NONE, true)},
                                           {SB_KROW(STYLER_ACT_RELU, false), SB_KROW(STYLER_ACT_RELU, true)},
                                           {SB_KROW(STYLER_ACT_TANH, false), SB_KROW(STYLER_ACT_TANH, true)},
                                           {SB_KROW(STYLER_ACT_LRELU, false), SB_KROW(STYLER_ACT_LRELU, true)}};
#undef SB_KROW
#undef SB_K
  static DeviceFlags attr_set[4][2][2][2];
  const int ia = a.act, il = has_ln ? 1 : 0;
  const int ifast = (stage_out && a.vt == nullptr && a.out_f32 == nullptr && a.dot_w == nullptr &&
                     (a.residual == nullptr || stage_res)) ? 1 : 0;
  const int ip = persist ? 1 : 0;
  SB_REQUIRE(!persist || ifast == 1, "conv1d_tc: internal: persistent form chosen for a non-staged epilogue");
  if constexpr (es == 2) {
    if (cg2) {
      SB_REQUIRE(ifast == 1, "conv1d_tc: internal: CTA-pair form chosen for a non-staged epilogue");
      static const KernFn table2[4][2] = {
          {conv1d_tc_kernel<T, STYLER_ACT_NONE, false, true, false, true>, conv1d_tc_kernel<T, STYLER_ACT_NONE, true, true, false, true>},
          {conv1d_tc_kernel<T, STYLER_ACT_RELU, false, true, false, true>, conv1d_tc_kernel<T, STYLER_ACT_RELU, true, true, false, true>},
          {conv1d_tc_kernel<T, STYLER_ACT_TANH, false, true, false, true>, conv1d_tc_kernel<T, STYLER_ACT_TANH, true, true, false, true>},
          {conv1d_tc_kernel<T, STYLER_ACT_LRELU, false, true, false, true>, conv1d_tc_kernel<T, STYLER_ACT_LRELU, true, true, false, true>}};
      static DeviceFlags attr_set2[4][2];
      KernFn kern2 = table2[ia][il];
      SB_OPT_IN_SMEM(attr_set2[ia][il], kern2, 227 * 1024);
      // grid = one CTA per 128-row tile as before, launched as clusters of two; the kernel counts tiles in PAIRS
      SB_CUDA_OK(launch_cluster2(kern2, dim3(total_tiles), dim3(kThreads), smem, stream, tmA, tmB, tmOut, tmRes, ep, a.T, n_tiles,
                                 tiles_per_utt, a.KS, a.pad, kb_per_tap, BN, stages, total_tiles / 2, a.dilation > 1 ? a.dilation : 1));
      SB_LAUNCH_OK();
      return 0;
    }
  }
  KernFn kern = table[ia][il][ifast][ip];
  SB_OPT_IN_SMEM(attr_set[ia][il][ifast][ip], kern, 227 * 1024);
  const int grid = persist ? (total_tiles < ctas_per_sm * num_sms() ? total_tiles : ctas_per_sm * num_sms()) : total_tiles;
  SB_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, tmA, tmB, tmOut, tmRes, ep, a.T, n_tiles,
                        tiles_per_utt, a.KS, a.pad, kb_per_tap, BN, stages, total_tiles, a.dilation > 1 ? a.dilation : 1));
  SB_LAUNCH_OK();
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool conv1d_tc_supported(const styler_conv1d_args& a, const char** why) {
  const int es = a.dtype != STYLER_F32 ? 2 : 4;
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
  if (a.out2_f32 != nullptr && (a.out_f32 == nullptr || !aligned16(a.out2_f32))) return fail("out2_f32 needs out_f32 and 16-byte alignment");
  if (a.gn_partial != nullptr && (a.ln_gamma != nullptr || a.dot_w != nullptr || a.vt != nullptr || a.N % 16 != 0))
    return fail("gn_partial: plain conv epilogues with N % 16 == 0 only");
  const int res_es = a.residual_is_f32 ? 4 : es;
  if (a.residual != nullptr && (!aligned16(a.residual) || (static_cast<int64_t>(a.r_ld) * res_es) % 16 != 0 || (a.r_bstride * res_es) % 16 != 0))
    return fail("residual not 16-byte aligned/strided");
  if (a.vt != nullptr && a.vt_col0 % 16 != 0) return fail("vt_col0 not a multiple of 16");
  if (a.T < 1 || a.B < 1) return fail("empty problem");
  if (a.act < 0 || a.act > 3) return fail("bad act");
  if (a.act2 != STYLER_ACT_NONE && a.act2 != STYLER_ACT_RELU && a.act2 != STYLER_ACT_LRELU)
    return fail("act2 must be none|relu|lrelu on the tensor-core path");
  return true;
}

void set_phase_buffer(long long* buf, int cap) { g_phase_buf = buf; g_phase_cap = cap; }

int conv1d_tc(const styler_conv1d_args& a, cudaStream_t s) {
  const char* why = nullptr;
  SB_REQUIRE(conv1d_tc_supported(a, &why), "conv1d_tc: unsupported arguments: %s", why ? why : "?");
  if (a.gn_partial == nullptr && conv1d_win_supported(a)) return conv1d_win(a, s);
  if (a.dtype == STYLER_BF16) return launch<__nv_bfloat16>(a, s);
  if (a.dtype == STYLER_F16) return launch<__half>(a, s);
  return launch<float>(a, s);
}

}  // namespace sb
