// Window form of the implicit-GEMM Conv1d for FEW channels and MANY taps (HiFi-GAN resblocks at 64 / 32 channels,
// k = 3/7/11, dilation 1/3/5; hifigan/models.py:20-98).
//
// conv_tc.cu loads one shifted [128 x Cin] A tile per tap.  With N <= 64 the MMA of a k-block lasts 16-32 cycles while its
// operands are 12-24 KB: the per-tap re-load makes the kernel L2-bandwidth-bound (measured 506 us for a 32-channel k=11
// conv that moves 0.4 GB of HBM).  Here a persistent CTA
//   * keeps ALL taps of the weights resident in smem (KS x N x 128 B <= 90 KB; loaded once),
//   * loads the input WINDOW of a 128-step tile once: rows t0-pad .. t0-pad+127+(KS-1)*dil, one TMA box, double-buffered,
//   * issues tap j's MMAs with the A descriptor advanced by j*dil ROWS inside that window: the 128B swizzle is a function
//     of the absolute shared-memory address, so a row offset - like the usual +32 B K advance - lands on the bytes TMA
//     wrote (verified on the B200: descriptor base-offset field left 0; setting it to row & 7 gives wrong results),
//   * accumulates in one of two TMEM accumulators while the 8 epilogue warps drain the other
//     (bias, leaky ReLU, activated-residual add, leaky ReLU; 16-byte global loads/stores).
// Traffic per tile drops from KS x 12 KB to one 11-23 KB window; measured: HiFi-GAN V1 forward 14.2 -> 9.2 ms (B=8 x 1024 frames).
// bf16 only; Cin in {16..64}, N in {16..64}; everything else stays on conv_tc.cu.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <stdlib.h>

namespace sb {
namespace {

constexpr int kWM = 128;          // output time steps per tile
constexpr int kWThreads = 320;    // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

struct WinParams {
  const float* bias;
  int act, act2;
  float slope, inv_slope;
  const __nv_bfloat16* residual; long long r_bstride; int r_ld; int res_inv;
  const int64_t* lens;
  __nv_bfloat16* out; long long o_bstride; int o_ld;
};

__global__ void __launch_bounds__(kWThreads, 2) conv1d_win_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                  const __grid_constant__ CUtensorMap tmW, const WinParams ep,
                                                                  int Tlen, int tiles_per_utt, int total_tiles, int KS,
                                                                  int pad, int dil, int Cin, int N, int win_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int w_tap_bytes = N * 128;                       // one tap of the weights: [N][64 ch] bf16, 128B-swizzled
  const int win_bytes = win_rows * 128;                  // one window: [win_rows][64 ch] bf16 (win_rows % 8 == 0)
  uint8_t* sW = smem;
  uint8_t* sWin = sW + KS * w_tap_bytes;                 // KS * N * 128 is a multiple of 1024 for N % 8 == 0
  uint8_t* tail = sWin + 2 * win_bytes;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* win_full = w_full + 1;                       // [2]
  uint64_t* win_empty = win_full + 2;                    // [2]
  uint64_t* tmem_full = win_empty + 2;                   // [2]
  uint64_t* tmem_empty = tmem_full + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 128);  // [N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t acc_cols = tmem_cols_pow2(N);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&win_full[i], 1); mbar_init(&win_empty[i], 1);
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * acc_cols);
  if (threadIdx.x >= 64 && static_cast<int>(threadIdx.x) - 64 < N) s_bias[threadIdx.x - 64] = ep.bias != nullptr ? ep.bias[threadIdx.x - 64] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, KS * w_tap_bytes);
      for (int tap = 0; tap < KS; ++tap) tma_load_3d(sW + tap * w_tap_bytes, &tmW, w_full, 0, 0, tap);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int b = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * kWM;
        mbar_wait(&win_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&win_full[buf], win_bytes);
        tma_load_3d(sWin + buf * win_bytes, &tmX, &win_full[buf], 0, t0 - pad, b);   // rows outside [0,T) arrive as zeros
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_BF16, kWM, N);
      const int ksteps = (Cin + 15) / 16;                // 16-channel MMA steps that hold real channels (2 for Cin=32)
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc = tmem_base + buf * acc_cols;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_wait(&win_full[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t win_addr = smem_u32(sWin + buf * win_bytes);
        const uint32_t w_addr = smem_u32(sW);
        for (int tap = 0; tap < KS; ++tap) {
          const uint32_t row = static_cast<uint32_t>(tap * dil);
          for (int k = 0; k < ksteps; ++k) {
            umma_ss<false>(acc, umma_desc_k_sw128(win_addr + row * 128 + k * 32),
                           umma_desc_k_sw128(w_addr + tap * w_tap_bytes + k * 32), idesc, (tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&win_empty[buf]);
        umma_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ncol = N / 2;                              // columns per thread (N % 32 == 0 -> ncol % 16 == 0)
    const int c0 = hf * ncol;
    const bool relu2 = ep.act2 == STYLER_ACT_RELU || ep.act2 == STYLER_ACT_LRELU;
    const float slope2 = ep.act2 == STYLER_ACT_LRELU ? ep.slope : 0.f;
    const float slope1 = ep.act == STYLER_ACT_LRELU ? ep.slope : 0.f;
    const bool relu1 = ep.act == STYLER_ACT_RELU || ep.act == STYLER_ACT_LRELU;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * kWM;
      const int t = t0 + r;
      const bool row_ok = t < Tlen;
      const bool masked = ep.lens != nullptr && row_ok && t >= static_cast<int>(ep.lens[b]);
      const uint32_t taddr = tmem_base + buf * acc_cols + (static_cast<uint32_t>(q * 32) << 16) + c0;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < ncol; c += 16) {
        uint32_t ra[16];
        tmem_ld16(taddr + c, ra);
        float rr[16];
        const bool do_res = ep.residual != nullptr && row_ok;
        if (do_res) {
          const __nv_bfloat16* rp = ep.residual + b * ep.r_bstride + static_cast<long long>(t) * ep.r_ld + c0 + c;
          float a8[8], b8[8];
          load8(rp, a8);
          load8(rp + 8, b8);
#pragma unroll
          for (int i = 0; i < 8; ++i) { rr[i] = a8[i]; rr[8 + i] = b8[i]; }
          if (ep.res_inv != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) rr[i] = rr[i] < 0.f ? rr[i] * ep.inv_slope : rr[i];
          }
        }
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(ra[i]) + s_bias[c0 + c + i];
          if (relu1) x = fmaxf(x, x * slope1);
          if (do_res) x += rr[i];
          if (relu2) x = fmaxf(x, x * slope2);
          v[i] = masked ? 0.f : x;
        }
        if (row_ok) {
          __nv_bfloat16* op = ep.out + b * ep.o_bstride + static_cast<long long>(t) * ep.o_ld + c0 + c;
          float a8[8], b8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a8[i] = v[i]; b8[i] = v[8 + i]; }
          store8(op, a8);
          store8(op + 8, b8);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * acc_cols);
}

int win_mode() {   // STYLER_CONV_WIN=0 sends these shapes back to conv_tc.cu (A/B measurements)
  return tuning(TUNE_CONV_WIN);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool conv1d_win_supported(const styler_conv1d_args& a) {
  if (win_mode() == 0 || a.dtype != STYLER_BF16) return false;
  if (a.KS < 2 || a.Cin < 16 || a.Cin > 64 || a.Cin % 8 != 0 || a.N < 32 || a.N > 64 || a.N % 32 != 0) return false;
  if (a.ln_gamma != nullptr || a.dot_w != nullptr || a.vt != nullptr || a.out_f32 != nullptr || a.out == nullptr) return false;
  if (a.residual != nullptr && (a.residual_is_f32 || a.r_ld == 0)) return false;
  if (a.act == STYLER_ACT_TANH || a.act2 == STYLER_ACT_TANH) return false;
  const int dil = a.dilation > 1 ? a.dilation : 1;
  const int win_rows = (kWM + (a.KS - 1) * dil + 7) / 8 * 8;
  if (win_rows > 256) return false;
  const size_t smem = static_cast<size_t>(a.KS) * a.N * 128 + 2ull * win_rows * 128 + 1024 + 128 + 256;
  if (smem > 200 * 1024) return false;
  if (!al16(a.x) || !al16(a.w) || !al16(a.out) || (a.x_ld * 2) % 16 != 0 || (a.x_bstride * 2) % 16 != 0 || (a.o_ld * 2) % 16 != 0 ||
      (a.o_bstride * 2) % 16 != 0)
    return false;
  if (a.residual != nullptr && (!al16(a.residual) || (a.r_ld * 2) % 16 != 0 || (a.r_bstride * 2) % 16 != 0)) return false;
  return static_cast<long long>(a.B) * ceil_div(a.T, kWM) >= 2LL * num_sms();   // persistent: needs enough tiles
}

int conv1d_win(const styler_conv1d_args& a, cudaStream_t stream) {
  const int dil = a.dilation > 1 ? a.dilation : 1;
  const int win_rows = (kWM + (a.KS - 1) * dil + 7) / 8 * 8;
  const int tiles_per_utt = ceil_div(a.T, kWM);
  const int total_tiles = a.B * tiles_per_utt;
  const size_t smem = static_cast<size_t>(a.KS) * a.N * 128 + 2ull * win_rows * 128 + 1024 /*align*/ + 128 /*barriers*/ + 256 /*bias*/;
  CUtensorMap tmX, tmW;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.Cin), static_cast<uint64_t>(a.T), static_cast<uint64_t>(a.B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.x_ld) * 2,
                                 static_cast<uint64_t>(a.B > 1 ? a.x_bstride : static_cast<int64_t>(a.x_ld) * a.T) * 2};
    const uint32_t box[3] = {64, static_cast<uint32_t>(win_rows), 1};
    int rc = make_tmap(&tmX, a.x, 1, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.Cin), static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.KS)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.Cin) * 2, static_cast<uint64_t>(a.Cin) * a.N * 2};
    const uint32_t box[3] = {64, static_cast<uint32_t>(a.N), 1};
    int rc = make_tmap(&tmW, a.w, 1, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  WinParams ep;
  ep.bias = a.bias; ep.act = a.act; ep.act2 = a.act2;
  ep.slope = a.act_slope; ep.inv_slope = a.act_slope > 0.f ? 1.0f / a.act_slope : 1.0f;
  ep.residual = static_cast<const __nv_bfloat16*>(a.residual); ep.r_bstride = a.r_bstride; ep.r_ld = a.r_ld;
  ep.res_inv = a.residual_inv_lrelu;
  ep.lens = a.lens;
  ep.out = static_cast<__nv_bfloat16*>(a.out); ep.o_bstride = a.o_bstride; ep.o_ld = a.o_ld;
  static DeviceFlags attr_set;
  SB_OPT_IN_SMEM(attr_set, conv1d_win_kernel, 200 * 1024);
  const int ctas_per_sm = smem <= 112 * 1024 ? 2 : 1;
  const int cap = ctas_per_sm * num_sms();
  const int grid = total_tiles < cap ? total_tiles : cap;
  conv1d_win_kernel<<<grid, kWThreads, smem, stream>>>(tmX, tmW, ep, a.T, tiles_per_utt, total_tiles, a.KS, a.pad, dil, a.Cin,
                                                       a.N, win_rows);
  SB_LAUNCH_OK();
  return 0;
}

}  // namespace sb
