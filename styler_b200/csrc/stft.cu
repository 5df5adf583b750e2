// TacotronSTFT.mel_spectrogram (audio/stft.py:51-79,141-160; audio_processing.py:80-86 of the reference).
// The reference evaluates a dense 1026x1024 DFT as a strided Conv1d (2.1 MFLOP/frame) and bounces through
// the host.  Here persistent CTAs (three per SM) walk work items = 16 consecutive frames of one utterance, one frame per
// warp at a time:
//   the frame's 1024 reflect-padded samples go straight from global memory into the registers of pass 1 (16 coalesced
//   8-byte loads per lane; the 4x overlap between frames is served by L1/L2) -- or, in the staged shapes, from a block
//   window staged once in smem --, periodic-Hann window from an smem table, 1024-point real FFT as a 512-point complex
//   Stockham FFT (three radix-8 passes, 16 points per lane in registers, warp-private XOR-swizzled smem exchange) + split
//   post-pass, magnitude for the 513 bins, energy = ||mag||_2, mel = log(max(basis @ mag, 1e-5)) over the non-zero band of
//   each filter row (bands cached in smem once per CTA), results staged in a double-buffered smem tile and stored as runs
//   of 16 frames per mel row.
// Bounding roofline: HBM by contract (464,580 algorithmic bytes per 4 s utterance); in practice the kernel is latency /
// issue bound (1.7 k warp-instructions per frame, issue slots 61 % busy at 6 warps per scheduler): profiles/ncu_stft_r2_final.md.
#include "common.cuh"

namespace sb {
namespace {

constexpr int NFFT = 1024, HOP = 256, NBINS = 513, HALF = 512;
constexpr int WBUF = HALF;                           // complex work buffer per warp, XOR-swizzled (slot() below): every 8-byte
                                                     // access pattern of the three passes hits 16 distinct banks per half-warp
constexpr int MAGLD = NBINS + 3;
// Shapes of the same arithmetic (bitwise-equal results, tests/test_kernels_gpu.py; STFT_OCC selects):
//   0 <32, 8, 2, staged>: 32 frames per item, separate magnitude buffer, 109 KB of smem -> 2 CTAs (16 warps) per SM, 126 registers;
//   1 <16, 8, 3, staged>: 16 frames per item, the magnitudes overwrite the warp's FFT exchange buffer (the split pass keeps its 18
//                         results in registers across one __syncwarp), 74 KB of smem and <= 85 registers -> 3 CTAs (24 warps) per SM;
//   2 <24, 12, 2, staged>: 24 frames per item, two CTAs of twelve warps;
//   3 <16, 8, 3, direct>: shape 1 without the staged block window (58 KB): every warp reads its frame from global memory itself.
// Measured at configs[3] (256 x 4 s): 0.251 / 0.241 / 0.241 / 0.226 ms.  The kernel is latency-bound: the extra warps are the point.
template <int FPB> struct StftShape {
  static constexpr int NSAMP = (FPB - 1) * HOP + NFFT;
  static constexpr int MELLD = FPB + 1;
};
constexpr int kBasisCap = 1536;                      // floats of smem for the non-zero bands of the mel basis (727 used by the
                                                     // reference's 80-mel Slaney basis); larger bases are read from global

__global__ void mel_band_kernel(const float* __restrict__ basis, int n_mels, int32_t* __restrict__ band) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= n_mels) return;
  int lo = NBINS, hi = -1;
  for (int k = lane; k < NBINS; k += 32)
    if (basis[static_cast<long long>(m) * NBINS + k] != 0.f) { lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { band[2 * m] = hi >= 0 ? lo : 0; band[2 * m + 1] = hi >= 0 ? hi + 1 : 0; }
}

__device__ __forceinline__ float sqrt_approx(float x) {   // MUFU.SQRT (2 ulp): the IEEE sqrtf costs ~8 instructions per magnitude
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Complex arithmetic on packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: one issue slot per COMPLEX add, two per complex
// multiply).  ptxas folds the component swaps and sign flips below into operand modifiers (.F32 broadcast, .F32x2.LO_HI, .NP):
// no register moves.  The scalar round-1 kernel spent 57 % of its 2.1 k warp-instructions per frame on FADD/FMUL/FFMA.
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {       // (a.x w.x - a.y w.y, a.x w.y + a.y w.x)
  const float2 t = __fmul2_rn(make_float2(a.x, a.x), w);
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(-w.y, w.x), t);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i): folded into the consumer

// 4-point forward DFT, natural order in and out
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = mul_mi(csub(a1, a3));
  a0 = cadd(s0, s2); a1 = cadd(s1, s3); a2 = csub(s0, s2); a3 = csub(s1, s3);
}
// 8-point forward DFT in registers; v[] natural order in, natural order out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  constexpr float kR = 0.70710678118654752f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = __fmul2_rn(cadd(b1, mul_mi(b1)), make_float2(kR, kR));        // * (1 - i)/sqrt(2):  ((x + y) r, (y - x) r)
  b2 = mul_mi(b2);
  b3 = __fmul2_rn(csub(b3, mul_mi(b3)), make_float2(-kR, -kR));      // * (-1 - i)/sqrt(2): ((y - x) r, -(x + y) r)
  dft4(a0, a1, a2, a3);
  dft4(b0, b1, b2, b3);
  v[0] = a0; v[1] = b0; v[2] = a1; v[3] = b1; v[4] = a2; v[5] = b2; v[6] = a3; v[7] = b3;
}

// One warp = one frame: 512-point complex Stockham FFT as three radix-8 passes (Ns = 1, 8, 64), each lane owning two
// 8-point butterflies per pass held in registers; the passes exchange data through a per-warp smem buffer, so the only
// synchronisation inside the transform is __syncwarp().
template <int FPB, int KW, int MINB, bool ALIAS_MAG, bool DIRECT>
__global__ void __launch_bounds__(KW * 32, MINB) stft_mel_kernel(const float* __restrict__ y, int N, int F,
                                                                  const float* __restrict__ basis,
                                                                  const int32_t* __restrict__ band, int n_mels,
                                                                  float* __restrict__ mel, float* __restrict__ energy,
                                                                  float in_scale, int clamp, int32_t* __restrict__ clip_flag,
                                                                  int frame_major, float* __restrict__ e_input, float e_min,
                                                                  float e_inv_range, const int64_t* __restrict__ n_samples,
                                                                  int n_items) {
  constexpr int NSAMP = StftShape<FPB>::NSAMP, MELLD = StftShape<FPB>::MELLD;
  extern __shared__ __align__(16) uint8_t stft_smem[];
  float2* tw = reinterpret_cast<float2*>(stft_smem);                       // W_1024^k, k < 512
  float2* twA = tw + HALF;                                                 // W_64^k,  k < 8   (pass-2 base twiddles)
  float2* twB = twA + 8;                                                   // W_512^k, k < 64  (pass-3 base twiddles)
  float2* win = twB + 64;                                                  // periodic Hann window as pairs (w[2n], w[2n+1]), n < 512
  float2* wbuf = win + HALF;                                               // [KW][WBUF]
  float* samp = reinterpret_cast<float*>(wbuf + KW * WBUF);                // [NSAMP] (absent when DIRECT)
  float* mag = samp + (DIRECT ? 0 : NSAMP);                                // [KW][MAGLD] (absent when ALIAS_MAG)
  float* s_en0 = mag + (ALIAS_MAG ? 0 : KW * MAGLD);                       // [FPB], two buffers when DIRECT
  float* s_basis = s_en0 + (DIRECT ? 2 : 1) * FPB;                         // [kBasisCap] bands of the mel basis, back to back
  int* s_lo = reinterpret_cast<int*>(s_basis + kBasisCap);                 // [n_mels] first bin of the band
  int* s_hi = s_lo + n_mels;                                               // [n_mels] one past the last bin
  int* s_off = s_hi + n_mels;                                              // [n_mels + 1] offset of the band in s_basis
  float* s_mel0 = reinterpret_cast<float*>(s_off + n_mels + 1);            // [n_mels][MELLD], two buffers when DIRECT
  // ---- once per CTA: twiddles, window, band tables, basis bands.  The CTA is PERSISTENT: it walks work items (utterance,
  // block of FPB frames) item = blockIdx.x + i * gridDim.x.  With one item per CTA this set-up (512 sincospif, the band scan, ten
  // dependent global reads per warp for the basis bands, three block barriers) was 21 % of the kernel's instructions and 47 %
  // of its warp-state samples (profiles/ncu_stft_r3e.md).
  const int N_row = N;
  for (int k = threadIdx.x; k < HALF; k += KW * 32) {
    float sn, cs;
    sincospif(-static_cast<float>(k) / 512.0f, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  if (threadIdx.x < 72) {
    const int k = threadIdx.x < 8 ? threadIdx.x : threadIdx.x - 8;
    float sn, cs;
    sincospif(-static_cast<float>(k) / (threadIdx.x < 8 ? 32.0f : 256.0f), &sn, &cs);
    (threadIdx.x < 8 ? twA : twB)[k] = make_float2(cs, sn);
  }
  for (int m = threadIdx.x; m < n_mels; m += KW * 32) { s_lo[m] = band[2 * m]; s_hi[m] = band[2 * m + 1]; }
  __syncthreads();
  for (int n = threadIdx.x; n < HALF; n += KW * 32) {
    // periodic Hann: 0.5 - 0.5 cos(2 pi i / 1024), cos(2 pi i / 1024) = Re W^i = -Re W^(i-512); one 8-byte table read per
    // sample pair in pass 1 (reading the twiddle table there cost a 16-byte read + an FFMA2 per pair)
    const float4 c = reinterpret_cast<const float4*>(tw)[n & (HALF / 2 - 1)];   // (Re W^2n, Im W^2n, Re W^(2n+1), Im W^(2n+1))
    const float sg = n < HALF / 2 ? -0.5f : 0.5f;
    win[n] = __ffma2_rn(make_float2(sg, sg), make_float2(c.x, c.z), make_float2(0.5f, 0.5f));
  }
  if (threadIdx.x < 32) {   // exclusive prefix of the band widths: warp scan, 32 rows at a time
    int base = 0;
    for (int m0 = 0; m0 < n_mels; m0 += 32) {
      const int m = m0 + threadIdx.x;
      const int w = m < n_mels ? s_hi[m] - s_lo[m] : 0;
      int incl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (static_cast<int>(threadIdx.x) >= o) incl += t;
      }
      if (m < n_mels) s_off[m] = base + incl - w;
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_off[n_mels] = base;
  }
  __syncthreads();
  const bool basis_in_smem = s_off[n_mels] <= kBasisCap;
  if (basis_in_smem) {
    for (int m = threadIdx.x >> 5; m < n_mels; m += KW) {
      const int lo = s_lo[m], w = s_hi[m] - lo;
      for (int i = threadIdx.x & 31; i < w; i += 32) s_basis[s_off[m] + i] = basis[static_cast<long long>(m) * NBINS + lo + i];
    }
  }
  if (DIRECT) __syncthreads();   // (staged form: the first item's post-staging barrier publishes the tables above)
  int parity = 0;                // DIRECT: which output tile (s_mel / s_en) this item fills

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks_per_utt = (F + FPB - 1) / FPB;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int b = item / blocks_per_utt, f0 = (item - b * blocks_per_utt) * FPB;
  const float* yb = y + static_cast<long long>(b) * N_row;
  // Per-utterance length (audio/tools.py:37-55 runs every utterance alone): row b of the zero-padded batch holds Nb valid
  // samples, is reflected around ITS OWN end and yields Fb = 1 + Nb/hop frames; frames >= Fb are written as zeros (the
  // collation padding of dataset.py:160-166).  `N_row` stays the row stride and `F` the padded frame count of the outputs.
  N = N_row;
  if (n_samples != nullptr) {
    const long long nb = n_samples[b];
    N = nb < N ? (nb > NFFT / 2 ? static_cast<int>(nb) : NFFT / 2 + 1) : N;
  }
  const int Fb = n_samples != nullptr ? min(F, 1 + N / HOP) : F;
  if (f0 >= Fb) {   // block-uniform: this whole block of frames is padding (no barrier is skipped by part of the CTA)
    const int nf = min(FPB, F - f0);
    for (int i = threadIdx.x; i < n_mels * nf; i += KW * 32) {
      if (frame_major) mel[(static_cast<long long>(b) * F + f0) * n_mels + i] = 0.f;
      else mel[(static_cast<long long>(b) * n_mels + i / nf) * F + f0 + i % nf] = 0.f;
    }
    if (threadIdx.x < nf) {
      energy[static_cast<long long>(b) * F + f0 + threadIdx.x] = 0.f;
      if (e_input != nullptr) e_input[static_cast<long long>(b) * F + f0 + threadIdx.x] = 0.f;
    }
    continue;
  }
  float* s_mel = s_mel0 + (DIRECT ? parity * n_mels * MELLD : 0);
  float* s_en = s_en0 + (DIRECT ? parity * FPB : 0);
  bool clipped = false;
  auto scale_clamp = [&](float x) {
    x *= in_scale;
    if (clamp) {   // get_mel_from_wav(norm=False), audio/tools.py:44-49: clamp to [-1,1]; the flag only sees the NEGATIVE side
      clipped |= x < -1.f;
      x = fminf(fmaxf(x, -1.f), 1.f);
    }
    return x;
  };
  // Sample window of the block.  Interior blocks (no reflection, 16-byte aligned start: all but the first / last two blocks of an
  // utterance) issue ALL of their loads as float4 before the first store: one memory round trip per item (the scalar loop --
  // four loads in flight, NSAMP / 1024 round trips -- was 10-13 % of the kernel's warp-state samples on long-scoreboard waits,
  // profiles/ncu_stft_r2l_packed.md).  `samp` is free here: every warp has passed the barrier behind the previous item's frames.
  constexpr int NV = NSAMP / 4, PER = (NV + KW * 32 - 1) / (KW * 32);
  const int first = f0 * HOP - NFFT / 2;
  const bool interior = first >= 0 && first + NSAMP <= N && (reinterpret_cast<uintptr_t>(yb + (first >= 0 ? first : 0)) & 15u) == 0;
  if constexpr (!DIRECT) {
  if (interior) {
    float4 v4[PER];
    const float4* src4 = reinterpret_cast<const float4*>(yb + first);
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = u * KW * 32 + threadIdx.x;
      if (idx < NV) v4[u] = __ldg(src4 + idx);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = u * KW * 32 + threadIdx.x;
      if (idx < NV) {
        float4 x = v4[u];
        x.x = scale_clamp(x.x); x.y = scale_clamp(x.y); x.z = scale_clamp(x.z); x.w = scale_clamp(x.w);
        reinterpret_cast<float4*>(samp)[idx] = x;
      }
    }
  } else {
    for (int i0 = 0; i0 < NSAMP; i0 += 4 * KW * 32) {   // reflect-padded sample window (F.pad mode='reflect', stft.py:58-62)
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                         // four independent loads in flight before anything is stored
        const int i = i0 + u * KW * 32 + threadIdx.x;
        int src = first + i;
        if (src < 0) src = -src;
        if (src >= N) src = 2 * (N - 1) - src;
        v[u] = (i < NSAMP && src >= 0 && src < N) ? __ldg(yb + src) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * KW * 32 + threadIdx.x;
        const float x = scale_clamp(v[u]);
        if (i < NSAMP) samp[i] = x;
      }
    }
  }
  if (clipped && clip_flag != nullptr) clip_flag[b] = 1;   // one (benign, same-value) store per thread after the staging
  __syncthreads();   // samples staged (first item: tables too); also orders the previous item's output stores (reads of s_mel /
                     // s_en) before this item's frames overwrite them
  }  // !DIRECT

  float2* wb = wbuf + warp * WBUF;
  float* mg = ALIAS_MAG ? reinterpret_cast<float*>(wb) : mag + warp * MAGLD;   // 513 floats; wb holds 2 * WBUF
  // bank (8-byte units, 16 per half-warp wavefront) = low four index bits ^ (i6, i6, i5, i4): conflict-free for the pass-1 stores
  // (i = 8 j + r), the strided loads (i = j + 64 r), the pass-2 stores (i = 64 a + 8 r + k: the additive skew i + (i >> 4) of the
  // first version was two-way conflicted there) and the pass-3 stores / split reads (consecutive i)
  auto slot = [](int i) { const int t = (i >> 4) & 7; return i ^ (t | ((t & 4) << 1)); };
  for (int fl = warp; fl < FPB; fl += KW) {
    if (f0 + fl >= Fb) break;                   // warp-uniform
    float2 v[2][8];
    // ---- pass 1 (Ns = 1): window, pack z[n] = x[2n] + i x[2n+1], butterfly, no twiddles
    {
      if constexpr (DIRECT) {
        // DIRECT: the frame's 1024 samples come straight from global memory (16 coalesced 8-byte loads per lane, all in flight
        // together; the 4x overlap between neighbouring frames is served by L1 / L2): no staging pass, no block barrier before
        // the frames, and the load latency stalls one warp instead of the CTA.  Same arithmetic as the staged form
        // (x * in_scale, clamp, * window), so the results are bitwise equal.
        const int start = (f0 + fl) * HOP - NFFT / 2;
        const bool whole = start >= 0 && start + NFFT <= N && (reinterpret_cast<uintptr_t>(yb + (start >= 0 ? start : 0)) & 7u) == 0;
        if (whole) {
          const float2* gp = reinterpret_cast<const float2*>(yb + start);
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int r = 0; r < 8; ++r) v[u][r] = __ldg(gp + lane + 32 * u + 64 * r);
        } else {                                 // frames that touch either end of the utterance: reflect (stft.py:58-62)
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              float xs[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                int src = start + 2 * (lane + 32 * u + 64 * r) + e;
                if (src < 0) src = -src;
                if (src >= N) src = 2 * (N - 1) - src;
                xs[e] = (src >= 0 && src < N) ? __ldg(yb + src) : 0.f;
              }
              v[u][r] = make_float2(xs[0], xs[1]);
            }
        }
        if (in_scale != 1.0f || clamp) {         // block-uniform; x * 1.0f is exact, so skipping it changes nothing
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int r = 0; r < 8; ++r) { v[u][r].x = scale_clamp(v[u][r].x); v[u][r].y = scale_clamp(v[u][r].y); }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int r = 0; r < 8; ++r) v[u][r] = __fmul2_rn(v[u][r], win[lane + 32 * u + 64 * r]);
      } else {
        const float2* sp = reinterpret_cast<const float2*>(samp + fl * HOP);   // fl*HOP is even -> 8-byte aligned
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = lane + 32 * u + 64 * r;
            v[u][r] = __fmul2_rn(sp[n], win[n]);
          }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        dft8(v[u]);
        const int j0 = (lane + 32 * u) * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) wb[slot(j0 + r)] = v[u][r];
      }
      __syncwarp();
    }
    // ---- passes 2, 3 (Ns = 8, 64): twiddle exp(-2*pi*i*r*k/(8*Ns)) = W_1024^(r*k*128/Ns)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int Ns = pass == 0 ? 8 : 64;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
#pragma unroll
        for (int r = 0; r < 8; ++r) v[u][r] = wb[slot(j + 64 * r)];
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        const int k = j & (Ns - 1);
        // twiddles w^r, w = exp(-2*pi*i*k/(8*Ns)), r = 1..7: one conflict-free table read + a product tree
        const float2 w1 = pass == 0 ? twA[k] : twB[k];
        const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
        v[u][1] = cmul(v[u][1], w1); v[u][2] = cmul(v[u][2], w2); v[u][3] = cmul(v[u][3], w3); v[u][4] = cmul(v[u][4], w4);
        v[u][5] = cmul(v[u][5], cmul(w4, w1)); v[u][6] = cmul(v[u][6], cmul(w3, w3)); v[u][7] = cmul(v[u][7], cmul(w4, w3));
        dft8(v[u]);
        const int j0 = ((j - k) << 3) + k;
#pragma unroll
        for (int r = 0; r < 8; ++r) wb[slot(j0 + r * Ns)] = v[u][r];
      }
      __syncwarp();
    }
    // ---- split post-pass: X[k] = E + W^k * O.  Bins k and 512-k share Z[k], Z[512-k] and the twiddle
    // (E(512-k) = conj E(k), O(512-k) = conj O(k), W^(512-k) = -conj W^k), so each step yields two magnitudes.
    float esum = 0.f;
    float mk[9], mm[9];                          // ALIAS_MAG: magnitudes of bins k / 512-k, parked until every lane has read wb
#pragma unroll
    for (int it = 0; it < 9; ++it) {
      const int k = lane + 32 * it;
      if (k <= HALF / 2) {
        const float2 zk = wb[slot(k & (HALF - 1))];
        const float2 zr = wb[slot((HALF - k) & (HALF - 1))];
        const float2 zc = make_float2(zr.x, -zr.y);                                        // conj(Z[512-k])
        const float2 e = __fmul2_rn(cadd(zk, zc), make_float2(0.5f, 0.5f));                // E(k)
        const float2 o = __fmul2_rn(mul_mi(csub(zk, zc)), make_float2(0.5f, 0.5f));        // O(k) = -i/2 (Z[k] - conj(Z[512-k]))
        const float2 wo = cmul(o, tw[k]);
        const float2 X = cadd(e, wo);                                                      // X[k]
        const float2 Y = __fadd2_rn(make_float2(e.x, -e.y), make_float2(-wo.x, wo.y));     // X[512-k] = conj(E) - conj(W^k O)
        const float2 pp = __ffma2_rn(make_float2(X.y, Y.y), make_float2(X.y, Y.y), __fmul2_rn(make_float2(X.x, Y.x), make_float2(X.x, Y.x)));
        const float p0 = pp.x, p1 = pp.y;
        if (ALIAS_MAG) { mk[it] = sqrt_approx(p0); mm[it] = sqrt_approx(p1); }
        else { mg[k] = sqrt_approx(p0); mg[HALF - k] = sqrt_approx(p1); }
        esum += k == HALF / 2 ? p0 : p0 + p1;     // bin 256 is its own mirror
      }
    }
    if (ALIAS_MAG) {
      __syncwarp();                              // all of wb consumed: the magnitudes may land on it
#pragma unroll
      for (int it = 0; it < 9; ++it) {
        const int k = lane + 32 * it;
        if (k <= HALF / 2) { mg[k] = mk[it]; mg[HALF - k] = mm[it]; }
      }
    }
    esum = warp_sum(esum);                      // energy: L2 norm over the 513 bins (stft.py:158)
    if (lane == 0) s_en[fl] = sqrtf(esum);
    __syncwarp();
    // ---- mel projection + log compression (stft.py:156-157) over the non-zero band of each filter row
    for (int m = lane; m < n_mels; m += 32) {
      const int lo = s_lo[m], hi = s_hi[m];
      float acc = 0.f;
      if (basis_in_smem) {
        const float* br = s_basis + s_off[m] - lo;
        for (int k = lo; k < hi; ++k) acc = fmaf(br[k], mg[k], acc);
      } else {
        const float* br = basis + static_cast<long long>(m) * NBINS;
        for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(br + k), mg[k], acc);
      }
      s_mel[m * MELLD + fl] = __logf(fmaxf(acc, 1e-5f));   // MUFU.LG2 form: |error| < 2e-6 on [1e-5, 1e4]
    }
    __syncwarp();                               // mg / wb are reused by this warp's next frame
  }
  if (DIRECT && clipped && clip_flag != nullptr) clip_flag[b] = 1;
  __syncthreads();   // DIRECT: the only barrier of an item.  The tile read below is not written again before the NEXT item's
  parity ^= 1;       // barrier (the next item fills the other tile), by which time every warp has finished these stores
  // coalesced stores: FPB consecutive frames of one mel row are contiguous in mel[b][m][:]
  const int nf = min(FPB, F - f0);
  const int nv = min(FPB, Fb - f0);            // valid frames of this block; [nv, nf) is per-utterance padding -> zeros
  if (frame_major) {   // [B][F][n_mels]: the channel-last layout STYLER.forward takes as mel_target (the reference stores mel.T)
    for (int i = threadIdx.x; i < n_mels * nf; i += KW * 32) {
      const int fl = i / n_mels, m = i % n_mels;
      mel[(static_cast<long long>(b) * F + f0 + fl) * n_mels + m] = fl < nv ? s_mel[m * MELLD + fl] : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < n_mels * FPB; i += KW * 32) {
      const int m = i / FPB, fl = i % FPB;
      if (fl < nf) mel[(static_cast<long long>(b) * n_mels + m) * F + f0 + fl] = fl < nv ? s_mel[m * MELLD + fl] : 0.f;
    }
  }
  if (threadIdx.x < nf) {
    const bool ok = static_cast<int>(threadIdx.x) < nv;
    const float e = ok ? s_en[threadIdx.x] : 0.f;
    energy[static_cast<long long>(b) * F + f0 + threadIdx.x] = e;
    if (e_input != nullptr)   // energy_rescaling, utils.py:412-416
      e_input[static_cast<long long>(b) * F + f0 + threadIdx.x] = ok ? fminf(fmaxf((e - e_min) * e_inv_range, 0.f), 1.f) : 0.f;
  }
  }   // item loop
}


template <int FPB, int KW, int MINB, bool ALIAS_MAG, bool DIRECT = false> struct StftLaunch {
  static size_t smem_bytes(int n_mels) {
    return sizeof(float2) * (2 * HALF + 8 + 64) + sizeof(float2) * KW * WBUF + (DIRECT ? 0 : sizeof(float) * StftShape<FPB>::NSAMP) +
           (ALIAS_MAG ? 0 : sizeof(float) * KW * MAGLD) + (DIRECT ? 2 : 1) * sizeof(float) * FPB + sizeof(float) * kBasisCap +
           sizeof(int) * (3 * n_mels + 1) + (DIRECT ? 2 : 1) * sizeof(float) * n_mels * StftShape<FPB>::MELLD;
  }
  static constexpr size_t kSmemCap = (228 * 1024 - MINB * 1024) / MINB / 1024 * 1024;   // MINB CTAs per SM, 1 KB reserved each
  static int run(const float* y, int B, int N, int F, const float* basis, const int32_t* band, int n_mels, float* mel,
                 float* energy, float in_scale, int clamp, int32_t* clip_flag, int frame_major, float* e_input, float e_min,
                 float e_inv, const int64_t* n_samples, cudaStream_t s) {
    static DeviceFlags attr_set;
    SB_OPT_IN_SMEM(attr_set, (stft_mel_kernel<FPB, KW, MINB, ALIAS_MAG, DIRECT>), kSmemCap);
    const long long items = static_cast<long long>(ceil_div(F, FPB)) * B;
    SB_REQUIRE(items < (1ll << 31), "stft_mel: too many frame blocks (%lld)", items);
    const int grid = static_cast<int>(items < static_cast<long long>(MINB) * num_sms() ? items : static_cast<long long>(MINB) * num_sms());
    stft_mel_kernel<FPB, KW, MINB, ALIAS_MAG, DIRECT><<<grid, KW * 32, smem_bytes(n_mels), s>>>(
        y, N, F, basis, band, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min, e_inv, n_samples,
        static_cast<int>(items));
    SB_LAUNCH_OK();
    return 0;
  }
};

}  // namespace
}  // namespace sb

extern "C" int styler_stft_mel_ex_fwd(const float* y, int32_t B, int32_t N, const float* mel_basis, int32_t n_mels,
                                      int32_t* band_ws, float* mel, float* energy, float in_scale, int32_t clamp,
                                      int32_t* clip_flag, int32_t frame_major, float* e_input, float e_min, float e_max,
                                      const int64_t* n_samples, void* stream) {
  using namespace sb;
  SB_REQUIRE(y && mel_basis && band_ws && mel && energy, "stft_mel: null pointer");
  SB_REQUIRE(B > 0 && N > NFFT / 2 && n_mels > 0 && n_mels <= 256, "stft_mel: bad shape (B=%d N=%d n_mels=%d)", B, N, n_mels);
  SB_REQUIRE(e_input == nullptr || e_max > e_min, "stft_mel: energy rescaling needs e_max > e_min");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int F = 1 + N / HOP;
  if (clip_flag != nullptr) SB_CUDA_OK(cudaMemsetAsync(clip_flag, 0, sizeof(int32_t) * B, s));
  mel_band_kernel<<<ceil_div(n_mels, 8), 256, 0, s>>>(mel_basis, n_mels, band_ws);
  SB_LAUNCH_OK();
  const float e_inv = e_input != nullptr ? 1.0f / (e_max - e_min) : 0.f;
  using Two = StftLaunch<32, 8, 2, false>;     // 2 CTAs x 8 warps per SM
  using Three = StftLaunch<16, 8, 3, true>;    // 3 CTAs x 8 warps per SM
  using Wide = StftLaunch<24, 12, 2, true>;    // 2 CTAs x 12 warps per SM (prologue amortised over 24 frames)
  using Direct = StftLaunch<16, 8, 3, true, true>;   // 3 CTAs x 8 warps per SM, samples read from global by the frame's own warp
  const int occ = tuning(TUNE_STFT_OCC);
  // (rows that are not 8-byte aligned -- odd N -- would take the direct form's scalar edge path for every frame: they keep the
  // staged form, whose block window is staged once either way)
  const bool rows_aligned = (N % 2) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0;
  if (occ == 3 && !rows_aligned) {
    if (Three::smem_bytes(n_mels) <= Three::kSmemCap)
      return Three::run(y, B, N, F, mel_basis, band_ws, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min,
                        e_inv, n_samples, s);
  } else if (occ == 3 && Direct::smem_bytes(n_mels) <= Direct::kSmemCap)
    return Direct::run(y, B, N, F, mel_basis, band_ws, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min,
                       e_inv, n_samples, s);
  if (occ == 1 && Three::smem_bytes(n_mels) <= Three::kSmemCap)
    return Three::run(y, B, N, F, mel_basis, band_ws, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min,
                      e_inv, n_samples, s);
  if (occ == 2 && Wide::smem_bytes(n_mels) <= Wide::kSmemCap)
    return Wide::run(y, B, N, F, mel_basis, band_ws, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min,
                     e_inv, n_samples, s);
  SB_REQUIRE(Two::smem_bytes(n_mels) <= Two::kSmemCap, "stft_mel: n_mels=%d needs %zu bytes of shared memory", n_mels,
             Two::smem_bytes(n_mels));
  return Two::run(y, B, N, F, mel_basis, band_ws, n_mels, mel, energy, in_scale, clamp, clip_flag, frame_major, e_input, e_min, e_inv,
                  n_samples, s);
}

extern "C" int styler_stft_mel_fwd(const float* y, int32_t B, int32_t N, const float* mel_basis, int32_t n_mels,
                                   int32_t* band_ws, float* mel, float* energy, void* stream) {
  return styler_stft_mel_ex_fwd(y, B, N, mel_basis, n_mels, band_ws, mel, energy, 1.0f, 0, nullptr, 0, nullptr, 0.f, 1.f, nullptr, stream);
}
