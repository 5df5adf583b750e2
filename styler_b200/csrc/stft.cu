// TacotronSTFT.mel_spectrogram (audio/stft.py:51-79,141-160; audio_processing.py:80-86 of the reference).
// The reference evaluates a dense 1026x1024 DFT as a strided Conv1d (2.1 MFLOP/frame) and bounces through
// the host; here one CTA handles 8 consecutive frames of one utterance:
//   stage the 8*256+768 reflect-padded samples in smem once (the 4x frame overlap is served from smem),
//   periodic-Hann window, 1024-point real FFT as a 512-point complex Stockham FFT (4 radix-4 + 1 radix-2 passes) + split post-pass
//   (64 threads per frame, 4 frames in flight), magnitude for the 513 bins -> smem,
//   energy = ||mag||_2, mel = log(max(basis @ mag, 1e-5)) using the non-zero band of each filter row.
// Bounding roofline: HBM (464,580 algorithmic bytes per 4 s utterance); the FFT stage is smem/ALU work.
#include "common.cuh"

namespace sb {
namespace {

constexpr int NFFT = 1024, HOP = 256, NBINS = 513, FPB = 8, HALF = 512;
constexpr int NSAMP = (FPB - 1) * HOP + NFFT;

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

__global__ void __launch_bounds__(256) stft_mel_kernel(const float* __restrict__ y, int N, int F,
                                                       const float* __restrict__ basis, const int32_t* __restrict__ band,
                                                       int n_mels, float* __restrict__ mel, float* __restrict__ energy) {
  extern __shared__ __align__(16) uint8_t stft_smem[];
  float2 (*buf)[2][HALF] = reinterpret_cast<float2 (*)[2][HALF]>(stft_smem);             // [4][2][512]
  float2* tw = reinterpret_cast<float2*>(stft_smem + sizeof(float2) * 4 * 2 * HALF);      // W_1024^k, k < 512
  float (*mag)[NBINS + 3] = reinterpret_cast<float (*)[NBINS + 3]>(tw + HALF);            // [FPB][516]
  float* samp = reinterpret_cast<float*>(mag + FPB);                                      // [NSAMP]
  float* hann = samp + NSAMP;                                                             // [NFFT] periodic Hann
  const int b = blockIdx.y, f0 = blockIdx.x * FPB;
  const float* yb = y + static_cast<long long>(b) * N;
  {   // all loads of the window are issued before the first store (11 independent global loads in flight per thread)
    constexpr int kIter = (NSAMP + 255) / 256;
    float v[kIter];
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int i = threadIdx.x + j * 256;
      int src = f0 * HOP + i - NFFT / 2;         // reflect padding (F.pad mode='reflect', stft.py:58-62)
      if (src < 0) src = -src;
      if (src >= N) src = 2 * (N - 1) - src;
      v[j] = (i < NSAMP && src >= 0 && src < N) ? __ldg(yb + src) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int i = threadIdx.x + j * 256;
      if (i < NSAMP) samp[i] = v[j];
    }
  }
  for (int k = threadIdx.x; k < HALF; k += 256) {
    float s, c;
    sincospif(-static_cast<float>(k) / 512.0f, &s, &c);
    tw[k] = make_float2(c, s);
    hann[k] = 0.5f - 0.5f * c;                 // cos(2*pi*k/1024) = Re W^k ;  cos(2*pi*(k+512)/1024) = -Re W^k
    hann[k + HALF] = 0.5f + 0.5f * c;
  }
  __syncthreads();

  const int grp = threadIdx.x >> 6, lt = threadIdx.x & 63;   // 4 frames in flight, 64 threads each
  for (int round = 0; round < FPB / 4; ++round) {
    const int fl = round * 4 + grp;                          // local frame index
    float2* d0 = buf[grp][0];
    float2* d1 = buf[grp][1];
    // window + pack real pairs into complex: z[n] = x[2n] + i x[2n+1]
    {
      const float2* sp = reinterpret_cast<const float2*>(samp + fl * HOP);   // fl*HOP is even -> 8-byte aligned
      const float2* hp2 = reinterpret_cast<const float2*>(hann);
#pragma unroll
      for (int n = lt; n < HALF; n += 64) {
        const float2 x = sp[n], w = hp2[n];
        d0[n] = make_float2(x.x * w.x, x.y * w.y);
      }
    }
    // 512-point Stockham FFT: four radix-4 passes (Ns = 1, 4, 16, 64; 128 butterflies each, 2 per thread) and one radix-2
    // pass (Ns = 256).  Only the 64 threads of this frame group synchronise (named barrier 1+grp), not the whole CTA.
    auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory"); };
    auto cmul = [](float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); };
    group_sync();
#pragma unroll 1
    for (int s4 = 0; s4 < 4; ++s4) {
      const int Ns = 1 << (2 * s4);
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int jdx = lt + qq * 64;                 // 0..127
        const int k = jdx & (Ns - 1);
        // twiddles exp(-2*pi*i*r*k/(4*Ns)) = W_1024^(r*k*256/Ns), r = 1..3 (index < 768: fold with W^(512+m) = -W^m)
        const int tstep = k * (256 / Ns);
        float2 v0 = d0[jdx], v1 = d0[jdx + 128], v2 = d0[jdx + 256], v3 = d0[jdx + 384];
        const float2 w1 = tw[tstep], w2 = tw[2 * tstep];
        const int t3 = 3 * tstep;
        const float2 w3 = t3 < HALF ? tw[t3] : make_float2(-tw[t3 - HALF].x, -tw[t3 - HALF].y);
        v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
        // radix-4 butterfly (forward transform: -i rotation)
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const int dst = ((jdx - k) << 2) + k;
        d1[dst] = make_float2(a0.x + a2.x, a0.y + a2.y);
        d1[dst + Ns] = make_float2(a1.x + a3.y, a1.y - a3.x);
        d1[dst + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        d1[dst + 3 * Ns] = make_float2(a1.x - a3.y, a1.y + a3.x);
      }
      group_sync();
      float2* tmp = d0; d0 = d1; d1 = tmp;
    }
    {
      const int Ns = 256;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int jdx = lt + qq * 64;                 // 0..255
        const int k = jdx & (Ns - 1);
        const float2 w = tw[k * (HALF / Ns)];
        const float2 a = d0[jdx], bw = cmul(d0[jdx + 256], w);
        const int dst = ((jdx - k) << 1) + k;
        d1[dst] = make_float2(a.x + bw.x, a.y + bw.y);
        d1[dst + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      group_sync();
      float2* tmp = d0; d0 = d1; d1 = tmp;
    }
    // split post-pass: X[k] = E + W^k * O.  Bins k and 512-k share Z[k], Z[512-k] and the twiddle
    // (E(512-k) = conj E(k), O(512-k) = conj O(k), W^(512-k) = -conj W^k), so each thread produces two magnitudes.
    for (int k = lt; k <= HALF / 2; k += 64) {
      const float2 zk = d0[k & (HALF - 1)];
      const float2 zr = d0[(HALF - k) & (HALF - 1)];
      const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
      const float2 o = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));        // -i/2 * (zk - conj(zr))
      const float2 w = tw[k];
      const float2 wo = make_float2(o.x * w.x - o.y * w.y, o.x * w.y + o.y * w.x);
      const float xr = e.x + wo.x, xi = e.y + wo.y;                                      // X[k]
      const float yr = e.x - wo.x, yi = -e.y + wo.y;                                     // X[512-k] = conj(E) - conj(W^k O)
      mag[fl][k] = sqrtf(xr * xr + xi * xi);
      mag[fl][HALF - k] = sqrtf(yr * yr + yi * yi);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");   // the group's buffers are reused by its next frame
  }
  __syncthreads();

  // energy: L2 norm over the 513 bins (stft.py:158); warp per frame
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < FPB && f0 + warp < F) {
    float s = 0.f;
    for (int k = lane; k < NBINS; k += 32) s += mag[warp][k] * mag[warp][k];
    s = warp_sum(s);
    if (lane == 0) energy[static_cast<long long>(b) * F + f0 + warp] = sqrtf(s);
  }
  // mel projection + log compression (stft.py:156-157)
  for (int i = threadIdx.x; i < n_mels * FPB; i += 256) {
    const int m = i / FPB, fl = i % FPB;
    if (f0 + fl >= F) continue;
    const int lo = band[2 * m], hi = band[2 * m + 1];
    const float* br = basis + static_cast<long long>(m) * NBINS;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(br[k], mag[fl][k], acc);
    mel[(static_cast<long long>(b) * n_mels + m) * F + f0 + fl] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace
}  // namespace sb

extern "C" int styler_stft_mel_fwd(const float* y, int32_t B, int32_t N, const float* mel_basis, int32_t n_mels,
                                   int32_t* band_ws, float* mel, float* energy, void* stream) {
  using namespace sb;
  SB_REQUIRE(y && mel_basis && band_ws && mel && energy, "stft_mel: null pointer");
  SB_REQUIRE(B > 0 && N > NFFT / 2 && n_mels > 0 && n_mels <= 256, "stft_mel: bad shape (B=%d N=%d n_mels=%d)", B, N, n_mels);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int F = 1 + N / HOP;
  mel_band_kernel<<<ceil_div(n_mels, 8), 256, 0, s>>>(mel_basis, n_mels, band_ws);
  SB_LAUNCH_OK();
  dim3 grid(ceil_div(F, FPB), B);
  constexpr size_t smem = sizeof(float2) * 4 * 2 * HALF + sizeof(float2) * HALF + sizeof(float) * FPB * (NBINS + 3) +
                          sizeof(float) * NSAMP + sizeof(float) * NFFT;
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA_OK(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  stft_mel_kernel<<<grid, 256, smem, s>>>(y, N, F, mel_basis, band_ws, n_mels, mel, energy);
  SB_LAUNCH_OK();
  return 0;
}
