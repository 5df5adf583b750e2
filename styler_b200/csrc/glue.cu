// HBM-bound glue kernels of the STYLER forward: embedding+position, adds, quantise, one-hot conv gather,
// GroupNorm+ReLU, Mel Calibrator, classifier tail, duration rounding, LengthRegulator (integer scan +
// vectorised gather-expand), bucketize+embedding+sum.  All are coalesced, 16-byte vectorised where the
// layout allows, fp32 math, activation I/O in the ABI dtype.  Reference lines are cited in include/styler_b200.h.
#include "common.cuh"

namespace sb {
namespace {

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------- embed + pos
template <typename T>
__global__ void embed_pos_kernel(const int64_t* __restrict__ seq, const float* __restrict__ emb, int vocab,
                                 const float* __restrict__ pos, T* __restrict__ out, int rows, int L, int D) {
  const int per_row = D / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(rows) * per_row) return;
  const int row = static_cast<int>(gid / per_row), c = static_cast<int>(gid % per_row) * 8;
  const int l = row % L;
  long long tok = seq[row];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  float e[8], p[8], v[8];
  load8(emb + tok * D + c, e);
  load8(pos + static_cast<long long>(l) * D + c, p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = e[i] + p[i];
  store8(out + static_cast<long long>(row) * D + c, v);
}

// ---------------------------------------------------------------------------------------- generic add
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, long long a_bs, int a_ld, const T* __restrict__ a2, long long a2_bs,
                           int a2_ld, const T* __restrict__ rowvec, int rv_ld, const float* __restrict__ pos,
                           T* __restrict__ out, long long o_bs, int o_ld, int B, int Tn, int C) {
  const int per_row = C / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(B) * Tn * per_row) return;
  const int c = static_cast<int>(gid % per_row) * 8;
  const long long row = gid / per_row;
  const int t = static_cast<int>(row % Tn), b = static_cast<int>(row / Tn);
  float v[8], w[8];
  if (a != nullptr) {
    load8(a + b * a_bs + static_cast<long long>(t) * a_ld + c, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  if (rowvec != nullptr) {
    load8(rowvec + static_cast<long long>(b) * rv_ld + c, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += w[i];
  }
  if (pos != nullptr) {
    load8(pos + static_cast<long long>(t) * C + c, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += w[i];
  }
  if (a2 != nullptr) {
    load8(a2 + b * a2_bs + static_cast<long long>(t) * a2_ld + c, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += w[i];
  }
  store8(out + b * o_bs + static_cast<long long>(t) * o_ld + c, v);
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) DT<T>::st(out + i, x[i]);
}

// HiFi-GAN multi-receptive-field fusion (hifigan/models.py:154-161): the resblock outputs arrive in activated form
// y_k = lrelu(x_k, slope_in) (the residual chain is stored that way, see styler_conv1d_args.residual_inv_lrelu); recover
// x_k, average, and apply the leaky ReLU that precedes the next layer.  8 elements (16 B of bf16) per thread.
template <typename T>
__global__ void lrelu_mean_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, float inv_slope_in,
                                  float scale, float slope_out, T* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float va[8], vb[8], vc[8], r[8];
  load8(a + i * 8, va);
  if (b != nullptr) load8(b + i * 8, vb);
  if (c != nullptr) load8(c + i * 8, vc);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = va[k] < 0.f ? va[k] * inv_slope_in : va[k];
    if (b != nullptr) s += vb[k] < 0.f ? vb[k] * inv_slope_in : vb[k];
    if (c != nullptr) s += vc[k] < 0.f ? vc[k] * inv_slope_in : vc[k];
    s *= scale;
    r[k] = fmaxf(s, s * slope_out);
  }
  store8(out + i * 8, r);
}

// f0_normalization / speaker_normalization (utils.py:387-409) for a padded batch of log-f0 contours: per utterance, over
// the voiced frames (f0 > -1e10) of its first lens[b] frames, z = (f0 - mean) / std / 4 clipped to [-1,1] and mapped to
// [0,1]; unvoiced frames keep their value; if the statistics are undefined (no voiced frame, std == 0 - the reference
// turns numpy's RuntimeWarning into an all-zero contour) the whole row is zero; frames >= lens[b] are zero (pad_1D).
// Statistics in fp64 like numpy's (float64 mean / population std).  One CTA per utterance.
__global__ void __launch_bounds__(256) f0_norm_kernel(const float* __restrict__ f0, const int64_t* __restrict__ lens,
                                                      float* __restrict__ out, int Tn) {
  __shared__ double s_sum[8], s_sq[8];
  __shared__ long long s_cnt[8];
  __shared__ double s_mean, s_std;
  __shared__ int s_ok;
  const int b = blockIdx.x;
  const float* x = f0 + static_cast<long long>(b) * Tn;
  float* o = out + static_cast<long long>(b) * Tn;
  int len = lens != nullptr ? static_cast<int>(lens[b]) : Tn;
  len = len < Tn ? len : Tn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto block_sum = [&](double v, double* sh) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    __syncthreads();
    return t;
  };
  double sum = 0.0;
  long long cnt = 0;
  for (int t = threadIdx.x; t < len; t += 256)
    if (x[t] > -1e10f) { sum += static_cast<double>(x[t]); ++cnt; }
  const double tot = block_sum(sum, s_sum);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  long long n = 0;
  for (int i = 0; i < 8; ++i) n += s_cnt[i];
  const double mean = n > 0 ? tot / static_cast<double>(n) : 0.0;
  double sq = 0.0;
  for (int t = threadIdx.x; t < len; t += 256)
    if (x[t] > -1e10f) { const double d = static_cast<double>(x[t]) - mean; sq += d * d; }
  const double var = block_sum(sq, s_sq);
  if (threadIdx.x == 0) {
    s_mean = mean;
    s_std = n > 0 ? sqrt(var / static_cast<double>(n)) : 0.0;
    s_ok = (n > 0 && s_std > 0.0) ? 1 : 0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Tn; t += 256) {
    float r = 0.f;
    if (t < len && s_ok) {
      const float v = x[t];
      if (v > -1e10f) {
        double z = (static_cast<double>(v) - s_mean) / s_std / 4.0;
        z = z < -1.0 ? -1.0 : (z > 1.0 ? 1.0 : z);
        r = static_cast<float>((z + 1.0) / 2.0);
      } else {
        r = v;
      }
    }
    o[t] = r;
  }
}

// ---------------------------------------------------------------------------------------- quantise / one-hot conv
__global__ void quantize_index_kernel(const float* __restrict__ x, int32_t* __restrict__ idx, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  idx[i] = v <= 0.f ? 0 : static_cast<int32_t>(rintf(__fmul_rn(v, 255.0f))) + 1;
}

template <typename T>
__global__ void onehot_conv_kernel(const int32_t* __restrict__ idx, const float* __restrict__ wg,
                                   const float* __restrict__ bias, T* __restrict__ out, int B, int Tn, int C, int nidx,
                                   int KS) {
  const int per_row = C / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(B) * Tn * per_row) return;
  const int c = static_cast<int>(gid % per_row) * 8;
  const long long row = gid / per_row;
  const int t = static_cast<int>(row % Tn), b = static_cast<int>(row / Tn);
  const int pad = (KS - 1) / 2;
  float acc[8];
  load8(bias + c, acc);
  for (int tap = 0; tap < KS; ++tap) {
    const int tt = t + tap - pad;
    if (tt < 0 || tt >= Tn) continue;
    int id = idx[static_cast<long long>(b) * Tn + tt];
    id = id < 0 ? 0 : (id >= nidx ? nidx - 1 : id);
    float w[8];
    load8(wg + (static_cast<long long>(tap) * nidx + id) * C + c, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += w[i];
  }
  store8(out + row * C + c, acc);
}

// ---------------------------------------------------------------------------------------- GroupNorm + ReLU
// stats: one CTA per (b, group); shifted sums (shift = first element) in fp32 per thread, fp64 block combine.
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, long long bs, int ld, float* __restrict__ stats,
                                                       int Tn, int groups, int cpg, float eps) {
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const T* base = x + b * bs + g * cpg;
  const float shift = DT<T>::ld(base);
  float s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const T* p = base + static_cast<long long>(t) * ld;
    for (int c = 0; c < cpg; c += 8) {
      float v[8];
      load8(p + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[i] - shift; s1 += d; s2 += d * d; }
    }
  }
  __shared__ double r1[8], r2[8];
  double d1 = warp_sum(s1), d2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { r1[warp] = d1; r2[warp] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, q = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { a += r1[i]; q += r2[i]; }
    const double n = static_cast<double>(Tn) * cpg;
    const double dm = a / n;
    double var = q / n - dm * dm;
    var = var < 0 ? 0 : var;
    stats[2 * blockIdx.x] = static_cast<float>(shift + dm);
    stats[2 * blockIdx.x + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}
// finalize from per-tile partial sums written by the conv epilogue: one thread per (b, group), fp64 combine in tile order
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int B, int groups, int n_part,
                                   int Tn, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  const int b = i / groups, g = i % groups;
  double a = 0, q = 0;
  for (int p = 0; p < n_part; ++p) {
    const float* src = partial + ((static_cast<long long>(b) * n_part + p) * groups + g) * 2;
    a += static_cast<double>(src[0]);
    q += static_cast<double>(src[1]);
  }
  const double n = static_cast<double>(Tn) * 16.0;
  const double mean = a / n;
  double var = q / n - mean * mean;
  var = var < 0 ? 0 : var;
  stats[2 * i] = static_cast<float>(mean);
  stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}
template <typename T>
__global__ void gn_apply_relu_kernel(T* __restrict__ x, long long bs, int ld, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int B, int Tn,
                                     int C, int cpg) {
  const int per_row = C / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(B) * Tn * per_row) return;
  const int c = static_cast<int>(gid % per_row) * 8;
  const long long row = gid / per_row;
  const int t = static_cast<int>(row % Tn), b = static_cast<int>(row / Tn);
  const int g = c / cpg;
  const float mean = stats[2 * (b * (C / cpg) + g)], rstd = stats[2 * (b * (C / cpg) + g) + 1];
  T* p = x + b * bs + static_cast<long long>(t) * ld + c;
  float v[8], ga[8], be[8];
  load8(p, v);
  load8(gamma + c, ga);
  load8(beta + c, be);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = fmaxf((v[i] - mean) * rstd * ga[i] + be[i], 0.f);
  store8(p, v);
}

// ---------------------------------------------------------------------------------------- Mel Calibrator
// GN: x is the RAW output of the branch's last convolution; GroupNorm (16 channels per group, statistics from gn_finalize) +
// affine + ReLU (modules.py:113-117) are applied to every frame as it is read, so the normalised [B,Tr,C] tensor is never
// written (one read of x instead of a read-modify-write pass plus this read; frames beyond mel_len are never touched).
template <typename T, bool GN>
__global__ void mel_calibrator_kernel(const T* __restrict__ x, long long x_bs, int x_ld, const int64_t* __restrict__ mel_len,
                                      const int64_t* __restrict__ src_len, T* __restrict__ out, long long o_bs, int o_ld,
                                      int B, int Tr, int L, int C, const float* __restrict__ stats,
                                      const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int per_row = C / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(B) * L * per_row) return;
  const int c = static_cast<int>(gid % per_row) * 8;
  const long long row = gid / per_row;
  const int l = static_cast<int>(row % L), b = static_cast<int>(row / L);
  int ml = static_cast<int>(mel_len[b]), sl = static_cast<int>(src_len[b]);
  ml = ml > Tr ? Tr : ml;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const T* xb = x + b * x_bs + c;
  float sc[8], sh[8];                            // GN: v -> relu(v * sc + sh), sc = rstd * gamma, sh = beta - mean * rstd * gamma
  if constexpr (GN) {
    const int g = b * (C / 16) + c / 16;
    const float mean = stats[2 * g], rstd = stats[2 * g + 1];
    load8(gamma + c, sc);
    load8(beta + c, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] *= rstd; sh[i] = fmaf(-mean, sc[i], sh[i]); }
  }
  auto load_frame = [&](int t, float (&v)[8]) {
    load8(xb + static_cast<long long>(t) * x_ld, v);
    if constexpr (GN) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
    }
  };
  if (l < sl && ml > 0) {
    if (ml == sl) {
      load_frame(l, acc);
    } else if (ml > sl) {                        // compression: mean over segment l (get_scale(ml, sl))
      const int q = ml / sl, r = ml % sl;
      const int start = l * q + (l < r ? l : r), size = q + (l < r ? 1 : 0);
      int k = 0;
      for (; k + 4 <= size; k += 4) {            // four frames in flight per thread (the loads are independent), summed in frame order
        float v0[8], v1[8], v2[8], v3[8];
        load_frame(start + k, v0);
        load_frame(start + k + 1, v1);
        load_frame(start + k + 2, v2);
        load_frame(start + k + 3, v3);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = (((acc[i] + v0[i]) + v1[i]) + v2[i]) + v3[i];
      }
      for (; k < size; ++k) {
        float v[8];
        load_frame(start + k, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
      const float fs = static_cast<float>(size);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = acc[i] / fs;
    } else {                                     // expansion: frame i repeated get_scale(sl, ml)[i] times
      const int q = sl / ml, r = sl % ml;
      const int boundary = r * (q + 1);
      const int i = l < boundary ? l / (q + 1) : r + (l - boundary) / q;
      load_frame(i, acc);
    }
  }
  store8(out + b * o_bs + static_cast<long long>(l) * o_ld + c, acc);
}

// ---------------------------------------------------------------------------------------- classifier tail
// One CTA per utterance, one warp per row (4 rows in flight per warp): every lane holds 8 consecutive channels of the two
// weight rows in registers and reads its 8 activations with one 16-byte load.  (Round 1 read 2-byte elements one at a
// time and re-read the weights from global for every row: 40 us for a 64-CTA toy op.)
template <typename T>
__global__ void __launch_bounds__(256) classifier_tail_kernel(const T* __restrict__ h, long long h_bs, int h_ld,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ out, int L, int C) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float b0 = bias[0], b1 = bias[1];
  float a0 = 0.f, a1 = 0.f;
  auto finish_row = [&](float z0, float z1) {
    z0 = warp_sum(z0) + b0;
    z1 = warp_sum(z1) + b1;
    const float m = fmaxf(z0, z1);
    const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
    a0 += z0 - lse;
    a1 += z1 - lse;
  };
  if (C == 256) {
    float w0[8], w1[8];
    load8(w + lane * 8, w0);
    load8(w + C + lane * 8, w1);
    constexpr int U = 4;
    for (int l0 = warp; l0 < L; l0 += 8 * U) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int l = l0 + 8 * u;
        if (l < L) load8(h + b * h_bs + static_cast<long long>(l) * h_ld + lane * 8, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (l0 + 8 * u >= L) break;       // warp-uniform
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { z0 = fmaf(v[u][i], w0[i], z0); z1 = fmaf(v[u][i], w1[i], z1); }
        finish_row(z0, z1);
      }
    }
  } else {
    for (int l = warp; l < L; l += 8) {
      const T* row = h + b * h_bs + static_cast<long long>(l) * h_ld;
      float z0 = 0.f, z1 = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = DT<T>::ld(row + c);
        z0 = fmaf(v, w[c], z0);
        z1 = fmaf(v, w[C + c], z1);
      }
      finish_row(z0, z1);
    }
  }
  __shared__ float s0[8], s1[8];
  if (lane == 0) { s0[warp] = a0; s1[warp] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int i = 0; i < 8; ++i) { t0 += s0[i]; t1 += s1[i]; }
    out[2 * b] = t0 / L;
    out[2 * b + 1] = t1 / L;
  }
}

// ---------------------------------------------------------------------------------------- duration / LengthRegulator
__global__ void duration_round_kernel(const float* __restrict__ log_d, float* __restrict__ dur, long long n, float off,
                                      float ctl) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dur[i] = fmaxf(__fmul_rn(rintf(__fsub_rn(expf(log_d[i]), off)), ctl), 0.f);
}

// inclusive scan of int(duration) per utterance; block per utterance, L <= 4096
__global__ void __launch_bounds__(1024) lr_scan_kernel(const int64_t* __restrict__ d64, const float* __restrict__ d32,
                                                       int32_t* __restrict__ cum, int64_t* __restrict__ mel_len, int L) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < L; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int v = 0;
    if (i < L) {
      if (d64 != nullptr) {
        const long long d = d64[static_cast<long long>(b) * L + i];
        v = d < 0 ? 0 : (d > 1000000 ? 1000000 : static_cast<int>(d));
      } else {
        const float f = d32[static_cast<long long>(b) * L + i];
        v = f > 0.f ? (f > 1.0e6f ? 1000000 : static_cast<int>(f)) : 0;   // int(x.item()): truncation
      }
    }
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += n; }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += n; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int incl = s + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
    if (i < L) cum[static_cast<long long>(b) * L + i] = incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) mel_len[b] = carry;
}

// warp per output frame: binary search the scan, copy the source row with 16-byte accesses
__global__ void __launch_bounds__(256) lr_expand_kernel(const uint4* __restrict__ x, long long x_bs16, int x_ld16,
                                                        const int32_t* __restrict__ cum, uint4* __restrict__ out,
                                                        long long o_bs16, int o_ld16, int L, int Tmax, int row16) {
  extern __shared__ int32_t scum[];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < L; i += blockDim.x) scum[i] = cum[static_cast<long long>(b) * L + i];
  __syncthreads();
  const int total = L > 0 ? scum[L - 1] : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= Tmax) return;
  uint4* orow = out + b * o_bs16 + static_cast<long long>(t) * o_ld16;
  if (t < total) {
    int lo = 0, hi = L - 1;                     // first i with cum[i] > t
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (scum[mid] > t) hi = mid; else lo = mid + 1; }
    const uint4* srow = x + b * x_bs16 + static_cast<long long>(lo) * x_ld16;
    for (int i = lane; i < row16; i += 32) orow[i] = __ldg(srow + i);
  } else {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < row16; i += 32) orow[i] = z;
  }
}

// ---------------------------------------------------------------------------------------- bucketize + embed + sum
template <typename T>
__global__ void __launch_bounds__(256) bucket_embed_sum_kernel(
    const T* __restrict__ text, const T* __restrict__ spk, const T* __restrict__ noise, long long in_bs, int in_ld,
    const float* __restrict__ p_val, const float* __restrict__ e_val, float p_scale, float e_scale,
    const float* __restrict__ pbins, const float* __restrict__ ebins, int nbins, const float* __restrict__ pemb,
    const float* __restrict__ eemb, T* __restrict__ out, T* __restrict__ out_noisy, long long o_bs, int o_ld,
    int32_t* __restrict__ p_idx, int32_t* __restrict__ e_idx, float* __restrict__ p_scaled, float* __restrict__ e_scaled,
    T* __restrict__ pemb_out, T* __restrict__ eemb_out, const float* __restrict__ pos, int B, int Tn, int C) {
  extern __shared__ float sbins[];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) { sbins[i] = pbins[i]; sbins[nbins + i] = ebins[i]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (row >= static_cast<long long>(B) * Tn) return;
  const int t = static_cast<int>(row % Tn), b = static_cast<int>(row / Tn);
  const float pv = __fmul_rn(p_val[row], p_scale), ev = __fmul_rn(e_val[row], e_scale);
  auto lower = [&](const float* bins, float v) {   // #bins < v  == torch.bucketize(v, bins, right=False)
    int lo = 0, hi = nbins;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (bins[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
  };
  const int pi = lower(sbins, pv), ei = lower(sbins + nbins, ev);
  if (lane == 0) {   // the inputs are never written: scaled predictions (modules.py:370,380) go to their own outputs
    if (p_idx != nullptr) p_idx[row] = pi;
    if (e_idx != nullptr) e_idx[row] = ei;
    if (p_scaled != nullptr) p_scaled[row] = pv;
    if (e_scaled != nullptr) e_scaled[row] = ev;
  }
  const long long in_off = b * in_bs + static_cast<long long>(t) * in_ld;
  const long long o_off = b * o_bs + static_cast<long long>(t) * o_ld;
  for (int c = lane * 8; c < C; c += 256) {
    float pe[8], ee[8];
    load8(pemb + static_cast<long long>(pi) * C + c, pe);
    load8(eemb + static_cast<long long>(ei) * C + c, ee);
    if (pemb_out != nullptr) store8(pemb_out + row * C + c, pe);     // the embedding rows themselves (predict_inference)
    if (eemb_out != nullptr) store8(eemb_out + row * C + c, ee);
    if (out == nullptr) continue;
    float tx[8], sp[8], v[8], ps[8];
    load8(text + in_off + c, tx);
    load8(spk + in_off + c, sp);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = ((tx[i] + pe[i]) + sp[i]) + ee[i]; ps[i] = 0.f; }
    // optional: the decoder's position rows (transformer/Models.py:124-125) added on the way out, in the reference's order
    // (x + pos, (x + noise) + pos): the decoder then starts on this buffer without its own full pass over it
    if (pos != nullptr) load8(pos + static_cast<long long>(t) * C + c, ps);
    {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pos != nullptr ? v[i] + ps[i] : v[i];
      store8(out + o_off + c, o);
    }
    if (out_noisy != nullptr) {
      float nz[8], o[8];
      load8(noise + in_off + c, nz);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float vn = v[i] + nz[i]; o[i] = pos != nullptr ? vn + ps[i] : vn; }
      store8(out_noisy + o_off + c, o);
    }
  }
}

inline int blocks_for(long long n, int per) { return static_cast<int>((n + per - 1) / per); }

}  // namespace
}  // namespace sb

using namespace sb;

extern "C" int styler_embed_pos_fwd(const int64_t* src_seq, const float* emb, int32_t vocab, const float* pos, void* out,
                                    int32_t B, int32_t L, int32_t D, int32_t dtype, void* stream) {
  sb::TraceScope trace__("embed_pos", stream);
  SB_REQUIRE(src_seq && emb && pos && out, "embed_pos: null pointer");
  SB_REQUIRE(B > 0 && L > 0 && D % 8 == 0, "embed_pos: bad shape");
  const long long n = static_cast<long long>(B) * L * (D / 8);
  SB_DISPATCH_DTYPE(dtype, T, (embed_pos_kernel<T><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                  src_seq, emb, vocab, pos, static_cast<T*>(out), B * L, L, D)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_add_fwd(const void* a, int64_t a_bstride, int32_t a_ld, const void* a2, int64_t a2_bstride,
                              int32_t a2_ld, const void* rowvec, int32_t rowvec_ld, const float* pos, void* out,
                              int64_t o_bstride, int32_t o_ld, int32_t B, int32_t T, int32_t C, int32_t dtype,
                              void* stream) {
  sb::TraceScope trace__("add", stream);
  SB_REQUIRE(out, "add: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && C % 8 == 0 && a_ld % 8 == 0 && o_ld % 8 == 0 && a_bstride % 8 == 0 && o_bstride % 8 == 0 &&
                 (a2 == nullptr || (a2_ld % 8 == 0 && a2_bstride % 8 == 0)) && (rowvec == nullptr || rowvec_ld % 8 == 0),
             "add: shapes/strides must be multiples of 8 elements");
  SB_REQUIRE((a == nullptr || al16(a)) && al16(out) && (a2 == nullptr || al16(a2)) && (rowvec == nullptr || al16(rowvec)) &&
                 (pos == nullptr || al16(pos)), "add: pointers must be 16-byte aligned");
  const long long n = static_cast<long long>(B) * T * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (add_kernel<TT><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                   static_cast<const TT*>(a), a_bstride, a_ld, static_cast<const TT*>(a2), a2_bstride, a2_ld,
                                   static_cast<const TT*>(rowvec), rowvec_ld, pos, static_cast<TT*>(out), o_bstride, o_ld, B,
                                   T, C)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_cast_fwd(const float* x, void* out, int64_t n, int32_t dtype, void* stream) {
  sb::TraceScope trace__("cast", stream);
  SB_REQUIRE(x && out && n > 0, "cast: bad arguments");
  SB_DISPATCH_DTYPE(dtype, T, (cast_kernel<T><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                  x, static_cast<T*>(out), n)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_lrelu_mean_fwd(const void* a, const void* b, const void* c, float slope_in, float slope_out, void* out,
                                     int64_t n, int32_t dtype, void* stream) {
  sb::TraceScope trace__("lrelu_mean", stream);
  SB_REQUIRE(a && out && n > 0 && n % 8 == 0, "lrelu_mean: bad arguments (n must be a multiple of 8)");
  SB_REQUIRE(slope_in > 0.f && slope_in <= 1.f && slope_out >= 0.f && slope_out <= 1.f, "lrelu_mean: bad slopes");
  SB_REQUIRE(b != nullptr || c == nullptr, "lrelu_mean: pass inputs in order (a, b, c)");
  const int cnt = 1 + (b != nullptr) + (c != nullptr);
  const long long n8 = n / 8;
  SB_DISPATCH_DTYPE(dtype, T, (lrelu_mean_kernel<T><<<blocks_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                  static_cast<const T*>(a), static_cast<const T*>(b), static_cast<const T*>(c), 1.0f / slope_in,
                                  1.0f / cnt, slope_out, static_cast<T*>(out), n8)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_f0_norm_fwd(const float* f0, const int64_t* lens, float* out, int32_t B, int32_t T, void* stream) {
  SB_REQUIRE(f0 && out && B > 0 && T > 0, "f0_norm: bad arguments");
  f0_norm_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(f0, lens, out, T);
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_quantize_index_fwd(const float* x, int32_t* idx, int64_t n, void* stream) {
  sb::TraceScope trace__("quantize_index", stream);
  SB_REQUIRE(x && idx && n > 0, "quantize_index: bad arguments");
  quantize_index_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, idx, n);
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_onehot_conv_fwd(const int32_t* idx, const float* wg, const float* bias, void* out, int32_t B,
                                      int32_t T, int32_t C, int32_t nidx, int32_t KS, int32_t dtype, void* stream) {
  sb::TraceScope trace__("onehot_conv", stream);
  SB_REQUIRE(idx && wg && bias && out, "onehot_conv: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && C % 8 == 0 && nidx > 0 && KS > 0, "onehot_conv: bad shape");
  const long long n = static_cast<long long>(B) * T * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (onehot_conv_kernel<TT><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                   idx, wg, bias, static_cast<TT*>(out), B, T, C, nidx, KS)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_groupnorm_relu_fwd(void* x, int64_t bstride, int32_t ld, const float* gamma, const float* beta,
                                         float* stats_ws, int32_t B, int32_t T, int32_t C, int32_t ch_per_group, float eps,
                                         int32_t dtype, void* stream) {
  sb::TraceScope trace__("groupnorm_relu", stream);
  SB_REQUIRE(x && gamma && beta && stats_ws, "groupnorm: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && ch_per_group % 8 == 0 && C % ch_per_group == 0 && ld % 8 == 0 && bstride % 8 == 0 && al16(x),
             "groupnorm: bad shape/alignment");
  const int groups = C / ch_per_group;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB_DISPATCH_DTYPE(dtype, TT, (gn_stats_kernel<TT><<<B * groups, 256, 0, s>>>(static_cast<const TT*>(x), bstride, ld, stats_ws,
                                                                               T, groups, ch_per_group, eps)));
  SB_LAUNCH_OK();
  const long long n = static_cast<long long>(B) * T * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (gn_apply_relu_kernel<TT><<<blocks_for(n, 256), 256, 0, s>>>(
                                   static_cast<TT*>(x), bstride, ld, stats_ws, gamma, beta, B, T, C, ch_per_group)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_groupnorm_relu_partial_fwd(void* x, int64_t bstride, int32_t ld, const float* gamma, const float* beta,
                                                 const float* partial, int32_t n_part, float* stats_ws, int32_t B, int32_t T,
                                                 int32_t C, float eps, int32_t dtype, void* stream) {
  sb::TraceScope trace__("groupnorm_relu_partial", stream);
  SB_REQUIRE(x && gamma && beta && partial && stats_ws, "groupnorm_partial: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && C % 16 == 0 && ld % 8 == 0 && bstride % 8 == 0 && al16(x) && n_part == (T + 127) / 128,
             "groupnorm_partial: bad shape/alignment (n_part must be ceil(T/128))");
  const int groups = C / 16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gn_finalize_kernel<<<blocks_for(static_cast<long long>(B) * groups, 128), 128, 0, s>>>(partial, stats_ws, B, groups, n_part, T, eps);
  SB_LAUNCH_OK();
  const long long n = static_cast<long long>(B) * T * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (gn_apply_relu_kernel<TT><<<blocks_for(n, 256), 256, 0, s>>>(
                                   static_cast<TT*>(x), bstride, ld, stats_ws, gamma, beta, B, T, C, 16)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_mel_calibrator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const int64_t* mel_len,
                                         const int64_t* src_len, void* out, int64_t o_bstride, int32_t o_ld, int32_t B,
                                         int32_t Tr, int32_t L, int32_t C, int32_t dtype, void* stream) {
  sb::TraceScope trace__("mel_calibrator", stream);
  SB_REQUIRE(x && mel_len && src_len && out, "mel_calibrator: null pointer");
  SB_REQUIRE(B > 0 && Tr > 0 && L > 0 && C % 8 == 0 && x_ld % 8 == 0 && o_ld % 8 == 0 && x_bstride % 8 == 0 &&
                 o_bstride % 8 == 0 && al16(x) && al16(out), "mel_calibrator: bad shape/alignment");
  const long long n = static_cast<long long>(B) * L * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (mel_calibrator_kernel<TT, false><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                   static_cast<const TT*>(x), x_bstride, x_ld, mel_len, src_len, static_cast<TT*>(out),
                                   o_bstride, o_ld, B, Tr, L, C, nullptr, nullptr, nullptr)));
  SB_LAUNCH_OK();
  return 0;
}

// GroupNorm (statistics from the producing conv's gn_partial sums) + ReLU + Mel Calibrator in one pass over the raw conv output.
extern "C" int styler_gn_calibrator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const float* gamma, const float* beta,
                                        const float* partial, int32_t n_part, float* stats_ws, const int64_t* mel_len,
                                        const int64_t* src_len, void* out, int64_t o_bstride, int32_t o_ld, int32_t B, int32_t Tr,
                                        int32_t L, int32_t C, float eps, int32_t dtype, void* stream) {
  sb::TraceScope trace__("gn_calibrator", stream);
  SB_REQUIRE(x && gamma && beta && partial && stats_ws && mel_len && src_len && out, "gn_calibrator: null pointer");
  SB_REQUIRE(B > 0 && Tr > 0 && L > 0 && C % 16 == 0 && x_ld % 8 == 0 && o_ld % 8 == 0 && x_bstride % 8 == 0 && o_bstride % 8 == 0 &&
                 al16(x) && al16(out) && al16(gamma) && al16(beta) && n_part == (Tr + 127) / 128,
             "gn_calibrator: bad shape/alignment (n_part must be ceil(Tr/128))");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = C / 16;
  gn_finalize_kernel<<<blocks_for(static_cast<long long>(B) * groups, 128), 128, 0, s>>>(partial, stats_ws, B, groups, n_part, Tr, eps);
  SB_LAUNCH_OK();
  const long long n = static_cast<long long>(B) * L * (C / 8);
  SB_DISPATCH_DTYPE(dtype, TT, (mel_calibrator_kernel<TT, true><<<blocks_for(n, 256), 256, 0, s>>>(
                                   static_cast<const TT*>(x), x_bstride, x_ld, mel_len, src_len, static_cast<TT*>(out), o_bstride, o_ld,
                                   B, Tr, L, C, stats_ws, gamma, beta)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_classifier_tail_fwd(const void* h, int64_t h_bstride, int32_t h_ld, const float* w, const float* b,
                                          float* out, int32_t B, int32_t L, int32_t C, int32_t dtype, void* stream) {
  sb::TraceScope trace__("classifier_tail", stream);
  SB_REQUIRE(h && w && b && out && B > 0 && L > 0 && C > 0, "classifier_tail: bad arguments");
  SB_REQUIRE(C != 256 || (al16(h) && h_ld % 8 == 0 && h_bstride % 8 == 0 && al16(w)), "classifier_tail: rows must be 16-byte aligned");
  SB_DISPATCH_DTYPE(dtype, TT, (classifier_tail_kernel<TT><<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                   static_cast<const TT*>(h), h_bstride, h_ld, w, b, out, L, C)));
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_duration_round_fwd(const float* log_d, float* dur, int64_t n, float log_offset, float d_control,
                                         void* stream) {
  sb::TraceScope trace__("duration_round", stream);
  SB_REQUIRE(log_d && dur && n > 0, "duration_round: bad arguments");
  duration_round_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(log_d, dur, n, log_offset,
                                                                                           d_control);
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_length_regulator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const int64_t* dur_i64,
                                           const float* dur_f32, void* out, int64_t o_bstride, int32_t o_ld,
                                           int64_t* mel_len, int32_t* cum_ws, int32_t B, int32_t L, int32_t Tmax, int32_t C,
                                           int32_t dtype, void* stream) {
  sb::TraceScope trace__("length_regulator", stream);
  SB_REQUIRE(x && out && mel_len && cum_ws, "length_regulator: null pointer");
  SB_REQUIRE((dur_i64 != nullptr) != (dur_f32 != nullptr), "length_regulator: exactly one of dur_i64 / dur_f32");
  SB_REQUIRE(sb::dtype_ok(dtype), "length_regulator: bad dtype");
  const int es = dtype != STYLER_F32 ? 2 : 4;
  SB_REQUIRE(B > 0 && L > 0 && L <= 8192 && Tmax >= 0 && C > 0, "length_regulator: bad shape");
  SB_REQUIRE((C * es) % 16 == 0 && (static_cast<int64_t>(x_ld) * es) % 16 == 0 && (static_cast<int64_t>(o_ld) * es) % 16 == 0 &&
                 (x_bstride * es) % 16 == 0 && (o_bstride * es) % 16 == 0 && al16(x) && al16(out),
             "length_regulator: rows must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = L >= 1024 ? 1024 : ((L + 31) / 32) * 32;
  lr_scan_kernel<<<B, threads, 0, s>>>(dur_i64, dur_f32, cum_ws, mel_len, L);
  SB_LAUNCH_OK();
  if (Tmax > 0) {
    dim3 grid(ceil_div(Tmax, 8), B);
    lr_expand_kernel<<<grid, 256, L * sizeof(int32_t), s>>>(static_cast<const uint4*>(x), x_bstride * es / 16, x_ld * es / 16,
                                                            cum_ws, static_cast<uint4*>(out), o_bstride * es / 16,
                                                            o_ld * es / 16, L, Tmax, C * es / 16);
    SB_LAUNCH_OK();
  }
  return 0;
}

extern "C" int styler_bucket_embed_sum_fwd(const void* text, const void* spk, const void* noise, int64_t in_bstride,
                                           int32_t in_ld, const float* p_val, const float* e_val, float p_scale,
                                           float e_scale, const float* pitch_bins, const float* energy_bins, int32_t nbins,
                                           const float* pitch_emb, const float* energy_emb, void* out, void* out_noisy,
                                           int64_t o_bstride, int32_t o_ld, int32_t* p_idx, int32_t* e_idx,
                                           float* p_scaled, float* e_scaled, void* pitch_emb_out, void* energy_emb_out,
                                           const float* pos, int32_t B, int32_t T, int32_t C, int32_t dtype, void* stream) {
  sb::TraceScope trace__("bucket_embed_sum", stream);
  SB_REQUIRE(p_val && e_val && pitch_bins && energy_bins && pitch_emb && energy_emb, "bucket_embed_sum: null pointer");
  SB_REQUIRE(out == nullptr || (text != nullptr && spk != nullptr), "bucket_embed_sum: out needs text and spk");
  SB_REQUIRE(out != nullptr || out_noisy == nullptr, "bucket_embed_sum: out_noisy needs out");
  SB_REQUIRE(out_noisy == nullptr || noise != nullptr, "bucket_embed_sum: out_noisy needs noise");
  SB_REQUIRE(p_scaled != p_val && e_scaled != e_val, "bucket_embed_sum: the scaled outputs must not alias the inputs");
  SB_REQUIRE(B > 0 && T > 0 && C % 8 == 0 && in_ld % 8 == 0 && o_ld % 8 == 0 && in_bstride % 8 == 0 && o_bstride % 8 == 0 &&
                 nbins > 0 && nbins <= 4096, "bucket_embed_sum: bad shape");
  SB_REQUIRE((text == nullptr || al16(text)) && (spk == nullptr || al16(spk)) && (out == nullptr || al16(out)) &&
                 (noise == nullptr || al16(noise)) && (out_noisy == nullptr || al16(out_noisy)) && al16(pitch_emb) &&
                 al16(energy_emb) && (pitch_emb_out == nullptr || al16(pitch_emb_out)) &&
                 (energy_emb_out == nullptr || al16(energy_emb_out)) && (pos == nullptr || al16(pos)),
             "bucket_embed_sum: pointers must be 16-byte aligned");
  const long long rows = static_cast<long long>(B) * T;
  SB_DISPATCH_DTYPE(dtype, TT,
                    (bucket_embed_sum_kernel<TT><<<blocks_for(rows, 8), 256, 2 * nbins * sizeof(float),
                                                   static_cast<cudaStream_t>(stream)>>>(
                        static_cast<const TT*>(text), static_cast<const TT*>(spk), static_cast<const TT*>(noise), in_bstride,
                        in_ld, p_val, e_val, p_scale, e_scale, pitch_bins,
                        energy_bins, nbins, pitch_emb, energy_emb, static_cast<TT*>(out), static_cast<TT*>(out_noisy),
                        o_bstride, o_ld, p_idx, e_idx, p_scaled, e_scaled, static_cast<TT*>(pitch_emb_out),
                        static_cast<TT*>(energy_emb_out), pos, B, T, C)));
  SB_LAUNCH_OK();
  return 0;
}
