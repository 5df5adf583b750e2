// STYLERLoss.forward / cal_mel_loss / DomainAdversarialTrainingLoss.forward (loss.py:16-66 of the reference) as one
// deterministic two-stage reduction: the masked means of (mel - target)^2, (mel_postnet - target)^2, |log_d - log_D|,
// |p - f0|, |e - energy| and the three NLL means.  The reference materialises seven masked_select copies per call
// (evaluate.py:88-93, train.py:139-146); here every tensor is read once, nothing is written but 8 floats.
// Bounding roofline: HBM (3 * B*T*80 * 4 bytes of mels dominate: 63 MB at B = 64, T = 1024).
#include "common.cuh"

namespace sb {
namespace {

constexpr int kLossAcc = 8;   // mel_sq, post_sq, n_mel_rows, d_abs, n_src, p_abs, e_abs, (unused)

struct LossArgs {
  const float* mel; const float* post; const float* target; const uint8_t* mel_keep;
  const float* d_pred; const float* d_tgt; const uint8_t* src_keep;
  const float* p_pred; const float* p_tgt; const float* e_pred; const float* e_tgt;
  long long rows_mel, rows_src; int n_mel;
};

__global__ void __launch_bounds__(256) loss_partial_kernel(const LossArgs a, float* __restrict__ partial) {
  float acc[kLossAcc];
#pragma unroll
  for (int i = 0; i < kLossAcc; ++i) acc[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wstride = static_cast<long long>(gridDim.x) * 8;
  // one warp per kept (b, t) row of the mels: n_mel contiguous floats of each of the three tensors
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < a.rows_mel; r += wstride) {
    if (a.mel_keep[r] == 0) continue;                      // warp-uniform
    if (a.mel != nullptr) {
      const float* m = a.mel + r * a.n_mel;
      const float* p = a.post + r * a.n_mel;
      const float* t = a.target + r * a.n_mel;
      for (int c = lane; c < a.n_mel; c += 32) {
        const float tv = t[c], dm = m[c] - tv, dp = p[c] - tv;
        acc[0] = fmaf(dm, dm, acc[0]);
        acc[1] = fmaf(dp, dp, acc[1]);
      }
    }
    if (lane == 0) {
      acc[2] += 1.f;
      if (a.p_pred != nullptr) acc[5] += fabsf(a.p_pred[r] - a.p_tgt[r]);
      if (a.e_pred != nullptr) acc[6] += fabsf(a.e_pred[r] - a.e_tgt[r]);
    }
  }
  if (a.d_pred != nullptr) {
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < a.rows_src; i += static_cast<long long>(gridDim.x) * 256) {
      if (a.src_keep[i] != 0) { acc[3] += fabsf(a.d_pred[i] - a.d_tgt[i]); acc[4] += 1.f; }
    }
  }
  __shared__ float s[8][kLossAcc];
#pragma unroll
  for (int i = 0; i < kLossAcc; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) s[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLossAcc) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s[w][threadIdx.x];      // fixed order: deterministic
    partial[static_cast<long long>(blockIdx.x) * kLossAcc + threadIdx.x] = v;
  }
}

// out[0..5] = mel_loss, mel_postnet_loss, d_loss, p_loss, e_loss, classifier_loss;  out[6], out[7] = kept mel rows, kept src positions
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ partial, int n_blocks, int n_mel,
                                                            const float* __restrict__ post_d, const float* __restrict__ post_p,
                                                            const float* __restrict__ post_e, const int64_t* __restrict__ label,
                                                            int B, float* __restrict__ out) {
  __shared__ double s[kLossAcc + 1][256];
  double acc[kLossAcc + 1];
  for (int i = 0; i <= kLossAcc; ++i) acc[i] = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += 256)
    for (int i = 0; i < kLossAcc; ++i) acc[i] += static_cast<double>(partial[static_cast<long long>(b) * kLossAcc + i]);
  if (post_d != nullptr)
    for (int b = threadIdx.x; b < B; b += 256) {
      const int64_t y = label[b];
      acc[kLossAcc] -= static_cast<double>(post_d[2 * b + y]) + static_cast<double>(post_p[2 * b + y]) + static_cast<double>(post_e[2 * b + y]);
    }
  for (int i = 0; i <= kLossAcc; ++i) s[i][threadIdx.x] = acc[i];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o)
      for (int i = 0; i <= kLossAcc; ++i) s[i][threadIdx.x] += s[i][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double nm = s[2][0], ns = s[4][0];
    // empty selections: nn.MSELoss / nn.L1Loss of an empty tensor is nan in the reference as well
    out[0] = static_cast<float>(s[0][0] / (nm * n_mel));
    out[1] = static_cast<float>(s[1][0] / (nm * n_mel));
    out[2] = static_cast<float>(s[3][0] / ns);
    out[3] = static_cast<float>(s[5][0] / nm);
    out[4] = static_cast<float>(s[6][0] / nm);
    out[5] = post_d != nullptr ? static_cast<float>(s[kLossAcc][0] / B) : 0.f;
    out[6] = static_cast<float>(nm);
    out[7] = static_cast<float>(ns);
  }
}

}  // namespace
}  // namespace sb

extern "C" int64_t styler_loss_workspace_bytes(void) { return static_cast<int64_t>(sizeof(float)) * sb::kLossAcc * 4 * 148 + 256; }

extern "C" int styler_loss_fwd(const float* mel, const float* mel_postnet, const float* mel_target, const uint8_t* mel_keep,
                               const float* log_d_pred, const float* log_d_target, const uint8_t* src_keep, const float* p_pred,
                               const float* p_target, const float* e_pred, const float* e_target, const float* post_d,
                               const float* post_p, const float* post_e, const int64_t* aug_label, int32_t B, int32_t T,
                               int32_t L, int32_t n_mel, void* workspace, int64_t workspace_bytes, float* out, void* stream) {
  using namespace sb;
  sb::TraceScope trace__("loss", stream, B, T, L, n_mel);
  SB_REQUIRE(out != nullptr && workspace != nullptr, "loss: null out / workspace");
  SB_REQUIRE(B > 0 && T >= 0 && L >= 0 && n_mel > 0, "loss: bad shape");
  SB_REQUIRE(workspace_bytes >= styler_loss_workspace_bytes(), "loss: workspace too small");
  SB_REQUIRE((mel == nullptr) == (mel_postnet == nullptr) && (mel == nullptr || mel_target != nullptr), "loss: mel, mel_postnet and mel_target come together");
  SB_REQUIRE(mel_keep != nullptr || (mel == nullptr && p_pred == nullptr && e_pred == nullptr), "loss: mel_keep required");
  SB_REQUIRE((log_d_pred == nullptr) == (log_d_target == nullptr) && (log_d_pred == nullptr || src_keep != nullptr), "loss: duration inputs incomplete");
  SB_REQUIRE((p_pred == nullptr) == (p_target == nullptr) && (e_pred == nullptr) == (e_target == nullptr), "loss: pitch / energy inputs incomplete");
  SB_REQUIRE((post_d == nullptr) == (post_p == nullptr) && (post_d == nullptr) == (post_e == nullptr) && (post_d == nullptr || aug_label != nullptr),
             "loss: the three posteriors and the label come together");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LossArgs a;
  a.mel = mel; a.post = mel_postnet; a.target = mel_target; a.mel_keep = mel_keep;
  a.d_pred = log_d_pred; a.d_tgt = log_d_target; a.src_keep = src_keep;
  a.p_pred = p_pred; a.p_tgt = p_target; a.e_pred = e_pred; a.e_tgt = e_target;
  a.rows_mel = mel_keep != nullptr ? static_cast<long long>(B) * T : 0;
  a.rows_src = static_cast<long long>(B) * L;
  a.n_mel = n_mel;
  const long long work = a.rows_mel > a.rows_src / 32 ? a.rows_mel : a.rows_src / 32;
  int blocks = static_cast<int>((work + 7) / 8);
  const int cap = 4 * num_sms() < 4 * 148 ? 4 * num_sms() : 4 * 148;
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  float* partial = static_cast<float*>(workspace);
  loss_partial_kernel<<<blocks, 256, 0, s>>>(a, partial);
  SB_LAUNCH_OK();
  loss_finalize_kernel<<<1, 256, 0, s>>>(partial, blocks, n_mel, post_d, post_p, post_e, aug_label, B, out);
  SB_LAUNCH_OK();
  return 0;
}
