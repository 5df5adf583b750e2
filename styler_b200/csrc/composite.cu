// Composite C-ABI entries: whole sub-graphs of the forward enqueued from native code in one call, every
// intermediate carved from one caller-provided workspace -- same kernels as the single-op entries, one host call instead of
// many (the Python host spends ~25 us per kernel call; a single-utterance forward is launch-bound without them).
//   styler_predictor_fwd : StylePredictor.forward                      (modules.py:457-465)
//   styler_postnet_fwd   : PostNet.forward + the caller's residual add   (transformer/Layers.py:121-130, styler.py:34)
//   styler_decoder_fwd   : Decoder.forward + mel_linear + PostNet = STYLER.decode   (transformer/Models.py:111-135, styler.py:29-37)
#include <string.h>

#include "common.cuh"

namespace sb {
namespace {
inline size_t align_up256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
inline size_t esz(int dtype) { return dtype != STYLER_F32 ? 2 : 4; }
}  // namespace
}  // namespace sb

using namespace sb;

extern "C" int64_t styler_predictor_workspace_bytes(int32_t B, int32_t T, int32_t C, int32_t dtype) {
  return static_cast<int64_t>(2 * align_up256(static_cast<size_t>(B) * T * C * esz(dtype)));
}

extern "C" int styler_predictor_fwd(const styler_predictor_weights* w, const void* x, int64_t x_bstride, int32_t x_ld,
                                    const int64_t* lens, float* out, int32_t B, int32_t T, int32_t dtype, int32_t impl,
                                    void* workspace, int64_t ws_bytes, void* stream) {
  SB_REQUIRE(w && x && out && workspace, "predictor: null pointer");
  SB_REQUIRE(B > 0 && T > 0 && w->channels > 0 && w->ks > 0, "predictor: bad shape");
  SB_REQUIRE(sb::dtype_ok(dtype), "predictor: bad dtype %d", dtype);
  const int C = w->channels;
  SB_REQUIRE(ws_bytes >= styler_predictor_workspace_bytes(B, T, C, dtype), "predictor: workspace too small");
  uint8_t* p = static_cast<uint8_t*>(workspace);
  void* h1 = p;
  void* h2 = p + align_up256(static_cast<size_t>(B) * T * C * esz(dtype));
  int rc;
  {   // Conv k + ReLU + LayerNorm (no masking between the layers: the halo reads un-masked padded rows, modules.py:438-453)
    styler_conv1d_args a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.x_bstride = x_bstride; a.x_ld = x_ld; a.B = B; a.T = T; a.Cin = w->c_in;
    a.w = w->w1; a.N = C; a.KS = w->ks; a.pad = (w->ks - 1) / 2; a.bias = w->b1; a.act = STYLER_ACT_RELU;
    a.ln_gamma = w->ln1_gamma; a.ln_beta = w->ln1_beta; a.ln_eps = w->ln_eps;
    a.out = h1; a.o_bstride = static_cast<int64_t>(T) * C; a.o_ld = C; a.dtype = dtype; a.impl = impl;
    if ((rc = styler_conv1d_fwd(&a, stream)) != 0) return rc;
  }
  {   // Conv k + ReLU + LayerNorm + Linear(C -> 1) + masked_fill
    styler_conv1d_args a;
    memset(&a, 0, sizeof(a));
    a.x = h1; a.x_bstride = static_cast<int64_t>(T) * C; a.x_ld = C; a.B = B; a.T = T; a.Cin = C;
    a.w = w->w2; a.N = C; a.KS = w->ks; a.pad = (w->ks - 1) / 2; a.bias = w->b2; a.act = STYLER_ACT_RELU;
    a.ln_gamma = w->ln2_gamma; a.ln_beta = w->ln2_beta; a.ln_eps = w->ln_eps; a.lens = lens;
    a.dot_w = w->lin_w; a.dot_b = w->lin_b; a.dot_out = out;
    // the CUDA-core path stages the pre-LayerNorm rows in the output buffer; the tensor-core path needs none
    const bool simt = impl == STYLER_IMPL_SIMT || (impl == STYLER_IMPL_AUTO && static_cast<int64_t>(B) * T < 64);
    if (simt) { a.out = h2; a.o_bstride = static_cast<int64_t>(T) * C; a.o_ld = C; }
    a.dtype = dtype; a.impl = impl;
    if ((rc = styler_conv1d_fwd(&a, stream)) != 0) return rc;
  }
  return 0;
}

extern "C" int64_t styler_postnet_workspace_bytes(int32_t B, int32_t T, int32_t channels, int32_t dtype) {
  return static_cast<int64_t>(2 * align_up256(static_cast<size_t>(B) * T * channels * esz(dtype)));
}

// mel_act: the mel in the activation dtype (conv input); mel_f32: the fp32 mel (residual); post_out = postnet(mel) + mel (fp32).
extern "C" int styler_postnet_fwd(const styler_postnet_weights* w, const void* mel_act, const float* mel_f32, float* post_out,
                                  float* post_out2, int32_t B, int32_t T, int32_t dtype, int32_t impl, void* workspace,
                                  int64_t ws_bytes, void* stream) {
  SB_REQUIRE(w && mel_act && mel_f32 && post_out && workspace, "postnet: null pointer");
  SB_REQUIRE(w->n_layers >= 2 && w->n_layers <= 8 && w->channels > 0 && w->n_mel > 0 && w->ks > 0, "postnet: bad weights");
  SB_REQUIRE(sb::dtype_ok(dtype), "postnet: bad dtype %d", dtype);
  const int CH = w->channels, NM = w->n_mel, pad = (w->ks - 1) / 2;
  SB_REQUIRE(ws_bytes >= styler_postnet_workspace_bytes(B, T, CH, dtype), "postnet: workspace too small");
  uint8_t* p = static_cast<uint8_t*>(workspace);
  void* buf[2] = {p, p + align_up256(static_cast<size_t>(B) * T * CH * esz(dtype))};
  const void* cur = mel_act;
  int cur_c = NM, rc;
  for (int j = 0; j + 1 < w->n_layers; ++j) {   // Conv1d k5 + folded eval BatchNorm + tanh (Layers.py:72-119)
    styler_conv1d_args a;
    memset(&a, 0, sizeof(a));
    a.x = cur; a.x_bstride = static_cast<int64_t>(T) * cur_c; a.x_ld = cur_c; a.B = B; a.T = T; a.Cin = cur_c;
    a.w = w->w[j]; a.N = CH; a.KS = w->ks; a.pad = pad; a.bias = w->b[j]; a.act = STYLER_ACT_TANH;
    a.out = buf[j & 1]; a.o_bstride = static_cast<int64_t>(T) * CH; a.o_ld = CH; a.dtype = dtype; a.impl = impl;
    if ((rc = styler_conv1d_fwd(&a, stream)) != 0) return rc;
    cur = buf[j & 1];
    cur_c = CH;
  }
  {   // last conv + folded BatchNorm, no activation, + the residual mel (styler.py:34), fp32 out
    const int j = w->n_layers - 1;
    styler_conv1d_args a;
    memset(&a, 0, sizeof(a));
    a.x = cur; a.x_bstride = static_cast<int64_t>(T) * cur_c; a.x_ld = cur_c; a.B = B; a.T = T; a.Cin = cur_c;
    a.w = w->w[j]; a.N = NM; a.KS = w->ks; a.pad = pad; a.bias = w->b[j];
    a.residual = mel_f32; a.r_bstride = static_cast<int64_t>(T) * NM; a.r_ld = NM; a.residual_is_f32 = 1;
    if (dtype == STYLER_F32) { a.out = post_out; a.o_bstride = static_cast<int64_t>(T) * NM; a.o_ld = NM; }
    else { a.out_f32 = post_out; a.of_bstride = static_cast<int64_t>(T) * NM; a.of_ld = NM; a.out2_f32 = post_out2; }
    a.dtype = dtype; a.impl = impl;
    if ((rc = styler_conv1d_fwd(&a, stream)) != 0) return rc;
    if (dtype == STYLER_F32 && post_out2 != nullptr)
      SB_CUDA_OK(cudaMemcpyAsync(post_out2, post_out, static_cast<size_t>(B) * T * NM * sizeof(float), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

extern "C" int64_t styler_decoder_workspace_bytes(const styler_decoder_weights* w, int32_t B, int32_t T, int32_t dtype) {
  if (w == nullptr || w->n_layers < 1) return -1;
  const int D = w->layers[0].d_model, DI = w->layers[0].d_inner;
  size_t n = 2 * align_up256(static_cast<size_t>(B) * T * D * esz(dtype));                 // ping-pong activations
  n += static_cast<size_t>(styler_fftblock_workspace_bytes(B, T, D, DI, dtype));
  n += align_up256(static_cast<size_t>(B) * T * w->n_mel * esz(dtype));                     // mel in the activation dtype
  if (w->postnet != nullptr) n += static_cast<size_t>(styler_postnet_workspace_bytes(B, T, w->postnet->channels, dtype));
  return static_cast<int64_t>(n);
}

extern "C" int styler_decoder_fwd(const styler_decoder_weights* w, const void* x, const float* pos, const int64_t* lens,
                                  float* mel_out, float* post_out, float* mel_out2, float* post_out2, int32_t B, int32_t T,
                                  int32_t dtype, int32_t impl, void* workspace, int64_t ws_bytes, void* stream) {
  SB_REQUIRE(w && x && mel_out && workspace, "decoder: null pointer");
  SB_REQUIRE(w->n_layers >= 1 && w->n_layers <= 16 && w->layers != nullptr && w->mel_w != nullptr, "decoder: bad weights");
  SB_REQUIRE(w->postnet == nullptr || post_out != nullptr, "decoder: post_out required with a PostNet");
  SB_REQUIRE(B > 0 && T > 0, "decoder: bad shape");
  SB_REQUIRE(ws_bytes >= styler_decoder_workspace_bytes(w, B, T, dtype), "decoder: workspace too small (%lld bytes)",
             static_cast<long long>(ws_bytes));
  const int D = w->layers[0].d_model, DI = w->layers[0].d_inner, NM = w->n_mel;
  const size_t es = esz(dtype);
  uint8_t* p = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += align_up256(bytes); return r; };
  void* act[2] = {take(static_cast<size_t>(B) * T * D * es), take(static_cast<size_t>(B) * T * D * es)};
  const int64_t fft_ws = styler_fftblock_workspace_bytes(B, T, D, DI, dtype);
  void* ws_fft = take(static_cast<size_t>(fft_ws));
  void* mel_act = take(static_cast<size_t>(B) * T * NM * es);
  const int64_t xs = static_cast<int64_t>(T) * D;
  int rc;
  // x + position rows (transformer/Models.py:124-125; the table beyond max_seq_len is the caller's, :120-122).  pos == NULL: the
  // caller's x already carries them (styler_bucket_embed_sum_fwd with pos): the first block reads x in place, no pass over it here
  const void* in = x;
  int cur = 1;                                   // act[cur ^ 1] receives the next block's output
  if (pos != nullptr) {
    if ((rc = styler_add_fwd(x, xs, D, nullptr, 0, 0, nullptr, 0, pos, act[0], xs, D, B, T, D, dtype, stream)) != 0) return rc;
    in = act[0];
    cur = 0;
  }
  for (int l = 0; l < w->n_layers; ++l) {
    if ((rc = styler_fftblock_fwd(&w->layers[l], in, xs, D, act[cur ^ 1], xs, D, lens, B, T, dtype, impl, ws_fft, fft_ws,
                                  stream)) != 0)
      return rc;
    cur ^= 1;
    in = act[cur];
  }
  {   // mel_linear (styler.py:31): fp32 result (+ optional second destination) and the activation-dtype copy the PostNet reads
    styler_conv1d_args a;
    memset(&a, 0, sizeof(a));
    a.x = act[cur]; a.x_bstride = xs; a.x_ld = D; a.B = B; a.T = T; a.Cin = D;
    a.w = w->mel_w; a.N = NM; a.KS = 1; a.bias = w->mel_b; a.dtype = dtype; a.impl = impl;
    if (dtype == STYLER_F32) { a.out = mel_out; a.o_bstride = static_cast<int64_t>(T) * NM; a.o_ld = NM; }
    else {
      a.out = mel_act; a.o_bstride = static_cast<int64_t>(T) * NM; a.o_ld = NM;
      a.out_f32 = mel_out; a.of_bstride = static_cast<int64_t>(T) * NM; a.of_ld = NM; a.out2_f32 = mel_out2;
    }
    if ((rc = styler_conv1d_fwd(&a, stream)) != 0) return rc;
    if (dtype == STYLER_F32 && mel_out2 != nullptr)
      SB_CUDA_OK(cudaMemcpyAsync(mel_out2, mel_out, static_cast<size_t>(B) * T * NM * sizeof(float), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
  }
  if (w->postnet == nullptr) return 0;     // use_postnet=False (styler.py:33-36): the caller returns the mel twice
  return styler_postnet_fwd(w->postnet, dtype == STYLER_F32 ? static_cast<const void*>(mel_out) : mel_act, mel_out, post_out,
                            post_out2, B, T, dtype, impl, p, ws_bytes - static_cast<int64_t>(p - static_cast<uint8_t*>(workspace)),
                            stream);
}
