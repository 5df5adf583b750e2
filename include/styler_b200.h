/* styler_b200 -- C ABI of the B200-native STYLER hot path (libstyler_b200.so).
 *
 * The reference (keonlee9420/STYLER) is pure Python/PyTorch and has no FFI layer; its boundary for this
 * path is the nn.Module API (styler.py:13-58, audio/stft.py:120-160).  This header is the native boundary
 * the Python host mirror (the styler_b200 package) binds with ctypes; each entry point names the reference code it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocates); the library never allocates
 *     or frees device memory and keeps no state between calls except a cache of TMA descriptors;
 *   - activations are channel-last [B][T][C]; element (b,t,c) lives at base + b*bstride + t*ld + c (ELEMENTS);
 *   - `dtype` is the activation/weight storage type: STYLER_F32 (tcgen05 kind::tf32 or fp32 SIMT) or
 *     STYLER_BF16 / STYLER_F16 (tcgen05 kind::f16 with bf16 / f16 operands); accumulation, LayerNorm, softmax, LSTM state are always fp32;
 *   - lengths are int64 (as the reference's src_len/mel_len), masks are derived from lengths;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises;
 *   - return 0 on success, negative = invalid argument, positive = CUDA error code; message via
 *     styler_last_error() (thread-local).  No C++ exception crosses the ABI.
 */
#ifndef STYLER_B200_H_
#define STYLER_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STYLER_F32 0
#define STYLER_BF16 1
#define STYLER_F16 2 /* IEEE half storage (tcgen05 kind::f16, f16 operands): tf32-class accuracy at the bf16 rate */

#define STYLER_IMPL_AUTO 0
#define STYLER_IMPL_SIMT 1   /* fp32 CUDA-core kernels (exact-fp32 parity mode, odd shapes) */
#define STYLER_IMPL_TC 2     /* tcgen05 + TMA kernels */

#define STYLER_ACT_NONE 0
#define STYLER_ACT_RELU 1
#define STYLER_ACT_TANH 2
#define STYLER_ACT_LRELU 3   /* max(v, act_slope * v), 0 < act_slope < 1 (hifigan/models.py:7,93-97,152) */

int styler_version(void);
const char* styler_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t styler_launch_count(void);
/* Debug timeline (tools/timeline.py): styler_debug_trace(1) clears and starts recording a CUDA event pair around every leaf
 * entry point's kernels (also when called from inside a composite entry), (0) stops.  styler_debug_trace_dump synchronises
 * the device and writes one text line per call -- "name a b c d stream start_ms end_ms", times relative to the first
 * record -- into buf; returns the bytes needed. */
int styler_debug_trace(int32_t on);
int64_t styler_debug_trace_dump(char* buf, int64_t cap);

/* ---- Conv1d / Linear over channel-last activations with fused epilogue ----------------------------------
 * y[b,t,n] = act2( LN( act( sum_{tap,c} x[b,t+tap*dilation-pad,c] * w[tap][n][c] + bias[n] ) + residual[b,t,n] ) )
 * then rows t >= lens[b] are zeroed (if lens), then optional row-dot.  Replaces nn.Linear / nn.Conv1d (+ReLU /
 * tanh / folded BatchNorm / residual + LayerNorm + masked_fill) at: transformer/SubLayers.py:41-43,58-59,72-76,
 * 84-87; Layers.py:29,32,121-130; modules.py:30-36,103-160,211-216,250-271,438-465,502-507; styler.py:31.
 * Time-edge taps read zeros (Conv1d zero padding); tiles never cross utterances. */
typedef struct {
  const void* x; int64_t x_bstride; int32_t x_ld;       /* input  [B][T][Cin], dtype */
  int32_t B, T, Cin;
  const void* w;                                         /* packed weights [KS][N][Cin], dtype */
  int32_t N, KS, pad;
  const float* bias;                                     /* [N] fp32 or NULL */
  int32_t act;                                           /* STYLER_ACT_* applied before residual/LN */
  const void* residual; int64_t r_bstride; int32_t r_ld; /* dtype, or NULL; r_ld==0 broadcasts one row over t */
  int32_t residual_is_f32;                               /* 1: `residual` is fp32 whatever `dtype` is */
  const float* ln_gamma; const float* ln_beta; float ln_eps; /* LayerNorm over the N outputs if ln_gamma != NULL */
  int32_t act2;                                          /* STYLER_ACT_* applied after LN */
  const int64_t* lens;                                   /* [B] or NULL */
  const float* dot_w; float dot_b; float* dot_out;       /* optional: dot_out[b*T+t] = <y[b,t,:],dot_w>+dot_b (0 if masked) */
  void* out; int64_t o_bstride; int32_t o_ld;            /* dtype output or NULL */
  float* out_f32; int64_t of_bstride; int32_t of_ld;     /* optional fp32 copy of the output or NULL */
  void* vt; int32_t vt_col0; int64_t vt_bstride; int32_t vt_ld; /* optional: columns n >= vt_col0 are stored
                                                            transposed, vt[b][n-vt_col0][t] (dtype), instead of `out` */
  int32_t dtype;                                         /* STYLER_F32 | STYLER_BF16 | STYLER_F16 */
  int32_t impl;                                          /* STYLER_IMPL_* */
  int32_t dilation;                                      /* tap spacing in time steps; 0 or 1 = dense (nn.Conv1d dilation) */
  float act_slope;                                       /* negative-side slope of STYLER_ACT_LRELU (act and act2) */
  int32_t residual_inv_lrelu;                            /* 1: `residual` holds lrelu(r) (slope act_slope); r is recovered as
                                                            (v < 0 ? v / act_slope : v) before the add, so a HiFi-GAN residual
                                                            chain can be stored in activated form only */
  float* out2_f32;                                       /* optional SECOND fp32 destination, same strides as out_f32 (requires
                                                            out_f32): e.g. the slice of rank 0's peer-mapped gather buffer, so
                                                            the result crosses NVLink from the producing epilogue */
  float* gn_partial;                                     /* optional (tensor-core path, no LN / dot / vt, N % 16 == 0): GroupNorm
                                                            partial statistics of the OUTPUT, fp32 [B][ceil(T/128)][N/16][2] =
                                                            (sum, sum of squares) over 16 channels x the 128-row tile; finished by
                                                            styler_groupnorm_relu_partial_fwd (no separate statistics pass) */
} styler_conv1d_args;
int styler_conv1d_fwd(const styler_conv1d_args* a, void* stream);

/* ---- One FFT block (transformer/Layers.py:26-34: MultiHeadAttention SubLayers.py:31-61 + PositionwiseFeedForward
 * SubLayers.py:81-89) in one call: y = LN2(conv_k2(relu(conv_k1(y1))) + y1) masked, y1 = LN1(fc(attn(qkv(x))) + x) masked.
 * Weights are the packed forms styler_conv1d_fwd takes ([KS][N][Cin], activation dtype; wqkv = W_q/temperature | W_k | W_v
 * stacked on N).  x, y: [B][T][d_model] with explicit strides (y may alias a slice of a wider buffer, not x).
 * workspace: styler_fftblock_workspace_bytes() bytes of device memory owned by the caller (256-byte aligned). */
typedef struct {
  int32_t d_model, d_inner, n_head;                      /* d_model == n_head * 64 */
  const void* wqkv; const float* bqkv;                   /* [1][3*d_model][d_model], [3*d_model] */
  const void* wfc; const float* bfc;                     /* [1][d_model][d_model] */
  const float* ln1_gamma; const float* ln1_beta;
  const void* w1; const float* b1; int32_t ks1;          /* [ks1][d_inner][d_model] */
  const void* w2; const float* b2; int32_t ks2;          /* [ks2][d_model][d_inner] */
  const float* ln2_gamma; const float* ln2_beta;
  float ln_eps;
} styler_fft_weights;
int64_t styler_fftblock_workspace_bytes(int32_t B, int32_t T, int32_t d_model, int32_t d_inner, int32_t dtype);
int styler_fftblock_fwd(const styler_fft_weights* w, const void* x, int64_t x_bstride, int32_t x_ld, void* y,
                        int64_t y_bstride, int32_t y_ld, const int64_t* lens, int32_t B, int32_t T, int32_t dtype,
                        int32_t impl, void* workspace, int64_t ws_bytes, void* stream);
/* bench.py's roofline leg: CUDA events around every FFN first-conv launch issued by styler_fftblock_fwd with T >= min_T;
 * _read synchronises them, returns count / summed ms / (B, T) of the last one and resets. */
int styler_debug_ffn1_timing(int32_t enable, int32_t min_T);
int styler_debug_ffn1_timing_read(float* total_ms, int32_t* launches, int64_t* last_B, int64_t* last_T);

/* ---- Composite entries: whole sub-graphs of the forward in one native call, intermediates in one caller-owned workspace.
 * StylePredictor.forward (modules.py:457-465): Conv(k)+ReLU+LN, Conv(k)+ReLU+LN, Linear(C->1), masked_fill -> out fp32 [B][T]. */
typedef struct {
  int32_t c_in, channels, ks;                             /* 256, 256, 3 */
  const void* w1; const float* b1; const float* ln1_gamma; const float* ln1_beta;   /* [ks][channels][c_in] */
  const void* w2; const float* b2; const float* ln2_gamma; const float* ln2_beta;   /* [ks][channels][channels] */
  const float* lin_w; float lin_b;                        /* Linear(channels -> 1) */
  float ln_eps;
} styler_predictor_weights;
int64_t styler_predictor_workspace_bytes(int32_t B, int32_t T, int32_t channels, int32_t dtype);
int styler_predictor_fwd(const styler_predictor_weights* w, const void* x, int64_t x_bstride, int32_t x_ld,
                         const int64_t* lens, float* out, int32_t B, int32_t T, int32_t dtype, int32_t impl,
                         void* workspace, int64_t ws_bytes, void* stream);
/* PostNet.forward + the residual add of styler.py:34 (transformer/Layers.py:121-130; eval BatchNorm folded into w/b at pack
 * time): mel_act = the mel in the activation dtype [B][T][n_mel] (conv input), mel_f32 = the fp32 mel (residual);
 * post_out (and post_out2 if not NULL, e.g. a peer-mapped gather slice) = postnet(mel) + mel, fp32 [B][T][n_mel]. */
typedef struct {
  int32_t n_layers, n_mel, channels, ks;                  /* 5, 80, 512, 5 */
  const void* w[8]; const float* b[8];                    /* layer j: [ks][channels or n_mel][n_mel or channels] */
} styler_postnet_weights;
int64_t styler_postnet_workspace_bytes(int32_t B, int32_t T, int32_t channels, int32_t dtype);
int styler_postnet_fwd(const styler_postnet_weights* w, const void* mel_act, const float* mel_f32, float* post_out,
                       float* post_out2, int32_t B, int32_t T, int32_t dtype, int32_t impl, void* workspace,
                       int64_t ws_bytes, void* stream);
/* STYLER.decode (styler.py:29-37) = Decoder.forward (transformer/Models.py:111-135: x + pos rows, n_layers FFT blocks) +
 * mel_linear + PostNet + residual.  x: [B][T][d_model] contiguous, activation dtype; pos: fp32 [T][d_model] (the caller
 * supplies rows beyond max_seq_len, Models.py:120-122), or NULL when x already carries the position rows
 * (styler_bucket_embed_sum_fwd with pos): x is then read in place by the first block; mel_out / post_out: fp32 [B][T][n_mel] contiguous; mel_out2 /
 * post_out2: optional second destinations (same layout).  postnet == NULL: use_postnet=False, post_out is not written. */
typedef struct {
  int32_t n_layers; const styler_fft_weights* layers;     /* array of n_layers */
  const void* mel_w; const float* mel_b; int32_t n_mel;   /* [1][n_mel][d_model] */
  const styler_postnet_weights* postnet;                  /* or NULL */
} styler_decoder_weights;
int64_t styler_decoder_workspace_bytes(const styler_decoder_weights* w, int32_t B, int32_t T, int32_t dtype);
int styler_decoder_fwd(const styler_decoder_weights* w, const void* x, const float* pos, const int64_t* lens, float* mel_out,
                       float* post_out, float* mel_out2, float* post_out2, int32_t B, int32_t T, int32_t dtype, int32_t impl,
                       void* workspace, int64_t ws_bytes, void* stream);

/* ---- Scaled-dot-product multi-head self-attention (transformer/Modules.py:14-25, SubLayers.py:44-56) ----
 * qk: [B][T][2*H*64] (Q columns then K columns, head h at h*64; 1/temperature already folded into Q),
 * vt: [B][H*64][vt_ld] (V transposed) -- or NULL, in which case V is read row-major from qk's columns [2*H*64, 3*H*64)
 * (qk is then the fused [B][T][3*H*64] QKV projection) and fed to the tensor core as an MN-major operand;
 * ctx: [B][T][H*64].  Keys >= lens[b] are masked (-inf); padded query rows are computed like the reference.
 * The attention matrix is never written. */
int styler_attention_fwd(const void* qk, int64_t qk_bstride, int32_t qk_ld, const void* vt, int64_t vt_bstride,
                         int32_t vt_ld, const int64_t* lens, void* ctx, int64_t ctx_bstride, int32_t ctx_ld,
                         int32_t B, int32_t T, int32_t H, int32_t dtype, int32_t impl, void* stream);

/* ---- Embedding + sinusoid position (transformer/Models.py:52-55,73-74) and position add (:124-125) ---- */
int styler_embed_pos_fwd(const int64_t* src_seq, const float* emb, int32_t vocab, const float* pos, void* out,
                         int32_t B, int32_t L, int32_t D, int32_t dtype, void* stream);
/* out[b,t,:] = (a ? a[b,t,:] : 0) + (rowvec ? rowvec[b,:] : 0) + (pos ? pos[t,:] : 0) + (a2 ? a2[b,t,:] : 0)   (dtype io) */
int styler_add_fwd(const void* a, int64_t a_bstride, int32_t a_ld, const void* a2, int64_t a2_bstride,
                   int32_t a2_ld, const void* rowvec, int32_t rowvec_ld, const float* pos, void* out,
                   int64_t o_bstride, int32_t o_ld, int32_t B, int32_t T, int32_t C, int32_t dtype, void* stream);
int styler_cast_fwd(const float* x, void* out, int64_t n, int32_t dtype, void* stream);

/* ---- HiFi-GAN multi-receptive-field fusion + the leaky ReLU before the next layer (hifigan/models.py:152-162):
 * inputs a, b, c (b, c optional) hold y_k = lrelu(x_k, slope_in); out = lrelu(mean_k x_k, slope_out), dtype io,
 * n contiguous elements (multiple of 8). */
int styler_lrelu_mean_fwd(const void* a, const void* b, const void* c, float slope_in, float slope_out, void* out,
                          int64_t n, int32_t dtype, void* stream);

/* ---- quantize_1D_torch index (utils.py:417-429): 0 if x<=0 else rint(x*255)+1; bit-exact integers ---- */
int styler_quantize_index_fwd(const float* x, int32_t* idx, int64_t n, void* stream);
/* First audio-encoder conv on a 257-way one-hot input (modules.py:218-223 + :119-146) as a 5-tap weight gather:
 * out[b,t,c] = bias[c] + sum_tap wg[tap][idx[b,t+tap-2]][c]   (out-of-range taps contribute 0) */
int styler_onehot_conv_fwd(const int32_t* idx, const float* wg, const float* bias, void* out, int32_t B, int32_t T,
                           int32_t C, int32_t nidx, int32_t KS, int32_t dtype, void* stream);

/* Same normalisation from the partial statistics a convolution left in `partial` (styler_conv1d_args.gn_partial, 16 channels
 * per group, n_part = ceil(T/128) tiles per utterance): a B x groups finalize (fp64 combine, fixed order) + the apply pass.
 * stats_ws: fp32 [B][C/16][2]. */
int styler_groupnorm_relu_partial_fwd(void* x, int64_t bstride, int32_t ld, const float* gamma, const float* beta,
                                      const float* partial, int32_t n_part, float* stats_ws, int32_t B, int32_t T, int32_t C,
                                      float eps, int32_t dtype, void* stream);
/* ---- GroupNorm(C/16 groups) over (16 channels x ALL T incl. padding) + ReLU, in place (modules.py:113,168-172) */
int styler_groupnorm_relu_fwd(void* x, int64_t bstride, int32_t ld, const float* gamma, const float* beta,
                              float* stats_ws /* [B*C/16*2] */, int32_t B, int32_t T, int32_t C, int32_t ch_per_group,
                              float eps, int32_t dtype, void* stream);

/* ---- Mel Calibrator (utils.py:351-384): per utterance resample mel_len[b] frames to src_len[b] rows by
 * segment mean (sizes q+1 for the first r segments, q after) or repeat; rows >= src_len[b] are zero. */
int styler_mel_calibrator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const int64_t* mel_len,
                              const int64_t* src_len, void* out, int64_t o_bstride, int32_t o_ld, int32_t B,
                              int32_t Tr, int32_t L, int32_t C, int32_t dtype, void* stream);
/* GroupNorm (+affine, ReLU; statistics from the producing conv's gn_partial sums, as styler_groupnorm_relu_partial_fwd) applied
 * while the Mel Calibrator reads the frames: x is the RAW conv output and is not modified; the normalised tensor never exists. */
int styler_gn_calibrator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const float* gamma, const float* beta,
                             const float* partial, int32_t n_part, float* stats_ws, const int64_t* mel_len, const int64_t* src_len,
                             void* out, int64_t o_bstride, int32_t o_ld, int32_t B, int32_t Tr, int32_t L, int32_t C, float eps,
                             int32_t dtype, void* stream);

/* ---- One layer of a bidirectional LSTM over the padded grid (modules.py:179-182; nn.LSTM, gates i,f,g,o).
 * gx: fp32 [B][L][2][H][4] = x @ W_ih^T + (b_ih + b_hh) in QUAD order: direction (fwd, rev), hidden unit, gate (i,f,g,o)
 * innermost -- i.e. the rows of [W_ih_fwd ; W_ih_rev] permuted from PyTorch's [dir][gate][unit] to [dir][unit][gate] before the
 * projection (produced by styler_conv1d_fwd; 16-byte aligned);
 * whh: fp32 [2][4H][H] in PyTorch row order; out: [B][L][2H] (fwd | rev), dtype. */
int styler_bilstm_layer_fwd(const float* gx, const float* whh, void* out, int64_t o_bstride, int32_t o_ld,
                            int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream);

/* ---- AugmentationClassifier tail (modules.py:34-45): Linear(256->2) + LogSoftmax + mean over L ---- */
int styler_classifier_tail_fwd(const void* h, int64_t h_bstride, int32_t h_ld, const float* w /*[2][C]*/,
                               const float* b /*[2]*/, float* out /*[B][2]*/, int32_t B, int32_t L, int32_t C,
                               int32_t dtype, void* stream);

/* ---- Duration rounding (modules.py:357-358): clamp(rint(exp(log_d) - offset) * d_control, min 0) ---- */
int styler_duration_round_fwd(const float* log_d, float* dur, int64_t n, float log_offset, float d_control,
                              void* stream);
/* ---- LengthRegulator (modules.py:396-423, utils.py:332-348): integer scan + gather-expand.
 * dur_i64 (teacher) or dur_f32 (rounded prediction; truncated toward zero like int(x.item())).
 * mel_len[b] = un-cropped total; out rows >= min(mel_len[b], Tmax) are zero.  Integers are bit-exact.
 * cum_ws: int32 [B][L] workspace (inclusive scan, kept for tests). */
int styler_length_regulator_fwd(const void* x, int64_t x_bstride, int32_t x_ld, const int64_t* dur_i64,
                                const float* dur_f32, void* out, int64_t o_bstride, int32_t o_ld, int64_t* mel_len,
                                int32_t* cum_ws, int32_t B, int32_t L, int32_t Tmax, int32_t C, int32_t dtype,
                                void* stream);

/* ---- bucketize + embedding + 4-way sum (modules.py:365-385, :299-307):
 * x[b,t,:] = text[b,t,:] + pitch_emb[bucket(p[b,t]*p_scale, pitch_bins)] + spk[b,t,:] + energy_emb[bucket(e*e_scale)]
 * (same summation order as the reference); optionally also x_noisy = x + noise; bucket = #bins < value
 * (torch.bucketize right=False); indices optionally returned (int32) for bit-exact tests.  The inputs are never written:
 * the scaled predictions p*p_control / e*e_control the reference returns (modules.py:370,380) go to p_scaled / e_scaled
 * (fp32 [B][T], may be NULL); pitch_emb_out / energy_emb_out (activation dtype, contiguous [B][T][C], may be NULL) receive
 * the two embedding rows on their own, which is what predict_inference returns (modules.py:299-309).  out may be NULL
 * (then text / spk / noise are not read).  pos (fp32 [T][C], may be NULL): the decoder's position rows, added to out and
 * out_noisy on the way out -- styler_decoder_fwd is then called with pos = NULL and skips its own pass. */
int styler_bucket_embed_sum_fwd(const void* text, const void* spk, const void* noise, int64_t in_bstride,
                                int32_t in_ld, const float* p_val, const float* e_val, float p_scale, float e_scale,
                                const float* pitch_bins, const float* energy_bins, int32_t nbins,
                                const float* pitch_emb, const float* energy_emb, void* out, void* out_noisy,
                                int64_t o_bstride, int32_t o_ld, int32_t* p_idx, int32_t* e_idx, float* p_scaled,
                                float* e_scaled, void* pitch_emb_out, void* energy_emb_out, const float* pos, int32_t B, int32_t T,
                                int32_t C, int32_t dtype, void* stream);

/* ---- TacotronSTFT.mel_spectrogram (audio/stft.py:51-79,141-160; audio_processing.py:80-86):
 * reflect pad n_fft/2, periodic Hann, 1024-point real FFT per hop, magnitude, mel_basis matmul,
 * log(clamp(.,1e-5)), energy = L2 norm over the 513 bins.  y fp32 [B][N]; mel fp32 [B][n_mels][F];
 * energy fp32 [B][F]; F = 1 + N/hop.  mel_basis fp32 [n_mels][n_fft/2+1]. */
int styler_stft_mel_fwd(const float* y, int32_t B, int32_t N, const float* mel_basis, int32_t n_mels,
                        int32_t* band_ws /* [2*n_mels] workspace: non-zero band of each filter row */, float* mel,
                        float* energy, void* stream);
/* ---- Fused preprocessing front end (audio/tools.py:37-55 get_mel_from_wav; utils.py:412-416 energy_rescaling;
 * the data/ preprocessing scripts store mel.T): the same kernel with the waveform scaling (y * in_scale, e.g. 1/max_wav_value) or the
 * norm=False clamp to [-1,1] (clip_flag[b] = 1 if any sample was < -1, exactly what the reference's `clipt` detects)
 * applied while the samples are staged, the mel written frame-major [B][F][n_mels] when frame_major != 0 (the layout
 * STYLER.forward takes), and e_input[b][f] = clip((energy - e_min) / (e_max - e_min), 0, 1) written next to the raw
 * energy when e_input != NULL.  clip_flag / e_input may be NULL.
 * n_samples (int64 [B], device, may be NULL = every row holds N samples): the reference transforms every utterance ALONE
 * (audio/tools.py:37-55 -> audio/stft.py:58-62), so row b of the zero-padded [B][N] batch is reflected around its own end
 * n_samples[b] (must exceed n_fft/2) and yields 1 + n_samples[b]/hop frames; the remaining frames of the padded
 * [.., F = 1 + N/hop] outputs are written as zeros (the collation padding of dataset.py:160-166). */
int styler_stft_mel_ex_fwd(const float* y, int32_t B, int32_t N, const float* mel_basis, int32_t n_mels,
                           int32_t* band_ws, float* mel, float* energy, float in_scale, int32_t clamp, int32_t* clip_flag,
                           int32_t frame_major, float* e_input, float e_min, float e_max, const int64_t* n_samples,
                           void* stream);
/* ---- f0_normalization / speaker_normalization (utils.py:387-409) over a padded batch of log-f0 contours [B][T]
 * (unvoiced frames marked <= -1e10 keep their value; rows with undefined statistics and frames >= lens[b] are zero). */
int styler_f0_norm_fwd(const float* f0, const int64_t* lens, float* out, int32_t B, int32_t T, void* stream);

/* ---- Peer memory for the fused compute + gather (one process per GPU; replaces nn.DataParallel's gather, train.py:33):
 * peer_alloc: cudaMalloc'ed, zeroed region + its 64-byte CUDA IPC handle; peer_open maps another process's region (NVLink
 * P2P, peer access enabled lazily); the fp32 outputs of the last convolutions are then written straight into the mapped
 * slice (out_f32 / out2_f32 of styler_conv1d_fwd).  peer_signal enqueues a one-thread kernel that publishes `value` at
 * `flag` (device or peer memory) with system-scope release semantics after everything enqueued before it on `stream`;
 * peer_wait enqueues a kernel that spins until flags[i * stride] >= value for all i < n (system-scope acquire, ~10 s
 * watchdog that traps instead of hanging). */
int styler_peer_alloc(int64_t bytes, void** dptr, void* handle64);
int styler_peer_open(const void* handle64, void** dptr);
int styler_peer_close(void* dptr);
int styler_peer_free(void* dptr);
int styler_peer_signal(void* flag, uint64_t value, void* stream);
int styler_peer_wait(const void* flags, int32_t n, int64_t stride, uint64_t value, void* stream);

/* ---- A/B switches of the library (each also read once from the environment as STYLER_<NAME>): "TC_2CTA" (CTA pairs /
 * tcgen05 cta_group::2 for the big bf16 convolutions: 0 off, 1 where it pays, 2 wherever legal), "TC_PERSIST", "CONV_WIN",
 * "TC_BN", "TC_SMEM_KB", "PDL", "ATTN_PERSIST".  value < 0 restores the environment / built-in default.  Affects only
 * which kernel variant runs, never the results' contract. */
int styler_set_tuning(const char* name, int32_t value);

/* ---- Debug/tuning hook: when a device buffer of capacity_ctas*8 int64 is set, every tcgen05 conv launch with at most
 * capacity_ctas CTAs writes 8 clock64() phase stamps per CTA into it (tools/phase_timing.py); NULL disables. */
int styler_debug_set_phase_buffer(int64_t* buf, int32_t capacity_ctas);

/* ---- Losses of evaluate.py / train.py (loss.py:16-66): STYLERLoss.forward, cal_mel_loss, DomainAdversarialTrainingLoss ----
 * One deterministic two-stage reduction over tensors read in place (the reference materialises seven masked_select copies):
 *   out[0] = mean over kept (b,t), all n_mel channels of (mel - mel_target)^2          (nn.MSELoss, loss.py:21)
 *   out[1] = the same for mel_postnet                                                  (loss.py:22)
 *   out[2] = mean over kept (b,l) of |log_d_pred - log_d_target|                       (nn.L1Loss, loss.py:41)
 *   out[3], out[4] = mean over kept (b,t) of |p_pred - p_target|, |e_pred - e_target|  (loss.py:42-43)
 *   out[5] = sum of the three nn.NLLLoss means, -mean_b post_x[b][aug_label[b]]        (loss.py:45-47, 61-64)
 *   out[6], out[7] = number of kept mel rows / source positions
 * mel* fp32 [B][T][n_mel] contiguous; *_keep bool (uint8) [B][T] / [B][L] with 1 = KEEP (the callers pass ~mel_mask, ~src_mask);
 * log_d*, fp32 [B][L]; p*, e* fp32 [B][T]; post_* fp32 [B][2] log-probabilities; aug_label int64 [B].  Every group may be NULL
 * (cal_mel_loss: only the mel group; DAT loss: only the posteriors); its outputs are then 0 / nan as an empty selection is in torch.
 * workspace: styler_loss_workspace_bytes() bytes. */
int64_t styler_loss_workspace_bytes(void);
int styler_loss_fwd(const float* mel, const float* mel_postnet, const float* mel_target, const uint8_t* mel_keep,
                    const float* log_d_pred, const float* log_d_target, const uint8_t* src_keep, const float* p_pred,
                    const float* p_target, const float* e_pred, const float* e_target, const float* post_d, const float* post_p,
                    const float* post_e, const int64_t* aug_label, int32_t B, int32_t T, int32_t L, int32_t n_mel,
                    void* workspace, int64_t workspace_bytes, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STYLER_B200_H_ */
