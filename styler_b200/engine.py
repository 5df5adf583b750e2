"""Host-side orchestration of the STYLER forward over the C-ABI kernels (no torch compute on the hot path).

`Engine` packs a reference-format state_dict once into kernel layouts (conv weights as [taps][N][Cin], fused QKV with
1/temperature folded into W_q, BatchNorm folded into the PostNet convs, one-hot conv tables, stacked LSTM weights) and
then runs the eval-mode forward of styler.py:39-58 / modules.py:311-387 as a sequence of kernel launches on the current
CUDA stream.  Precision modes:
  "bf16": bf16 activations/weights, tcgen05 kind::f16, fp32 accumulate / LayerNorm / softmax / LSTM state (throughput mode)
  "fp16": IEEE-half activations/weights, tcgen05 kind::f16 with f16 operands: the bf16 kernels and speed with an 11-bit
          significand (tf32-class accuracy: meets the 1e-3 fp32-parity tolerance on the mels); accurate ex2/rcp tanh / sigmoid
  "tf32": fp32 activations/weights, tcgen05 kind::tf32 (fp32-parity mode on tensor cores)
  "fp32": fp32 activations/weights, CUDA-core fp32 kernels only (exact-fp32 parity mode, slow)
"""
import math

import numpy as np
import torch

from . import ops
from .ops import ACT_RELU, ACT_TANH, IMPL_AUTO, IMPL_SIMT, IMPL_TC

MAX_SEQ_LEN = 1000  # hparams.max_seq_len (Models.py:69,120)


def sinusoid_table(n_position, d_hid=256):
    """transformer/Models.py:11-30 (float64 table cast to float32), vectorised."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    angle = pos / np.power(10000.0, 2.0 * (j // 2) / d_hid)[None, :]
    tab = angle.copy()
    tab[:, 0::2] = np.sin(angle[:, 0::2])
    tab[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.from_numpy(tab).float()


N_MEL = 80


def packed_nbytes(B, T):
    return 4 * B * T * N_MEL * 4 + B * 8


def packed_views(B, T, device, buf=None):
    """One contiguous byte buffer holding a forward's results: fp32 [2B,T,80] (mel, mel_noisy) | fp32 [2B,T,80] (postnet,
    postnet_noisy) | int64 [B] mel_len.  Returns ((buf, len_view), mel_view, post_view)."""
    half = 2 * B * T * N_MEL * 4
    if buf is None:
        buf = torch.empty(packed_nbytes(B, T), device=device, dtype=torch.uint8)
    mel = buf[:half].view(torch.float32).view(2 * B, T, N_MEL)
    post = buf[half:2 * half].view(torch.float32).view(2 * B, T, N_MEL)
    lens = buf[2 * half:2 * half + 8 * B].view(torch.int64)
    return (buf, lens), mel, post


def unpack_results(buf, B, T):
    """Inverse of packed_views for a received buffer: (mel, mel_noisy, postnet, postnet_noisy, mel_len)."""
    (_, lens), mel, post = packed_views(B, T, buf.device, buf)
    return mel[:B], mel[B:], post[:B], post[B:], lens


class _NS(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class Engine:
    def __init__(self, state_dict, device, precision="bf16"):
        if precision not in ("bf16", "fp16", "tf32", "fp32"):
            raise ValueError("precision must be bf16 | fp16 | tf32 | fp32")
        self.precision = precision
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("styler_b200.Engine needs a CUDA device: the product path has no CPU implementation")
        self.dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision, torch.float32)
        self.impl = IMPL_SIMT if precision == "fp32" else IMPL_AUTO
        self.attn_impl = IMPL_SIMT if precision == "fp32" else IMPL_TC
        self._pos_cache = {}
        self.inter = {}
        self.last_packed = None
        self.result_mirror = None            # uint8 tensor of packed_nbytes(B, T) in PEER memory: forward() also writes its results there
        self.v_rowmajor = precision in ("bf16", "fp16")   # attention reads V row-major from the fused QKV buffer (MN-major B operand;
                                               # validated for kind::f16 only -- the fp32 modes keep the transposed-V layout)
        import os
        self.use_streams = os.environ.get("STYLER_NO_STREAMS", "0") != "1"   # four audio-encoder branches on side streams
        self._streams = None
        self._aux_used = set()
        self._aux = None                     # three more side streams: the duration / energy / pitch predictors
        self._branch_events = []
        self._post = [None, None, None]
        self._pack({k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()})

    # ------------------------------------------------------------------------------------------ packing
    # Packing runs on the HOST (permutes, folds, casts), each packed tensor then crosses to the device as one memcpy: model
    # load launches no ATen kernels on the GPU (round 1: ~180 copy / cat launches per load).
    def _f32(self, t):
        return t.detach().to("cpu", torch.float32).contiguous().to(self.device)

    def _w(self, t):
        """Conv1d weight [N,Cin,KS] (or Linear [N,Cin]) -> [KS,N,Cin] in the activation dtype."""
        t = t.detach().to("cpu", torch.float32)
        if t.dim() == 2:
            t = t.unsqueeze(-1)
        return t.permute(2, 0, 1).contiguous().to(self.dt).to(self.device)

    def _pack_fft(self, sd, p):
        W = _NS()
        a = p + "slf_attn."
        inv_temp = 1.0 / math.sqrt(64.0)   # temperature sqrt(d_k) (SubLayers.py:24); power of two -> exact fold
        wq = sd[a + "w_qs.weight"].float() * inv_temp
        bq = sd[a + "w_qs.bias"].float() * inv_temp
        W.wqkv = self._w(torch.cat([wq, sd[a + "w_ks.weight"].float(), sd[a + "w_vs.weight"].float()], 0))
        W.bqkv = self._f32(torch.cat([bq, sd[a + "w_ks.bias"].float(), sd[a + "w_vs.bias"].float()], 0))
        W.wfc, W.bfc = self._w(sd[a + "fc.weight"]), self._f32(sd[a + "fc.bias"])
        W.ln1 = (self._f32(sd[a + "layer_norm.weight"]), self._f32(sd[a + "layer_norm.bias"]))
        f = p + "pos_ffn."
        W.w1, W.b1 = self._w(sd[f + "w_1.weight"]), self._f32(sd[f + "w_1.bias"])
        W.w2, W.b2 = self._w(sd[f + "w_2.weight"]), self._f32(sd[f + "w_2.bias"])
        W.pad1 = (sd[f + "w_1.weight"].shape[2] - 1) // 2
        W.ln2 = (self._f32(sd[f + "layer_norm.weight"]), self._f32(sd[f + "layer_norm.bias"]))
        W.c_struct = ops.make_fft_weights(W.wqkv, W.bqkv, W.wfc, W.bfc, W.ln1, W.w1, W.b1, W.w2, W.b2, W.ln2)   # styler_fft_weights
        return W

    def _pack_predictor(self, sd, p):
        W = _NS()
        q = p + "conv_layer."
        W.c1, W.b1 = self._w(sd[q + "conv1d_1.conv.weight"]), self._f32(sd[q + "conv1d_1.conv.bias"])
        W.c2, W.b2 = self._w(sd[q + "conv1d_2.conv.weight"]), self._f32(sd[q + "conv1d_2.conv.bias"])
        W.ln1 = (self._f32(sd[q + "layer_norm_1.weight"]), self._f32(sd[q + "layer_norm_1.bias"]))
        W.ln2 = (self._f32(sd[q + "layer_norm_2.weight"]), self._f32(sd[q + "layer_norm_2.bias"]))
        W.lw = self._f32(sd[p + "linear_layer.weight"].reshape(-1))
        W.lb = float(sd[p + "linear_layer.bias"].reshape(-1)[0].item())
        W.c_struct = ops.make_predictor_weights(W.c1, W.b1, W.ln1, W.c2, W.b2, W.ln2, W.lw, W.lb)   # styler_predictor_weights
        return W

    def _pack(self, sd):
        P, SE = "style_modeling.", "style_modeling.style_encoder."
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}        # one D2H per tensor if the parameters live on the GPU
        w = _NS()
        te = SE + "text_encoder."
        w.emb = self._f32(sd[te + "src_word_emb.weight"])
        w.enc_pos = self._f32(sd[te + "position_enc"][0])
        w.dec_pos = self._f32(sd["decoder.position_enc"][0])
        w.enc_layers = [self._pack_fft(sd, "%slayer_stack.%d." % (te, i)) for i in range(2)]
        w.dec_layers = [self._pack_fft(sd, "decoder.layer_stack.%d." % i) for i in range(4)]
        ae = SE + "audio_encoder."
        w.branches = []
        for n in (1, 2, 3, 4):
            br = _NS(convs=[], onehot=n in (2, 3))
            for j in range(3):
                q = "%sconvolutions_%d.%d." % (ae, n, j)
                cw = sd[q + "0.conv.weight"]
                if j == 0 and br.onehot:   # [C,257,5] -> gather table [5,257,C] fp32
                    cwp = self._f32(cw.permute(2, 1, 0))
                else:
                    cwp = self._w(cw)
                br.convs.append((cwp, self._f32(sd[q + "0.conv.bias"]), self._f32(sd[q + "1.weight"]),
                                 self._f32(sd[q + "1.bias"])))
            L = "%slstm_%d." % (ae, n)
            br.lstm = []
            for layer in range(2):
                k, kr = "l%d" % layer, "l%d_reverse" % layer
                wih = torch.cat([sd[L + "weight_ih_" + k], sd[L + "weight_ih_" + kr]], 0)
                bias = torch.cat([sd[L + "bias_ih_" + k] + sd[L + "bias_hh_" + k],
                                  sd[L + "bias_ih_" + kr] + sd[L + "bias_hh_" + kr]], 0)
                perm = ops.lstm_quad_order(wih.shape[0] // 8).to(wih.device)   # [dir][gate][unit] -> [dir][unit][gate]
                wih, bias = wih[perm], bias[perm]
                whh = torch.stack([sd[L + "weight_hh_" + k], sd[L + "weight_hh_" + kr]], 0)
                br.lstm.append((self._w(wih), self._f32(bias), self._f32(whh)))
            w.branches.append(br)
        w.tld = (self._w(sd[SE + "text_linear_down.0.weight"]), self._f32(sd[SE + "text_linear_down.0.bias"]))
        w.slp = (self._w(sd[SE + "speaker_linear_p.0.weight"]), self._f32(sd[SE + "speaker_linear_p.0.bias"]))
        w.sl = (self._w(sd[SE + "speaker_linear.0.weight"]), self._f32(sd[SE + "speaker_linear.0.bias"]))
        w.cls = {}
        for n in ("d", "p", "e"):
            q = "%saugmentation_classifier_%s.classifier." % (P, n)
            w.cls[n] = (self._w(sd[q + "d_fc1.weight"]), self._f32(sd[q + "d_fc1.bias"]),
                        (self._f32(sd[q + "d_bn1.weight"]), self._f32(sd[q + "d_bn1.bias"])),
                        self._f32(sd[q + "d_fc2.weight"]), self._f32(sd[q + "d_fc2.bias"]))
        w.mlp = {}
        for n in ("duration", "pitch", "energy", "residual", "pitch_norm"):
            q = "%s%s_linear." % (P, n)
            w.mlp[n] = (self._w(sd[q + "0.weight"]), self._f32(sd[q + "0.bias"]), self._w(sd[q + "2.weight"]),
                        self._f32(sd[q + "2.bias"]))
        w.tlu = (self._w(sd[P + "text_linear_up.0.weight"]), self._f32(sd[P + "text_linear_up.0.bias"]))
        # the same packed weights by reference-module name: model.KernelMLP (the nn.Sequential sub-modules synthesize.py
        # calls directly, :116-128,194-196) runs them through the GEMM kernels
        w.seq = {"text_linear_down": [w.tld], "speaker_linear_p": [w.slp], "speaker_linear": [w.sl], "text_linear_up": [w.tlu]}
        for n in w.mlp:
            w.seq[n + "_linear"] = [w.mlp[n][0:2], w.mlp[n][2:4]]
        w.pred = {n: self._pack_predictor(sd, "%s%s_predictor." % (P, n)) for n in ("duration", "pitch", "energy")}
        w.pitch_bins, w.energy_bins = self._f32(sd[P + "pitch_bins"]), self._f32(sd[P + "energy_bins"])
        w.pitch_emb, w.energy_emb = self._f32(sd[P + "pitch_embedding.weight"]), self._f32(sd[P + "energy_embedding.weight"])
        w.mel = (self._w(sd["mel_linear.weight"]), self._f32(sd["mel_linear.bias"]))
        w.postnet = []
        for j in range(5 if "postnet.convolutions.0.0.conv.weight" in sd else 0):   # (absent when use_postnet=False, styler.py:24-26)
            # fold eval-mode BatchNorm1d into the conv (Layers.py:91-119,121-130)
            q = "postnet.convolutions.%d." % j
            cw, cb = sd[q + "0.conv.weight"].float(), sd[q + "0.conv.bias"].float()
            g, be = sd[q + "1.weight"].float(), sd[q + "1.bias"].float()
            rm, rv = sd[q + "1.running_mean"].float(), sd[q + "1.running_var"].float()
            scale = g / torch.sqrt(rv + 1e-5)
            w.postnet.append((self._w(cw * scale[:, None, None]), self._f32((cb - rm) * scale + be)))
        w.dec_struct = ops.make_decoder_weights([W.c_struct for W in w.dec_layers], w.mel[0], w.mel[1], w.postnet)
        self.w = w

    # ------------------------------------------------------------------------------------------ building blocks
    def _pos(self, which, n):
        """Position rows [n,256] fp32: stored table up to max_seq_len, rebuilt (and cached) beyond (Models.py:69-74)."""
        base = self.w.enc_pos if which == "enc" else self.w.dec_pos
        if n <= MAX_SEQ_LEN:
            return base[:n]
        if n not in self._pos_cache:
            self._pos_cache[n] = sinusoid_table(n).to(self.device)
        return self._pos_cache[n]

    def fft_block(self, x, lens, W, out=None):
        """transformer/Layers.py:26-34: MHA -> zero padded rows -> Conv1d FFN -> zero padded rows.  One native call
        (styler_fftblock_fwd) enqueues the five kernels; intermediates live in one workspace allocation."""
        return ops.fftblock(x, W.c_struct, lens, out=out, impl=self.impl)

    def text_encoder(self, src_seq, src_len, out=None):
        """transformer/Models.py:60-84."""
        L = src_seq.shape[1]
        x = ops.embed_pos(src_seq, self.w.emb, self._pos("enc", L), self.dt)
        x = self.fft_block(x, src_len, self.w.enc_layers[0])
        return self.fft_block(x, src_len, self.w.enc_layers[1], out=out)

    def predictor(self, x, lens, W):
        """modules.py:457-465: Conv k3 + ReLU + LN, Conv k3 + ReLU + LN, Linear(256->1), masked_fill -- one native call."""
        return ops.predictor(x, W.c_struct, lens, impl=self.impl)

    def _audio_branch(self, br, xin, mel_len, src_len, L):
        x = None
        c = None
        last = len(br.convs) - 1
        for j, (cw, cb, g, be) in enumerate(br.convs):
            if j == 0 and br.onehot:
                x = ops.onehot_conv(xin, cw, cb, self.dt)
                ops.groupnorm_relu_(x, g, be, 16, 1e-5)
                continue
            src = xin if j == 0 else x
            B, Tr, N = src.shape[0], src.shape[1], cw.shape[1]
            if self.impl != IMPL_SIMT and B * Tr >= 64:
                # tensor-core conv: the GroupNorm statistics come out of its epilogue (no separate pass over the tensor)
                part = torch.empty(B, (Tr + 127) // 128, N // 16, 2, device=src.device, dtype=torch.float32)
                x = ops.conv1d(src, cw, cb, pad=2, impl=IMPL_TC, gn_partial=part)
                if j == last:      # the last GroupNorm + ReLU is applied by the Mel Calibrator as it reads the frames
                    c = ops.gn_calibrator(x, g, be, part, mel_len, src_len, L, 1e-5)
                else:
                    ops.groupnorm_relu_partial_(x, g, be, part, 1e-5)
            else:
                x = ops.conv1d(src, cw, cb, pad=2, impl=self.impl)
                ops.groupnorm_relu_(x, g, be, 16, 1e-5)
        if c is None:
            c = ops.mel_calibrator(x, mel_len, src_len, L)
        for (wih, bias, whh) in br.lstm:
            B = c.shape[0]
            gx = torch.empty(B, L, wih.shape[1], device=c.device, dtype=torch.float32)
            ops.conv1d(c, wih, bias, out_f32=gx, want_out=False, impl=self.impl)
            c = ops.bilstm_layer(gx, whh, self.dt)
        return c

    def audio_encoder(self, mel_target, p_idx, e_idx, mel_aug, mel_len, src_len, L, join=True, classify=False, tails=None):
        """modules.py:164-201 on the padded grid: conv stacks + GroupNorm + ReLU @Tr, Mel Calibrator, 2-layer BiLSTMs @L.
        The four style-factor branches are independent: each runs on its own CUDA stream so the latency-bound BiLSTM
        recurrences overlap the other branches' convolutions instead of serialising.
        classify: also run the three augmentation classifiers (modules.py:319-321) at the tail of their branch's stream.  Their
        posteriors are forward OUTPUTS that nothing downstream consumes, so the main stream only waits for the branch encodings
        (`join_branches`, per-branch events) and picks the posteriors up at the very end of the forward (`join_audio_streams`).
        tails: optional per-branch callables c -> tensor(s), run on the branch's stream right behind its encoding (the per-factor
        up-projection MLPs of modules.py:334-347: four independent chains instead of eight serial launches on the main stream);
        their results are returned as a second list and are covered by the branch events."""
        ins = (mel_target, p_idx, e_idx, mel_aug)
        names = ("d", "p", "e", None)
        self._post = [None, None, None]
        prep = lambda x: x() if callable(x) else x      # an input may be a thunk: its cast / quantisation then runs on ITS branch's stream
        if not self.use_streams:
            outs = [self._audio_branch(br, prep(xin), mel_len, src_len, L) for br, xin in zip(self.w.branches, ins)]
            ups = [t(c) for t, c in zip(tails, outs)] if tails is not None else None
            if classify:
                self._post = [self._classifier(outs[i], self.w.cls[names[i]]) for i in range(3)]
            return outs if tails is None else (outs, ups)
        main = torch.cuda.current_stream(self.device)
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(4)]
        start = torch.cuda.Event()
        start.record(main)
        outs, ups, self._branch_events = [], [], []
        for i, (st, br, xin) in enumerate(zip(self._streams, self.w.branches, ins)):
            st.wait_event(start)
            with torch.cuda.stream(st):
                c = self._audio_branch(br, prep(xin), mel_len, src_len, L)
                if tails is not None:
                    up = tails[i](c)
                    for t_ in (up if isinstance(up, (tuple, list)) else (up,)):
                        t_.record_stream(main)
                    ups.append(up)
                ev = torch.cuda.Event()
                ev.record(st)
                if classify and names[i] is not None:
                    self._post[i] = self._classifier(c, self.w.cls[names[i]])
                    self._post[i].record_stream(main)
            c.record_stream(main)
            outs.append(c)
            self._branch_events.append(ev)
        if join:
            self.join_audio_streams()
        return outs if tails is None else (outs, ups)

    def join_branches(self):
        """Main stream waits for the four branch ENCODINGS only (not for the classifiers queued behind them)."""
        if self.use_streams and self._streams is not None:
            main = torch.cuda.current_stream(self.device)
            for ev in self._branch_events:
                main.wait_event(ev)

    def _aux_stream(self, i):
        if self._aux is None:
            self._aux = [torch.cuda.Stream(device=self.device) for _ in range(3)]
        self._aux_used.add(i)                # only streams forked in THIS forward are joined (a CUDA-graph capture must not wait
        return self._aux[i]                  # for a stream that is not part of it)

    def join_audio_streams(self):
        if self.use_streams and self._aux is not None:
            main = torch.cuda.current_stream(self.device)
            for i in sorted(self._aux_used):
                main.wait_stream(self._aux[i])
            self._aux_used.clear()
        if self.use_streams and self._streams is not None:
            main = torch.cuda.current_stream(self.device)
            for st in self._streams:
                main.wait_stream(st)

    def _mlp2(self, x, W, out=None):
        h = ops.conv1d(x, W[0], W[1], act=ACT_RELU, impl=self.impl)
        return ops.conv1d(h, W[2], W[3], act=ACT_RELU, out=out, impl=self.impl)

    def _classifier(self, x, W):
        h = ops.conv1d(x, W[0], W[1], ln=W[2], act2=ACT_RELU, impl=self.impl)
        return ops.classifier_tail(h, W[3], W[4])

    def _act(self, t):
        """fp32 user tensor -> activation dtype (no-op view in the fp32 modes)."""
        t = t.to(self.device, torch.float32).contiguous()
        return t if self.dt == torch.float32 else ops.cast(t, self.dt)

    # ------------------------------------------------------------------------------------------ decode (styler.py:29-37)
    def decode(self, x, mel_lens, mel_out=None, post_out=None, mel_mirror=None, post_mirror=None, has_pos=False):
        """x [B,T,256] (activation dtype) -> (mel fp32 [B,T,80], mel_postnet fp32 [B,T,80]); the results are written into
        mel_out / post_out when given (slices of the packed result buffer).  mel_mirror / post_mirror: second, WRITE-ONLY
        destinations -- this rank's slice of rank 0's gather buffer mapped over NVLink (dist.PeerGather): mel_linear and the last
        PostNet convolution store their fp32 results there from their own epilogues (`out2_f32`), so the gather is fused into
        the tensor-core kernels that produce the mels."""
        T = x.shape[1]         # has_pos: x already carries the decoder's position rows (variance_adapt adds them in its embedding sum)
        mel, post = ops.decoder(x.contiguous(), self.w.dec_struct, None if has_pos else self._pos("dec", T), mel_lens, mel_out=mel_out, post_out=post_out,
                                mel_out2=mel_mirror, post_out2=post_mirror, impl=self.impl)
        if post is None:                             # use_postnet=False (styler.py:33-36): mel_output_postnet = mel_output
            if post_mirror is not None:
                post_mirror.copy_(mel)
            return mel, mel
        return mel, post

    # ------------------------------------------------------------------------------------------ style modeling
    def encode(self, src_seq, speaker_embed, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, dur_async=False):
        """StyleEncoder.forward + the L-level part of StyleModeling.forward (modules.py:225-235,311-353).
        Returns the [B,L,1280] concatenated encodings, log-duration prediction and the DAT posteriors.
        dur_async: the caller does not consume log_d inside the forward (teacher-forced durations): the duration predictor then
        runs on a side stream that is joined at the end of the forward."""
        w = self.w
        B, L = src_seq.shape
        dev = self.device
        enc = torch.empty(B, L, 1280, device=dev, dtype=self.dt)
        # the audio-encoder branches (side streams) overlap the text encoder and speaker projections (this stream)
        # each branch's own input preparation (fp32 -> activation dtype cast of a mel, utils.py:417-429 quantisation of p / e) is the
        # first thing on that branch's stream: nothing the branches need runs on this stream first
        p_idx, e_idx = (lambda: ops.quantize_index(p_norm)), (lambda: ops.quantize_index(e_input))
        mel_t, mel_a = (lambda: self._act(mel_target)), (lambda: self._act(mel_aug))
        spk_in = self._act(speaker_embed).unsqueeze(0)                                                  # [1,B,512]

        def pitch_tail(p_enc):                         # modules.py:332,338-340 on the pitch branch's own stream
            spk_p = ops.conv1d(spk_in, w.slp[0], w.slp[1], act=ACT_RELU, impl=self.impl)[0]             # [B,128]
            return self._mlp2(ops.add(p_enc, rowvec=spk_p), w.mlp["pitch"]), spk_p

        tails = (lambda c: self._mlp2(c, w.mlp["duration"]), pitch_tail, lambda c: self._mlp2(c, w.mlp["energy"]),
                 lambda c: self._mlp2(c, w.mlp["residual"], out=enc[..., 1024:1280]))
        (d_enc, p_enc, e_enc, n_enc), (d_up, (p_up, spk_p), e_up, n_up) = self.audio_encoder(
            mel_t, p_idx, e_idx, mel_a, mel_len, src_len, L, join=False, classify=True, tails=tails)
        text = self.text_encoder(src_seq, src_len, out=enc[..., 0:256])
        neck = ops.conv1d(text, w.tld[0], w.tld[1], act=ACT_RELU, impl=self.impl)                      # [B,L,4]
        spk = ops.conv1d(spk_in, w.sl[0], w.sl[1], act=ACT_RELU, impl=self.impl)[0]                     # [B,256]
        neck_up = ops.conv1d(neck, w.tlu[0], w.tlu[1], act=ACT_RELU, impl=self.impl)                    # [B,L,256]
        ops.add(None, rowvec=spk, out=enc[..., 512:768])
        self.join_branches()
        post = tuple(self._post)                       # produced on the side streams; valid after join_audio_streams() (end of forward)
        ops.add(neck_up, p_up, out=enc[..., 256:512])                                                   # modules.py:350
        ops.add(neck_up, e_up, out=enc[..., 768:1024])
        if dur_async and self.use_streams:
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            a = self._aux_stream(0)
            a.wait_event(fork)
            with torch.cuda.stream(a):
                dur_in = ops.add(neck_up, d_up)
                log_d = self.predictor(dur_in, src_len, w.pred["duration"])                             # modules.py:353
            for t_ in (neck_up, d_up):
                t_.record_stream(a)
            log_d.record_stream(main)
        else:
            dur_in = ops.add(neck_up, d_up)
            log_d = self.predictor(dur_in, src_len, w.pred["duration"])                                 # modules.py:353
        self.inter = dict(text_encoding=text, text_encoding_neck=neck_up, pitch_encoding=p_enc, speaker_encoding=spk,
                          speaker_encoding_p=spk_p, duration_encoding=d_up, energy_encoding=e_up, noise_encoding=n_up,
                          pitch_up=p_up, max_seq_len=L)
        return enc, log_d, post

    def variance_adapt(self, enc, log_d, T, mel_lens_for_mask, d_target=None, p_target=None, e_target=None,
                       d_control=1.0, p_control=1.0, e_control=1.0, duration=None):
        """LengthRegulator + pitch/energy predictors + bucketize/embedding sum (modules.py:355-385).
        `duration` (already rounded, fp32) or d_target (int64) drives the expand to T frames."""
        w = self.w
        dur = d_target if d_target is not None else duration
        encT, mel_len, cum = ops.length_regulator(enc, dur, T)
        lens = mel_lens_for_mask if mel_lens_for_mask is not None else mel_len
        # The two predictors are independent.  Teacher-forced (p_target and e_target given: evaluate.py / train.py) their outputs
        # are only RETURNED -- the embedding sum below reads the targets -- so both run on side streams under the first decoder
        # kernels and are joined at the end of the forward; free-running, the energy predictor overlaps the pitch predictor.
        teacher = p_target is not None and e_target is not None
        if self.use_streams:
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            a0 = self._aux_stream(1)
            a0.wait_event(fork)
            with torch.cuda.stream(a0):
                e_pred = self.predictor(encT[..., 768:1024], lens, w.pred["energy"])
                e_done = torch.cuda.Event()
                e_done.record(a0)
            encT.record_stream(a0)
            e_pred.record_stream(main)
            a1 = self._aux_stream(2) if teacher else main
            if teacher:
                a1.wait_event(fork)
            with torch.cuda.stream(a1):
                p_in = ops.add(encT[..., 256:512], encT[..., 512:768])
                p_pred = self.predictor(p_in, lens, w.pred["pitch"])
            if teacher:
                encT.record_stream(a1)
                p_pred.record_stream(main)
            else:
                main.wait_event(e_done)
        else:
            e_pred = self.predictor(encT[..., 768:1024], lens, w.pred["energy"])
            p_in = ops.add(encT[..., 256:512], encT[..., 512:768])
            p_pred = self.predictor(p_in, lens, w.pred["pitch"])
        p_val, p_scale = (p_target.to(self.device, torch.float32).contiguous(), 1.0) if p_target is not None else (p_pred, float(p_control))
        e_val, e_scale = (e_target.to(self.device, torch.float32).contiguous(), 1.0) if e_target is not None else (e_pred, float(e_control))
        # clean and noisy decoder inputs land in one [2B,T,256] buffer so both decodes run as a single batched pass
        B = encT.shape[0]
        xx = torch.empty(2 * B, T, 256, device=encT.device, dtype=self.dt)
        scaled = p_scale != 1.0 or e_scale != 1.0      # the reference returns prediction * control (modules.py:370,380)
        res = ops.bucket_embed_sum(encT[..., 0:256], encT[..., 512:768], encT[..., 1024:1280], p_val, e_val,
                                   p_scale, e_scale, w.pitch_bins, w.energy_bins, w.pitch_emb, w.energy_emb,
                                   want_noisy=True, out=xx[:B], out_noisy=xx[B:], want_scaled=scaled, pos=self._pos("dec", T))
        x, x_noisy = res[0], res[1]
        if scaled:
            p_pred = res[4] if p_target is None else p_pred
            e_pred = res[5] if e_target is None else e_pred
        self._xx = xx
        return x, x_noisy, encT, p_pred, e_pred, mel_len

    def forward(self, src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target=None, p_target=None,
                e_target=None, max_src_len=None, max_mel_len=None, speaker_embed=None, d_control=1.0, p_control=1.0,
                e_control=1.0):
        """styler.py:39-58.  Returns the reference's 9-tuple (mels/predictions fp32, masks bool, lengths int64)."""
        ctx = self.forward_front(src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target, p_target, e_target,
                                 max_src_len, max_mel_len, speaker_embed, d_control, p_control, e_control, join=False)
        return self.forward_back(ctx)

    def forward_front(self, src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target=None, p_target=None,
                      e_target=None, max_src_len=None, max_mel_len=None, speaker_embed=None, d_control=1.0, p_control=1.0,
                      e_control=1.0, join=True):
        """First stage of the forward: style encoders + variance adaptor (styler.py:41-49) up to the decoder input.  Returns a
        context for `forward_back`.  The two stages are separate so that a serving loop can run stage one of batch i+1 under
        stage two of batch i (model.PipelinedSTYLER): stage one is latency-bound (BiLSTM recurrences, many small kernels) and
        leaves most SMs idle, stage two is four fifths of the FLOPs.
        join: wait here for the side streams (the DAT posteriors are then complete when this stage is)."""
        dev = self.device
        src_seq = src_seq.to(dev).contiguous()
        src_len = src_len.to(dev, torch.int64).contiguous()
        mel_len = mel_len.to(dev, torch.int64).contiguous()
        p_norm = p_norm.to(dev, torch.float32).contiguous()
        e_input = e_input.to(dev, torch.float32).contiguous()
        B, L = src_seq.shape
        if max_src_len is not None and max_src_len != L:
            raise ValueError("max_src_len (%d) must equal the padded text length (%d)" % (max_src_len, L))
        enc, log_d, post = self.encode(src_seq, speaker_embed, mel_target, mel_aug, p_norm, e_input, src_len, mel_len,
                                       dur_async=d_target is not None)
        if d_target is not None:                                   # teacher forcing (styler.py:44-46)
            d_target = d_target.to(dev, torch.int64).contiguous()
            T = int(max_mel_len) if max_mel_len else int(mel_len.max().item())
            x, x_noisy, _, p_pred, e_pred, _ = self.variance_adapt(enc, log_d, T, mel_len, d_target, p_target, e_target,
                                                                   d_control, p_control, e_control)
            out_len = mel_len
        else:                                                      # free running (styler.py:47-49, modules.py:357-360)
            duration = ops.duration_round(log_d, 1.0, float(d_control))
            tot, _ = ops.length_regulator_scan(duration)
            T = int(max_mel_len) if max_mel_len else int(tot.max().item())     # the one host sync of the path
            x, x_noisy, _, p_pred, e_pred, out_len = self.variance_adapt(enc, log_d, T, None, None, p_target, e_target,
                                                                         d_control, p_control, e_control, duration=duration)
        if join:
            self.join_audio_streams()
        ctx = _NS(xx=self._xx, out_len=out_len, B=B, L=L, T=T, src_len=src_len, log_d=log_d, p_pred=p_pred, e_pred=e_pred,
                  post=post, joined=join)
        self._xx = None
        return ctx

    def forward_back(self, ctx):
        """Second stage: clean + noisy decode, mel_linear, PostNet (styler.py:52-57) and the packed result buffer."""
        dev = self.device
        B, L, T, out_len = ctx.B, ctx.L, ctx.T, ctx.out_len
        # styler.py:52,55: clean decode and noisy decode (x.detach() + noise_encoding) -- batched as one [2B] pass.
        # The four mel tensors and the lengths land in ONE contiguous buffer [mel | mel_noisy | postnet | postnet_noisy | len]
        # so that the data-parallel gather to rank 0 (dist.AsyncGather.launch_packed) is a single NCCL operation.
        packed, mel_out, post_out = (None, None, None) if not self.w.postnet else packed_views(B, T, dev)
        mirror, mel_m, post_m = None, None, None
        if self.result_mirror is not None:             # this rank's slice of rank 0's receive region (dist.PeerGather)
            if packed is None or self.result_mirror.numel() != packed_nbytes(B, T):
                raise ValueError("result_mirror must hold exactly %d bytes for B=%d, T=%d (and needs the PostNet)" % (packed_nbytes(B, T), B, T))
            mirror, mel_m, post_m = packed_views(B, T, dev, self.result_mirror)
        mel2, post2 = self.decode(ctx.xx, out_len.repeat(2), mel_out, post_out, mel_m, post_m, has_pos=True)
        post = ctx.post
        if not ctx.joined:
            self.join_audio_streams()                  # the DAT posteriors (side streams) are outputs of this forward
        if mirror is not None:
            mirror[1].copy_(out_len)
        if packed is not None:
            packed[1].copy_(out_len)
        self.last_packed = packed[0] if packed is not None else None
        mel, mel_n, post_mel, post_mel_n = mel2[:B], mel2[B:], post2[:B], post2[B:]
        ar = torch.arange(L, device=dev)
        src_mask = ar.unsqueeze(0) >= ctx.src_len.unsqueeze(1)
        mel_mask = torch.arange(T, device=dev).unsqueeze(0) >= out_len.unsqueeze(1)
        return (mel, mel_n), (post_mel, post_mel_n), ctx.log_d, ctx.p_pred, ctx.e_pred, src_mask, mel_mask, out_len, post
