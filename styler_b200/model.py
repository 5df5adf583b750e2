"""Drop-in `STYLER` nn.Module: the reference's constructor, forward()/decode() signatures, submodule tree and
328-key state_dict (styler.py:13-58, modules.py, transformer/*.py), with the eval-mode forward executed by
hand-written sm_100a kernels through the C ABI (styler_b200.engine.Engine).

The torch.nn modules below are PARAMETER CONTAINERS only (they give `load_state_dict` / `state_dict` / `.to()` the
reference's exact key surface, Appendix D of SURVEY.md); none of their torch forward code runs on the product path.
Training (backward, dropout, batch-stat BatchNorm) is out of scope: calling forward in training mode raises.
"""
import weakref
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import hparams as hp
from .engine import Engine, sinusoid_table

N_SRC_VOCAB = 152  # len(text.symbols) + 1 (transformer/Models.py:37)


# --------------------------------------------------------------------------------------------- containers
class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("%s is a parameter container; the computation runs in styler_b200.engine" % type(self).__name__)


class MultiHeadAttention(_Container):      # transformer/SubLayers.py:9-29
    def __init__(self, n_head, d_model, d_k, d_v):
        super().__init__()
        self.w_qs, self.w_ks, self.w_vs = nn.Linear(d_model, n_head * d_k), nn.Linear(d_model, n_head * d_k), nn.Linear(d_model, n_head * d_v)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)


class PositionwiseFeedForward(_Container):  # transformer/SubLayers.py:64-79
    def __init__(self, d_in, d_hid):
        super().__init__()
        k = hp.fft_conv1d_kernel_size
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size=k[0], padding=(k[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size=k[1], padding=(k[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)


class FFTBlock(_Container):                 # transformer/Layers.py:10-24
    def __init__(self, d_model, d_inner, n_head, d_k, d_v):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner)


def _fft_stack(n_layers, d_model, n_head):
    return nn.ModuleList([FFTBlock(d_model, hp.fft_conv1d_filter_size, n_head, d_model // n_head, d_model // n_head)
                          for _ in range(n_layers)])


class Encoder(_Container):                  # transformer/Models.py:33-58
    def __init__(self):
        super().__init__()
        self.src_word_emb = nn.Embedding(N_SRC_VOCAB, hp.encoder_hidden, padding_idx=0)
        self.position_enc = nn.Parameter(sinusoid_table(hp.max_seq_len + 1, hp.encoder_hidden).unsqueeze(0), requires_grad=False)
        self.layer_stack = _fft_stack(hp.encoder_layer, hp.encoder_hidden, hp.encoder_head)


class Decoder(_Container):                  # transformer/Models.py:87-109
    def __init__(self):
        super().__init__()
        self.position_enc = nn.Parameter(sinusoid_table(hp.max_seq_len + 1, hp.decoder_hidden).unsqueeze(0), requires_grad=False)
        self.layer_stack = _fft_stack(hp.decoder_layer, hp.decoder_hidden, hp.decoder_head)


class ConvNorm(_Container):                 # transformer/Layers.py:37-64
    def __init__(self, cin, cout, kernel_size):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=kernel_size, padding=(kernel_size - 1) // 2)


class PostNet(_Container):                  # transformer/Layers.py:67-119
    def __init__(self, n_mel=80, dim=512, ks=5, n=5):
        super().__init__()
        chans = [n_mel] + [dim] * (n - 1) + [n_mel]
        self.convolutions = nn.ModuleList([nn.Sequential(ConvNorm(chans[i], chans[i + 1], ks), nn.BatchNorm1d(chans[i + 1]))
                                           for i in range(n)])


class KernelMLP(nn.Sequential):
    """`nn.Sequential(Linear, ReLU[, Linear, ReLU])` of the reference (modules.py:207-216,250-265) with the SAME state_dict
    keys (`0.weight`, `2.weight`, ...) whose forward runs the library's GEMM kernels (Linear + bias + ReLU fused) instead of
    torch's: synthesize.py calls these modules directly when it recombines encodings (`self_.pitch_linear(p_down + s_down)`
    :118-119,196; `style_encoder.speaker_linear_p / speaker_linear(speaker_embed)` :194-195).  Takes [.., Cin] fp32 or
    activation-dtype tensors, returns fp32 like the reference."""

    def __init__(self, owner, key, dims):
        mods = []
        for i, o in zip(dims[:-1], dims[1:]):
            mods += [nn.Linear(i, o), nn.ReLU()]
        super().__init__(*mods)
        object.__setattr__(self, "_owner", owner)
        self._key = key

    def forward(self, x):
        from . import ops
        eng = self._owner._engine_for(x)
        layers = eng.w.seq[self._key]
        lead = x.shape[:-1]
        h = x.to(eng.device).reshape(1, -1, x.shape[-1]) if x.dim() != 3 else x.to(eng.device)
        h = eng._act(h) if h.dtype != eng.dt else h.contiguous()
        with torch.cuda.device(eng.device):
            for j, (w, b) in enumerate(layers):
                if j + 1 < len(layers):
                    h = ops.conv1d(h, w, b, act=ops.ACT_RELU, impl=eng.impl)
                else:
                    out = torch.empty(h.shape[0], h.shape[1], w.shape[1], device=eng.device, dtype=torch.float32)
                    if eng.dt == torch.float32:
                        ops.conv1d(h, w, b, act=ops.ACT_RELU, out=out, impl=eng.impl)
                    else:
                        ops.conv1d(h, w, b, act=ops.ACT_RELU, out_f32=out, want_out=False, impl=eng.impl)
        return out.reshape(*lead, out.shape[-1])


class EncoderInput(tuple):
    """What `StyleEncoder.encoder_input_cat` returns here: (mel_target, p_index, e_index, mel_aug) instead of the
    reference's dense [B,674,Tr] one-hot tensor (modules.py:218-223); `AudioEncoder.forward` consumes it."""


class AudioEncoder(_Container):             # modules.py:84-162
    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        dims = ((hp.n_mel_channels, hp.va_enc_dim_d, hp.va_neck_hidden_d), (hp.va_dim_f0, hp.va_enc_dim_p, hp.va_neck_hidden_p),
                (hp.va_dim_energy, hp.va_enc_dim_e, hp.va_neck_hidden_e), (hp.n_mel_channels, hp.va_enc_dim_r, hp.va_neck_hidden_r))
        for n, (cin, c, h) in enumerate(dims, start=1):
            setattr(self, "convolutions_%d" % n, nn.ModuleList(
                [nn.Sequential(ConvNorm(cin if j == 0 else c, c, 5), nn.GroupNorm(c // hp.va_chs_grp, c)) for j in range(3)]))
            setattr(self, "lstm_%d" % n, nn.LSTM(c, h, 2, batch_first=True, bidirectional=True))

    def forward(self, cat, len_org, seq_len, mask=None):
        """modules.py:164-201; `cat` is the EncoderInput from StyleEncoder.encoder_input_cat."""
        eng = self._owner._engine_for(cat[0])
        mel_t, p_idx, e_idx, mel_a = cat
        L = int(seq_len.max().item())
        return tuple(eng.audio_encoder(eng._act(mel_t), p_idx, e_idx, eng._act(mel_a), len_org.to(eng.device), seq_len.to(eng.device), L))


class StyleEncoder(_Container):             # modules.py:204-216
    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        self.text_encoder = Encoder()
        self.audio_encoder = AudioEncoder(owner)
        self.text_linear_down = KernelMLP(owner, "text_linear_down", (hp.encoder_hidden, hp.va_neck_hidden_t))
        self.speaker_linear_p = KernelMLP(owner, "speaker_linear_p", (hp.speaker_embed_dim, hp.va_neck_hidden_p * 2))
        self.speaker_linear = KernelMLP(owner, "speaker_linear", (hp.speaker_embed_dim, hp.encoder_hidden))

    def encoder_input_cat(self, mel_target, p_norm, e_input, mel_aug):
        from . import ops
        eng = self._owner._engine_for(mel_target)
        return EncoderInput((mel_target, ops.quantize_index(p_norm.to(eng.device, torch.float32)),
                             ops.quantize_index(e_input.to(eng.device, torch.float32)), mel_aug))


class AugmentationClassifier(_Container):   # modules.py:23-36
    def __init__(self, input_dim):
        super().__init__()
        self.classifier = nn.Sequential(OrderedDict([
            ("d_fc1", nn.Linear(input_dim, hp.encoder_hidden)), ("d_bn1", nn.LayerNorm(hp.encoder_hidden)),
            ("d_relu1", nn.ReLU()), ("d_fc2", nn.Linear(hp.encoder_hidden, 2)), ("d_softmax", nn.LogSoftmax(dim=-1))]))


class Conv(_Container):                     # modules.py:468-500
    def __init__(self, cin, cout, kernel_size, padding):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=kernel_size, padding=padding)


class StylePredictor(_Container):           # modules.py:426-455
    def __init__(self):
        super().__init__()
        f, k = hp.style_predictor_filter_size, hp.style_predictor_kernel_size
        self.conv_layer = nn.Sequential(OrderedDict([
            ("conv1d_1", Conv(hp.encoder_hidden, f, k, (k - 1) // 2)), ("relu_1", nn.ReLU()), ("layer_norm_1", nn.LayerNorm(f)),
            ("dropout_1", nn.Dropout(hp.style_predictor_dropout)),
            ("conv1d_2", Conv(f, f, k, 1)), ("relu_2", nn.ReLU()), ("layer_norm_2", nn.LayerNorm(f)),
            ("dropout_2", nn.Dropout(hp.style_predictor_dropout))]))
        self.linear_layer = nn.Linear(f, 1)


class LengthRegulator(_Container):          # modules.py:390-423
    def forward(self, x, duration, max_len):
        from . import ops
        T = int(max_len) if max_len else int(ops.length_regulator_scan(duration.contiguous())[0].max().item())
        out, mel_len, _ = ops.length_regulator(x, duration, T)
        return out, mel_len


class _Inspection:
    """Inspection tensor left on StyleModeling by forward() (modules.py:328-331,342-348).  The engine keeps it in the
    activation dtype; callers (synthesize.py:114-144,180-205) feed it to fp32 torch sub-modules, so it is exposed as fp32,
    converted lazily on first access (nothing is converted on the forward hot path).  Assignable like a plain attribute."""

    def __init__(self, name):
        self.slot = "_insp_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        v = obj.__dict__.get(self.slot)
        if torch.is_tensor(v) and v.dtype != torch.float32:
            v = v.float()
            obj.__dict__[self.slot] = v
        return v

    def __set__(self, obj, value):
        obj.__dict__[self.slot] = value


class StyleModeling(_Container):            # modules.py:238-283
    pitch_encoding = _Inspection("pitch_encoding")
    speaker_encoding = _Inspection("speaker_encoding")
    speaker_encoding_p = _Inspection("speaker_encoding_p")
    text_encoding_neck = _Inspection("text_encoding_neck")
    duration_encoding = _Inspection("duration_encoding")
    energy_encoding = _Inspection("energy_encoding")
    noise_encoding = _Inspection("noise_encoding")
    text_encoding = _Inspection("text_encoding")

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        H = hp.encoder_hidden
        self.style_encoder = StyleEncoder(owner)
        self.augmentation_classifier_d = AugmentationClassifier(hp.va_neck_hidden_d * 2)
        self.augmentation_classifier_p = AugmentationClassifier(hp.va_neck_hidden_p * 2)
        self.augmentation_classifier_e = AugmentationClassifier(hp.va_neck_hidden_e * 2)
        self.duration_linear = KernelMLP(owner, "duration_linear", (hp.va_neck_hidden_d * 2, H, H))
        self.pitch_norm_linear = KernelMLP(owner, "pitch_norm_linear", (hp.va_neck_hidden_p * 2, H, H))   # allocated, saved, never
        self.pitch_linear = KernelMLP(owner, "pitch_linear", (hp.va_neck_hidden_p * 2, H, H))             # used by forward (modules.py:254-257)
        self.energy_linear = KernelMLP(owner, "energy_linear", (hp.va_neck_hidden_e * 2, H, H))
        self.residual_linear = KernelMLP(owner, "residual_linear", (hp.va_neck_hidden_r * 2, H, H))
        self.text_linear_up = KernelMLP(owner, "text_linear_up", (hp.va_neck_hidden_t, H))
        self.duration_predictor = StylePredictor()
        self.length_regulator = LengthRegulator()
        self.pitch_predictor = StylePredictor()
        self.energy_predictor = StylePredictor()
        self.pitch_bins = nn.Parameter(torch.exp(torch.linspace(np.log(hp.f0_min), np.log(hp.f0_max), hp.n_bins - 1)), requires_grad=False)
        self.energy_bins = nn.Parameter(torch.linspace(hp.energy_min, hp.energy_max, hp.n_bins - 1), requires_grad=False)
        self.pitch_embedding = nn.Embedding(hp.n_bins, H)
        self.energy_embedding = nn.Embedding(hp.n_bins, H)

    def _store_inspection(self, eng, src_mask, max_len):
        """modules.py:328-331,342-348: tensors the reference leaves on `self` for synthesize.py's inspection mode."""
        it = eng.inter
        L = it["max_seq_len"]
        self.max_seq_len = L
        self.pitch_encoding = it["pitch_encoding"]
        self.speaker_encoding = it["speaker_encoding"].unsqueeze(1).expand(-1, L, -1)
        self.speaker_encoding_p = it["speaker_encoding_p"].unsqueeze(1).expand(-1, L, -1)
        self.text_encoding_neck = it["text_encoding_neck"]
        self.duration_encoding = it["duration_encoding"]
        self.energy_encoding = it["energy_encoding"]
        self.noise_encoding = it["noise_encoding"]
        self.text_encoding = it["text_encoding"]
        self.src_mask = src_mask
        self.max_len = max_len

    def predict_inference(self, text_encoding, pitch_encoding, energy_encoding, duration_encoding, speaker_encoding,
                          noise_encoding, src_mask, max_len, speaker_normalized=True, d_control=1.0, p_control=1.0,
                          e_control=1.0):
        """modules.py:285-309 (used by synthesize.py:171): recombine stored encodings, predict, expand, embed."""
        from . import ops
        eng = self._owner._engine_for(text_encoding)
        dt = eng.dt
        with torch.cuda.device(eng.device):
            parts = [t.to(eng.device, dt) for t in (text_encoding, pitch_encoding, speaker_encoding, energy_encoding, noise_encoding)]
            enc = torch.cat([p.expand(parts[0].shape[0], parts[0].shape[1], -1) for p in parts], dim=-1).contiguous()
            src_len = (~src_mask).sum(1).to(eng.device, torch.int64)
            log_d = eng.predictor(duration_encoding.to(eng.device, dt).contiguous(), src_len, eng.w.pred["duration"])
            duration = ops.duration_round(log_d, hp.log_offset, float(d_control))
            tot, _ = ops.length_regulator_scan(duration)
            T = int(max_len) if max_len else int(tot.max().item())
            encT, mel_len, _ = ops.length_regulator(enc, duration, T)
            e_pred = eng.predictor(encT[..., 768:1024], mel_len, eng.w.pred["energy"])
            p_in = encT[..., 256:512] if speaker_normalized else ops.add(encT[..., 256:512], encT[..., 512:768])
            p_pred = eng.predictor(p_in, mel_len, eng.w.pred["pitch"])
            # one launch: predictions * control (modules.py:299,305), bucketize, and the two embedding rows as separate outputs
            w = eng.w
            _, _, _, _, p_scaled, e_scaled, p_emb, e_emb = ops.bucket_embed_sum(
                None, None, None, p_pred, e_pred, float(p_control), float(e_control), w.pitch_bins, w.energy_bins, w.pitch_emb,
                w.energy_emb, want_noisy=False, want_scaled=True, want_emb=True, want_sum=False, emb_dtype=dt)
            mel_mask = torch.arange(T, device=eng.device).unsqueeze(0) >= mel_len.unsqueeze(1)
        return (encT[..., 0:256], p_emb, encT[..., 512:768], e_emb, encT[..., 1024:1280], log_d, p_scaled, e_scaled, mel_mask)


# --------------------------------------------------------------------------------------------- the model
def _invalidate_after_load(module, incompatible_keys):
    module._invalidate()


class STYLER(nn.Module):
    """Drop-in for the reference `styler.STYLER` (styler.py:13-58), eval-mode forward on B200 kernels.

    Extra constructor argument `precision` ("bf16" | "tf32" | "fp32", see engine.py) selects the compute mode; the
    default follows BASELINE.json's full-forward configuration (bf16).
    """

    def __init__(self, use_postnet=True, precision="bf16"):
        super().__init__()
        self.precision = precision
        self.style_modeling = StyleModeling(self)
        self.decoder = Decoder()
        self.mel_linear = nn.Linear(hp.decoder_hidden, hp.n_mel_channels)
        self.use_postnet = use_postnet
        if self.use_postnet:                      # styler.py:24-26: without it decode() returns the mel twice
            self.postnet = PostNet()
        # One packed Engine per device.  nn.DataParallel (train.py:33, synthesize.py:62) makes shallow replicas whose
        # parameters are plain per-device tensors (their `_parameters` is empty, so `replica.state_dict()` has no weights):
        # the dict and the weak reference below are shared through the copied `__dict__`, so replica i finds (or builds once)
        # the engine of ITS device, packed from the master's state_dict -- the values the replicas were broadcast.
        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_master", weakref.ref(self))
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    # -- engine lifecycle ------------------------------------------------------------------------------------
    def __getstate__(self):                       # copy.deepcopy / torch.save(model): engines and the weak self-reference
        st = self.__dict__.copy()                 # are per-instance runtime state, rebuilt lazily
        st.pop("_engines", None)
        st.pop("_master", None)
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_master", weakref.ref(self))

    def _invalidate(self):
        self._engines.clear()

    @property
    def _engine(self):
        """The engine of the device the parameters live on (None before the first forward)."""
        return self._engines.get((self.mel_linear.weight.device, self.precision))

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision):
        self.precision = precision
        self._invalidate()
        return self

    def _engine_for(self, like=None):
        """Engine for the device of `like` (a CUDA tensor handed to a sub-module entry point) or of the parameters."""
        dev = like.device if torch.is_tensor(like) and like.is_cuda else self.mel_linear.weight.device
        if self.training:
            raise RuntimeError("styler_b200.STYLER implements the eval-mode forward only; call .eval()")
        if dev.type != "cuda":
            raise RuntimeError("styler_b200.STYLER runs only on a CUDA (sm_100a) device; call .cuda() first -- "
                               "there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        key = (dev, self.precision)
        eng = self._engines.get(key)
        if eng is None:
            master = (self._master() or self) if getattr(self, "_is_replica", False) else self
            with torch.no_grad(), torch.cuda.device(dev):
                eng = Engine(master.state_dict(), dev, self.precision)
            self._engines[key] = eng
        return eng

    # -- reference API ----------------------------------------------------------------------------------------
    def decode(self, style_modeling_output, mel_mask):
        """styler.py:29-37: (mel_output, mel_output_postnet), fp32 [B,T,80]."""
        eng = self._engine_for(style_modeling_output)
        with torch.cuda.device(eng.device):
            x = style_modeling_output.to(eng.device, eng.dt).contiguous()
            lens = (~mel_mask.to(eng.device)).sum(1).to(torch.int64)
            return eng.decode(x, lens)

    @torch.no_grad()
    def forward(self, src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target=None, p_target=None,
                e_target=None, max_src_len=None, max_mel_len=None, speaker_embed=None, d_control=1.0, p_control=1.0,
                e_control=1.0):
        eng = self._engine_for()
        with torch.cuda.device(eng.device):      # kernels, streams and allocations all on the parameters' device
            out = eng.forward(src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target, p_target, e_target,
                              max_src_len, max_mel_len, speaker_embed, d_control, p_control, e_control)
        self.style_modeling._store_inspection(eng, out[5], max_mel_len)
        return out


class GraphedSTYLER:
    """CUDA-graph replay of one forward geometry (shapes, teacher forcing and padded lengths fixed).

    The 130-odd kernel launches of a forward (plus the side-stream fork/join of the audio-encoder branches) are
    captured once into a CUDA graph; `__call__` copies the new inputs into the captured static buffers and replays the
    graph with a single launch, which removes the per-launch host cost that dominates small batches (single-utterance
    latency).  Requirements: teacher-forced durations or an explicit `max_mel_len` (the free-running branch needs one
    host read of max(mel_len), styler.py:47-49 / modules.py:360, which cannot be captured) -- for free-running inference
    capture with the `max_mel_len` you are willing to pad to.
    """

    def __init__(self, model, example_args, example_kwargs, warmup=2, result_mirror=None):
        """result_mirror: uint8 tensor of engine.packed_nbytes(B, T) that ALSO receives the four mels + lengths, written by the
        epilogues of the producing kernels: this rank's slice of rank 0's peer-mapped gather region (dist.PeerGather)."""
        self.model = model
        eng = model._engine_for()
        self._eng = eng          # the graph holds raw pointers into this engine's packed weights / tables / side streams
        dev = eng.device
        if example_kwargs.get("d_target") is None and not example_kwargs.get("max_mel_len"):
            raise ValueError("GraphedSTYLER needs d_target or max_mel_len (no host sync may happen inside a CUDA graph)")
        self.static_args = [a.to(dev).clone() if torch.is_tensor(a) else a for a in example_args]
        self.static_kwargs = {k: (v.to(dev).clone() if torch.is_tensor(v) else v) for k, v in example_kwargs.items()}
        T = self.static_kwargs.get("max_mel_len") or int(self.static_kwargs["d_target"].sum(1).max().item())
        self.static_kwargs["max_mel_len"] = T
        eng._pos("dec", T)                                   # position tables beyond max_seq_len are built on the host: do it now
        eng._pos("enc", self.static_args[0].shape[1])
        eng.result_mirror = result_mirror
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    model(*self.static_args, **self.static_kwargs)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = model(*self.static_args, **self.static_kwargs)
        finally:
            eng.result_mirror = None
        self.packed = eng.last_packed      # the captured results as one byte buffer (dist.AsyncGather.launch_packed)

    def _check(self, kwargs):
        if self.model._engine is not self._eng:
            raise RuntimeError("GraphedSTYLER: the model was moved / reloaded / re-precisioned after capture (its packed "
                               "weights were released); build a new GraphedSTYLER")
        for k, v in kwargs.items():
            if not torch.is_tensor(v) and k in self.static_kwargs and self.static_kwargs[k] != v and \
                    not (k == "max_mel_len" and v is None):
                raise ValueError("GraphedSTYLER: %s=%r differs from the captured %r (scalars are baked into the graph)"
                                 % (k, v, self.static_kwargs[k]))

    def load_inputs(self, *args, **kwargs):
        """Copy new inputs (device or pinned-host tensors) into the captured static buffers on the current stream."""
        self._check(kwargs)
        for dst, src in zip(self.static_args, args):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        for k, v in kwargs.items():
            dst = self.static_kwargs.get(k)
            if torch.is_tensor(dst):
                dst.copy_(v, non_blocking=True)

    def replay(self):
        """Replay the captured forward on the current stream over whatever the static buffers hold; returns the static outputs
        (overwritten by the next replay)."""
        self._check({})
        self.graph.replay()
        return self.static_out

    def __call__(self, *args, **kwargs):
        self.load_inputs(*args, **kwargs)
        return self.replay()


class PipelinedSTYLER:
    """Two-stage software pipeline over consecutive batches of one geometry (a serving loop's view of the forward).

    Stage one of a forward (style encoders + variance adaptor, styler.py:41-49) is latency-bound: BiLSTM recurrences, ~90 small
    kernels, most SMs idle for about a quarter of the step.  Stage two (decoder, mel_linear, PostNet, styler.py:52-57) is four
    fifths of the FLOPs in a handful of machine-filling kernels.  Here each stage is its own CUDA graph per slot: stage one runs
    on a LOW-priority stream, stage two on a HIGH-priority stream, and `submit` of batch i+1 enqueues its stage one while stage
    two of batch i is still running -- the block scheduler hands SM slots to the decoder's CTAs first and fills what they leave
    (kernel tails, the gaps between launches) with the next batch's encoder work.  Results are bitwise those of `STYLER.forward`.

        pipe = PipelinedSTYLER(model, args, kwargs)          # capture (fixed geometry, as GraphedSTYLER)
        k = pipe.submit(*args, **kwargs)                     # enqueue one batch; returns its slot
        out = pipe.outputs(k); pipe.done(k).synchronize()    # the 9-tuple of styler.py:58 (static buffers of slot k)
    A slot's outputs are overwritten by the submit `slots` batches later; that submit waits (stream order) for the events
    passed as `after=` -- record one behind whatever still reads the slot (D2H copy, gather).
    """

    def __init__(self, model, example_args, example_kwargs, slots=2, result_mirrors=None, back_priority=-1):
        self.model = model
        eng = model._engine_for()
        self._eng = eng
        dev = eng.device
        if example_kwargs.get("d_target") is None and not example_kwargs.get("max_mel_len"):
            raise ValueError("PipelinedSTYLER needs d_target or max_mel_len (no host sync may happen inside a CUDA graph)")
        self.slots = int(slots)
        self.s_front = torch.cuda.Stream(device=dev, priority=0)
        self.s_back = torch.cuda.Stream(device=dev, priority=back_priority)
        self.static_args, self.static_kwargs, self.g_front, self.g_back, self.ctx, self.out, self.packed = [], [], [], [], [], [], []
        self.ev_front = [None] * self.slots
        self.ev_back = [None] * self.slots
        self._n = 0
        with torch.cuda.device(dev):
            for k in range(self.slots):
                sa = [a.to(dev).clone() if torch.is_tensor(a) else a for a in example_args]
                skw = {n: (v.to(dev).clone() if torch.is_tensor(v) else v) for n, v in example_kwargs.items()}
                T = skw.get("max_mel_len") or int(skw["d_target"].sum(1).max().item())
                skw["max_mel_len"] = T
                eng._pos("dec", T)
                eng._pos("enc", sa[0].shape[1])
                if k == 0:                                   # eager warm-up on the capture streams: allocator pools, side streams
                    self.s_front.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(self.s_front):
                        model(*sa, **skw)
                    torch.cuda.synchronize(dev)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, stream=self.s_front):
                    ctx = eng.forward_front(*sa, **skw, join=True)
                eng.result_mirror = result_mirrors[k] if result_mirrors is not None else None
                try:
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2, stream=self.s_back):
                        out = eng.forward_back(ctx)
                finally:
                    eng.result_mirror = None
                self.static_args.append(sa)
                self.static_kwargs.append(skw)
                self.g_front.append(g1)
                self.g_back.append(g2)
                self.ctx.append(ctx)
                self.out.append(out)
                self.packed.append(eng.last_packed)
            torch.cuda.synchronize(dev)

    def _check(self, k, kwargs):
        if self.model._engine is not self._eng:
            raise RuntimeError("PipelinedSTYLER: the model was moved / reloaded / re-precisioned after capture; build a new one")
        for n, v in kwargs.items():
            sk = self.static_kwargs[k]
            if not torch.is_tensor(v) and n in sk and sk[n] != v and not (n == "max_mel_len" and v is None):
                raise ValueError("PipelinedSTYLER: %s=%r differs from the captured %r" % (n, v, sk[n]))

    def load_inputs(self, k, *args, **kwargs):
        """Copy a batch (device or pinned-host tensors) into slot k's static inputs on the CURRENT stream, after the stage-one
        graph that last read them."""
        self._check(k, kwargs)
        cur = torch.cuda.current_stream(self._eng.device)
        if self.ev_front[k] is not None:
            cur.wait_event(self.ev_front[k])
        for dst, src in zip(self.static_args[k], args):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        for n, v in kwargs.items():
            dst = self.static_kwargs[k].get(n)
            if torch.is_tensor(dst):
                dst.copy_(v, non_blocking=True)

    def next_slot(self):
        return self._n % self.slots

    def run(self, k=None, after=(), pre_back=None, post_back=None):
        """Enqueue both stages of slot k over whatever its static inputs hold (loaded on the current stream).
        after: events the stage-two graph must wait for before it overwrites slot k's results.
        pre_back / post_back: callables run with the stage-two stream current, right before / after its graph launch
        (gather flow control, dist.AsyncPeerGather.begin / launch_packed)."""
        dev = self._eng.device
        if k is None:
            k = self.next_slot()
        self._n += 1
        cur = torch.cuda.current_stream(dev)
        ev_in = torch.cuda.Event()
        ev_in.record(cur)
        self.s_front.wait_event(ev_in)
        if self.ev_back[k] is not None:
            self.s_front.wait_event(self.ev_back[k])         # stage two of the batch that last used this slot has read its input
        with torch.cuda.stream(self.s_front):
            self.g_front[k].replay()
            ev = torch.cuda.Event()
            ev.record(self.s_front)
        self.ev_front[k] = ev
        self.s_back.wait_event(ev)
        for e in after:
            if e is not None:
                self.s_back.wait_event(e)
        with torch.cuda.stream(self.s_back):
            if pre_back is not None:
                pre_back()
            self.g_back[k].replay()
            ev2 = torch.cuda.Event()
            ev2.record(self.s_back)
            self.ev_back[k] = ev2
            if post_back is not None:
                post_back()
        return k

    def submit(self, *args, after=(), **kwargs):
        k = self.next_slot()
        self.load_inputs(k, *args, **kwargs)
        return self.run(k, after=after)

    def outputs(self, k):
        return self.out[k]

    def done(self, k):
        return self.ev_back[k]

    def join(self):
        """Make the current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self._eng.device)
        cur.wait_stream(self.s_front)
        cur.wait_stream(self.s_back)
