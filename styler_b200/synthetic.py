"""Seeded synthetic weights and input batches with the reference's shapes.

Neither /root/reference nor its checkpoints exist on the GPU box, so tests, smoke() and bench.py build the model from this
deterministic generator (torch CPU generator => identical values on every machine; tests/golden pins a SHA-256 of the
result).  Pure data generation: no model arithmetic lives here.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401

from .engine import sinusoid_table

HP = dict(n_bins=256, f0_min=71.0, f0_max=797.9, energy_min=0.1, energy_max=525.43, n_src_vocab=152)


def mask_from_lengths(lengths, max_len=None):
    if max_len is None:
        max_len = int(lengths.max().item())
    return torch.arange(0, max_len).unsqueeze(0) >= lengths.unsqueeze(1)


def _u(g, shape, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2 - 1) * b


def make_state_dict(seed=0):
    """328-tensor state_dict with the reference's keys/shapes, seeded values."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(k, o, i):
        sd[k + ".weight"] = _u(g, (o, i), i)
        sd[k + ".bias"] = _u(g, (o,), i)

    def conv(k, o, i, ks):
        sd[k + ".weight"] = _u(g, (o, i, ks), i * ks)
        sd[k + ".bias"] = _u(g, (o,), i * ks)

    def affine(k, n):
        sd[k + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[k + ".bias"] = 0.1 * torch.randn(n, generator=g)

    def fft(p):
        for n in ("w_qs", "w_ks", "w_vs", "fc"):
            lin(p + "slf_attn." + n, 256, 256)
        affine(p + "slf_attn.layer_norm", 256)
        conv(p + "pos_ffn.w_1", 1024, 256, 9)
        conv(p + "pos_ffn.w_2", 256, 1024, 1)
        affine(p + "pos_ffn.layer_norm", 256)

    P = "style_modeling."
    SE = P + "style_encoder."
    sd[P + "pitch_bins"] = torch.exp(torch.linspace(np.log(HP["f0_min"]), np.log(HP["f0_max"]), HP["n_bins"] - 1))
    sd[P + "energy_bins"] = torch.linspace(HP["energy_min"], HP["energy_max"], HP["n_bins"] - 1)
    sd[SE + "text_encoder.position_enc"] = sinusoid_table(1001).unsqueeze(0)
    emb = torch.randn(HP["n_src_vocab"], 256, generator=g)
    emb[0] = 0                                                      # padding_idx=0 (Models.py:52-53)
    sd[SE + "text_encoder.src_word_emb.weight"] = emb
    for i in range(2):
        fft("%stext_encoder.layer_stack.%d." % (SE, i))
    AE = SE + "audio_encoder."
    for n, (cin, c) in enumerate(((80, 256), (257, 320), (257, 320), (80, 256)), start=1):
        for j in range(3):
            conv("%sconvolutions_%d.%d.0.conv" % (AE, n, j), c, cin if j == 0 else c, 5)
            affine("%sconvolutions_%d.%d.1" % (AE, n, j), c)
    for n, (cin, H) in enumerate(((256, 80), (320, 64), (320, 64), (256, 64)), start=1):
        for layer in range(2):
            for suffix in ("", "_reverse"):
                k = "l%d%s" % (layer, suffix)
                i = cin if layer == 0 else 2 * H
                sd["%slstm_%d.weight_ih_%s" % (AE, n, k)] = _u(g, (4 * H, i), H)
                sd["%slstm_%d.weight_hh_%s" % (AE, n, k)] = _u(g, (4 * H, H), H)
                sd["%slstm_%d.bias_ih_%s" % (AE, n, k)] = _u(g, (4 * H,), H)
                sd["%slstm_%d.bias_hh_%s" % (AE, n, k)] = _u(g, (4 * H,), H)
    lin(SE + "text_linear_down.0", 4, 256)
    lin(SE + "speaker_linear_p.0", 128, 512)
    lin(SE + "speaker_linear.0", 256, 512)
    for n, i in (("d", 160), ("p", 128), ("e", 128)):
        q = "%saugmentation_classifier_%s.classifier." % (P, n)
        lin(q + "d_fc1", 256, i)
        affine(q + "d_bn1", 256)
        lin(q + "d_fc2", 2, 256)
    for n, i in (("duration", 160), ("pitch_norm", 128), ("pitch", 128), ("energy", 128), ("residual", 128)):
        lin("%s%s_linear.0" % (P, n), 256, i)
        lin("%s%s_linear.2" % (P, n), 256, 256)
    lin(P + "text_linear_up.0", 256, 4)
    for n in ("duration", "pitch", "energy"):
        q = "%s%s_predictor." % (P, n)
        for i in (1, 2):
            conv("%sconv_layer.conv1d_%d.conv" % (q, i), 256, 256, 3)
            affine("%sconv_layer.layer_norm_%d" % (q, i), 256)
        lin(q + "linear_layer", 1, 256)
    sd[P + "pitch_embedding.weight"] = torch.randn(256, 256, generator=g)
    sd[P + "energy_embedding.weight"] = torch.randn(256, 256, generator=g)
    sd["decoder.position_enc"] = sinusoid_table(1001).unsqueeze(0)
    for i in range(4):
        fft("decoder.layer_stack.%d." % i)
    lin("mel_linear", 80, 256)
    for j, (cin, c) in enumerate(((80, 512), (512, 512), (512, 512), (512, 512), (512, 80))):
        conv("postnet.convolutions.%d.0.conv" % j, c, cin, 5)
        affine("postnet.convolutions.%d.1" % j, c)
        sd["postnet.convolutions.%d.1.running_mean" % j] = 0.1 * torch.randn(c, generator=g)
        sd["postnet.convolutions.%d.1.running_var" % j] = 1.0 + 0.2 * torch.rand(c, generator=g)
        sd["postnet.convolutions.%d.1.num_batches_tracked" % j] = torch.tensor(0, dtype=torch.int64)
    return sd


def set_duration_bias(sd, frames_per_phoneme=8):
    """make the free-running branch predict a fixed duration:
    linear weight 0, bias log(frames+1) -> round(exp(log_d) - 1) == frames."""
    k = "style_modeling.duration_predictor.linear_layer."
    sd[k + "weight"] = torch.zeros_like(sd[k + "weight"])
    sd[k + "bias"] = torch.full_like(sd[k + "bias"], math.log(frames_per_phoneme + 1.0))
    return sd


def make_inputs(B, L, Tr=None, seed=1234, ragged=False, d_mode="const8", frames=8):
    """Seeded synthetic batch.

    d_mode 'const' (alias 'const8'): teacher-forced, `frames` frames per phoneme; 'ragged': teacher-forced,
    d ~ randint(0,13); None: free running.  In the teacher-forced modes the reference's contract
    (train.py:135, dataset.py:210-226) is mel_len == d_target.sum(1) and Tr == max(mel_len), so Tr and
    mel_len are derived from the durations; in free-running mode Tr is the argument and mel_len is ragged
    when `ragged`.
    """
    g = torch.Generator().manual_seed(seed)
    src_len = torch.full((B,), L, dtype=torch.int64)
    if ragged and B > 1:
        src_len[1:] = torch.randint(max(1, L // 2), L + 1, (B - 1,), generator=g)
    pad_src = mask_from_lengths(src_len, L)
    src_seq = torch.randint(1, HP["n_src_vocab"], (B, L), generator=g).masked_fill(pad_src, 0)
    d = None
    if d_mode in ("const", "const8"):
        d = torch.full((B, L), frames, dtype=torch.int64).masked_fill(pad_src, 0)
    elif d_mode == "ragged":
        d = torch.randint(0, 13, (B, L), generator=g).masked_fill(pad_src, 0)
    if d is not None:
        mel_len = d.sum(dim=1)
        Tr = int(mel_len.max().item())
    else:
        mel_len = torch.full((B,), Tr, dtype=torch.int64)
        if ragged and B > 1:
            mel_len[1:] = torch.randint(max(1, Tr // 2), Tr + 1, (B - 1,), generator=g)
    pad_mel = mask_from_lengths(mel_len, Tr)
    mel = torch.randn(B, Tr, 80, generator=g).masked_fill(pad_mel.unsqueeze(-1), 0)
    p_norm = torch.rand(B, Tr, generator=g)
    p_norm[torch.rand(B, Tr, generator=g) < 0.2] = 0.0              # unvoiced frames -> index 0
    p_norm = p_norm.masked_fill(pad_mel, 0)
    e_in = torch.rand(B, Tr, generator=g).masked_fill(pad_mel, 0)
    spk = torch.randn(B, 512, generator=g)
    batch = dict(src_seq=src_seq, mel_target=mel, mel_aug=mel + 0.05 * torch.randn(B, Tr, 80, generator=g).masked_fill(pad_mel.unsqueeze(-1), 0),
                 p_norm=p_norm, e_input=e_in, src_len=src_len, mel_len=mel_len, speaker_embed=spk, max_src_len=L)
    if d is not None:
        batch.update(d_target=d, max_mel_len=Tr)
        batch["p_target"] = (torch.rand(B, Tr, generator=g) * 900.0).masked_fill(pad_mel, 0)
        batch["e_target"] = (torch.rand(B, Tr, generator=g) * 600.0).masked_fill(pad_mel, 0)
    return batch


# ---------------------------------------------------------------------------------------------------------------- vocoder
def make_vocoder_state_dict(seed=0, h=None):
    """Seeded HiFi-GAN generator weights in plain (remove_weight_norm) form with O(1) activations; same keys/shapes as
    hifigan/models.py:104-148.  (The reference's own N(0, 0.01) init makes every output vanish.)"""
    from .vocoder import CONFIG_V1
    h = CONFIG_V1 if h is None else h
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def put(name, shape, fan_in, nbias, scale=1.0):
        sd[name + ".weight"] = torch.randn(*shape, generator=g) / math.sqrt(fan_in) * scale
        sd[name + ".bias"] = torch.randn(nbias, generator=g) * 0.05

    ch = h["upsample_initial_channel"]
    put("conv_pre", (ch, h["num_mels"], 7), h["num_mels"] * 7, ch)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        put("ups.%d" % i, (ch, ch // 2, k), ch * k / u, ch // 2)
        ch //= 2
        for j, ks in enumerate(h["resblock_kernel_sizes"]):
            for c in range(3):
                put("resblocks.%d.convs1.%d" % (i * nk + j, c), (ch, ch, ks), ch * ks, ch)
                put("resblocks.%d.convs2.%d" % (i * nk + j, c), (ch, ch, ks), ch * ks, ch)
    put("conv_post", (1, ch, 7), ch * 7, 1, scale=0.5)
    return sd


def make_mel(B, T, seed=0):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(B, 80, T, generator=g) * 2.0 - 4.0      # log-mel-like range
