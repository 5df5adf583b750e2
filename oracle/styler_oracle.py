"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the STYLER non-AR forward.

Functional, state_dict-driven restatement of the reference's eval-mode forward.  Every function cites
the reference lines it follows (paths relative to /root/reference).  It is pinned against the live
reference by ``oracle/make_golden.py`` -> ``tests/golden`` (see oracle/__init__.py).

Conventions: ``sd`` is a state_dict with the reference's 328 keys (no ``module.`` prefix);
masks are bool with True = padding; lengths int64; activations are channel-last [B, T, C].
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# hparams.py:23-76,105 -- only the values the forward path reads
HP = dict(
    n_mel_channels=80, n_bins=256, f0_min=71.0, f0_max=797.9, energy_min=0.1, energy_max=525.43,
    encoder_layer=2, encoder_head=4, encoder_hidden=256, decoder_layer=4, decoder_head=4,
    decoder_hidden=256, fft_conv1d_filter_size=1024, fft_conv1d_kernel_size=(9, 1),
    style_predictor_filter_size=256, style_predictor_kernel_size=3, max_seq_len=1000,
    va_neck_hidden_t=4, va_neck_hidden_r=64, va_neck_hidden_d=80, va_neck_hidden_p=64,
    va_neck_hidden_e=64, va_enc_dim_r=256, va_enc_dim_d=256, va_enc_dim_p=320, va_enc_dim_e=320,
    va_dim_f0=257, va_dim_energy=257, va_chs_grp=16, speaker_embed_dim=512, log_offset=1.0,
    n_src_vocab=152,
)


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def sinusoid_table(n_position, d_hid=256):
    """transformer/Models.py:11-30 -- float64 table cast to float32."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    angle = pos / np.power(10000.0, 2.0 * (j // 2) / d_hid)[None, :]
    tab = angle.copy()
    tab[:, 0::2] = np.sin(angle[:, 0::2])
    tab[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.from_numpy(tab).float()


def mask_from_lengths(lengths, max_len=None):
    """utils.py:223-232 -- True where position >= length."""
    if max_len is None:
        max_len = int(lengths.max().item())
    ids = torch.arange(0, max_len).unsqueeze(0)
    return ids >= lengths.unsqueeze(1)


def _position_rows(sd, key, n):
    """Models.py:69-74,120-125 -- stored table for n <= max_seq_len, rebuilt on the fly beyond."""
    if n > HP["max_seq_len"]:
        return sinusoid_table(n)[:n]
    return sd[key][0, :n]


# ------------------------------------------------------------------------------------------------
# FFT block (transformer/Layers.py:26-34, SubLayers.py:31-61,81-89, Modules.py:14-25)
# ------------------------------------------------------------------------------------------------
def multi_head_attention(sd, p, x, key_pad_mask, n_head=4):
    B, T, D = x.shape
    dk = D // n_head
    q = F.linear(x, sd[p + "w_qs.weight"], sd[p + "w_qs.bias"]).view(B, T, n_head, dk)
    k = F.linear(x, sd[p + "w_ks.weight"], sd[p + "w_ks.bias"]).view(B, T, n_head, dk)
    v = F.linear(x, sd[p + "w_vs.weight"], sd[p + "w_vs.bias"]).view(B, T, n_head, dk)
    q = q.permute(2, 0, 1, 3).reshape(n_head * B, T, dk)          # SubLayers.py:44-49
    k = k.permute(2, 0, 1, 3).reshape(n_head * B, T, dk)
    v = v.permute(2, 0, 1, 3).reshape(n_head * B, T, dk)
    attn = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(dk)         # Modules.py:16-17
    km = key_pad_mask.unsqueeze(1).expand(-1, T, -1).repeat(n_head, 1, 1)
    attn = attn.masked_fill(km, float("-inf"))                     # keys only (Modules.py:19-20)
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, v).view(n_head, B, T, dk).permute(1, 2, 0, 3).reshape(B, T, D)
    out = F.linear(out, sd[p + "fc.weight"], sd[p + "fc.bias"])    # SubLayers.py:58
    return F.layer_norm(out + x, (D,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)


def position_wise_ffn(sd, p, x):
    h = F.conv1d(x.transpose(1, 2), sd[p + "w_1.weight"], sd[p + "w_1.bias"],
                 padding=(sd[p + "w_1.weight"].shape[2] - 1) // 2)
    h = F.conv1d(F.relu(h), sd[p + "w_2.weight"], sd[p + "w_2.bias"],
                 padding=(sd[p + "w_2.weight"].shape[2] - 1) // 2).transpose(1, 2)
    D = x.shape[-1]
    return F.layer_norm(h + x, (D,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)


def fft_block(sd, p, x, pad_mask):
    """Layers.py:26-34 -- MHA, zero padded rows, FFN, zero padded rows."""
    y = multi_head_attention(sd, p + "slf_attn.", x, pad_mask)
    y = y.masked_fill(pad_mask.unsqueeze(-1), 0)
    y = position_wise_ffn(sd, p + "pos_ffn.", y)
    return y.masked_fill(pad_mask.unsqueeze(-1), 0)


def text_encoder(sd, p, src_seq, src_mask):
    """Models.py:60-84."""
    L = src_seq.shape[1]
    x = F.embedding(src_seq, sd[p + "src_word_emb.weight"]) + _position_rows(sd, p + "position_enc", L).unsqueeze(0)
    for i in range(HP["encoder_layer"]):
        x = fft_block(sd, "%slayer_stack.%d." % (p, i), x, src_mask)
    return x


def decoder(sd, p, x, mel_mask):
    """Models.py:111-135."""
    T = x.shape[1]
    x = x + _position_rows(sd, p + "position_enc", T).unsqueeze(0)
    for i in range(HP["decoder_layer"]):
        x = fft_block(sd, "%slayer_stack.%d." % (p, i), x, mel_mask)
    return x


def postnet(sd, p, mel):
    """Layers.py:121-130 -- eval-mode BatchNorm (running stats), tanh on all but the last conv."""
    x = mel.transpose(1, 2)
    for i in range(5):
        q = "%sconvolutions.%d." % (p, i)
        x = F.conv1d(x, sd[q + "0.conv.weight"], sd[q + "0.conv.bias"], padding=2)
        x = F.batch_norm(x, sd[q + "1.running_mean"], sd[q + "1.running_var"], sd[q + "1.weight"],
                         sd[q + "1.bias"], training=False, eps=1e-5)
        if i < 4:
            x = torch.tanh(x)
    return x.transpose(1, 2)


def decode(sd, x, mel_mask):
    """styler.py:29-37."""
    dec = decoder(sd, "decoder.", x, mel_mask)
    mel = F.linear(dec, sd["mel_linear.weight"], sd["mel_linear.bias"])
    return mel, postnet(sd, "postnet.", mel) + mel


# ------------------------------------------------------------------------------------------------
# style encoders (modules.py:164-235, utils.py:351-384,417-429)
# ------------------------------------------------------------------------------------------------
def quantize_index(x, num_bins=256):
    """utils.py:417-429 -- index of the one-hot: 0 if x<=0 else round_half_even(x*255)+1."""
    x = x.clone().float()
    uv = x <= 0
    x[uv] = 0
    assert bool((x >= 0).all()) and bool((x <= 1).all())
    idx = torch.round(x * (num_bins - 1)) + 1
    idx[uv] = 0
    return idx.long()


def encoder_input_cat(mel_target, p_norm, e_input, mel_aug):
    """modules.py:218-223 -- [B, Tr, 80+257+257+80] (channel-last; the reference then transposes)."""
    p1 = F.one_hot(quantize_index(p_norm), 257).float()
    e1 = F.one_hot(quantize_index(e_input), 257).float()
    return torch.cat((mel_target, p1, e1, mel_aug), dim=-1)


def get_scale(src, tgt):
    """utils.py:351-352."""
    return [src // tgt + (1 if i < src % tgt else 0) for i in range(tgt)]


def mel_calibrator(x, mel_len, seq_len):
    """utils.py:355-384 -- per utterance resample ml frames to sl frames (segment mean / repeat)."""
    outs = []
    for b in range(x.shape[0]):
        ml, sl = int(mel_len[b]), int(seq_len[b])
        m = x[b, :ml]
        if ml > sl:
            sizes = get_scale(ml, sl)
            segs = torch.split(m, sizes, dim=0)
            m = torch.stack([s.sum(dim=0) / float(n) for s, n in zip(segs, sizes)])
        elif ml < sl:
            m = torch.repeat_interleave(m, torch.tensor(get_scale(sl, ml)), dim=0)
        outs.append(m)
    T = max(o.shape[0] for o in outs)
    return torch.stack([F.pad(o, (0, 0, 0, T - o.shape[0])) for o in outs])


def _lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM over the padded, un-packed grid; gate order i,f,g,o."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    gx = F.linear(x, w_ih, b_ih)                                   # [B,L,4H]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_zeros(B, L, H)
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        g = gx[:, t] + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def bilstm2(sd, p, x):
    """nn.LSTM(in, H, 2, batch_first=True, bidirectional=True) (modules.py:117,132,147,162,179-182)."""
    for layer in range(2):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            k = "l%d%s" % (layer, suffix)
            outs.append(_lstm_direction(x, sd[p + "weight_ih_" + k], sd[p + "weight_hh_" + k],
                                        sd[p + "bias_ih_" + k], sd[p + "bias_hh_" + k], rev))
        x = torch.cat(outs, dim=-1)
    return x


def audio_encoder(sd, p, cat, mel_len, src_len):
    """modules.py:164-201.  cat: [B, Tr, 674] channel-last.  Returns d[B,L,160], f0/e/noise[B,L,128]."""
    parts = torch.split(cat, [80, 257, 257, 80], dim=-1)
    outs = []
    for n, x in enumerate(parts, start=1):
        x = x.transpose(1, 2)
        for j in range(3):
            q = "%sconvolutions_%d.%d." % (p, n, j)
            x = F.conv1d(x, sd[q + "0.conv.weight"], sd[q + "0.conv.bias"], padding=2)
            C = x.shape[1]
            x = F.relu(F.group_norm(x, C // HP["va_chs_grp"], sd[q + "1.weight"], sd[q + "1.bias"], 1e-5))
        outs.append(x.transpose(1, 2))
    widths = [o.shape[-1] for o in outs]
    cal = mel_calibrator(torch.cat(outs, dim=-1), mel_len, src_len)
    cal = torch.split(cal, widths, dim=-1)
    return tuple(bilstm2(sd, "%slstm_%d." % (p, n), cal[n - 1]) for n in (1, 2, 3, 4))


def _mlp2(sd, p, x):
    x = F.relu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    return F.relu(F.linear(x, sd[p + "2.weight"], sd[p + "2.bias"]))


def augmentation_classifier(sd, p, x):
    """modules.py:38-45 -- GRL is identity in forward; mean over L includes padded positions."""
    q = p + "classifier."
    h = F.linear(x, sd[q + "d_fc1.weight"], sd[q + "d_fc1.bias"])
    h = F.relu(F.layer_norm(h, (h.shape[-1],), sd[q + "d_bn1.weight"], sd[q + "d_bn1.bias"], 1e-5))
    s = F.log_softmax(F.linear(h, sd[q + "d_fc2.weight"], sd[q + "d_fc2.bias"]), dim=-1)
    return s.mean(dim=1) if s.dim() > 2 else s


def style_predictor(sd, p, x, mask):
    """modules.py:457-465 (+ Conv :502-507): no masking between layers."""
    q = p + "conv_layer."
    for i in (1, 2):
        x = F.conv1d(x.transpose(1, 2), sd["%sconv1d_%d.conv.weight" % (q, i)],
                     sd["%sconv1d_%d.conv.bias" % (q, i)], padding=1).transpose(1, 2)
        x = F.layer_norm(F.relu(x), (x.shape[-1],), sd["%slayer_norm_%d.weight" % (q, i)],
                         sd["%slayer_norm_%d.bias" % (q, i)], 1e-5)
    out = F.linear(x, sd[p + "linear_layer.weight"], sd[p + "linear_layer.bias"]).squeeze(-1)
    return out.masked_fill(mask, 0.0) if mask is not None else out


def length_regulator(x, duration, max_len=None):
    """modules.py:396-423 + utils.py:332-348.

    Row i of utterance b is repeated int(duration[b,i]) times (truncation toward zero); mel_len is the
    un-cropped total; the result is zero padded to ``max_len`` (cropped if shorter; a falsy max_len
    means the batch maximum).  Returns (out f32[B,T,C], mel_len int64[B]).
    """
    reps = duration.to(torch.float64).trunc().to(torch.int64) if duration.is_floating_point() else duration.to(torch.int64)
    outs = [torch.repeat_interleave(x[b], reps[b], dim=0) for b in range(x.shape[0])]
    mel_len = torch.tensor([o.shape[0] for o in outs], dtype=torch.int64)
    T = max_len if max_len else int(mel_len.max().item())
    outs = [o[:T] if o.shape[0] >= T else F.pad(o, (0, 0, 0, T - o.shape[0])) for o in outs]
    return torch.stack(outs), mel_len


def duration_from_log(log_d, d_control=1.0):
    """modules.py:290-291,357-358 -- round (half-even) BEFORE scaling by d_control, clamp >= 0."""
    return torch.clamp(torch.round(torch.exp(log_d) - HP["log_offset"]) * d_control, min=0)


# ------------------------------------------------------------------------------------------------
# StyleModeling.forward (modules.py:311-387) and STYLER.forward (styler.py:39-58)
# ------------------------------------------------------------------------------------------------
def style_modeling(sd, src_seq, speaker_embed, mel_target, mel_aug, p_norm, e_input, src_len, mel_len,
                   src_mask, mel_mask=None, d_target=None, p_target=None, e_target=None, max_len=None,
                   d_control=1.0, p_control=1.0, e_control=1.0, want_intermediates=False):
    P = "style_modeling."
    SE = P + "style_encoder."
    # StyleEncoder.forward (modules.py:225-235)
    text = text_encoder(sd, SE + "text_encoder.", src_seq, src_mask)
    text_neck = F.relu(F.linear(text, sd[SE + "text_linear_down.0.weight"], sd[SE + "text_linear_down.0.bias"]))
    spk_p = F.relu(F.linear(speaker_embed, sd[SE + "speaker_linear_p.0.weight"], sd[SE + "speaker_linear_p.0.bias"]))
    spk = F.relu(F.linear(speaker_embed, sd[SE + "speaker_linear.0.weight"], sd[SE + "speaker_linear.0.bias"]))
    cat = encoder_input_cat(mel_target, p_norm, e_input, mel_aug)
    d_enc, p_enc, e_enc, n_enc = audio_encoder(sd, SE + "audio_encoder.", cat, mel_len, src_len)
    L = text.shape[1]

    post_d = augmentation_classifier(sd, P + "augmentation_classifier_d.", d_enc)
    post_p = augmentation_classifier(sd, P + "augmentation_classifier_p.", p_enc)
    post_e = augmentation_classifier(sd, P + "augmentation_classifier_e.", e_enc)

    spk_L = spk.unsqueeze(1).repeat(1, L, 1)                      # modules.py:324-325
    spk_p_L = spk_p.unsqueeze(1).repeat(1, L, 1)
    p_enc = p_enc + spk_p_L                                        # :332

    d_up = _mlp2(sd, P + "duration_linear.", d_enc)                # :335-339
    p_up = _mlp2(sd, P + "pitch_linear.", p_enc)
    e_up = _mlp2(sd, P + "energy_linear.", e_enc)
    n_up = _mlp2(sd, P + "residual_linear.", n_enc)[:, :L]
    neck_up = F.relu(F.linear(text_neck, sd[P + "text_linear_up.0.weight"], sd[P + "text_linear_up.0.bias"]))

    enc = torch.cat((text, neck_up + p_up, spk_L, neck_up + e_up, n_up), dim=-1)   # :350
    log_d = style_predictor(sd, P + "duration_predictor.", neck_up + d_up, src_mask)  # :353
    if d_target is not None:
        enc, mel_len_out = length_regulator(enc, d_target, max_len)
    else:
        enc, mel_len_out = length_regulator(enc, duration_from_log(log_d, d_control), max_len)
        mel_mask = mask_from_lengths(mel_len_out)
    text_T, pitch_T, spk_T, energy_T, noise_T = torch.split(enc, 256, dim=-1)

    e_pred = style_predictor(sd, P + "energy_predictor.", energy_T, mel_mask)         # :365-372
    if e_target is not None:
        e_idx = torch.bucketize(e_target, sd[P + "energy_bins"])
    else:
        e_pred = e_pred * e_control
        e_idx = torch.bucketize(e_pred, sd[P + "energy_bins"])
    p_pred = style_predictor(sd, P + "pitch_predictor.", pitch_T + spk_T, mel_mask)   # :375-382
    if p_target is not None:
        p_idx = torch.bucketize(p_target, sd[P + "pitch_bins"])
    else:
        p_pred = p_pred * p_control
        p_idx = torch.bucketize(p_pred, sd[P + "pitch_bins"])
    out = text_T + F.embedding(p_idx, sd[P + "pitch_embedding.weight"]) + spk_T \
        + F.embedding(e_idx, sd[P + "energy_embedding.weight"])                       # :385
    res = (out, noise_T, log_d, p_pred, e_pred, mel_len_out, mel_mask, (post_d, post_p, post_e))
    if want_intermediates:
        inter = dict(text=text, text_neck=text_neck, spk=spk, spk_p=spk_p, d_enc=d_enc, p_enc_raw=p_enc - spk_p_L,
                     e_enc=e_enc, n_enc=n_enc, d_up=d_up, p_up=p_up, e_up=e_up, n_up=n_up, neck_up=neck_up,
                     enc_T=enc, p_idx=p_idx, e_idx=e_idx)
        return res, inter
    return res


def predict_inference(sd, text_encoding, pitch_encoding, energy_encoding, duration_encoding, speaker_encoding,
                      noise_encoding, src_mask, max_len, speaker_normalized=True, d_control=1.0, p_control=1.0,
                      e_control=1.0):
    """modules.py:285-309 (StyleModeling.predict_inference, used by synthesize.py:171)."""
    P = "style_modeling."
    enc = torch.cat((text_encoding, pitch_encoding, speaker_encoding, energy_encoding, noise_encoding), dim=-1)
    log_d = style_predictor(sd, P + "duration_predictor.", duration_encoding, src_mask)
    enc, mel_len = length_regulator(enc, duration_from_log(log_d, d_control), max_len)
    mel_mask = mask_from_lengths(mel_len)
    text_T, pitch_T, spk_T, energy_T, noise_T = torch.split(enc, 256, dim=-1)
    e_pred = style_predictor(sd, P + "energy_predictor.", energy_T, mel_mask) * e_control
    e_emb = F.embedding(torch.bucketize(e_pred, sd[P + "energy_bins"]), sd[P + "energy_embedding.weight"])
    p_pred = style_predictor(sd, P + "pitch_predictor.", pitch_T if speaker_normalized else pitch_T + spk_T, mel_mask) * p_control
    p_emb = F.embedding(torch.bucketize(p_pred, sd[P + "pitch_bins"]), sd[P + "pitch_embedding.weight"])
    return text_T, p_emb, spk_T, e_emb, noise_T, log_d, p_pred, e_pred, mel_mask


def styler_forward(sd, src_seq, mel_target, mel_aug, p_norm, e_input, src_len, mel_len, d_target=None,
                   p_target=None, e_target=None, max_src_len=None, max_mel_len=None, speaker_embed=None,
                   d_control=1.0, p_control=1.0, e_control=1.0):
    """styler.py:39-58.  Returns the reference's 9-tuple."""
    src_mask = mask_from_lengths(src_len, max_src_len)
    mel_mask = mask_from_lengths(mel_len, max_mel_len)
    sm = style_modeling(sd, src_seq, speaker_embed, mel_target, mel_aug, p_norm, e_input, src_len, mel_len,
                        src_mask, mel_mask, d_target, p_target, e_target, max_mel_len, d_control, p_control,
                        e_control)
    x, noise, log_d, p_pred, e_pred, mel_len_lr, mel_mask_lr, post = sm
    if d_target is None:                                            # styler.py:47-49
        mel_len, mel_mask = mel_len_lr, mel_mask_lr
    mel, mel_post = decode(sd, x, mel_mask)
    mel_n, mel_post_n = decode(sd, x + noise, mel_mask)
    return (mel, mel_n), (mel_post, mel_post_n), log_d, p_pred, e_pred, src_mask, mel_mask, mel_len, post


# ------------------------------------------------------------------------------------------------
# seeded synthetic weights / inputs live in the product package (styler_b200/synthetic.py) so that bench.py's GPU arm never
# imports oracle/; re-exported here for the tests
# ------------------------------------------------------------------------------------------------
from styler_b200.synthetic import make_inputs, make_state_dict, set_duration_bias  # noqa: E402,F401
