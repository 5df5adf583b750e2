"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the reference's HiFi-GAN generator forward, the step that
follows the mel-synthesis path in every caller (synthesize.py:366,375; utils.py:250-262,276-293).  Only tests/, smoke() and
bench.py's CPU baseline may import this module; the product path (styler_b200/vocoder.py) never does.

Follows hifigan/models.py of the reference:
  * get_padding                      models.py:16-17
  * ResBlock.forward                 models.py:91-98   (leaky_relu 0.1 -> conv(dilated) -> leaky_relu 0.1 -> conv -> + x, three times)
  * Generator.__init__ layer shapes  models.py:108-148 (conv_pre 80->512 k7; ConvTranspose1d k=2u, stride u, pad (k-u)/2;
                                                        resblock kernels 3/7/11 x dilations 1/3/5; conv_post ->1 k7)
  * Generator.forward                models.py:150-166 (note the LAST leaky_relu uses torch's default slope 0.01, line 163)
  * weight_norm (dim=0) folding      torch.nn.utils.weight_norm as applied at models.py:24-76,114-147: w = g * v / ||v||
Pinned by oracle/make_golden_hifigan.py against the unmodified reference module (tests/golden/hifigan_*.pt).
"""
import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1   # models.py:7

# hifigan/config.json of the reference (HiFi-GAN V1); only the generator-shape entries
CONFIG_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                 upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def layer_shapes(h=CONFIG_V1):
    """name -> (weight shape, kind) for every convolution, in state_dict naming (after remove_weight_norm)."""
    shapes = {"conv_pre": ((h["upsample_initial_channel"], h["num_mels"], 7), "conv")}
    ch = h["upsample_initial_channel"]
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        shapes["ups.%d" % i] = ((ch, ch // 2, k), "convT")
        ch //= 2
        for j, ks in enumerate(h["resblock_kernel_sizes"]):
            for c in range(3):
                shapes["resblocks.%d.convs1.%d" % (i * len(h["resblock_kernel_sizes"]) + j, c)] = ((ch, ch, ks), "conv")
                shapes["resblocks.%d.convs2.%d" % (i * len(h["resblock_kernel_sizes"]) + j, c)] = ((ch, ch, ks), "conv")
    shapes["conv_post"] = ((1, ch, 7), "conv")
    return shapes


def make_state_dict(seed=0, h=CONFIG_V1, weight_norm=False):
    """Seeded synthetic generator weights with O(1) activations (the reference's N(0, 0.01) init makes every output ~0).
    weight_norm=True returns the checkpoint form (weight_g / weight_v, as saved by the HiFi-GAN trainer)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (shape, kind) in layer_shapes(h).items():
        fan_in = shape[1] * shape[2] if kind == "conv" else shape[0] * shape[2] / h["upsample_rates"][int(name.split(".")[1])]
        w = torch.randn(*shape, generator=g) / math.sqrt(fan_in)
        if name == "conv_post":
            w = w * 0.5
        b = torch.randn(shape[0] if kind == "conv" else shape[1], generator=g) * 0.05
        if weight_norm:
            norm = w.flatten(1).norm(dim=1).view(-1, 1, 1)
            gain = 1.0 + 0.1 * torch.randn(shape[0], 1, 1, generator=g)
            sd[name + ".weight_v"] = w * 1.7                    # any positive rescaling of v must fold away
            sd[name + ".weight_g"] = norm * gain
        else:
            sd[name + ".weight"] = w
        sd[name + ".bias"] = b
    return sd


def fold_weight_norm(sd):
    """weight_g / weight_v -> weight (torch.nn.utils.weight_norm, dim=0: one gain per index of dimension 0)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            base = k[:-len(".weight_v")]
            g = sd[base + ".weight_g"]
            out[base + ".weight"] = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))
        elif not k.endswith(".weight_g"):
            out[k] = v
    return out


def generator_forward(sd, mel, h=CONFIG_V1):
    """mel [B, 80, T] -> wav [B, 1, T * prod(upsample_rates)]   (Generator.forward, models.py:150-166)."""
    sd = fold_weight_norm(sd)
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd["ups.%d.weight" % i], sd["ups.%d.bias" % i], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (ks, dils) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = "resblocks.%d." % (i * nk + j)
            y = x
            for c, d in enumerate(dils):                      # ResBlock.forward, models.py:91-98
                xt = F.leaky_relu(y, LRELU_SLOPE)
                xt = F.conv1d(xt, sd[p + "convs1.%d.weight" % c], sd[p + "convs1.%d.bias" % c], dilation=d,
                              padding=get_padding(ks, d))
                xt = F.leaky_relu(xt, LRELU_SLOPE)
                xt = F.conv1d(xt, sd[p + "convs2.%d.weight" % c], sd[p + "convs2.%d.bias" % c], padding=get_padding(ks, 1))
                y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                       # default slope 0.01 (models.py:163)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def make_mel(B, T, seed=0):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(B, 80, T, generator=g) * 2.0 - 4.0      # log-mel-like range
