"""Drop-in HiFi-GAN `Generator` (mel -> waveform), the step that follows the mel-synthesis path in every caller of the
reference (utils.get_vocoder / vocoder_infer, utils.py:250-262,276-293; synthesize.py:366,375).  Same constructor
argument (the AttrDict of hifigan/config.json), submodule tree and state_dict keys as hifigan/models.py:104-173 -- in
checkpoint form (weight_g / weight_v, folded on load) or after remove_weight_norm() (weight) -- with the forward executed
by the hand-written sm_100a kernels of libstyler_b200.so.  There is no CPU path: a CPU tensor raises.

B200 design (everything channel-last [B][T][C], bf16 or fp32 storage):
  * every Conv1d is the tcgen05 implicit GEMM `styler_conv1d_fwd` with the tap spacing (dilation) applied to the TMA
    time coordinate; `conv_post` (N = 1) runs on the CUDA-core kernel;
  * ConvTranspose1d(k = 2u, stride u, pad u/2) is ONE 3-tap implicit GEMM with N = u * Cout: output phase r of frame t
    only touches x[t-1], x[t], x[t+1], and the [B, T, u*Cout] result IS the channel-last [B, T*u, Cout] signal;
  * the residual stream is kept in activated form only: y = lrelu(x).  Every convolution input is y (what the reference
    feeds its convs), the residual add recovers x = (y < 0 ? y / 0.1 : y) inside the epilogue, and the epilogue writes
    lrelu(x_new) -- so no stand-alone leaky_relu pass and no second copy of the 17 MB-per-utterance stage tensors exists;
  * the three resblock outputs of a stage are averaged (and re-activated with 0.1, or torch's default 0.01 before
    conv_post, models.py:163) by `styler_lrelu_mean_fwd`.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, IMPL_AUTO, IMPL_SIMT, IMPL_TC

LRELU_SLOPE = 0.1          # hifigan/models.py:7
FINAL_SLOPE = 0.01         # F.leaky_relu default, hifigan/models.py:163

# hifigan/config.json of the reference (HiFi-GAN V1), generator entries only
CONFIG_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                 upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


def get_padding(kernel_size, dilation=1):      # hifigan/models.py:16-17
    return int((kernel_size * dilation - dilation) / 2)


def _cfg(h, key):
    if h is None:
        return CONFIG_V1[key]
    v = h.get(key) if isinstance(h, dict) else getattr(h, key, None)
    return CONFIG_V1[key] if v is None else v


class _ConvParams(nn.Module):
    """Parameter container with nn.Conv1d / nn.ConvTranspose1d's `weight` / `bias` keys (no torch compute)."""

    def __init__(self, wshape, nbias):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(*wshape), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(nbias), requires_grad=False)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the computation runs in styler_b200.vocoder.Generator.forward")


class ResBlock(nn.Module):                     # hifigan/models.py:20-102 (parameters only)
    def __init__(self, channels, kernel_size, dilation):
        super().__init__()
        self.kernel_size, self.dilation = kernel_size, tuple(dilation)
        self.convs1 = nn.ModuleList([_ConvParams((channels, channels, kernel_size), channels) for _ in dilation])
        self.convs2 = nn.ModuleList([_ConvParams((channels, channels, kernel_size), channels) for _ in dilation])

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container")


def pack_conv(w, dtype):
    """nn.Conv1d weight [N, Cin, k] -> [k][N][Cin] (K-major B operand of the implicit GEMM)."""
    return w.permute(2, 0, 1).contiguous().to(dtype)


def pack_conv_transpose(w, bias, stride, dtype):
    """nn.ConvTranspose1d weight [Cin, Cout, k] (k = 2*stride, padding = stride/2... generally (k-stride)/2) ->
    3-tap conv weight [3][stride*Cout][Cin] + bias [stride*Cout]:
        out[t*u + r, co] = b[co] + sum_tau sum_ci x[t + tau - 1, ci] * W[ci, co, (1 - tau)*u + r + p]   (taps outside [0,k) are 0)
    """
    cin, cout, k = w.shape
    u, p = stride, (k - stride) // 2
    assert k <= 2 * u + 2 * p and (k - u) % 2 == 0 and p < u, (k, u, p)
    wp = torch.zeros(3, u * cout, cin, dtype=torch.float32, device=w.device)
    for tau in range(3):
        for r in range(u):
            kidx = (1 - tau) * u + r + p
            if 0 <= kidx < k:
                wp[tau, r * cout:(r + 1) * cout, :] = w[:, :, kidx].t()
    return wp.to(dtype).contiguous(), bias.float().repeat(u).contiguous()


class Generator(nn.Module):
    """hifigan.Generator drop-in: `Generator(h)`, `load_state_dict(ckpt["generator"])`, `.eval()`,
    `.remove_weight_norm()`, `.to(device)`, `vocoder(mel[B, 80, T]) -> wav[B, 1, 256*T]` (fp32)."""

    def __init__(self, h=None, precision="bf16"):
        super().__init__()
        assert precision in ("bf16", "tf32", "fp32")
        self.h, self.precision = h, precision
        self.upsample_rates = list(_cfg(h, "upsample_rates"))
        self.upsample_kernel_sizes = list(_cfg(h, "upsample_kernel_sizes"))
        self.resblock_kernel_sizes = list(_cfg(h, "resblock_kernel_sizes"))
        self.resblock_dilation_sizes = [list(d) for d in _cfg(h, "resblock_dilation_sizes")]
        ch0 = int(_cfg(h, "upsample_initial_channel"))
        self.num_kernels, self.num_upsamples = len(self.resblock_kernel_sizes), len(self.upsample_rates)
        assert str(_cfg(h, "resblock")) == "1", "only ResBlock1 generators (the reference's config) are supported"
        assert 1 <= self.num_kernels <= 3
        self.conv_pre = _ConvParams((ch0, 80, 7), ch0)
        self.ups, self.resblocks = nn.ModuleList(), nn.ModuleList()
        ch = ch0
        for u, k in zip(self.upsample_rates, self.upsample_kernel_sizes):
            self.ups.append(_ConvParams((ch, ch // 2, k), ch // 2))
            ch //= 2
            for ks, d in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes):
                self.resblocks.append(ResBlock(ch, ks, d))
        self.conv_post = _ConvParams((1, ch, 7), 1)
        self._packed = None

    # ---- state_dict surface -------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        sd = {}
        for k, v in state_dict.items():
            k = k[len("module."):] if k.startswith("module.") else k
            if k.endswith(".weight_v"):        # torch.nn.utils.weight_norm(dim=0): w = g * v / ||v|| per index of dim 0
                g = state_dict.get(("module." if ("module." + k) in state_dict else "") + k[:-len("weight_v")] + "weight_g")
                assert g is not None, "weight_v without weight_g for %s" % k
                v = v.float()
                sd[k[:-len("_v")]] = v * (g.float() / v.flatten(1).norm(dim=1).view(-1, 1, 1))
            elif not k.endswith(".weight_g"):
                sd[k] = v
        self._packed = None
        return super().load_state_dict(sd, strict=strict)

    def remove_weight_norm(self):
        """Weight norm is folded when the checkpoint is loaded; kept for call-site compatibility (utils.py:258)."""
        return self

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # ---- weight packing (once per load / .to()) ------------------------------------------------------------------
    def _pack(self):
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        P = {"dtype": dt, "impl": IMPL_SIMT if self.precision == "fp32" else IMPL_TC}
        f32 = lambda t: t.detach().float().contiguous()   # noqa: E731
        P["pre"] = (pack_conv(self.conv_pre.weight.detach(), dt), f32(self.conv_pre.bias))
        P["ups"] = [pack_conv_transpose(m.weight.detach().float(), m.bias.detach(), u, dt)
                    for m, u in zip(self.ups, self.upsample_rates)]
        P["rb"] = [[(pack_conv(c1.weight.detach(), dt), f32(c1.bias), pack_conv(c2.weight.detach(), dt), f32(c2.bias))
                    for c1, c2 in zip(rb.convs1, rb.convs2)] for rb in self.resblocks]
        P["post"] = (pack_conv(self.conv_post.weight.detach(), dt), f32(self.conv_post.bias))
        self._packed = P
        return P

    # ---- forward ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x):
        """x: mel [B, 80, T] (the reference layout, hifigan/models.py:150) -> [B, 1, T * prod(upsample_rates)] fp32."""
        if not x.is_cuda:
            raise RuntimeError("styler_b200.vocoder: expected a CUDA tensor (there is no CPU implementation)")
        assert x.dim() == 3 and x.shape[1] == 80, x.shape
        P = self._packed if self._packed is not None else self._pack()
        dt, impl = P["dtype"], P["impl"]
        B = x.shape[0]
        y = x.detach().transpose(1, 2).to(dt).contiguous()                       # input layout conversion [B, T, 80]
        lre = dict(act=ACT_LRELU, act_slope=LRELU_SLOPE, impl=impl)
        y = ops.conv1d(y, P["pre"][0], P["pre"][1], pad=3, **lre)              # lrelu(conv_pre(x)): models.py:151,153
        for i, u in enumerate(self.upsample_rates):
            w, b = P["ups"][i]
            y = ops.conv1d(y, w, b, pad=1, **lre)                              # y0 = lrelu(ups[i](.)) as [B, T, u*C]
            y0 = y.view(B, y.shape[1] * u, y.shape[2] // u)                      # == channel-last [B, T*u, C]
            outs = []
            for j in range(self.num_kernels):
                rb = self.resblocks[i * self.num_kernels + j]
                yk = y0
                for c, d in enumerate(rb.dilation):                             # ResBlock.forward, models.py:91-98
                    w1, b1, w2, b2 = P["rb"][i * self.num_kernels + j][c]
                    z = ops.conv1d(yk, w1, b1, pad=get_padding(rb.kernel_size, d), dilation=d, **lre)
                    yk = ops.conv1d(z, w2, b2, pad=get_padding(rb.kernel_size, 1), residual=yk, residual_inv_lrelu=True,
                                    act=ACT_NONE, act2=ACT_LRELU, act_slope=LRELU_SLOPE, impl=impl)
                outs.append(yk)
            last = i == self.num_upsamples - 1
            y = ops.lrelu_mean(*outs, slope_in=LRELU_SLOPE, slope_out=FINAL_SLOPE if last else LRELU_SLOPE)
        wav = torch.empty(B, y.shape[1], 1, device=y.device, dtype=torch.float32)
        ops.conv1d(y, P["post"][0], P["post"][1], pad=3, act=ACT_TANH, out_f32=wav, want_out=False,
                   impl=IMPL_SIMT)                                             # N = 1: CUDA-core kernel, fp32 out
        return wav.view(B, 1, -1)
