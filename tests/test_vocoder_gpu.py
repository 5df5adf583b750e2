"""GPU parity tests of the HiFi-GAN vocoder row through the C ABI: the conv epilogue extensions it needs (dilation,
leaky ReLU, inverse-lrelu residual), the multi-receptive-field mean, and the full generator against the golden waveform
of the unmodified reference (tests/golden/hifigan_b2_t24.pt) in all three precision modes."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import hifigan_oracle as ho

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("shape", [(2, 700, 32, 32, 11, 5), (2, 300, 64, 64, 7, 3), (3, 200, 128, 128, 3, 1), (1, 260, 256, 256, 3, 5)],
                         ids=["c32_k11_d5", "c64_k7_d3", "c128_k3_d1", "c256_k3_d5"])
def test_conv1d_dilated_lrelu_residual(cuda, dtype, impl, shape):
    """One resblock step as the generator issues it: z = lrelu(conv_dil(y)); y' = lrelu(conv(z) + inv_lrelu(y))."""
    from styler_b200 import ops
    B, T, Cin, N, KS, dil = shape
    g = torch.Generator().manual_seed(KS * 100 + dil)
    y = torch.randn(B, T, Cin, generator=g)
    y = torch.where(y < 0, y * 0.1, y)                                   # an activated residual stream
    w1 = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    w2 = (torch.rand(KS, N, N, generator=g) * 2 - 1) / math.sqrt(N * KS)
    b1, b2 = torch.randn(N, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    q = lambda t: t.to(dtype).float()                                    # noqa: E731  operands as the kernel sees them
    lre = lambda t: torch.where(t < 0, t * 0.1, t)                       # noqa: E731
    conv = lambda x, w, b, d: F.conv1d(x.transpose(1, 2), q(w).permute(1, 2, 0).contiguous(), b, padding=(KS * d - d) // 2,  # noqa: E731
                                       dilation=d).transpose(1, 2)
    z_ref = lre(conv(q(y), w1, b1, dil))
    kimpl = ops.IMPL_SIMT if impl == "simt" else ops.IMPL_TC
    yd = y.to(cuda, dtype)
    z = ops.conv1d(yd, w1.to(cuda, dtype), b1.to(cuda), pad=(KS * dil - dil) // 2, dilation=dil, act=ops.ACT_LRELU,
                   act_slope=0.1, impl=kimpl)
    tol = 2e-3 if (dtype == torch.float32 and impl == "tc") else 2e-5
    tol += 8e-3 if dtype == torch.bfloat16 else 0.0
    assert rel_err(z, z_ref) < tol
    yq = q(y)
    y2_ref = lre(conv(z.float().cpu(), w2, b2, 1) + torch.where(yq < 0, yq / 0.1, yq))
    y2 = ops.conv1d(z, w2.to(cuda, dtype), b2.to(cuda), pad=(KS - 1) // 2, residual=yd, residual_inv_lrelu=True,
                    act2=ops.ACT_LRELU, act_slope=0.1, impl=kimpl)
    torch.cuda.synchronize()
    assert rel_err(y2, y2_ref) < tol


@pytest.mark.parametrize("shape", [(4, 10000, 32, 32, 11, 5), (4, 10000, 32, 32, 3, 1), (4, 10000, 64, 64, 7, 3), (5, 8100, 64, 64, 11, 5),
                                   (4, 10000, 64, 32, 7, 1)],
                         ids=["c32_k11_d5", "c32_k3_d1", "c64_k7_d3", "c64_k11_d5", "c64_n32_k7"])
def test_conv1d_window_form(cuda, shape):
    """Enough 128-step tiles (>= 2 per CTA slot) for the persistent window kernel (conv_win.cu): weights resident in smem,
    one TMA window per tile, per-tap A descriptors at row offsets; ragged T, both epilogue forms of a resblock step."""
    from styler_b200 import ops
    B, T, Cin, N, KS, dil = shape
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(KS * 10 + dil + Cin)
    y = torch.randn(B, T, Cin, generator=g)
    y = torch.where(y < 0, y * 0.1, y)
    w1 = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    b1 = torch.randn(N, generator=g) * 0.1
    q = lambda t: t.to(dtype).float()                                    # noqa: E731
    lre = lambda t: torch.where(t < 0, t * 0.1, t)                       # noqa: E731
    pad = (KS * dil - dil) // 2
    z_ref = lre(F.conv1d(q(y).transpose(1, 2), q(w1).permute(1, 2, 0).contiguous(), b1, padding=pad, dilation=dil).transpose(1, 2))
    yd = y.to(cuda, dtype)
    z = ops.conv1d(yd, w1.to(cuda, dtype), b1.to(cuda), pad=pad, dilation=dil, act=ops.ACT_LRELU, act_slope=0.1, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    for b in range(B):
        assert rel_err(z[b], z_ref[b]) < 1e-2, ("plain", b)
    # edges of every utterance (zero padding through TMA out-of-range rows) and an interior tile boundary
    for sl in (slice(0, 64), slice(T - 64, T), slice(128 * 3 - 32, 128 * 3 + 32)):
        assert rel_err(z[:, sl], z_ref[:, sl]) < 1e-2
    if Cin == N:
        w2 = (torch.rand(KS, N, N, generator=g) * 2 - 1) / math.sqrt(N * KS)
        b2 = torch.randn(N, generator=g) * 0.1
        yq = q(y)
        y2_ref = lre(F.conv1d(z.float().cpu().transpose(1, 2), q(w2).permute(1, 2, 0).contiguous(), b2,
                              padding=(KS - 1) // 2).transpose(1, 2) + torch.where(yq < 0, yq / 0.1, yq))
        y2 = ops.conv1d(z, w2.to(cuda, dtype), b2.to(cuda), pad=(KS - 1) // 2, residual=yd, residual_inv_lrelu=True,
                        act2=ops.ACT_LRELU, act_slope=0.1, impl=ops.IMPL_TC)
        torch.cuda.synchronize()
        for b in range(B):
            assert rel_err(y2[b], y2_ref[b]) < 1e-2, ("residual", b)
        y3 = ops.conv1d(z, w2.to(cuda, dtype), b2.to(cuda), pad=(KS - 1) // 2, residual=yd, residual_inv_lrelu=True,
                        act2=ops.ACT_LRELU, act_slope=0.1, impl=ops.IMPL_TC)
        torch.cuda.synchronize()
        assert torch.equal(y2, y3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(2, 700, 32, 1, 7, 1, 2), (3, 300, 64, 4, 3, 2, 0), (1, 1000, 128, 2, 5, 3, 3)],
                         ids=["conv_post_32_1_k7_tanh", "c64_n4_k3_d2", "c128_n2_k5_d3_lrelu"])
def test_conv1d_small_n(cuda, dtype, shape):
    """The one-thread-per-time-step CUDA-core kernel used for N <= 4 (HiFi-GAN conv_post + tanh, models.py:164-165)."""
    from styler_b200 import ops
    B, T, Cin, N, KS, dil, act = shape
    g = torch.Generator().manual_seed(Cin + N)
    x = torch.randn(B, T, Cin, generator=g)
    w = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    bias = torch.randn(N, generator=g) * 0.1
    q = lambda t: t.to(dtype).float()                                    # noqa: E731
    pad = (KS * dil - dil) // 2
    ref = F.conv1d(q(x).transpose(1, 2), q(w).permute(1, 2, 0).contiguous(), bias, padding=pad, dilation=dil).transpose(1, 2)
    ref = {0: ref, 2: torch.tanh(ref), 3: torch.where(ref < 0, ref * 0.1, ref)}[act]
    out_f32 = torch.empty(B, T, N, device=cuda)
    ops.conv1d(x.to(cuda, dtype), w.to(cuda, dtype), bias.to(cuda), pad=pad, dilation=dil, act=act, act_slope=0.1,
               out_f32=out_f32, want_out=False, impl=ops.IMPL_SIMT)
    torch.cuda.synchronize()
    assert rel_err(out_f32, ref) < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_lrelu_mean(cuda, dtype):
    from styler_b200 import ops
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(2, 100, 64, generator=g) * 3 for _ in range(3)]
    ys = [torch.where(x < 0, x * 0.1, x).to(dtype) for x in xs]
    inv = lambda t: torch.where(t < 0, t / 0.1, t)                        # noqa: E731
    for n, so_ in ((3, 0.1), (3, 0.01), (2, 0.1), (1, 0.01)):
        m = sum(inv(t.float()) for t in ys[:n]) / n
        ref = torch.where(m < 0, m * so_, m)
        got = ops.lrelu_mean(*[t.to(cuda) for t in ys[:n]], slope_in=0.1, slope_out=so_)
        assert rel_err(got, ref) < (8e-3 if dtype == torch.bfloat16 else 1e-6)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tf32", 1e-2), ("bf16", 8e-2)])
def test_generator_matches_reference_golden(cuda, precision, tol):
    from styler_b200.vocoder import Generator
    gold = torch.load(os.path.join(GOLD, "hifigan_b2_t24.pt"))
    sd = ho.make_state_dict(seed=gold["seed"], weight_norm=True)
    mel = ho.make_mel(gold["B"], gold["T"], seed=gold["seed"])
    voc = Generator(precision=precision)
    voc.load_state_dict(sd)
    voc = voc.eval().to(cuda)
    voc.remove_weight_norm()
    wav = voc(mel.to(cuda))
    torch.cuda.synchronize()
    assert wav.shape == gold["wav"].shape and wav.dtype == torch.float32
    assert torch.isfinite(wav).all()
    err = rel_err(wav, gold["wav"])
    rms = ((wav.cpu() - gold["wav"]).pow(2).mean().sqrt() / gold["wav"].pow(2).mean().sqrt()).item()
    print("hifigan %s: max-normalised err %.3e, relative rms err %.3e" % (precision, err, rms))
    assert err < tol, (precision, err)


def test_generator_longer_utterance_against_oracle(cuda):
    """A 100-frame, 3-utterance batch (25,600 samples each) against the CPU oracle, tf32 storage."""
    from styler_b200.vocoder import Generator
    sd = ho.make_state_dict(seed=11, weight_norm=False)
    mel = ho.make_mel(3, 100, seed=11)
    with torch.no_grad():
        ref = ho.generator_forward(sd, mel)
    voc = Generator(precision="tf32")
    voc.load_state_dict(sd)
    wav = voc.to(cuda)(mel.to(cuda))
    torch.cuda.synchronize()
    assert rel_err(wav, ref) < 1e-2
