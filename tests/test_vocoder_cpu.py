"""CPU tests of the HiFi-GAN vocoder row: the oracle against the golden waveform produced by
the unmodified reference generator (oracle/make_golden_hifigan.py), and the host-side logic of the drop-in `Generator`
(state_dict surface, weight-norm folding, ConvTranspose1d phase packing, the activated-residual formulation)."""
import os

import torch
import torch.nn.functional as F

from oracle import hifigan_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_hifigan_oracle_matches_reference_golden():
    gold = torch.load(os.path.join(GOLD, "hifigan_b2_t24.pt"))
    sd = ho.make_state_dict(seed=gold["seed"], weight_norm=True)
    assert len(sd) == gold["n_state_tensors"]
    mel = ho.make_mel(gold["B"], gold["T"], seed=gold["seed"])
    with torch.no_grad():
        wav = ho.generator_forward(sd, mel)
    assert wav.shape == gold["wav"].shape == (gold["B"], 1, gold["T"] * 256)
    assert (wav - gold["wav"]).abs().max().item() < 1e-5


def test_generator_state_dict_surface_and_weight_norm_folding():
    from styler_b200.vocoder import Generator
    g = Generator()
    shapes = ho.layer_shapes()
    keys = set(g.state_dict().keys())
    assert keys == {n + s for n in shapes for s in (".weight", ".bias")}
    for n, (shape, _) in shapes.items():
        assert tuple(g.state_dict()[n + ".weight"].shape) == tuple(shape), n
    sd_wn = ho.make_state_dict(seed=3, weight_norm=True)
    g.load_state_dict({"module." + k: v for k, v in sd_wn.items()})        # DataParallel-style prefix is accepted too
    folded = ho.fold_weight_norm(sd_wn)
    for k, v in g.state_dict().items():
        assert torch.allclose(v, folded[k], atol=1e-6), k
    g.load_state_dict(folded)                                              # plain (remove_weight_norm) form
    assert g.remove_weight_norm() is g


def test_conv_transpose_phase_packing():
    from styler_b200.vocoder import pack_conv_transpose
    g = torch.Generator().manual_seed(0)
    for cin, cout, k, u in [(16, 8, 16, 8), (8, 4, 4, 2), (6, 3, 8, 4)]:
        w, b = torch.randn(cin, cout, k, generator=g), torch.randn(cout, generator=g)
        x = torch.randn(2, cin, 13, generator=g)
        ref = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        wp, bp = pack_conv_transpose(w, b, u, torch.float32)
        y = F.conv1d(x, wp.permute(1, 2, 0).contiguous(), bp, padding=1)     # 3-tap conv, N = u*cout
        y = y.transpose(1, 2).reshape(2, 13 * u, cout).transpose(1, 2)       # [B, T, u*C] read as [B, T*u, C]
        assert (y - ref).abs().max().item() < 1e-5


def test_activated_residual_formulation_matches_oracle():
    """The exact sequence of kernel contracts Generator.forward issues (packed weights, lrelu epilogues, inverse-lrelu
    residual, lrelu_mean), evaluated with torch on the CPU, reproduces the oracle waveform."""
    from styler_b200 import vocoder as V
    sd = ho.make_state_dict(seed=5, weight_norm=True)
    mel = ho.make_mel(2, 10, seed=5)
    with torch.no_grad():
        ref = ho.generator_forward(sd, mel)
    g = V.Generator(precision="fp32")
    g.load_state_dict(sd)
    P = g._pack()
    lre = lambda t, s=0.1: torch.where(t < 0, t * s, t)      # noqa: E731
    inv = lambda t: torch.where(t < 0, t / 0.1, t)            # noqa: E731

    def conv(x, wp, b, pad, dil=1):
        return F.conv1d(x.transpose(1, 2), wp.permute(1, 2, 0).contiguous(), b, padding=pad, dilation=dil).transpose(1, 2)

    y = lre(conv(mel.transpose(1, 2), *P["pre"], 3))
    for i, u in enumerate(g.upsample_rates):
        y = lre(conv(y, *P["ups"][i], 1))
        y0 = y.reshape(y.shape[0], y.shape[1] * u, y.shape[2] // u)
        outs = []
        for j in range(g.num_kernels):
            rb, yk = g.resblocks[i * g.num_kernels + j], y0
            for c, d in enumerate(rb.dilation):
                w1, b1, w2, b2 = P["rb"][i * g.num_kernels + j][c]
                z = lre(conv(yk, w1, b1, V.get_padding(rb.kernel_size, d), d))
                yk = lre(conv(z, w2, b2, V.get_padding(rb.kernel_size, 1)) + inv(yk))
            outs.append(yk)
        y = lre(sum(inv(o) for o in outs) / g.num_kernels, 0.01 if i == g.num_upsamples - 1 else 0.1)
    wav = torch.tanh(conv(y, *P["post"], 3)).reshape(2, 1, -1)
    assert (wav - ref).abs().max().item() < 1e-5


def test_generator_refuses_cpu_tensors():
    import pytest
    from styler_b200.vocoder import Generator
    with pytest.raises(RuntimeError):
        Generator()(torch.zeros(1, 80, 4))


def _real_checkpoint(tmp_path):
    """The reference ships a pretrained generator (hifigan/generator_LJSpeech.pth.tar.zip); dev container only."""
    import zipfile
    z = "/root/reference/hifigan/generator_LJSpeech.pth.tar.zip"
    if not os.path.isfile(z):
        import pytest
        pytest.skip("reference tree (and its pretrained vocoder checkpoint) not present")
    zipfile.ZipFile(z).extractall(tmp_path)
    return torch.load(os.path.join(tmp_path, "generator_LJSpeech.pth.tar"), map_location="cpu", weights_only=False)["generator"]


def test_pretrained_checkpoint_loads_and_oracle_matches_live_reference(tmp_path):
    """With the REAL LJSpeech weights (realistic dynamic range): the drop-in Generator accepts the checkpoint as saved
    (234 weight_g / weight_v / bias tensors), folds it to exactly what the reference's remove_weight_norm() yields, and the
    oracle reproduces the live reference generator on it."""
    import contextlib
    import io
    import sys
    sd = _real_checkpoint(tmp_path)
    assert len(sd) == 234
    sys.path.insert(0, "/root/reference")
    import hifigan                                      # the reference package (torch only)
    ref = hifigan.Generator(hifigan.AttrDict(ho.CONFIG_V1)).eval()
    ref.load_state_dict(sd)
    with contextlib.redirect_stdout(io.StringIO()):
        ref.remove_weight_norm()
    from styler_b200.vocoder import Generator
    g = Generator()
    g.load_state_dict(sd)
    for k, v in ref.state_dict().items():
        assert torch.allclose(g.state_dict()[k], v, rtol=1e-5, atol=1e-7), k
    mel = ho.make_mel(1, 12, seed=3)
    with torch.no_grad():
        y_ref = ref(mel)
        y_orc = ho.generator_forward(sd, mel)
    assert (y_ref - y_orc).abs().max().item() < 1e-5 * max(1.0, y_ref.abs().max().item())
