"""CPU tests (-m "not gpu"): the oracle restatement against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py), the seeded generators, and -- only where /root/reference exists (dev container) -- a live
comparison against the reference itself."""
import os

import pytest
import torch

from oracle import make_golden as mg
from oracle import ref_shim, stft_oracle
from oracle import styler_oracle as so

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = ["tf_const_b2_l16", "tf_ragged_b3_l24", "free_b2_l12_tr50", "free_ragged_b3_l20_tr90"]


@pytest.mark.parametrize("case", SMALL)
def test_oracle_matches_reference_golden(case):
    gold = torch.load(os.path.join(GOLD, case + ".pt"))
    sd, batch = mg.build_case(case)
    assert mg.sd_checksum(sd) == gold["_weights_sha256"]
    args, kw = mg.call_kwargs(batch)
    with torch.no_grad():
        out = mg.flatten_outputs(so.styler_forward(sd, *args, **kw))
    for k, v in out.items():
        if v.dtype in (torch.bool, torch.int64):
            assert torch.equal(v, gold[k]), k                      # lengths / masks: bit exact
        else:
            err = (v - gold[k]).abs().max() / gold[k].abs().max().clamp_min(1e-12)
            assert err < 2e-5, (k, float(err))


def test_state_dict_surface():
    sd = so.make_state_dict(0)
    assert len(sd) == 328
    assert sum(v.numel() for k, v in sd.items()) == 29_999_984
    assert sd["style_modeling.style_encoder.text_encoder.position_enc"].shape == (1, 1001, 256)
    assert torch.equal(sd["decoder.position_enc"][0], so.sinusoid_table(1001))


def test_stft_oracle_matches_reference_golden():
    gold = torch.load(os.path.join(GOLD, "stft_b3_n6000.pt"))
    g = torch.Generator().manual_seed(gold["seed"])
    y = (torch.rand(*gold["shape"], generator=g) * 2 - 1) * 0.5
    for dense in (True, False):
        mel, energy = stft_oracle.mel_spectrogram(y, dense=dense)
        assert (mel - gold["mel"]).abs().max() < 1e-4
        assert ((energy - gold["energy"]).abs() / gold["energy"]).max() < 1e-5
    basis = torch.from_numpy(stft_oracle.slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0))
    assert torch.allclose(basis[0, :8], gold["mel_basis_row0"]) and torch.allclose(basis.sum(), gold["mel_basis_sum"])
    assert int((basis != 0).sum()) == 727 and int(basis.nonzero()[:, 1].max()) == 371


def test_slaney_mel_doc_example():
    """librosa.filters.mel documentation example: mel(22050, 2048)[0,:4] ~= [0, 0.0162, 0.0324, 0.029]."""
    m = stft_oracle.slaney_mel_basis(22050, 2048, 128)
    assert abs(m[0, 0]) < 1e-9 and abs(m[0, 1] - 0.016182) < 2e-4 and abs(m[0, 2] - 0.032364) < 2e-4
    from styler_b200.stft import mel_filterbank
    assert (mel_filterbank(22050, 1024, 80, 0.0, 8000.0) == stft_oracle.slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0)).all()


def test_length_regulator_semantics():
    """modules.py:396-423: truncation of float durations, un-cropped mel_len, crop/pad to max_len."""
    x = torch.arange(12, dtype=torch.float32).view(1, 4, 3)
    out, mel_len = so.length_regulator(x, torch.tensor([[2.6, 0.0, 1.2, 3.9]]), None)
    assert mel_len.tolist() == [6] and out.shape == (1, 6, 3)
    assert torch.equal(out[0, :, 0], torch.tensor([0., 0., 6., 9., 9., 9.]))
    out, mel_len = so.length_regulator(x, torch.tensor([[2, 0, 1, 3]]), 4)
    assert mel_len.tolist() == [6] and out.shape == (1, 4, 3)       # cropped output, un-cropped length
    out, _ = so.length_regulator(x, torch.tensor([[1, 0, 0, 0]]), 5)
    assert torch.equal(out[0, 1:], torch.zeros(4, 3))


def test_mel_calibrator_and_quantise_semantics():
    assert so.get_scale(10, 4) == [3, 3, 2, 2] and so.get_scale(4, 10) == [1, 1, 1, 1] + [0] * 6
    x = torch.arange(10, dtype=torch.float32).view(1, 10, 1)
    c = so.mel_calibrator(x, torch.tensor([10]), torch.tensor([4]))
    assert torch.allclose(c[0, :, 0], torch.tensor([1.0, 4.0, 6.5, 8.5]))
    e = so.mel_calibrator(x[:, :3], torch.tensor([3]), torch.tensor([7]))
    assert e[0, :, 0].tolist() == [0, 0, 0, 1, 1, 2, 2]
    q = so.quantize_index(torch.tensor([0.0, -1.0, 1.0, 0.5 / 255, 1.5 / 255, 2.5 / 255]))
    assert q.tolist() == [0, 0, 256, 1, 3, 3]                        # half-to-even


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the dev container")
def test_oracle_matches_live_reference_submodules():
    """Dev-container only: FFTBlock (BASELINE configs[1] geometry, reduced batch) and StylePredictor vs the reference."""
    _, ref_modules, ref_models, ref_layers, _ = ref_shim.load_reference_modules()
    sd = so.make_state_dict(3)
    blk = ref_layers.FFTBlock(256, 1024, 4, 64, 64).eval()
    p = "decoder.layer_stack.1."
    blk.load_state_dict({k[len(p):]: v for k, v in sd.items() if k.startswith(p)})
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 128, 256, generator=g)
    lens = torch.tensor([128, 64, 100, 77])
    mask = so.mask_from_lengths(lens, 128)
    with torch.no_grad():
        ref, _ = blk(x, mask=mask, slf_attn_mask=mask.unsqueeze(1).expand(-1, 128, -1))
        got = so.fft_block(sd, p, x, mask)
    assert (ref - got).abs().max() < 1e-5
    assert torch.equal(got[mask], torch.zeros_like(got[mask]))       # padded rows exactly zero
    pred = ref_modules.StylePredictor().eval()
    q = "style_modeling.pitch_predictor."
    pred.load_state_dict({k[len(q):]: v for k, v in sd.items() if k.startswith(q)})
    with torch.no_grad():
        assert (pred(x, mask) - so.style_predictor(sd, q, x, mask)).abs().max() < 1e-5
