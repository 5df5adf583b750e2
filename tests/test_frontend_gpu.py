"""GPU parity of the fused preprocessing front end (styler_b200.frontend.ReferenceFrontEnd, through
styler_stft_mel_ex_fwd / styler_f0_norm_fwd) against the golden values of the unmodified reference functions."""
import os

import pytest
import torch

from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "frontend_b3.pt")


def test_front_end_matches_reference_golden(cuda):
    from styler_b200.frontend import ReferenceFrontEnd
    gold = torch.load(GOLD)
    wav, n_samples, f0, frames = fo.make_case(seed=gold["seed"])
    fe = ReferenceFrontEnd().to(cuda)
    out = fe(wav, n_samples, f0, norm=True)
    torch.cuda.synchronize()
    B = wav.shape[0]
    assert torch.equal(out["mel_len"].cpu(), frames) and not out["clipt"].any()
    for b in range(B):
        fr = int(frames[b])
        # the reference transforms every utterance alone (audio/tools.py:37-55): the kernel reflects each row of the
        # zero-padded batch around its OWN end (n_samples[b]), so EVERY frame of every utterance must match, tail included
        mel = out["mel_target"][b, :fr].cpu()
        assert gold["mel"][b].shape[1] == fr
        assert (mel.t() - gold["mel"][b]).abs().max().item() < 1e-3
        en = out["energy"][b, :fr].cpu()
        assert ((en - gold["energy"][b]).abs() / gold["energy"][b].abs().clamp_min(1e-3)).max().item() < 1e-4
        assert (out["e_input"][b, :fr].cpu() - gold["e_input"][b]).abs().max().item() < 1e-5
        assert (out["mel_target"][b, fr:] == 0).all() and (out["p_norm"][b, fr:] == 0).all()
        assert (out["energy"][b, fr:] == 0).all() and (out["e_input"][b, fr:] == 0).all()
        p = out["p_norm"][b, :fr].cpu().double()
        g = gold["f0_norm"][b]
        voiced = g > -1e9
        assert torch.equal(p[~voiced].float(), g[~voiced].float())
        assert (p[voiced] - g[voiced]).abs().max().item() < 1e-6 if voiced.any() else True
    # single-utterance calls reproduce the reference exactly on every frame (no batch padding)
    for b in range(B):
        n = int(n_samples[b])
        mel, energy, e_in, clipt = fe.mel_energy_from_wav(wav[b:b + 1, :n], norm=True)
        assert (mel[0].cpu() - gold["mel"][b]).abs().max().item() < 1e-3
        assert ((energy[0].cpu() - gold["energy"][b]).abs() / gold["energy"][b].abs().clamp_min(1e-3)).max().item() < 1e-4
        assert (e_in[0].cpu() - gold["e_input"][b]).abs().max().item() < 1e-5
        _, _, _, c2 = fe.mel_energy_from_wav(wav[b:b + 1, :n] / 16384.0, norm=False)
        assert bool(c2[0]) == gold["clipt"][b]
    quiet = fe.mel_energy_from_wav(wav[:1, :8000] / 65536.0, norm=False)[3]
    assert not bool(quiet[0])


def test_front_end_feeds_the_model(cuda):
    """End to end: waveform -> front end -> STYLER.forward (teacher-free) runs and yields finite mels."""
    from styler_b200 import STYLER
    from styler_b200.frontend import ReferenceFrontEnd
    from oracle import styler_oracle as so
    wav, n_samples, f0, frames = fo.make_case(seed=1)
    fe = ReferenceFrontEnd().to(cuda)
    ref = fe(wav, n_samples, f0)
    model = STYLER(precision="tf32")
    model.load_state_dict(so.make_state_dict(0))
    model = model.to(cuda).eval()
    B, L = wav.shape[0], 20
    g = torch.Generator().manual_seed(0)
    src = torch.randint(1, 152, (B, L), generator=g).to(cuda)
    src_len = torch.full((B,), L, dtype=torch.int64, device=cuda)
    d_target = torch.full((B, L), 3, dtype=torch.int64, device=cuda)
    out = model(src, ref["mel_target"], ref["mel_target"], ref["p_norm"].clamp(0, 1), ref["e_input"], src_len, ref["mel_len"],
                d_target=d_target, p_target=torch.zeros(B, 3 * L, device=cuda), e_target=torch.zeros(B, 3 * L, device=cuda),
                max_src_len=L, max_mel_len=3 * L, speaker_embed=torch.randn(B, 512, generator=g).to(cuda))
    torch.cuda.synchronize()
    assert torch.isfinite(out[1][0]).all() and out[1][0].shape == (B, 3 * L, 80)
