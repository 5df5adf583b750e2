"""CPU tests of the preprocessing front-end row: the oracle restatement against the golden values
produced by the unmodified reference functions (oracle/make_golden_frontend.py)."""
import os

import numpy as np
import torch

from oracle import frontend_oracle as fo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frontend_b3.pt")


def test_frontend_oracle_matches_reference_golden():
    gold = torch.load(GOLD)
    wav, n_samples, f0, frames = fo.make_case(seed=gold["seed"])
    for b in range(wav.shape[0]):
        n, fr = int(n_samples[b]), int(frames[b])
        mel, energy, clipt = fo.get_mel_from_wav(wav[b, :n], norm=True)
        assert not clipt
        assert (mel - gold["mel"][b]).abs().max().item() < 2e-5
        assert ((energy - gold["energy"][b]).abs() / gold["energy"][b].abs().clamp_min(1e-6)).max().item() < 2e-5
        assert np.allclose(fo.energy_rescaling(energy.numpy()), gold["e_input"][b].numpy(), atol=2e-6)
        _, _, c2 = fo.get_mel_from_wav(wav[b, :n] / 16384.0, norm=False)
        assert c2 == gold["clipt"][b]
        assert np.array_equal(fo.f0_normalization(f0[b, :fr].numpy()), gold["f0_norm"][b].numpy())


def test_f0_normalization_edge_cases():
    f = np.array([5.0, -1e10, 5.5, 4.5, -1e10], dtype=np.float32)
    out = fo.f0_normalization(f)
    assert out[1] == -1e10 and out[4] == -1e10 and 0.0 <= out[0] <= 1.0
    assert (fo.f0_normalization(np.full(4, -1e10, dtype=np.float32)) == 0).all()      # no voiced frame -> Warning -> zeros
    assert (fo.f0_normalization(np.array([5.0, 5.0, -1e10], dtype=np.float32)) == 0).all()   # std == 0 -> zeros
