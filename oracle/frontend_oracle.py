"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-utterance preprocessing front end, the step
before the mel-synthesis path.  Only tests/, smoke() and bench.py's CPU leg may import it.

Follows the reference:
  * get_mel_from_wav          audio/tools.py:37-55   (x / max_wav_value, or clamp to [-1,1] with the `clipt` flag, then
                                                      TacotronSTFT.mel_spectrogram -> (mel [80,F], energy [F]))
  * speaker_normalization     utils.py:387-398       (float64; voiced = f0 > -1e10; (f0-mean)/std/4 -> clip[-1,1] -> (x+1)/2)
  * f0_normalization          utils.py:401-409       (any numpy Warning -> zeros_like(f0))
  * energy_rescaling          utils.py:412-416       ((e - energy_min)/(energy_max - energy_min) clipped to [0,1])
  * quantize_1D_torch index   utils.py:417-429       (in styler_oracle.quantize_index)
Pinned against the unmodified reference functions by tests/test_frontend_cpu.py (dev container) and the golden file
tests/golden/frontend_b3.pt written by oracle/make_golden_frontend.py.
"""
import warnings

import numpy as np
import torch

from . import stft_oracle

MAX_WAV_VALUE = 32768.0            # hparams.py:35
ENERGY_MIN, ENERGY_MAX = 0.1, 525.43   # hparams.py:25-26


def get_mel_from_wav(audio, norm=True):
    """audio: 1-D float tensor (int16 scale when norm=True) -> (mel [80,F], energy [F], clipt)."""
    clipt = False
    audio_norm = audio / MAX_WAV_VALUE if norm else audio
    audio_norm = audio_norm.unsqueeze(0)
    if not norm:
        pre_min = torch.min(audio_norm)
        audio_norm = torch.clamp(audio_norm, -1, 1)
        if pre_min != torch.min(audio_norm):
            clipt = True
    mel, energy = stft_oracle.mel_spectrogram(audio_norm)
    return mel.squeeze(0), energy.squeeze(0), clipt


def speaker_normalization(f0):
    f0 = np.asarray(f0).astype(float).copy()
    voiced = f0 > -1e10
    mean_f0, std_f0 = np.mean(f0[voiced]), np.std(f0[voiced])
    f0[voiced] = (f0[voiced] - mean_f0) / std_f0 / 4.0
    f0[voiced] = np.clip(f0[voiced], -1, 1)
    f0[voiced] = (f0[voiced] + 1) / 2.0
    return f0


def f0_normalization(f0):
    with warnings.catch_warnings():
        warnings.filterwarnings("error")
        try:
            return speaker_normalization(f0)
        except Warning:
            return np.zeros_like(np.asarray(f0))


def energy_rescaling(energy):
    e = (energy - ENERGY_MIN) / (ENERGY_MAX - ENERGY_MIN)
    return np.clip(e, 0, 1)


def make_case(seed=0, B=3, n_samples=(30000, 22050, 41000)):
    """Seeded synthetic front-end inputs: int16-scale waveforms of different lengths (padded with zeros to the longest,
    as a batched loader would), and log-f0 contours with unvoiced stretches (-1e10), one utterance fully unvoiced when B>=3."""
    g = torch.Generator().manual_seed(4000 + seed)
    N = max(n_samples[:B])
    wav = torch.zeros(B, N)
    for b in range(B):
        n = n_samples[b]
        t = torch.arange(n) / 22050.0
        wav[b, :n] = (0.4 * torch.sin(2 * np.pi * (110.0 * (b + 1)) * t) + 0.2 * torch.randn(n, generator=g)) * 20000.0
    wav = wav.round()
    frames = [1 + n // 256 for n in n_samples[:B]]
    T = max(frames)
    f0 = torch.zeros(B, T)
    for b in range(B):
        c = 5.0 + 0.3 * torch.randn(frames[b], generator=g)
        unv = torch.rand(frames[b], generator=g) < (1.1 if b == 2 else 0.3)   # utterance 2: everything unvoiced
        c[unv] = -1e10
        f0[b, :frames[b]] = c
    return wav, torch.tensor(n_samples[:B]), f0, torch.tensor(frames, dtype=torch.int64)
