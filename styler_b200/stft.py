"""Drop-in `TacotronSTFT` (audio/stft.py:120-160 of the reference): same constructor arguments and
`mel_spectrogram(y) -> (mel [B,n_mels,F], energy [B,F])`, computed by one fused CUDA kernel (reflect pad, Hann,
1024-point real FFT, magnitude, mel projection, log compression, frame energy).

The reference hard-codes `.cuda()` / `.cpu()` around its conv (stft.py:66-69) and therefore returns CPU tensors;
pass `return_on_cpu=True` to mirror that, the default keeps results on the device.  The mel filterbank is the
Slaney-scale, area-normalised bank `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` (librosa 0.7.2) restated
here (the dependency is not part of the reference tree)."""
import numpy as np
import torch
import torch.nn as nn

from . import ops


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, brk = 200.0 / 3.0, 1000.0
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        hi = brk / f_sp + np.log(np.maximum(f, 1e-30) / brk) / logstep
    return np.where(f >= brk, hi, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, brk = 200.0 / 3.0, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= brk / f_sp, brk * np.exp(logstep * (m - brk / f_sp)), m * f_sp)


def mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    fmax = sr / 2.0 if fmax is None else fmax
    bins = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    span = np.diff(edges)
    delta = edges[:, None] - bins[None, :]
    fb = np.zeros((n_mels, bins.size), dtype=np.float32)
    for i in range(n_mels):
        fb[i] = np.maximum(0.0, np.minimum(-delta[i] / span[i], delta[i + 2] / span[i + 1]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None].astype(np.float32)
    return fb


class TacotronSTFT(nn.Module):
    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0, return_on_cpu=False, validate=False):
        super().__init__()
        if (filter_length, hop_length, win_length) != (1024, 256, 1024):
            raise NotImplementedError("the STFT kernel is specialised for n_fft=1024, hop=256, win=1024 (hparams.py:31-33)")
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.return_on_cpu = return_on_cpu
        self.validate = validate   # the reference asserts y in [-1,1] (stft.py:151-152); that forces a host sync
        self.register_buffer("mel_basis", torch.from_numpy(mel_filterbank(sampling_rate, filter_length, n_mel_channels,
                                                                          mel_fmin, mel_fmax)).float())

    def mel_spectrogram(self, y):
        if self.mel_basis.device.type != "cuda":
            raise RuntimeError("styler_b200.TacotronSTFT runs only on CUDA; call .cuda() (no CPU fallback)")
        y = y.to(self.mel_basis.device, torch.float32)
        if self.validate:
            assert float(y.min()) >= -1 and float(y.max()) <= 1
        mel, energy = ops.stft_mel(y, self.mel_basis)
        if self.return_on_cpu:
            return mel.cpu(), energy.cpu()
        return mel, energy
