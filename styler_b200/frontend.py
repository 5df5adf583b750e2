"""Fused preprocessing front end: waveform (+ log-f0 contour) -> the reference-audio inputs of `STYLER.forward`
(`mel_target` [B,F,80], `p_norm` [B,F], `e_input` [B,F], `mel_len`), batched and on the device.

Replaces, for a batch, the per-utterance CPU<->GPU chain of the reference: `audio.tools.get_mel_from_wav`
(audio/tools.py:37-55: scaling or clamp + `clipt`, then TacotronSTFT), `utils.energy_rescaling` (utils.py:412-416),
`utils.f0_normalization` / `speaker_normalization` (utils.py:387-409), the `.T` / `.npy` round trip of data/*.py and the
`pad_1D` / `pad_2D` collation (dataset.py:160-166).  f0 EXTRACTION (pyworld / sptk, dataset.py:40-46) stays outside: it
is third-party CPU code; this module takes the log-f0 contour it produces (unvoiced frames marked -1e10).

One kernel launch produces mel (already frame-major), raw energy and rescaled energy; one more normalises f0.
There is no CPU path."""
import torch
import torch.nn as nn

from . import hparams as hp
from . import ops
from .stft import mel_filterbank


class ReferenceFrontEnd(nn.Module):
    def __init__(self, n_mel_channels=80, sampling_rate=22050, mel_fmin=0.0, mel_fmax=8000.0, max_wav_value=32768.0,
                 energy_min=None, energy_max=None):
        super().__init__()
        self.max_wav_value = float(max_wav_value)
        self.energy_min = float(getattr(hp, "energy_min", 0.1) if energy_min is None else energy_min)
        self.energy_max = float(getattr(hp, "energy_max", 525.43) if energy_max is None else energy_max)
        self.register_buffer("mel_basis", torch.from_numpy(mel_filterbank(sampling_rate, 1024, n_mel_channels, mel_fmin,
                                                                          mel_fmax)).float())

    def _dev(self):
        if self.mel_basis.device.type != "cuda":
            raise RuntimeError("styler_b200.ReferenceFrontEnd runs only on CUDA; call .cuda() (no CPU fallback)")
        return self.mel_basis.device

    @torch.no_grad()
    def mel_energy_from_wav(self, wav, norm=True, frame_major=False, n_samples=None):
        """Batched get_mel_from_wav: wav [B,N] (int16 scale if norm) -> (mel [B,80,F] or [B,F,80], energy [B,F],
        e_input [B,F], clipt bool [B]).  With `n_samples` (int64 [B]) every row of the zero-padded batch is transformed as
        the reference transforms it alone (audio/tools.py:37-55): reflected around its own end, 1 + n_samples[b] // 256
        frames, zeros beyond."""
        dev = self._dev()
        wav = wav.to(dev, torch.float32)
        if n_samples is not None:
            n_samples = torch.as_tensor(n_samples, dtype=torch.int64).to(dev)
        mel, energy, flag, e_in = ops.stft_mel_ex(wav, self.mel_basis, in_scale=(1.0 / self.max_wav_value) if norm else 1.0,
                                                  clamp=not norm, frame_major=frame_major,
                                                  energy_range=(self.energy_min, self.energy_max), n_samples=n_samples)
        clipt = flag.bool() if flag is not None else torch.zeros(wav.shape[0], dtype=torch.bool, device=dev)
        return mel, energy, e_in, clipt

    @torch.no_grad()
    def f0_normalization(self, logf0, lens=None):
        """Batched utils.f0_normalization over padded log-f0 contours [B,T] (unvoiced = -1e10)."""
        dev = self._dev()
        lens = lens.to(dev, torch.int64) if lens is not None else None
        return ops.f0_norm(logf0.to(dev, torch.float32), lens)

    @torch.no_grad()
    def forward(self, wav, n_samples, logf0, norm=True):
        """-> dict(mel_target [B,F,80], p_norm [B,F], e_input [B,F], energy [B,F], mel_len int64 [B], clipt bool [B])
        with frames >= mel_len[b] zeroed like pad_1D / pad_2D do; mel_len[b] = 1 + n_samples[b] // 256."""
        dev = self._dev()
        n_samples = torch.as_tensor(n_samples, dtype=torch.int64)
        mel_len = (1 + n_samples // 256).to(dev)
        if int(n_samples.min()) <= 512 or int(n_samples.max()) > wav.shape[1]:
            raise ValueError("n_samples must lie in (n_fft/2, N]: reflect padding needs more than 512 samples per utterance")
        # per-utterance reflect padding / frame count / zero tail (the collation padding of dataset.py:160-166) happen in the kernel
        mel, energy, e_in, clipt = self.mel_energy_from_wav(wav, norm=norm, frame_major=True, n_samples=n_samples)
        F = mel.shape[1]
        assert logf0.shape[1] <= F
        p_norm = torch.zeros(mel.shape[0], F, device=dev, dtype=torch.float32)
        p_norm[:, :logf0.shape[1]] = self.f0_normalization(logf0, mel_len.clamp(max=logf0.shape[1]))   # zero beyond lens[b]
        return dict(mel_target=mel, p_norm=p_norm, e_input=e_in, energy=energy, mel_len=mel_len, clipt=clipt)
