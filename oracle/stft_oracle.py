"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the TacotronSTFT mel front end.

Follows /root/reference/audio/stft.py:18-79 (STFT.__init__/transform), :120-160 (TacotronSTFT) and
audio/audio_processing.py:80-86 (dynamic_range_compression).  The reference evaluates the DFT as a
Conv1d with a [1026,1,1024] windowed Fourier basis (stft.py:26-47,64-69); this restatement offers both
that dense form (`dense=True`, used to pin the oracle against the live reference) and an rfft form
(mathematically identical, used for the big CPU-baseline runs).

Third-party arithmetic absent from /root/reference: ``librosa.filters.mel`` (pinned librosa==0.7.2 in
requirements.txt:6, call site audio/stft.py:128-129).  `slaney_mel_basis` restates its published
algorithm (Slaney auditory-toolbox mel scale, triangular filters, area normalisation) and is pinned by
the librosa documentation example (`mel(22050, 2048)[0,:4] ~= [0, 0.0162, 0.0324, 0.029]`); the
reference itself holds no golden vector for it ("parity unpinned" by the reference for this table).
"""
import numpy as np
import torch


# ---- Slaney mel scale (librosa.filters.mel, htk=False, norm=1) -------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    with np.errstate(divide="ignore"):
        log = _MIN_LOG_MEL + np.log(np.maximum(f, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * _F_SP
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def slaney_mel_basis(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """[n_mels, 1+n_fft//2] float32 filterbank (audio/stft.py:128-131)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fft_f = np.linspace(0.0, sr / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    dist = edges[:, None] - fft_f[None, :]          # edges minus bin frequency
    w = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        rising = -dist[i] / width[i]
        falling = dist[i + 2] / width[i + 1]
        w[i] = np.maximum(0.0, np.minimum(rising, falling))
    w *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None].astype(np.float32)
    return w


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True) (audio/stft.py:40)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def dense_fourier_basis(n_fft):
    """Windowed [2*(n_fft/2+1), n_fft] real/imag basis, float32 (audio/stft.py:26-47)."""
    cutoff = n_fft // 2 + 1
    fb = np.fft.fft(np.eye(n_fft))
    basis = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])])
    basis32 = torch.from_numpy(basis).float()
    win = torch.from_numpy(hann_periodic(n_fft)).float()
    return basis32 * win


def stft_magnitude(y, n_fft=1024, hop=256, dense=False):
    """y f32[B,N] -> magnitude f32[B, n_fft/2+1, 1+N//hop] (audio/stft.py:51-75)."""
    y = torch.as_tensor(y, dtype=torch.float32)
    pad = n_fft // 2
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect")  # stft.py:58-62
    cutoff = n_fft // 2 + 1
    if dense:
        basis = dense_fourier_basis(n_fft).unsqueeze(1)                       # [2*cutoff,1,n_fft]
        ft = torch.nn.functional.conv1d(yp, basis, stride=hop)                # stft.py:64-69
        re, im = ft[:, :cutoff], ft[:, cutoff:]
        return torch.sqrt(re * re + im * im)                                  # stft.py:75
    frames = yp.squeeze(1).unfold(1, n_fft, hop)                              # [B,F,n_fft]
    win = torch.from_numpy(hann_periodic(n_fft)).float()
    spec = torch.fft.rfft(frames * win, dim=-1)                               # [B,F,cutoff]
    return spec.abs().transpose(1, 2).contiguous()


def mel_spectrogram(y, n_fft=1024, hop=256, n_mels=80, sr=22050, fmin=0.0, fmax=8000.0,
                    dense=False, mel_basis=None):
    """TacotronSTFT.mel_spectrogram (audio/stft.py:141-160): -> (mel f32[B,80,F], energy f32[B,F])."""
    y = torch.as_tensor(y, dtype=torch.float32)
    assert float(y.min()) >= -1.0 and float(y.max()) <= 1.0                   # stft.py:151-152
    mag = stft_magnitude(y, n_fft, hop, dense=dense)
    if mel_basis is None:
        mel_basis = torch.from_numpy(slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax))
    mel = torch.matmul(mel_basis, mag)                                        # stft.py:156
    mel = torch.log(torch.clamp(mel, min=1e-5))                               # audio_processing.py:86
    energy = torch.norm(mag, dim=1)                                           # stft.py:158
    return mel, energy
