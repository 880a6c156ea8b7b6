"""Slaney mel filterbank, host side (construction time only).

The reference builds its `melmat` buffer with `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`
(losses/mel_loss.py:54-61; librosa==0.8.1 per requirements.txt:26, defaults htk=False,
norm='slaney', dtype=float32).  librosa is not a dependency of this package; the published
algorithm is restated here: triangular filters whose corner frequencies are equally spaced on
the Slaney mel scale (linear below 1 kHz at 200/3 Hz per mel, logarithmic above with
ln(6.4)/27 per mel), each scaled by 2 / (f_hi - f_lo), evaluated in float64, stored as float32.
"""
from __future__ import annotations

import numpy as np

_LIN_HZ_PER_MEL = 200.0 / 3.0
_KNEE_HZ = 1000.0
_KNEE_MEL = _KNEE_HZ / _LIN_HZ_PER_MEL
_LOG_STEP = np.log(6.4) / 27.0


def hz_to_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    safe = np.maximum(hz, _KNEE_HZ)          # keeps log() away from zero in the branch not taken
    return np.where(hz >= _KNEE_HZ, _KNEE_MEL + np.log(safe / _KNEE_HZ) / _LOG_STEP, hz / _LIN_HZ_PER_MEL)


def mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= _KNEE_MEL, _KNEE_HZ * np.exp(_LOG_STEP * (mel - _KNEE_MEL)), _LIN_HZ_PER_MEL * mel)


def mel_filterbank(sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """Returns the (n_mels, n_fft // 2 + 1) float32 matrix librosa.filters.mel would return."""
    if fmax is None:
        fmax = float(sr) / 2.0
    freqs = np.linspace(0.0, float(sr) / 2.0, n_fft // 2 + 1)
    corners = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    gaps = corners[1:] - corners[:-1]
    offs = corners[:, None] - freqs[None, :]                  # (n_mels + 2, K)
    up = -offs[:-2] / gaps[:-1, None]
    down = offs[2:] / gaps[1:, None]
    tri = np.clip(np.minimum(up, down), 0.0, None)
    tri *= (2.0 / (corners[2:] - corners[:-2]))[:, None]
    return tri.astype(np.float32)


def htk_mel_filterbank(sample_rate: int, n_fft: int, n_mels: int = 128, f_min: float = 0.0, f_max=None):
    """(n_fft // 2 + 1, n_mels) fp32 triangular filterbank of torchaudio.transforms.MelSpectrogram's defaults
    (mel_scale="htk", norm=None) -- the bank behind the reference's `Mel_L1` metric (mel_spectrogram.py:36-44).
    torchaudio (third party, pinned 2.1.1 in the reference's requirements.txt) computes it in FP32 with torch ops:
        all_freqs = linspace(0, sr // 2, n_freqs);  m = linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2)
        f = 700 (10^(m / 2595) - 1);  fb = max(0, min(-(f[:-2] - all_freqs) / (f[1:-1] - f[:-2]), (f[2:] - all_freqs) / (f[2:] - f[1:-1])))
    The same op sequence on fp32 torch tensors is restated here so that the weights agree with torchaudio's bit for bit
    (an fp64 evaluation differs from it by ~1e-5 relative: f ~ 1e4 Hz carries 1e-3 Hz of fp32 rounding)."""
    import math

    import torch

    f_max = float(sample_rate // 2) if f_max is None else float(f_max)
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (float(f_min) / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).numpy()
