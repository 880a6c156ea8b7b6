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
