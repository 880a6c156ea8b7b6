"""Import the reference's loss modules VERBATIM from /root/reference.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference is mounted (the authoring container); the GPU box does
not have it, so nothing under `-m gpu`, smoke() or bench.py may call this.  It is used by
tests/golden/make_golden.py to produce the committed fixtures and by the (skippable) CPU
test that compares the oracle restatement with the live reference.

losses/mel_loss.py does `import librosa` (mel_loss.py:14); librosa is not installed and
there is no network, so a module exposing only `librosa.filters.mel` is injected into
sys.modules.  Its arithmetic is oracle.spectral_oracle.slaney_mel_filterbank (a restatement
of librosa 0.8.1), cross-checked against torchaudio in tests/test_melfb.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SPECLOSS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "losses", "stft_loss.py"))


def _install_librosa_shim():
    if "librosa" in sys.modules and not getattr(sys.modules["librosa"], "_specloss_shim", False):
        return
    from oracle.spectral_oracle import slaney_mel_filterbank

    lib = types.ModuleType("librosa")
    lib._specloss_shim = True
    filt = types.ModuleType("librosa.filters")

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_unused):
        fmax = float(sr) / 2 if fmax is None else fmax
        return slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)

    filt.mel = mel
    lib.filters = filt
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = filt


def _load(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_losses():
    """Returns (stft_loss_module, mel_loss_module) executed from the reference's files."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_librosa_shim()
    return _load("_ref_stft_loss", "losses/stft_loss.py"), _load("_ref_mel_loss", "losses/mel_loss.py")


def load_fixture_pair(idx: int = 1):
    """(y_hat, y) = (noise{idx}.wav resampled 24->48 kHz, clean{idx}.wav / 32768), shapes (1,1,T)."""
    import warnings

    import numpy as np
    import torch
    import torchaudio
    from scipy.io import wavfile

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sr_c, clean = wavfile.read(os.path.join(REFERENCE_ROOT, "notebook_files", f"clean{idx}.wav"))
        sr_n, noise = wavfile.read(os.path.join(REFERENCE_ROOT, "notebook_files", f"noise{idx}.wav"))
    y = torch.from_numpy(clean.astype(np.float32) / 32768.0)
    n = torch.from_numpy(noise.astype(np.float32))
    if sr_n != sr_c:
        n = torchaudio.functional.resample(n, sr_n, sr_c)
    t = min(y.numel(), n.numel())
    return n[:t].reshape(1, 1, t).contiguous(), y[:t].reshape(1, 1, t).contiguous()
