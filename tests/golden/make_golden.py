#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference modules (authoring container only).

Run from the repo root:   python tests/golden/make_golden.py

Each fixture holds the inputs (so the GPU box never needs /root/reference or the same RNG),
the ctor kwargs, and what the reference's own MultiResolutionSTFTLoss / MultiMelSpectrogramLoss
(losses/stft_loss.py:120-170, losses/mel_loss.py:97-156, imported verbatim through
oracle/ref_loader.py) return on CPU in fp32 and, with .double(), in fp64:
    loss32 / loss64 : (sc, mag, mel)
    grad32 / grad64 : d(sc + mag + mel)/d y_hat, stored as float32, shape of y_hat
The fp64 run is the conditioning yardstick of SURVEY.md section 7 ("ill-conditioned gradients").
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.spectral_oracle import synth_pair  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

MEL48 = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
             num_mels=80, fmin=0, fmax=24000, log_base=None)            # config/denoise/symAD_vctk_48000_hop300.yaml:88-97
MEL24_LIBRITTS = dict(fs=24000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[2048], window="hann_window",
                      num_mels=80, fmin=0, fmax=12000, log_base=None)   # config/autoencoder/symAD_libritts_24000_hop300.yaml:85-94
MEL24_OVER_NYQ = dict(fs=24000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
                      num_mels=80, fmin=0, fmax=24000, log_base=None)   # config/denoise/symAD_24Mel.yaml:87-96 (fmax > fs/2)
MEL_DEFAULT = dict()                                                     # ctor defaults, losses/mel_loss.py:100-115 (3 res, log10)
MEL_LOG2 = dict(fs=22050, fft_sizes=[1024, 512], hop_sizes=[256, 128], win_lengths=[None, 400], log_base=2.0)
STFT_DEFAULT = dict()                                                    # losses/stft_loss.py:125-128


def run_reference(stft_mod, mel_mod, y_hat, y, stft_kw, mel_kw, dtype):
    x = y_hat.detach().to(dtype).clone().requires_grad_(True)
    t = y.detach().to(dtype)
    total = 0.0
    vals = [0.0, 0.0, 0.0]
    if stft_kw is not None:
        crit = stft_mod.MultiResolutionSTFTLoss(**stft_kw).to(dtype)
        sc, mag = crit(x, t)
        total = total + sc + mag
        vals[0], vals[1] = float(sc.detach()), float(mag.detach())
    if mel_kw is not None:
        crit = mel_mod.MultiMelSpectrogramLoss(**mel_kw).to(dtype)
        mel = crit(x, t)
        total = total + mel
        vals[2] = float(mel.detach())
    (g,) = torch.autograd.grad(total, x)
    return np.array(vals, dtype=np.float64), g.to(torch.float32).numpy()


def make(name, y_hat, y, stft_kw, mel_kw, stft_mod, mel_mod, extra=None):
    l32, g32 = run_reference(stft_mod, mel_mod, y_hat, y, stft_kw, mel_kw, torch.float32)
    l64, g64 = run_reference(stft_mod, mel_mod, y_hat, y, stft_kw, mel_kw, torch.float64)
    meta = dict(stft_kwargs=stft_kw, mel_kwargs=mel_kw, torch=torch.__version__)
    arrays = dict(loss32=l32, loss64=l64, grad32=g32, grad64=g64, meta=np.array(json.dumps(meta)))
    arrays.update(extra if extra is not None else dict(y_hat=y_hat.numpy(), y=y.numpy()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    rel = np.linalg.norm(g32.astype(np.float64) - g64) / np.linalg.norm(g64)
    print(f"{name:24s} loss32={l32} ref32-vs-ref64 grad rel-L2={rel:.2e}")


def main():
    torch.set_num_threads(os.cpu_count())
    stft_mod, mel_mod = ref_loader.load_reference_losses()

    # 1. Gaussian recipe, seed 1234, config-2 parameters at a size the oracle finishes in seconds
    yh, y = synth_pair(2, 4800, seed=1234)
    make("gauss_b2_t4800", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod)
    yh, y = synth_pair(2, 16000, seed=1234)
    make("gauss_b2_t16000", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod)

    # 2. ragged length (T % hop != 0 for every hop), 2-D (B, T) input
    yh, y = synth_pair(3, 5003, seed=7)
    make("ragged_b3_t5003_2d", yh[:, 0].contiguous(), y[:, 0].contiguous(), STFT_DEFAULT, MEL48, stft_mod, mel_mod)

    # 3. minimal length T = n_fft/2 + 1 for the largest FFT
    yh, y = synth_pair(1, 1025, seed=11)
    make("minlen_b1_t1025", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod)

    # 4. independent uniform[-1,1] pair, multi-channel (B, C, T)
    g = torch.Generator().manual_seed(99)
    yh = torch.rand(2, 2, 6000, generator=g) * 2 - 1
    y = torch.rand(2, 2, 6000, generator=g) * 2 - 1
    make("uniform_b2_c2_t6000", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod)

    # 5. digital silence: a leading block of exact zeros in prediction and/or target (clamp gates)
    yh, y = synth_pair(3, 6000, seed=5)
    yh = yh.clone(); y = y.clone()
    yh[0, :, :3500] = 0.0
    y[1, :, :3500] = 0.0
    yh[2, :, :3000] = 0.0; y[2, :, :3000] = 0.0
    make("silence_b3_t6000", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod)

    # 6. mel variants: ctor defaults (3 resolutions, log10, fmin 80/fmax 7600 @22.05k), log2, 24 kHz configs
    yh, y = synth_pair(2, 8000, seed=21)
    make("mel_default_b2_t8000", yh, y, None, MEL_DEFAULT, stft_mod, mel_mod)
    make("mel_log2_b2_t8000", yh, y, None, MEL_LOG2, stft_mod, mel_mod)
    make("mel24_libritts_b2_t8000", yh, y, None, MEL24_LIBRITTS, stft_mod, mel_mod)
    make("mel24_overnyq_b2_t8000", yh, y, None, MEL24_OVER_NYQ, stft_mod, mel_mod)

    # 7. STFT-only with non-default resolutions from a single STFTLoss-like setup
    make("stft_only_b2_t8000", yh, y, dict(fft_sizes=[512, 1024], hop_sizes=[128, 256], win_lengths=[512, 1024]),
         None, stft_mod, mel_mod)

    # 8. config 1: real audio, clean1.wav vs noise1.wav (24 kHz -> 48 kHz), full length
    import warnings
    from scipy.io import wavfile
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, clean = wavfile.read(os.path.join(ref_loader.REFERENCE_ROOT, "notebook_files", "clean1.wav"))
        sr_n, noise = wavfile.read(os.path.join(ref_loader.REFERENCE_ROOT, "notebook_files", "noise1.wav"))
    yh, y = ref_loader.load_fixture_pair(1)
    make("c1_clean1_noise1", yh, y, STFT_DEFAULT, MEL48, stft_mod, mel_mod,
         extra=dict(clean_int16=clean[:y.numel()], y_hat=yh.numpy()))   # resampled noise stored as is: conv order is thread-count dependent


if __name__ == "__main__":
    main()
