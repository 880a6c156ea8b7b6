#!/usr/bin/env python3
"""Golden vectors of the reference's Mel_L1 metric (mel_spectrogram.py:36-44): produced by the code the reference calls,
torchaudio.transforms.MelSpectrogram(48000) + nn.L1Loss, on CPU in fp32 and (module.double()) fp64.
    python tests/golden/make_golden_mel_l1.py      # rewrites tests/golden/mel_l1_*.npz
Runs wherever torchaudio is installed (it is third party for the reference: requirements.txt pins 2.1.1; here 2.11)."""
import os
import sys
import warnings

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def ref_mel_l1(pred, target, dtype):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                       # "at least one mel filterbank has all zero values"
        mel_spectrogram = torchaudio.transforms.MelSpectrogram(48000).to(dtype)
    mae = torch.nn.L1Loss()
    mp, mt = mel_spectrogram(pred.to(dtype)), mel_spectrogram(target.to(dtype))
    return float(mae(mp, mt)), mp.numpy(), mt.numpy()


def main():
    g = torch.Generator().manual_seed(2024)
    cases = {}
    clean = 0.1 * torch.randn(1, 9600, generator=g)
    cases["mel_l1_gauss_c1_t9600"] = (clean + 0.05 * torch.randn(1, 9600, generator=g), clean)
    clean = torch.rand(2, 5003, generator=g) * 2 - 1                        # ragged length, 2 channels, uniform
    cases["mel_l1_uniform_c2_t5003"] = (torch.rand(2, 5003, generator=g) * 2 - 1, clean)
    t = torch.arange(7200) / 48000.0                                        # tonal + silence tail: sparse spectra, exact zeros
    tone = 0.5 * torch.sin(2 * np.pi * 440.0 * t) + 0.2 * torch.sin(2 * np.pi * 9000.0 * t)
    tone[6000:] = 0.0
    cases["mel_l1_tone_silence_c1_t7200"] = ((tone * 0.8).reshape(1, -1), tone.reshape(1, -1))
    x = 0.1 * torch.randn(1, 201, generator=g)                              # minimal length: T = n_fft / 2 + 1
    cases["mel_l1_minlen_c1_t201"] = (x, 0.5 * x)
    for name, (pred, target) in cases.items():
        l32, mp32, mt32 = ref_mel_l1(pred, target, torch.float32)
        l64, mp64, mt64 = ref_mel_l1(pred, target, torch.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pred=pred.numpy(), target=target.numpy(),
                            loss32=np.float64(l32), loss64=np.float64(l64), mel_pred64=mp64.astype(np.float32),
                            mel_target64=mt64.astype(np.float32), torchaudio=np.array(torchaudio.__version__))
        print(f"{name}: loss32 {l32:.9g} loss64 {l64:.12g} mel shape {mp64.shape}")


if __name__ == "__main__":
    main()
