#!/usr/bin/env python3
"""Short profiling driver: a few fwd+bwd steps of BASELINE configs[1] (16 x 1 s @ 48 kHz) through the
drop-in modules.  Run plain first, then under ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402

B, T = int(os.environ.get("PROF_B", 16)), int(os.environ.get("PROF_T", 48000))
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
dev = torch.device("cuda:0")
stft = pkg.MultiResolutionSTFTLoss().to(dev)
mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
y = 0.1 * torch.randn(B, 1, T, device=dev, generator=g)
x = (y + 0.05 * torch.randn(B, 1, T, device=dev, generator=g)).requires_grad_(True)
for _ in range(int(os.environ.get("PROF_STEPS", 3))):
    x.grad = None
    ml = mel(x, y)
    sc, mag = stft(x, y)
    (sc + mag + ml).backward()
torch.cuda.synchronize()
print("ok", float(sc.detach()), float(mag.detach()), float(ml.detach()))
