#!/usr/bin/env python3
"""One small invocation of EVERY kernel of libspecloss.so, for `compute-sanitizer --tool memcheck|racecheck|initcheck`
(the shared-memory mirror-half exchange, the shape-backward tile and the mel scratch are the race-prone spots).
Sizes are small (sanitizers slow kernels down 10-100x) but cover edge frames (reflect padding), ragged T, B*F not a
multiple of the warps per CTA, and both GRAD / no-grad instantiations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window", num_mels=80, fmin=0,
              fmax=24000, log_base=None)
g = torch.Generator(device=dev).manual_seed(0)
for (b, t) in ((3, 5003), (2, 9600)):
    y = 0.1 * torch.randn(b, 1, t, device=dev, generator=g)
    x = (y + 0.05 * torch.randn(b, 1, t, device=dev, generator=g)).requires_grad_(True)
    stft = pkg.MultiResolutionSTFTLoss().to(dev)
    mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
    mel10 = pkg.MultiMelSpectrogramLoss(fs=24000, fft_sizes=[1024, 512], hop_sizes=[256, 128], win_lengths=[None, 400]).to(dev)   # generic-window kernels
    shape = pkg.MultiWindowShapeLoss().to(dev)
    shape2 = pkg.MultiWindowShapeLoss([300, 77]).to(dev)
    sc, mag = stft(x, y)
    total = sc + mag + mel(x, y) + mel10(x, y) + shape(x, y) + shape2(x, y)
    total.backward()
    with torch.no_grad():
        stft(x, y), mel(x, y)
    # explicit tensors and their backward
    x2 = x.detach().reshape(b, t).clone().requires_grad_(True)
    win = torch.hann_window(600, device=dev)
    a = pkg.stft(x2, 1024, 120, 600, win)
    a2 = pkg.stft(y.reshape(b, t), 1024, 120, 600, win)
    l = pkg.SpectralConvergenceLoss()(a, a2) + pkg.LogSTFTMagnitudeLoss()(a, a2)
    ms = pkg.MelSpectrogram(**{k: (v[0] if isinstance(v, list) else v) for k, v in
                               dict(fs=48000, fft_size=[2048], hop_size=[300], win_length=[None], num_mels=80, fmin=0, fmax=24000, log_base=None).items()}).to(dev)
    l = l + ms(x2).abs().mean()
    sp = pkg.spectrogram(x2, 300, torch.hann_window(600, device=dev), 1024, 120, 600, 1.0, False)
    (l + sp.mean()).backward()
    ml1 = pkg.MelL1(48000).to(dev)(x.detach(), y)
    torch.cuda.synchronize()
    print(f"B={b} T={t}: sc {float(sc):.6f} mag {float(mag):.6f} total {float(total):.6f} explicit {float(l):.6f} Mel_L1 {float(ml1):.6f} "
          f"|dx| {float(x.grad.norm()):.6f} {float(x2.grad.norm()):.6f}")
print("sanitize_step done")
