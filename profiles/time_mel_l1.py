#!/usr/bin/env python3
"""Mel_L1 metric timing on one B200: this repo's 400-point power-mel kernel vs
torchaudio.transforms.MelSpectrogram(48000) + nn.L1Loss (what mel_spectrogram.py:36-44 runs) on the same GPU.
Algorithmic bytes: 8 per sample (both signals read once)."""
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
crit = pkg.MelL1(48000).to(dev)
try:
    import torchaudio
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ta = torchaudio.transforms.MelSpectrogram(48000).to(dev)
except ImportError:
    ta = None
peak = 6543.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


for b, t in ((16, 48000), (32, 192000), (256, 192000)):
    g = torch.Generator(device=dev).manual_seed(0)
    y = 0.1 * torch.randn(b, 1, t, device=dev, generator=g)
    x = y + 0.05 * torch.randn(b, 1, t, device=dev, generator=g)
    us = timeit(lambda: crit(x, y))
    gbs = 8.0 * b * t / (us * 1e-6) / 1e9
    line = f"{b} x {t / 48000:g} s: Mel_L1 {us:8.1f} us = {gbs:7.1f} GB/s algorithmic = {gbs / peak:.3f} of HBM peak ({b * t / 48000 / (us * 1e-6):.3g} audio-s/s)"
    if ta is not None:
        us_ta = timeit(lambda: torch.nn.functional.l1_loss(ta(x), ta(y)))
        line += f"; torchaudio on this GPU {us_ta:8.1f} us ({us_ta / us:.1f}x)"
    print(line, flush=True)
