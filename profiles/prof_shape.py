#!/usr/bin/env python3
"""Profiling driver for the waveform-shape-loss kernels: a few fwd+bwd at the config-4 per-GPU share (32 x 4 s)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402

rows, t_len = int(os.environ.get("PROF_B", 32)), int(os.environ.get("PROF_T", 192000))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
y = 0.1 * torch.randn(rows, 1, t_len, device=dev, generator=g)
x = (y + 0.05 * torch.randn(rows, 1, t_len, device=dev, generator=g)).requires_grad_(True)
crit = pkg.MultiWindowShapeLoss().to(dev)
for _ in range(3):
    x.grad = None
    loss = crit(x, y)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss.detach()))
