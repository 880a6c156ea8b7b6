#!/usr/bin/env python3
"""Device timeline of the e2e loop of bench.py (pinned H2D of step i+1 under step i, D2H of the losses + host sync every step) at
256 x 4 s: where the 0.6 ms between the device-only step (7.49 ms) and the e2e step (8.09 ms) goes."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dl_speech_enhancement_b200 as pkg  # noqa: E402

B, T = int(os.environ.get("PROF_B", "256")), 192000
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
stft = pkg.MultiResolutionSTFTLoss().to(dev)
mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
g = torch.Generator().manual_seed(0)
host = []
for _ in range(2):
    y = 0.1 * torch.randn(B, 1, T, generator=g)
    host.append(((y + 0.05 * torch.randn(B, 1, T, generator=g)).pin_memory(), y.pin_memory()))
dbuf = [(torch.empty(B, 1, T, device=dev).requires_grad_(True), torch.empty(B, 1, T, device=dev)) for _ in range(2)]
copied = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]
copy_stream = torch.cuda.Stream()
host_out = torch.empty(3).pin_memory()


def loop(steps):
    cur = torch.cuda.current_stream()

    def fetch(i):
        hx, hy = host[i % 2]
        x, y = dbuf[i % 2]
        copy_stream.wait_event(consumed[i % 2])
        with torch.cuda.stream(copy_stream), torch.no_grad():
            x.copy_(hx, non_blocking=True)
            y.copy_(hy, non_blocking=True)
            copied[i % 2].record(copy_stream)
    consumed[0].record(cur)
    consumed[1].record(cur)
    fetch(0)
    for i in range(steps):
        x, y = dbuf[i % 2]
        fetch(i + 1)
        cur.wait_event(copied[i % 2])
        x.grad = None
        ml = mel(x, y)
        sc, mag = stft(x, y)
        (sc + mag + ml).backward()
        consumed[i % 2].record(cur)
        host_out.copy_(torch.stack([sc.detach(), mag.detach(), ml.detach()]), non_blocking=True)
        cur.synchronize()
    copy_stream.synchronize()


import time
loop(5)
t0 = time.perf_counter(); loop(10); print(f"e2e {1e3 * (time.perf_counter() - t0) / 10:.3f} ms/step (unprofiled)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]):
    loop(2)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    loop(3)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print(f"{'start':>10} {'dur':>9}  name")
for e in evs:
    print(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:9.1f}  {e.name[:90]}")
