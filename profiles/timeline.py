#!/usr/bin/env python3
"""Device timeline of eager steps through the drop-in modules (torch.profiler / CUPTI; no nsys on this image):
every kernel of the last profiled step with its start, duration, stream and the gap to the previous kernel's end,
plus the totals per kernel name.  Answers "where does the step go besides the transform and gather kernels".

    PROF_B=32 PROF_T=192000 python profiles/timeline.py
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dl_speech_enhancement_b200 as pkg  # noqa: E402

B = int(os.environ.get("PROF_B", "32"))
T = int(os.environ.get("PROF_T", "192000"))
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
stft = pkg.MultiResolutionSTFTLoss().to(dev)
mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
y = 0.1 * torch.randn(B, 1, T, device=dev, generator=gen)
x = (y + 0.05 * torch.randn(B, 1, T, device=dev, generator=gen)).requires_grad_(True)


def step():
    x.grad = None
    ml = mel(x, y)
    sc, mag = stft(x, y)
    (sc + mag + ml).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    step()
b.record()
torch.cuda.synchronize()
print(f"{B} x {T / 48000:g} s: {a.elapsed_time(b) / 20 * 1000:.1f} us per eager step (CUDA events, unprofiled)")

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]):      # CUPTI start-up outside the step looked at
    step()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
last = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
last.sort(key=lambda e: e.time_range.start)
t0 = last[0].time_range.start
end_prev = t0
print(f"last profiled step: {len(last)} device activities, span {last[-1].time_range.end - t0:.1f} us")
print(f"{'start':>9} {'dur':>9} {'gap':>7}  name")
busy = 0.0
cover_end = t0
for e in last:
    s, en = e.time_range.start, e.time_range.end
    gap = s - cover_end
    if en > cover_end:
        busy += en - max(s, cover_end)
        cover_end = en
    print(f"{s - t0:9.1f} {en - s:9.1f} {gap:7.1f}  {e.name[:110]}")
print(f"device busy (union of activities) {busy:.1f} us of {last[-1].time_range.end - t0:.1f} us span")
tot = {}
for e in last:
    k = e.name.split("<")[0][:60]
    tot[k] = tot.get(k, 0.0) + (e.time_range.end - e.time_range.start)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:9.1f} us  {k}")
