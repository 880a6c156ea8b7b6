#!/usr/bin/env python3
"""Device time of the waveform-shape-loss kernels (CUDA-graph replay, CUDA events) against the HBM roofline:
algorithmic bytes = 8 per sample forward (both signals once) + 4 per sample backward (dx once)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dl_speech_enhancement_b200.engine import cuda_engine  # noqa: E402

dev = torch.device("cuda:0")
eng = cuda_engine()
peak = 6543.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
WIN = [300, 200, 100]


def timeit(fn, inner=10, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(inner):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * inner) * 1e3


for rows, t_len, tag in ((16, 48000, "config 2"), (32, 192000, "config-4 share"), (256, 192000, "config 4 on one GPU")):
    g = torch.Generator(device=dev).manual_seed(0)
    y = 0.1 * torch.randn(rows, t_len, device=dev, generator=g)
    x = y + 0.05 * torch.randn(rows, t_len, device=dev, generator=g)
    one = torch.ones((), device=dev)
    fwd = timeit(lambda: eng.shape_forward(x, y, WIN))
    _, rec, rg = eng.shape_forward(x, y, WIN)
    bwd = timeit(lambda: eng.shape_backward(rec, rows, rg, t_len, WIN, one))
    n = rows * t_len
    print(f"{tag}: {rows} x {t_len}: forward (kernel + reduce + finalize) {fwd:7.1f} us = {8 * n / fwd / 1e3:7.1f} GB/s "
          f"({8 * n / fwd / 1e3 / peak:.1%} of {peak:.0f}), backward {bwd:7.1f} us = {4 * n / bwd / 1e3:7.1f} GB/s "
          f"({4 * n / bwd / 1e3 / peak:.1%})", flush=True)
