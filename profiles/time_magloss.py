#!/usr/bin/env python3
"""Device time of the explicit-magnitude loss kernels (spl_mag_loss_*) against the HBM roofline: forward reads 8 B per
element, backward (gradient w.r.t. x_mag only) reads 8 B and writes 4 B."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dl_speech_enhancement_b200.engine import cuda_engine  # noqa: E402

dev = torch.device("cuda:0")
eng = cuda_engine()
peak = 6543.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])


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


for shape, tag in (((16, 401, 513), "config 2, 1024-point resolution"), ((256, 1601, 513), "config 4, 1024-point resolution")):
    x = torch.rand(*shape, device=dev) + 0.01
    y = torch.rand(*shape, device=dev) + 0.01
    one = torch.ones((), device=dev)
    n = x.numel()
    fwd = timeit(lambda: eng.mag_loss_forward(x, y, True, True))
    _, _, sums = eng.mag_loss_forward(x, y, True, True)
    bwd = timeit(lambda: eng.mag_loss_backward(x, y, sums, one, one, True, False))
    print(f"{tag}: {n / 1e6:.1f} M elements: forward (sums + reduce + finalize) {fwd:7.1f} us = {8 * n / fwd / 1e3:7.1f} GB/s "
          f"({8 * n / fwd / 1e3 / peak:.1%} of {peak:.0f}), backward {bwd:7.1f} us = {12 * n / bwd / 1e3:7.1f} GB/s "
          f"({12 * n / bwd / 1e3 / peak:.1%})", flush=True)
