#!/usr/bin/env python3
"""Per-transform kernel timing (CUDA events) at BASELINE configs[1], for chunk-size sweeps:
   python profiles/time_kernels.py "1024:2,2048:2,512:2" "1024:4,2048:1,512:4" ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402
from dl_speech_enhancement_b200 import _abi  # noqa: E402
if os.environ.get("PROF_LIB"):          # profiling only: time an alternative build of the library
    _abi.LIB_PATH = os.path.abspath(os.environ["PROF_LIB"])
from dl_speech_enhancement_b200.engine import cuda_engine  # noqa: E402

B, T = int(os.environ.get("PROF_B", 16)), int(os.environ.get("PROF_T", 48000))
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
dev = torch.device("cuda:0")
stft = pkg.MultiResolutionSTFTLoss().to(dev)
mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
y = 0.1 * torch.randn(B, T, device=dev, generator=g)
x = y + 0.05 * torch.randn(B, T, device=dev, generator=g)
eng = cuda_engine()
one = torch.ones((), device=dev)
cases = [("stft1024", [stft.stft_losses[0].plan()]), ("stft2048", [stft.stft_losses[1].plan()]),
         ("stft512", [stft.stft_losses[2].plan()]), ("mel2048", mel.plans()),
         ("all4", stft.plans() + mel.plans())]


def timeit(fn, inner=10, reps=10):
    """Device time per call, free of host launch overhead: `inner` calls captured in one CUDA graph, replayed."""
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


for spec in (sys.argv[1:] or [""]):
    if spec:
        os.environ["SPECLOSS_FRAMES_PER_CHUNK"] = spec
    else:
        os.environ.pop("SPECLOSS_FRAMES_PER_CHUNK", None)
    row = []
    for name, plans in cases:
        fwd = timeit(lambda: eng.forward(plans, x, y, need_grad=True))
        nog = timeit(lambda: eng.forward(plans, x, y, need_grad=False))
        st = eng.forward(plans, x, y, need_grad=True)
        has_stft = any(p.kind == 0 for p in plans)
        bwd = timeit(lambda: eng.backward(st, one if has_stft else None, one if has_stft else None,
                                          None if name.startswith("stft") else one))
        row.append(f"{name}: fwd+grad {fwd:6.1f} us, fwd-only {nog:6.1f} us, combine {bwd:5.1f} us")
    print(f"[{spec or 'default'}]\n  " + "\n  ".join(row), flush=True)
