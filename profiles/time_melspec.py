#!/usr/bin/env python3
"""Explicit spectrogram / log-mel path (stft(), MelSpectrogram.forward): device time of the spectrogram kernel and
of the tcgen05 mel-projection GEMM, next to the fused mel-loss kernel that never materialises the spectrogram.
CUDA-graph replay timing (no host overhead).  PROF_B / PROF_T select the batch (default BASELINE configs[1])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dl_speech_enhancement_b200 as pkg  # noqa: E402
from dl_speech_enhancement_b200.engine import cuda_engine, gemm_ld  # noqa: E402

B, T = int(os.environ.get("PROF_B", 16)), int(os.environ.get("PROF_T", 48000))
dev = torch.device("cuda:0")
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
mel_loss = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
ms = mel_loss.mel_transfers[0]
stft = pkg.MultiResolutionSTFTLoss().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
y = 0.1 * torch.randn(B, T, device=dev, generator=g)
x = y + 0.05 * torch.randn(B, T, device=dev, generator=g)
eng = cuda_engine()


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


print(f"batch {B} x {T}")
for sl in stft.stft_losses:
    us = timeit(lambda: eng.spectrogram(x, sl.fft_size, sl.hop_size, sl.win_length, sl.window, sl._twiddle, 1e-7))
    print(f"  spectrogram {sl.fft_size}/{sl.hop_size}/{sl.win_length}: {us:7.1f} us")
ld = gemm_ld(2048)
us_spec = timeit(lambda: eng.spectrogram(x, 2048, 300, 2048, ms.window, ms._twiddle, 1e-10, ld=ld, split=True))
hi, lo = eng.spectrogram(x, 2048, 300, 2048, ms.window, ms._twiddle, 1e-10, ld=ld, split=True)
us_gemm = timeit(lambda: eng.mel_project(hi, lo, ms._w_hi, ms._w_lo, 80, 1e-10, 1.0))
rows = hi.shape[0] * hi.shape[1]
flops = 3 * 2.0 * rows * ld * 80
print(f"  mel 2048/300: spectrogram (TF32 split) {us_spec:7.1f} us, tcgen05 GEMM {us_gemm:7.1f} us "
      f"({rows} x {ld} x 80, 3xTF32: {flops / us_gemm / 1e6:.1f} TFLOP/s tensor, "
      f"{(2 * rows * ld * 4 + 2 * 80 * ld * 4 + rows * 80 * 4) / us_gemm / 1e3:.0f} GB/s operand traffic)")
us_ms = timeit(lambda: ms(x))
us_fused = timeit(lambda: eng.forward(mel_loss.plans(), x, y, need_grad=False))
print(f"  MelSpectrogram.forward(x) [spectrogram + GEMM]: {us_ms:7.1f} us;  x2 signals = {2 * us_ms:7.1f} us "
      f"(+ an L1 kernel) vs fused mel-loss forward (banded projection, both signals, no spectrogram in HBM): {us_fused:7.1f} us")

# backward of the explicit tensors (specgrad_kernel + combine): gradient of stft() / MelSpectrogram.forward w.r.t. x
print("  backward of the explicit tensors (recompute + adjoint transform, then the overlap-add gather):")
for sl in stft.stft_losses:
    plan = sl.plan()
    frames = 1 + T // sl.hop_size
    gout = torch.randn(B, frames, sl.fft_size // 2 + 1, device=dev)
    us = timeit(lambda: eng.spectrogram_backward(plan, x, gout))
    print(f"    stft {sl.fft_size}/{sl.hop_size}/{sl.win_length}: {us:7.1f} us")
gmel = torch.randn(B, 80, 1 + T // 300, device=dev)
us = timeit(lambda: eng.spectrogram_backward(ms.plan(), x, gmel))
print(f"    log-mel 2048/300 (banded transposed projection in the kernel): {us:7.1f} us")
