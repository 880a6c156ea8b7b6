import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """Spectral-loss fixtures (tests/golden/make_golden.py)."""
    return [n for n in _all_golden() if not n.startswith(("shape_", "mel_l1_"))]


def mel_l1_golden_names():
    """Mel_L1 fixtures, produced by torchaudio (tests/golden/make_golden_mel_l1.py)."""
    return [n for n in _all_golden() if n.startswith("mel_l1_")]


def load_mel_l1_golden(name):
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return dict(pred=torch.from_numpy(z["pred"]), target=torch.from_numpy(z["target"]), loss32=float(z["loss32"]),
                loss64=float(z["loss64"]), mel_pred64=z["mel_pred64"], mel_target64=z["mel_target64"])


def shape_golden_names():
    """Waveform-shape-loss fixtures (tests/golden/make_golden_shape.py)."""
    return [n for n in _all_golden() if n.startswith("shape_")]


def load_shape_golden(name):
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return dict(y_hat=torch.from_numpy(z["y_hat"]), y=torch.from_numpy(z["y"]), winlens=[int(w) for w in z["winlens"]],
                loss32=float(z["loss32"]), loss64=float(z["loss64"]), grad32=z["grad32"], grad64=z["grad64"])


def load_golden(name):
    """Returns dict with y_hat, y (torch fp32), loss32/64, grad32/64 (numpy), stft_kwargs, mel_kwargs."""
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    out = dict(loss32=z["loss32"], loss64=z["loss64"], grad32=z["grad32"], grad64=z["grad64"],
               stft_kwargs=meta["stft_kwargs"], mel_kwargs=meta["mel_kwargs"])
    out["y_hat"] = torch.from_numpy(z["y_hat"])
    if "y" in z.files:
        out["y"] = torch.from_numpy(z["y"])
    else:  # config 1: clean1.wav kept as int16 (reference scaling: / 32768)
        y = torch.from_numpy(z["clean_int16"].astype(np.float32) / 32768.0)
        out["y"] = y.reshape(1, 1, -1).contiguous()
    return out


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def block_median_rel(a, b, block=300):
    """Typical LOCAL accuracy of a waveform gradient a against the yardstick b (torch tensors, any shape): the median over
    blocks of `block` samples of ||a - b||_block, relative to the RMS block norm of b.  Unlike the global rel-L2 it is not
    decided by a handful of bins at the clamp floor of stft() (|X|^2 ~ eps), where the gate [|X|^2 >= eps] and
    sign(ln A_x - ln A_y) flip under ANY rounding change and one flipped bin moves one frame's span of the gradient by
    1 / sqrt(eps) = 3162 units -- every fp32 evaluation, the reference's own included, is such an outlier on some tensors."""
    import torch

    d = (a.double() - b.double()).reshape(-1)
    r = b.double().reshape(-1)
    n = d.numel() // block
    e = d[:n * block].reshape(n, block).norm(dim=1)
    s = r[:n * block].reshape(n, block).norm(dim=1)
    return float(e.median() / torch.sqrt((s ** 2).mean()).clamp_min(1e-300))


EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="session")
def emu_engine():
    """Host-side Engine over the CPU SIMT emulation of the kernels (tests/emu/specloss_emu.cpp).

    Test infrastructure: it runs the same device code and the same C-ABI host logic as
    libspecloss.so with host threads playing the lanes, so index maps, epilogues, overlap-add
    and the Python driver are exercised without a GPU.  The product never loads it."""
    import ctypes
    import subprocess

    from dl_speech_enhancement_b200 import _abi
    from dl_speech_enhancement_b200.engine import Engine

    so_path = os.path.join(EMU_DIR, "libspecloss_emu.so")
    srcs = [os.path.join(EMU_DIR, "specloss_emu.cpp"), os.path.join(EMU_DIR, "cuda_emu.h"),
            os.path.join(ROOT, "include", "specloss.h")] + \
        [os.path.join(_abi.CSRC, f) for f in ("specloss_kernels.cuh", "specloss_host.inl", "fft_codelets.cuh", "melpower.cuh", "transform_eo.cuh")]
    if not os.path.exists(so_path) or os.path.getmtime(so_path) < max(os.path.getmtime(s) for s in srcs):
        cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-I" + EMU_DIR, "-o", so_path, srcs[0]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    return Engine(_abi.bind(ctypes.CDLL(so_path)))


def plans_for(golden):
    from dl_speech_enhancement_b200 import modules

    plans = []
    if golden["stft_kwargs"] is not None:
        plans += modules.MultiResolutionSTFTLoss(**golden["stft_kwargs"]).plans()
    if golden["mel_kwargs"] is not None:
        plans += modules.MultiMelSpectrogramLoss(**golden["mel_kwargs"]).plans()
    return plans


def run_losses(engine, golden, device="cpu", weights=(1.0, 1.0, 1.0)):
    """Runs the product's autograd function on a golden case; returns ([sc, mag, mel], grad ndarray)."""
    import torch

    from dl_speech_enhancement_b200.functional import spectral_losses

    plans = plans_for(golden)
    if device != "cpu":
        for p in plans:
            p.window, p.twiddle = p.window.to(device), p.twiddle.to(device)
            p.tables = {k: v.to(device) for k, v in p.tables.items()}
    x = golden["y_hat"].to(device).clone().requires_grad_(True)
    y = golden["y"].to(device)
    outs = spectral_losses(x, y, plans, engine=engine)
    vals = [0.0, 0.0, 0.0]
    w = []
    outs = list(outs)
    if golden["stft_kwargs"] is not None:
        vals[0], vals[1] = float(outs[0].detach()), float(outs[1].detach())
        w += [weights[0], weights[1]]
    if golden["mel_kwargs"] is not None:
        vals[2] = float(outs[-1].detach())
        w += [weights[2]]
    total = sum(wi * o for wi, o in zip(w, outs))
    total.backward()
    return vals, x.grad.detach().cpu().numpy()
