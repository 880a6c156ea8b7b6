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


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


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
