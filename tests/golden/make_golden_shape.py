#!/usr/bin/env python3
"""Generate tests/golden/shape_*.npz from the UNMODIFIED reference losses/waveform_loss.py (authoring container only).

    python tests/golden/make_golden_shape.py

Each fixture: inputs, the window lengths, and what MultiWindowShapeLoss (waveform_loss.py:41-75) returns on CPU in
fp32 and fp64 together with d loss / d y_hat from autograd."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.spectral_oracle import synth_pair  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location(
        "_ref_waveform_loss", os.path.join(ref_loader.REFERENCE_ROOT, "losses", "waveform_loss.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = {
        "shape_default_b3_t9000": ([300, 200, 100], synth_pair(3, 9000, seed=21)),      # yaml shape_loss_params default
        "shape_ragged_b2_t5003": ([300, 200, 100], synth_pair(2, 5003, seed=22)),       # T not a multiple of any window
        "shape_single_b2_t4096": ([64], synth_pair(2, 4096, seed=23)),
    }
    for name, (winlens, (y_hat, y)) in cases.items():
        if "ragged" in name:
            y_hat = y_hat.clone()
            y_hat[0, 0, :900] = y[0, 0, :900]            # equal windows
            y_hat[1, 0, 1000:1700] = 0.0                 # silent prediction windows
        out = {}
        for tag, dt in (("32", torch.float32), ("64", torch.float64)):
            x = y_hat.to(dt).clone().requires_grad_(True)
            crit = ref.MultiWindowShapeLoss(winlen=winlens)
            loss = crit(x, y.to(dt))
            (g,) = torch.autograd.grad(loss, x)
            out["loss" + tag] = np.float64(loss.detach())
            out["grad" + tag] = g.to(torch.float32).numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y_hat=y_hat.numpy(), y=y.numpy(),
                            winlens=np.array(winlens, dtype=np.int64), **out)
        print(name, out["loss32"], out["loss64"])


if __name__ == "__main__":
    main()
