"""The oracle restatement is pinned against outputs of the reference's own modules
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, load_shape_golden, rel_l2, shape_golden_names
from oracle import spectral_oracle as so


def _cfg(g):
    stft = so.stft_from_kwargs(**g["stft_kwargs"]) if g["stft_kwargs"] is not None else []
    mel = so.mel_from_kwargs(**g["mel_kwargs"]) if g["mel_kwargs"] is not None else []
    return stft, mel


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("route", ["explicit", "aten"])
def test_oracle_fp32_matches_reference(name, route):
    g = load_golden(name)
    stft, mel = _cfg(g)
    losses, grad = so.losses_and_grad(g["y_hat"], g["y"], stft, mel, dtype=torch.float32,
                                      use_torch_stft=(route == "aten"))
    np.testing.assert_allclose(losses, g["loss32"], rtol=2e-6, atol=1e-7)
    # fp32 gradients of two fp32 routes differ by rounding only; the yardstick is the
    # reference's own fp32-vs-fp64 distance (ill-conditioned gradients)
    yard = rel_l2(g["grad32"], g["grad64"])
    tol = 1e-6 if route == "aten" else max(3.0 * yard, 5e-6)
    assert rel_l2(grad.numpy().reshape(g["grad32"].shape), g["grad32"]) <= tol


@pytest.mark.parametrize("name", golden_names())
def test_oracle_fp64_autograd_and_analytic(name):
    g = load_golden(name)
    stft, mel = _cfg(g)
    losses, grad = so.losses_and_grad(g["y_hat"], g["y"], stft, mel, dtype=torch.float64)
    np.testing.assert_allclose(losses, g["loss64"], rtol=1e-11, atol=1e-13)
    assert rel_l2(grad.numpy().reshape(g["grad64"].shape), g["grad64"]) < 1e-6   # golden stored as float32
    if g["y_hat"].shape[-1] > 20000:
        return
    (sc, mag, mel_l), dx, _ = so.analytic(g["y_hat"], g["y"], stft, mel, dtype=np.float64)
    np.testing.assert_allclose([sc, mag, mel_l], g["loss64"], rtol=1e-11, atol=1e-13)
    assert rel_l2(dx, grad.numpy().reshape(dx.shape)) < 1e-12


def test_anchor_values_config1():
    """SURVEY 8c / BASELINE.md anchors for clean1 vs noise1."""
    g = load_golden("c1_clean1_noise1")
    assert g["y_hat"].shape == (1, 1, 123008)
    np.testing.assert_allclose(g["loss32"], [3.015629768, 3.128701448, 3.465816498], rtol=2e-7)
    np.testing.assert_allclose(g["loss64"], [3.015577380260, 3.128702007808, 3.465816607950], rtol=1e-10)


@pytest.mark.parametrize("name", shape_golden_names())
def test_shape_oracle_matches_reference_module(name):
    """oracle.shape_loss_and_grad against the outputs of the reference's own MultiWindowShapeLoss."""
    from oracle import spectral_oracle as so

    g = load_shape_golden(name)
    loss, grad = so.shape_loss_and_grad(g["y_hat"].numpy(), g["y"].numpy(), g["winlens"])
    assert abs(loss - g["loss64"]) <= 1e-12 * abs(g["loss64"])
    np.testing.assert_allclose(grad, g["grad64"], rtol=1e-6, atol=1e-12)   # atol: cancelling window lengths
    assert abs(loss - g["loss32"]) <= 1e-6 * abs(g["loss32"])
