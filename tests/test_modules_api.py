"""Drop-in contract of the nn.Modules: ctor signatures, buffers, error behaviour.
CPU-only; arithmetic goes through the emulator where needed."""
import inspect

import numpy as np
import pytest
import torch

import dl_speech_enhancement_b200 as pkg
from conftest import load_golden
from oracle import ref_loader

YAML_MEL = [  # every distinct mel_loss_params block in the reference's config/ tree
    dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window", num_mels=80, fmin=0, fmax=24000, log_base=None),
    dict(fs=24000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window", num_mels=80, fmin=0, fmax=24000, log_base=None),
    dict(fs=24000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[2048], window="hann_window", num_mels=80, fmin=0, fmax=12000, log_base=None),
    dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[2048], window="hann_window", num_mels=80, fmin=0, fmax=24000, log_base=None),
]
YAML_STFT = dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240], window="hann_window")


def test_public_names():
    for n in ("stft", "SpectralConvergenceLoss", "LogSTFTMagnitudeLoss", "STFTLoss", "MultiResolutionSTFTLoss",
              "MelSpectrogram", "MultiMelSpectrogramLoss"):
        assert hasattr(pkg, n)


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")
@pytest.mark.parametrize("cls", ["STFTLoss", "MultiResolutionSTFTLoss", "MelSpectrogram", "MultiMelSpectrogramLoss"])
def test_ctor_signatures_equal_reference(cls):
    stft_mod, mel_mod = ref_loader.load_reference_losses()
    ref = getattr(stft_mod, cls, None) or getattr(mel_mod, cls)
    a, b = inspect.signature(ref.__init__), inspect.signature(getattr(pkg, cls).__init__)
    assert [(p.name, p.default) for p in a.parameters.values()] == [(p.name, p.default) for p in b.parameters.values()]


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")
def test_buffers_equal_reference():
    stft_mod, mel_mod = ref_loader.load_reference_losses()
    for kw in YAML_MEL + [dict()]:
        ours, ref = pkg.MultiMelSpectrogramLoss(**kw), mel_mod.MultiMelSpectrogramLoss(**kw)
        assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
        for k, v in ref.state_dict().items():
            torch.testing.assert_close(ours.state_dict()[k], v, rtol=0, atol=1e-9)
    ours, ref = pkg.MultiResolutionSTFTLoss(**YAML_STFT), stft_mod.MultiResolutionSTFTLoss(**YAML_STFT)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v)


def test_state_dict_keys():
    assert sorted(pkg.MultiResolutionSTFTLoss().state_dict()) == [f"stft_losses.{i}.window" for i in range(3)]
    assert sorted(pkg.MultiMelSpectrogramLoss(**YAML_MEL[0]).state_dict()) == ["mel_transfers.0.melmat", "mel_transfers.0.window"]


def test_ctor_errors():
    with pytest.raises(AssertionError):
        pkg.MultiResolutionSTFTLoss(fft_sizes=[1024, 512], hop_sizes=[120], win_lengths=[600, 240])
    with pytest.raises(AssertionError):
        pkg.MultiMelSpectrogramLoss(fft_sizes=[1024], hop_sizes=[120, 50], win_lengths=[600])
    with pytest.raises(ValueError):
        pkg.MelSpectrogram(log_base=3.0)
    with pytest.raises(NotImplementedError):          # outside the kernels' envelope: raise, never fall back
        pkg.STFTLoss(fft_size=400, hop_size=100, win_length=400)


def test_cpu_tensors_raise_no_fallback():
    crit = pkg.MultiResolutionSTFTLoss()
    x = torch.randn(2, 1, 4800)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        crit(x, x)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pkg.MultiMelSpectrogramLoss(**YAML_MEL[0])(x, x)


def test_short_input_raises_like_torch_stft(emu_engine):
    from dl_speech_enhancement_b200.functional import spectral_losses
    x = torch.randn(1, 1024)        # T == n_fft/2 for 2048: reflect pad impossible
    with pytest.raises(RuntimeError, match="reflect"):
        spectral_losses(x, x, pkg.MultiResolutionSTFTLoss().plans(), engine=emu_engine)


def test_outputs_are_independent_and_inplace_scalable(emu_engine):
    """trainerGAN.py:221,228-229 scale the returned losses in place before backward."""
    from dl_speech_enhancement_b200.functional import spectral_losses
    g = load_golden("minlen_b1_t1025")
    x = g["y_hat"].clone().requires_grad_(True)
    sc, mag = spectral_losses(x, g["y"], pkg.MultiResolutionSTFTLoss().plans(), engine=emu_engine)
    assert sc.dim() == 0 and mag.dim() == 0 and sc.data_ptr() != mag.data_ptr()
    sc *= 45.0
    mag *= 45.0
    (sc + mag).backward()
    assert x.grad.shape == x.shape and torch.isfinite(x.grad).all()
