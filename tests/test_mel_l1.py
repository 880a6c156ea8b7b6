"""the reference's `Mel_L1` evaluation metric (mel_spectrogram.py:36-44, sandbox.py:183-191) =
nn.L1Loss()(M(pred), M(target)), M = torchaudio.transforms.MelSpectrogram(48000) (n_fft 400, hop 200, 128 HTK mels,
power 2).  Golden vectors come from torchaudio itself (tests/golden/make_golden_mel_l1.py).  CPU tests pin the oracle
restatement and run the 400-point kernel through the SIMT emulator; `-m gpu` tests run the product module."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_mel_l1_golden, mel_l1_golden_names

REL = 1e-5            # parity with torchaudio's fp64 evaluation <= 1e-5 relative


def _mel_close(got, ref):
    """power-mel tensors: relative to the tensor's scale (empty HTK filters are exact zeros on both sides)."""
    scale = float(np.abs(ref).max())
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6 * scale)
    assert np.all(got[ref == 0.0] == 0.0)


@pytest.mark.parametrize("name", mel_l1_golden_names())
def test_oracle_restatement_matches_torchaudio(name):
    from oracle import spectral_oracle as so
    g = load_mel_l1_golden(name)
    l32 = float(so.mel_l1(g["pred"], g["target"]))
    l64 = float(so.mel_l1(g["pred"].double(), g["target"].double()))
    assert abs(l32 - g["loss32"]) <= 2e-6 * g["loss32"]
    assert abs(l64 - g["loss64"]) <= 1e-10 * g["loss64"]
    _mel_close(so.power_mel(g["pred"].double()).numpy(), g["mel_pred64"])


def test_htk_filterbank_equals_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    import warnings

    from dl_speech_enhancement_b200 import melfb
    from oracle import spectral_oracle as so
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for sr, n_mels in ((48000, 128), (24000, 128), (16000, 64)):
            ref = torchaudio.functional.melscale_fbanks(201, 0.0, float(sr // 2), n_mels, sr, norm=None, mel_scale="htk").numpy()
            assert np.array_equal(melfb.htk_mel_filterbank(sr, 400, n_mels), ref)
            assert np.array_equal(so.htk_fbanks(sr, 400, n_mels).numpy(), ref)
    fb = melfb.htk_mel_filterbank(48000, 400, 128)
    assert max(int((fb[k] != 0).sum()) for k in range(201)) <= 2          # banded: <= 2 mels per bin


def _emu_mel_l1(eng, pred, target, want_mels=True):
    from dl_speech_enhancement_b200 import modules
    crit = modules.MelL1(48000)
    x, y = pred.reshape(-1, pred.shape[-1]).contiguous(), target.reshape(-1, target.shape[-1]).contiguous()
    return eng.melpow_l1(x, y, 400, 200, crit.window, crit._twiddle, crit.n_mels, crit._mel_ptr, crit._mel_ent, want_mels)


@pytest.mark.parametrize("name", mel_l1_golden_names())
def test_emulated_kernel_matches_torchaudio(emu_engine, name):
    g = load_mel_l1_golden(name)
    loss, mx, my = _emu_mel_l1(emu_engine, g["pred"], g["target"])
    assert abs(float(loss) - g["loss64"]) <= REL * g["loss64"], (float(loss), g["loss64"])
    _mel_close(mx.numpy().reshape(g["mel_pred64"].shape), g["mel_pred64"])
    _mel_close(my.numpy().reshape(g["mel_target64"].shape), g["mel_target64"])


def test_emulated_identical_inputs_give_exact_zero(emu_engine):
    x = 0.3 * torch.randn(2, 1500, generator=torch.Generator().manual_seed(1))
    loss, mx, my = _emu_mel_l1(emu_engine, x, x.clone())
    assert float(loss) == 0.0 and torch.equal(mx, my)


def test_abi_rejects_other_transform_lengths(emu_engine):
    n = ctypes.c_int64()
    lib = emu_engine.lib
    assert lib.spl_melpow_geometry(2, 9600, 512, 200, ctypes.byref(n)) == -1
    assert "400" in lib.spl_last_error().decode()
    assert lib.spl_melpow_geometry(2, 200, 400, 200, ctypes.byref(n)) == -1          # T <= n_fft / 2: reflect pad impossible
    assert lib.spl_melpow_geometry(2, 9600, 400, 200, ctypes.byref(n)) == 0 and n.value > 0


def test_module_envelope():
    from dl_speech_enhancement_b200 import modules
    with pytest.raises(NotImplementedError):
        modules.MelL1(48000, n_fft=512)
    with pytest.raises(NotImplementedError):
        modules.MelL1(48000, mel_scale="slaney")
    crit = modules.MelL1(48000)
    assert crit.hop_length == 200 and crit.fb.shape == (201, 128) and crit.window.shape == (400,)
    with pytest.raises(RuntimeError, match="CUDA"):
        crit(torch.zeros(1, 9600), torch.zeros(1, 9600))                 # no CPU fallback


# ---- product path on the GPU --------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", mel_l1_golden_names())
def test_gpu_golden_vectors(name):
    import dl_speech_enhancement_b200 as pkg
    dev = torch.device("cuda:0")
    g = load_mel_l1_golden(name)
    crit = pkg.MelL1(48000).to(dev)
    loss = crit(g["pred"].to(dev), g["target"].to(dev))
    assert loss.dim() == 0 and not loss.requires_grad
    assert abs(float(loss) - g["loss64"]) <= REL * g["loss64"], (float(loss), g["loss64"])
    assert abs(float(pkg.Mel_L1(g["pred"].to(dev), g["target"].to(dev))) - g["loss64"]) <= REL * g["loss64"]
    mx, my = crit.mel_spectrograms(g["pred"].to(dev), g["target"].to(dev))
    assert tuple(mx.shape) == g["mel_pred64"].shape
    _mel_close(mx.cpu().numpy(), g["mel_pred64"])
    _mel_close(my.cpu().numpy(), g["mel_target64"])


@pytest.mark.gpu
def test_gpu_full_size_against_oracle_and_torchaudio():
    """16 x 1 s @ 48 kHz (B, 1, T) and the long-form 2 x 60 s: against the fp64 oracle; and against
    torchaudio.transforms.MelSpectrogram(48000) itself, evaluated in fp64 on the same GPU."""
    import warnings

    import dl_speech_enhancement_b200 as pkg
    from oracle import spectral_oracle as so
    dev = torch.device("cuda:0")
    crit = pkg.MelL1(48000).to(dev)
    for b, t in ((16, 48000), (2, 2880000)):
        y_hat, y = so.synth_pair(b, t, seed=21)
        loss = float(crit(y_hat.to(dev), y.to(dev)))
        ref = float(so.mel_l1(y_hat.double().to(dev), y.double().to(dev)))
        assert abs(loss - ref) <= REL * ref, (b, t, loss, ref)
        try:
            import torchaudio
        except ImportError:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = torchaudio.transforms.MelSpectrogram(48000).double().to(dev)
        ta = float(torch.nn.L1Loss()(m(y_hat.double().to(dev)), m(y.double().to(dev))))
        assert abs(loss - ta) <= REL * ta, (b, t, loss, ta)


@pytest.mark.gpu
def test_gpu_identical_inputs_and_errors():
    import dl_speech_enhancement_b200 as pkg
    dev = torch.device("cuda:0")
    crit = pkg.MelL1(48000).to(dev)
    x = torch.randn(3, 1, 7000, device=dev)
    assert float(crit(x, x.clone())) == 0.0
    with pytest.raises(RuntimeError):
        crit(x, x[:, :, :-1])
    with pytest.raises(RuntimeError):
        crit(torch.randn(1, 150, device=dev), torch.randn(1, 150, device=dev))     # T <= 200: torch.stft raises likewise
