"""Import the reference's modules VERBATIM.  TEST / BASELINE INFRASTRUCTURE ONLY -- never on the product path.

Where the reference comes from, in this order:
  1. $SPECLOSS_REFERENCE_ROOT,
  2. /root/reference (the authoring container),
  3. oracle/_ref/ -- the unmodified copies staged by oracle/make_ref.sh; git-ignored, but they travel to the GPU box
     with gpurun, so that the `-m gpu` trainer test, bench.py's reference arm and its cpu_baseline leg can run the
     reference's own code there (nothing at run time reads /root/reference on the box).

losses/mel_loss.py does `import librosa` (mel_loss.py:14); librosa is not installed and there is no network, so a module
exposing only `librosa.filters.mel` is injected into sys.modules.  Its arithmetic is
oracle.spectral_oracle.slaney_mel_filterbank (a restatement of librosa 0.8.1), cross-checked against torchaudio in
tests/test_melfb.py.  trainer/trainerGAN.py does `from tensorboardX import SummaryWriter` (:20): a no-op writer is
injected likewise.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")


def reference_root() -> str | None:
    for cand in (os.environ.get("SPECLOSS_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "losses", "stft_loss.py")):
            return cand
    return None


REFERENCE_ROOT = reference_root() or "/root/reference"


def available() -> bool:
    return reference_root() is not None


def live_mount_available() -> bool:
    """The full reference tree (wav fixtures included), i.e. the authoring container."""
    root = reference_root()
    return root is not None and os.path.isdir(os.path.join(root, "notebook_files"))


def trainer_available() -> bool:
    root = reference_root()
    return root is not None and os.path.isfile(os.path.join(root, "trainer", "denoise.py")) and \
        os.path.isfile(os.path.join(root, "models", "autoencoder", "AudioDec.py"))


def gan_trainers_available() -> bool:
    """The autoencoder and vocoder trainers with their models and shipped YAMLs, beside the denoise trainer."""
    root = reference_root()
    return trainer_available() and all(os.path.isfile(os.path.join(root, *p)) for p in (
        ("trainer", "autoencoder.py"), ("trainer", "vocoder.py"), ("models", "vocoder", "HiFiGAN.py"),
        ("models", "vocoder", "UnivNet.py"), ("config", "autoencoder", "symAD_vctk_48000_hop300.yaml"),
        ("config", "vocoder", "AudioDec_v1_symAD_vctk_48000_hop300_clean.yaml")))


def _install_librosa_shim():
    if "librosa" in sys.modules and not getattr(sys.modules["librosa"], "_specloss_shim", False):
        return
    from oracle.spectral_oracle import slaney_mel_filterbank

    lib = types.ModuleType("librosa")
    lib._specloss_shim = True
    filt = types.ModuleType("librosa.filters")

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_unused):
        fmax = float(sr) / 2 if fmax is None else fmax
        return slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)

    filt.mel = mel
    lib.filters = filt
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = filt


def _install_tensorboardx_shim():
    if "tensorboardX" in sys.modules:
        return
    try:
        importlib.import_module("tensorboardX")
        return
    except ImportError:
        pass
    tb = types.ModuleType("tensorboardX")
    tb._specloss_shim = True

    class SummaryWriter:                      # the trainer only calls add_scalar / close
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def close(self):
            pass

    tb.SummaryWriter = SummaryWriter
    sys.modules["tensorboardX"] = tb


def _load(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(reference_root(), rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_losses():
    """Returns (stft_loss_module, mel_loss_module) executed from the reference's files."""
    if not available():
        raise FileNotFoundError("reference not found (neither mounted nor staged under oracle/_ref)")
    _install_librosa_shim()
    return _load("_ref_stft_loss", "losses/stft_loss.py"), _load("_ref_mel_loss", "losses/mel_loss.py")


def load_reference_trainer():
    """Imports the reference's packages the way its own scripts do (top-level `losses`, `trainer`, `models`, `layers`
    with the reference root on sys.path; train_denoise.py:20-33, bin/train.py) and returns a namespace with
    Trainer (trainer.denoise), Generator (models.autoencoder.AudioDec), the loss classes and the YAML config dict of
    config/denoise/symAD_vctk_48000_hop300.yaml."""
    if not trainer_available():
        raise FileNotFoundError("reference trainer not found (run oracle/make_ref.sh where /root/reference is mounted)")
    import yaml

    root = reference_root()
    _install_librosa_shim()
    _install_tensorboardx_shim()
    for top in ("losses", "trainer", "models", "layers"):
        mod = sys.modules.get(top)
        if mod is not None and not str(getattr(mod, "__file__", "") or getattr(mod, "__path__", [""])[0]).startswith(root):
            raise RuntimeError(f"a different top-level module named {top!r} is already imported")
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace()
    ns.Trainer = importlib.import_module("trainer.denoise").Trainer
    ns.Generator = importlib.import_module("models.autoencoder.AudioDec").Generator
    losses = importlib.import_module("losses")
    ns.MultiMelSpectrogramLoss = losses.MultiMelSpectrogramLoss
    ns.MultiResolutionSTFTLoss = losses.MultiResolutionSTFTLoss
    ns.MultiWindowShapeLoss = losses.MultiWindowShapeLoss
    with open(os.path.join(root, "config", "denoise", "symAD_vctk_48000_hop300.yaml")) as f:
        ns.config = yaml.safe_load(f)
    ns.root = root
    ns.trainers = {"denoise": ns.Trainer}
    ns.configs = {"denoise/symAD_vctk_48000_hop300": ns.config}
    if gan_trainers_available():
        # trainer/autoencoder.py:19 (symmetric codec, stage 1 metric-only, stage 2 adversarial), trainer/vocoder.py:19 (HiFiGAN
        # decoder on the frozen analyzer's codes); models and criteria as the shipped YAMLs name them
        ns.trainers["autoencoder"] = importlib.import_module("trainer.autoencoder").Trainer
        ns.trainers["vocoder"] = importlib.import_module("trainer.vocoder").Trainer
        hifigan = importlib.import_module("models.vocoder.HiFiGAN")
        univnet = importlib.import_module("models.vocoder.UnivNet")
        ns.HiFiGANGenerator, ns.HiFiGANDiscriminator, ns.UnivNetDiscriminator = hifigan.Generator, hifigan.Discriminator, univnet.Discriminator
        ns.discriminator_module = importlib.import_module("models.vocoder.modules.discriminator")
        ns.GeneratorAdversarialLoss = losses.GeneratorAdversarialLoss
        ns.DiscriminatorAdversarialLoss = losses.DiscriminatorAdversarialLoss
        ns.FeatureMatchLoss = losses.FeatureMatchLoss
        for sub in ("autoencoder", "vocoder"):
            d = os.path.join(root, "config", sub)
            for name in sorted(os.listdir(d)):
                if name.endswith(".yaml"):
                    with open(os.path.join(d, name)) as f:
                        ns.configs[f"{sub}/{name[:-5]}"] = yaml.safe_load(f)
    return ns


def load_fixture_pair(idx: int = 1):
    """(y_hat, y) = (noise{idx}.wav resampled 24->48 kHz, clean{idx}.wav / 32768), shapes (1,1,T)."""
    import warnings

    import numpy as np
    import torch
    import torchaudio
    from scipy.io import wavfile

    root = reference_root()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sr_c, clean = wavfile.read(os.path.join(root, "notebook_files", f"clean{idx}.wav"))
        sr_n, noise = wavfile.read(os.path.join(root, "notebook_files", f"noise{idx}.wav"))
    y = torch.from_numpy(clean.astype(np.float32) / 32768.0)
    n = torch.from_numpy(noise.astype(np.float32))
    if sr_n != sr_c:
        n = torchaudio.functional.resample(n, sr_n, sr_c)
    t = min(y.numel(), n.numel())
    return n[:t].reshape(1, 1, t).contiguous(), y[:t].reshape(1, 1, t).contiguous()
