"""Drop-in torch.nn.Module replacements for the reference's spectral losses.

Same class names, constructor signatures (defaults included), buffer names and forward()
contracts as /root/reference/losses/stft_loss.py and losses/mel_loss.py, so
`criterion["stft"]` / `criterion["mel"]` (trainer/trainerGAN.py:220,227) and
`MultiMelSpectrogramLoss(**config["mel_loss_params"])` (train_denoise.py:121) work unchanged.
The arithmetic runs in libspecloss.so (hand-written sm_100a CUDA); tensors must be fp32 CUDA.

Differences from the reference, all raising rather than silently diverging:
  * CPU / fp64 / fp16 inputs raise RuntimeError (the reference runs anywhere torch.stft does);
  * fft_size must be 512, 1024 or 2048 and hop <= win_length (every shipped YAML satisfies this);
  * the gradient w.r.t. the *target* is not produced (NotImplementedError if requested).
"""
from __future__ import annotations

import math
from typing import List

import torch

from . import melfb
from ._abi import SPL_KIND_MEL, SPL_KIND_STFT
from .engine import TransformPlan, cuda_engine, fft_geometry, mel_gemm_weights, mel_tables, melpow_csr, melpow_twiddle_table, twiddle_eo_table, twiddle_table
from .functional import log_mel_spectrogram, magnitude_loss, shape_loss, spectral_losses
from .functional import spectrogram as _spectrogram_fn


def _cached_plans(owner, children) -> List[TransformPlan]:
    """Plans are rebuilt only when a buffer moved (module.to(device), load_state_dict)."""
    sig = tuple(c.window.data_ptr() for c in children) + tuple(c._twiddle.data_ptr() for c in children)
    cache = owner.__dict__.get("_plan_cache")
    if cache is None or cache[0] != sig:
        cache = (sig, [c.plan() for c in children])
        owner.__dict__["_plan_cache"] = cache
    return cache[1]


def _window(name: str, win_length: int) -> torch.Tensor:
    return getattr(torch, name)(win_length)          # as stft_loss.py:97 / mel_loss.py:49


_TWIDDLES = {}


def _twiddle_on(n_fft: int, device) -> torch.Tensor:
    key = (n_fft, str(device))
    if key not in _TWIDDLES:
        _TWIDDLES[key] = twiddle_table(n_fft).to(device)
    return _TWIDDLES[key]


def _explicit_input(x: torch.Tensor, what: str) -> torch.Tensor:
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError(f"{what}: fp32 CUDA tensors only (sm_100a kernels, no CPU fallback)")
    return x.contiguous()


def _target_needs_grad(y: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and y.requires_grad


def _flatten_channels(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.size(-1)) if x.dim() == 3 else x          # stft_loss.py:158-160


def stft(x, fft_size, hop_size, win_length, window, eps=1e-7):
    """Magnitude spectrogram (B, #frames, fft_size // 2 + 1) of x (B, T): losses/stft_loss.py:19-35, i.e.
    sqrt(clamp(|torch.stft(x, fft_size, hop_size, win_length, window)|^2, eps)).transpose(2, 1), computed by the
    sm_100a spectrogram kernel (two frames per complex FFT).  Differentiable w.r.t. x (not the window): backward
    recomputes the spectra and runs the adjoint transform (spl_spectrogram_backward).  The fused loss path never
    materialises this tensor."""
    x = _explicit_input(x, "stft()")
    if x.dim() != 2:
        raise RuntimeError(f"stft(): expected (B, T), got {tuple(x.shape)}")
    if window.numel() != win_length:
        raise RuntimeError("stft(): window must have win_length taps")
    if window.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("stft(): the gradient w.r.t. the window is not implemented (it is a constant buffer "
                                  "in every reference module, stft_loss.py:97)")
    window = window.detach().to(device=x.device, dtype=torch.float32).contiguous()
    plan = TransformPlan(SPL_KIND_STFT, fft_size, hop_size, win_length, eps, window, _twiddle_on(fft_size, x.device))
    plan.validate_explicit()
    return _spectrogram_fn(x, plan)


def spectrogram(waveform, pad, window, n_fft, hop_length, win_length, power, normalized, center=True,
                pad_mode="reflect", onesided=True, return_complex=None):
    """torchaudio.functional.spectrogram for the case the UnivNet multi-resolution spectral discriminator uses
    (models/vocoder/modules/discriminator.py:556-565: pad=win_length // 2, power=1.0, normalized=False): zero-pad the
    waveform by `pad` on both sides, then |STFT| (reflect-centred, onesided, no clamp).  waveform (..., T) fp32 CUDA ->
    (..., n_fft // 2 + 1, #frames), differentiable w.r.t. the waveform.  Runs on the sm_100a spectrogram kernels
    (spl_spectrogram / spl_spectrogram_backward with eps = 0); anything outside this case raises."""
    if power != 1.0 or normalized or not center or pad_mode != "reflect" or not onesided or return_complex:
        raise NotImplementedError("spectrogram(): only power=1.0, normalized=False, center=True, pad_mode='reflect', "
                                  "onesided=True (the UnivNet discriminator front-end) is implemented")
    x = _explicit_input(waveform, "spectrogram()")
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    if pad > 0:
        x = torch.nn.functional.pad(x, (pad, pad), "constant")
    mag = stft(x, n_fft, hop_length, win_length, window, eps=0.0)          # (B, F, K)
    return mag.reshape(lead + mag.shape[1:]).transpose(-1, -2)


class SpectralConvergenceLoss(torch.nn.Module):
    """||y_mag - x_mag||_F / ||y_mag||_F on explicit magnitudes (stft_loss.py:38-56), on the sm_100a streaming kernels
    (spl_mag_loss_*), differentiable w.r.t. both arguments.  The fused STFTLoss never materialises the magnitudes and
    does not call this; it is for callers that compose stft() + losses themselves."""

    def forward(self, x_mag, y_mag):
        return magnitude_loss(x_mag, y_mag, 0)


class LogSTFTMagnitudeLoss(torch.nn.Module):
    """mean |log y_mag - log x_mag| on explicit magnitudes (stft_loss.py:59-77), same kernels."""

    def forward(self, x_mag, y_mag):
        return magnitude_loss(x_mag, y_mag, 1)


class STFTLoss(torch.nn.Module):
    """One STFT resolution: forward(x (B,T), y (B,T)) -> (sc_loss, mag_loss).  stft_loss.py:80-117."""

    def __init__(self, fft_size=1024, hop_size=120, win_length=600, window="hann_window"):
        super().__init__()
        fft_geometry(fft_size)
        self.fft_size = fft_size
        self.hop_size = hop_size
        self.win_length = win_length
        self.spectral_convergence_loss = SpectralConvergenceLoss()
        self.log_stft_magnitude_loss = LogSTFTMagnitudeLoss()
        self.register_buffer("window", _window(window, win_length))
        self.register_buffer("_twiddle", twiddle_table(fft_size), persistent=False)
        if fft_size == 2048:
            self.register_buffer("_twiddle_eo", twiddle_eo_table(), persistent=False)

    def plan(self) -> TransformPlan:
        return TransformPlan(SPL_KIND_STFT, self.fft_size, self.hop_size, self.win_length, 1e-7,
                             self.window, self._twiddle, twiddle_eo=getattr(self, "_twiddle_eo", None))

    def forward(self, x, y):
        if _target_needs_grad(y):
            return self._forward_explicit(x, y, getattr(self, "process_group", None))
        return spectral_losses(x, y, _cached_plans(self, [self]), group=getattr(self, "process_group", None))

    def _forward_explicit(self, x, y, group=None):
        """The reference's own composition (stft_loss.py:99-117) on the explicit kernels -- stft() of both signals, then the
        two magnitude losses, each differentiable w.r.t. both arguments: the route taken when the TARGET requires a gradient
        (the fused kernels produce the prediction's only).  No caller in the reference needs it; it exists because the
        reference's modules would deliver that gradient."""
        if group is not None:
            raise NotImplementedError("a gradient w.r.t. the target is not available together with process_group (sharded mode)")
        x, y = _flatten_channels(x), _flatten_channels(y)
        x_mag = stft(x, self.fft_size, self.hop_size, self.win_length, self.window)
        y_mag = stft(y, self.fft_size, self.hop_size, self.win_length, self.window)
        return self.spectral_convergence_loss(x_mag, y_mag), self.log_stft_magnitude_loss(x_mag, y_mag)


class MultiResolutionSTFTLoss(torch.nn.Module):
    """forward(x, y) -> (sc_loss, mag_loss), means over resolutions.  stft_loss.py:120-170.
    All resolutions run in one fused launch sequence with a single backward kernel."""

    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                 window="hann_window"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = torch.nn.ModuleList()
        for fft_size, hop_size, win_length in zip(fft_sizes, hop_sizes, win_lengths):
            self.stft_losses += [STFTLoss(fft_size, hop_size, win_length, window)]
        self.process_group = None      # set to a torch.distributed group to shard the batch over ranks

    def plans(self) -> List[TransformPlan]:
        return _cached_plans(self, self.stft_losses)

    def forward(self, x, y):
        if _target_needs_grad(y):                # stft_loss.py:161-168 over the explicit route of every resolution
            sc_loss, mag_loss = 0.0, 0.0
            for f in self.stft_losses:
                sc_l, mag_l = f._forward_explicit(x, y, self.process_group)
                sc_loss = sc_loss + sc_l
                mag_loss = mag_loss + mag_l
            return sc_loss / len(self.stft_losses), mag_loss / len(self.stft_losses)
        return spectral_losses(x, y, self.plans(), group=self.process_group)


class MelSpectrogram(torch.nn.Module):
    """Log-mel spectrogram (mel_loss.py:19-94): same ctor, same `window` and `melmat` buffers.  forward(x) returns
    the explicit (B, num_mels, #frames) tensor: spectrogram kernel + mel projection as a tcgen05 tensor-core GEMM
    (3xTF32) with the clamp and the log fused into its epilogue.  Differentiable w.r.t. x: backward recomputes the
    spectra and applies the transposed (banded) projection inside the FFT kernel (spl_spectrogram_backward).  Inside
    MultiMelSpectrogramLoss the spectrogram is never materialised (banded projection inside the fused loss kernel)."""

    def __init__(self, fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann_window",
                 num_mels=80, fmin=80, fmax=7600, center=True, normalized=False, onesided=True,
                 eps=1e-10, log_base=10.0):
        super().__init__()
        fft_geometry(fft_size)
        self.fft_size = fft_size
        self.hop_size = hop_size
        self.win_length = win_length if win_length is not None else fft_size
        self.center = center
        self.normalized = normalized
        self.onesided = onesided
        self.register_buffer("window", _window(window, self.win_length))
        self.eps = eps
        fmin = 0 if fmin is None else fmin
        fmax = fs / 2 if fmax is None else fmax
        mel = melfb.mel_filterbank(sr=fs, n_fft=fft_size, n_mels=num_mels, fmin=fmin, fmax=fmax)
        self.register_buffer("melmat", torch.from_numpy(mel.T.copy()).float())
        self.log_base = log_base
        if self.log_base is None:
            self.log = torch.log
        elif self.log_base == 2.0:
            self.log = torch.log2
        elif self.log_base == 10.0:
            self.log = torch.log10
        else:
            raise ValueError(f"log_base: {log_base} is not supported.")
        self.num_mels = num_mels
        self.register_buffer("_twiddle", twiddle_table(fft_size), persistent=False)
        if fft_size == 2048:
            self.register_buffer("_twiddle_eo", twiddle_eo_table(), persistent=False)
        tabs = mel_tables(mel.T, fft_size)
        self._table_names = tuple(tabs.keys())
        for name, t in tabs.items():
            self.register_buffer("_" + name, t, persistent=False)
        if num_mels <= 128:
            w_hi, w_lo = mel_gemm_weights(mel.T, fft_size)
            self.register_buffer("_w_hi", w_hi, persistent=False)
            self.register_buffer("_w_lo", w_lo, persistent=False)

    def plan(self) -> TransformPlan:
        tables = {n: getattr(self, "_" + n) for n in self._table_names}
        inv_ln = 1.0 if self.log_base is None else 1.0 / math.log(self.log_base)
        return TransformPlan(SPL_KIND_MEL, self.fft_size, self.hop_size, self.win_length, self.eps,
                             self.window, self._twiddle, self.num_mels, inv_ln, tables,
                             twiddle_eo=getattr(self, "_twiddle_eo", None))

    def forward(self, x):
        from .engine import gemm_ld
        if x.dim() == 3:
            x = x.reshape(-1, x.size(2))             # mel_loss.py:84-85
        x = _explicit_input(x, "MelSpectrogram.forward()")
        if not hasattr(self, "_w_hi"):
            raise NotImplementedError("MelSpectrogram.forward(): the tensor-core projection supports num_mels <= 128")
        log_scale = 1.0 if self.log_base is None else 1.0 / math.log(self.log_base)
        return log_mel_spectrogram(x, self.plan(), self._w_hi, self._w_lo, gemm_ld(self.fft_size), log_scale)


class MultiMelSpectrogramLoss(torch.nn.Module):
    """forward(y_hat, y) -> mel_loss, mean over resolutions of the log-mel L1.  mel_loss.py:97-156."""

    def __init__(self, fs=22050, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                 window="hann_window", num_mels=80, fmin=80, fmax=7600, center=True, normalized=False,
                 onesided=True, eps=1e-10, log_base=10.0):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.mel_transfers = torch.nn.ModuleList()
        for fft_size, hop_size, win_length in zip(fft_sizes, hop_sizes, win_lengths):
            self.mel_transfers += [
                MelSpectrogram(fs=fs, fft_size=fft_size, hop_size=hop_size, win_length=win_length, window=window,
                               num_mels=num_mels, fmin=fmin, fmax=fmax, center=center, normalized=normalized,
                               onesided=onesided, eps=eps, log_base=log_base)
            ]
        self.process_group = None

    def plans(self) -> List[TransformPlan]:
        return _cached_plans(self, self.mel_transfers)

    def forward(self, y_hat, y):
        if _target_needs_grad(y):
            # mel_loss.py:151-154 on the explicit log-mel tensors (tensor-core projection forward, banded adjoint backward),
            # differentiable w.r.t. both signals: the route for a target that requires a gradient
            if self.process_group is not None:
                raise NotImplementedError("a gradient w.r.t. the target is not available together with process_group (sharded mode)")
            mel_loss = 0.0
            for f in self.mel_transfers:
                mel_loss = mel_loss + torch.nn.functional.l1_loss(f(y_hat), f(y))
            return mel_loss / len(self.mel_transfers)
        (mel,) = spectral_losses(y_hat, y, self.plans(), group=self.process_group)
        return mel


class SpectralLoss(torch.nn.Module):
    """MultiResolutionSTFTLoss + MultiMelSpectrogramLoss in ONE launch sequence (extension, not in the
    reference): forward(y_hat, y) -> (sc_loss, mag_loss, mel_loss).  Shares the waveform reads and the
    backward kernel between the two loss families."""

    def __init__(self, stft_loss_params=None, mel_loss_params=None):
        super().__init__()
        self.stft = MultiResolutionSTFTLoss(**(stft_loss_params or {}))
        self.mel = MultiMelSpectrogramLoss(**(mel_loss_params or {}))
        self.process_group = None

    def forward(self, y_hat, y):
        return spectral_losses(y_hat, y, self.stft.plans() + self.mel.plans(), group=self.process_group)


class WaveformShapeLoss(torch.nn.Module):
    """L1 between the max-pooled magnitudes of prediction and target (losses/waveform_loss.py:15-38): same ctor and
    forward(y_hat (B, 1, T), y (B, 1, T)) -> 0-dim loss, on the sm_100a shape-loss kernels (both signals read once,
    one 4-byte record per window for the backward)."""

    def __init__(self, winlen):
        super().__init__()
        self.winlen = winlen
        self.process_group = None

    def forward(self, y_hat, y):
        return shape_loss(y_hat, y, [self.winlen], group=self.process_group)


class MultiWindowShapeLoss(torch.nn.Module):
    """Mean over window lengths of WaveformShapeLoss (losses/waveform_loss.py:41-75; criterion["shape"] of
    trainer/trainerGAN.py:235-239).  All window lengths run in ONE pass over the two signals."""

    def __init__(self, winlen=[300, 200, 100]):
        super().__init__()
        self.shape_losses = torch.nn.ModuleList()
        for wl in winlen:
            self.shape_losses += [WaveformShapeLoss(wl)]
        self.process_group = None

    def forward(self, y_hat, y):
        return shape_loss(y_hat, y, [m.winlen for m in self.shape_losses], group=self.process_group)


class MelL1(torch.nn.Module):
    """The reference's `Mel_L1` evaluation metric (mel_spectrogram.py:36-44, duplicated at sandbox.py:183-191):
        nn.L1Loss()(M(pred), M(target)),  M = torchaudio.transforms.MelSpectrogram(sample_rate)
    with torchaudio's defaults: n_fft = win_length = 400 (periodic Hann), hop 200, reflect-centred, power 2, 128 HTK mel
    filters, no normalisation, no log.  Same constructor arguments as torchaudio.transforms.MelSpectrogram for the subset
    the metric uses; anything else raises.  Forward only (it is a printed metric, mel_spectrogram.py:47-64): the result
    carries no autograd graph.  Runs on the sm_100a power-mel kernel (csrc/melpower.cuh, a 25 x 16 = 400-point transform)."""

    def __init__(self, sample_rate=16000, n_fft=400, win_length=None, hop_length=None, f_min=0.0, f_max=None, pad=0,
                 n_mels=128, window_fn=torch.hann_window, power=2.0, normalized=False, center=True, pad_mode="reflect",
                 norm=None, mel_scale="htk"):
        super().__init__()
        win_length = n_fft if win_length is None else win_length
        if n_fft != 400 or win_length != n_fft:
            raise NotImplementedError("MelL1: the power-mel kernel implements torchaudio's default n_fft = win_length = 400")
        if pad != 0 or power != 2.0 or normalized or not center or pad_mode != "reflect" or norm is not None or mel_scale != "htk":
            raise NotImplementedError("MelL1: only torchaudio.transforms.MelSpectrogram's defaults (pad=0, power=2, "
                                      "normalized=False, center=True, reflect, norm=None, mel_scale='htk') are implemented")
        self.sample_rate, self.n_fft, self.win_length = sample_rate, n_fft, win_length
        self.hop_length = win_length // 2 if hop_length is None else hop_length
        self.n_mels = n_mels
        fb = melfb.htk_mel_filterbank(sample_rate, n_fft, n_mels, f_min, f_max)
        self.register_buffer("window", window_fn(win_length))
        self.register_buffer("fb", torch.from_numpy(fb.copy()))
        self.register_buffer("_twiddle", melpow_twiddle_table(), persistent=False)
        ptr, ent = melpow_csr(fb)
        self.register_buffer("_mel_ptr", ptr, persistent=False)
        self.register_buffer("_mel_ent", ent, persistent=False)

    def _run(self, pred, target, want_mels):
        if pred.shape != target.shape:
            raise RuntimeError(f"shape mismatch: {tuple(pred.shape)} vs {tuple(target.shape)}")
        x = _explicit_input(pred.detach(), "MelL1")
        y = _explicit_input(target.detach(), "MelL1")
        if x.device != y.device or x.device != self.window.device:
            raise RuntimeError("MelL1: inputs and module buffers must be on the same CUDA device (call .to(device))")
        lead = x.shape[:-1]
        x, y = x.reshape(-1, x.shape[-1]), y.reshape(-1, y.shape[-1])
        loss, mx, my = cuda_engine().melpow_l1(x, y, self.n_fft, self.hop_length, self.window, self._twiddle, self.n_mels,
                                               self._mel_ptr, self._mel_ent, want_mels)
        if want_mels:
            mx, my = mx.reshape(lead + mx.shape[1:]), my.reshape(lead + my.shape[1:])
        return loss, mx, my

    def forward(self, pred, target):
        """pred, target: (..., T) fp32 CUDA -> 0-dim L1 between their power-mel spectrograms."""
        return self._run(pred, target, False)[0]

    def mel_spectrograms(self, pred, target):
        """(M(pred), M(target)), each (..., n_mels, 1 + T // hop): the tensors the metric compares."""
        return self._run(pred, target, True)[1:]


_MEL_L1 = {}


def Mel_L1(pred, target, sample_rate=48000):
    """Drop-in for the reference's module-level `Mel_L1(pred, target)` (mel_spectrogram.py:36-44:
    `mel_spectrogram = transforms.MelSpectrogram(48000)` at import time, then L1 of the two outputs)."""
    key = (sample_rate, str(pred.device))
    crit = _MEL_L1.get(key)
    if crit is None:
        crit = _MEL_L1[key] = MelL1(sample_rate).to(pred.device)
    return crit(pred, target)
