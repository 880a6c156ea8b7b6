"""CPU oracle for the spectral-loss hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (dl_speech_enhancement_b200) never
does, and raises when its CUDA library is missing instead of falling back to this.

What it restates (all citations relative to /root/reference):
  * losses/stft_loss.py:19-35     stft() -> clamped magnitude            -> magnitude()
  * losses/stft_loss.py:38-56     spectral convergence                   -> stft_terms()/mr_stft_loss()
  * losses/stft_loss.py:59-77     log-magnitude L1                       -> stft_terms()/mr_stft_loss()
  * losses/stft_loss.py:120-170   multi-resolution mean                  -> mr_stft_loss()
  * losses/mel_loss.py:19-94      MelSpectrogram.forward                 -> log_mel()
  * losses/mel_loss.py:97-156     MultiMelSpectrogramLoss.forward        -> multi_mel_loss()
  * librosa.filters.mel (librosa==0.8.1, requirements.txt:26; third-party, absent from
    /root/reference and from this image) called at losses/mel_loss.py:54-60
                                                                         -> slaney_mel_filterbank()
  * torch.stft (torch==2.1.1, requirements.txt:75; third-party) called at
    losses/stft_loss.py:33 and losses/mel_loss.py:88 with center=True, reflect padding,
    window zero-padded to n_fft and centred, onesided, un-normalised     -> frames()/spectrum()

Two routes are provided on purpose:
  * the "explicit" route (frames() -> rfft) spells the framing out; it is what the
    analytic gradient analytic_grad() is checked against;
  * the "aten" route (use_torch_stft=True) issues the same ATen calls as the reference
    (torch.stft + elementwise ops + autograd) and is what the CPU baseline times.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's own modules, imported verbatim in
the authoring container by tests/golden/make_golden.py and committed as tests/golden/*.npz
(checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# third-party restatement: librosa 0.8.1 filters.mel(htk=False, norm='slaney', dtype=float32)
# --------------------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def _hz_to_mel(f: float) -> float:
    if f >= _MIN_LOG_HZ:
        return _MIN_LOG_MEL + math.log(f / _MIN_LOG_HZ) / _LOGSTEP
    return f / _F_SP


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    lin = _F_SP * m
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32 triangular Slaney filters, area-normalised."""
    n_bins = n_fft // 2 + 1
    bin_hz = np.linspace(0.0, float(sr) / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(float(fmin)), _hz_to_mel(float(fmax)), n_mels + 2))
    widths = np.diff(edges)
    out = np.zeros((n_mels, n_bins), dtype=np.float64)
    for m in range(n_mels):
        rising = (bin_hz - edges[m]) / widths[m]
        falling = (edges[m + 2] - bin_hz) / widths[m + 1]
        out[m] = np.maximum(0.0, np.minimum(rising, falling))
    out *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# configuration records (field names follow the reference's ctor kwargs)
# --------------------------------------------------------------------------------------
@dataclass
class StftRes:
    fft_size: int
    hop_size: int
    win_length: int
    window: str = "hann_window"
    eps: float = 1e-7          # losses/stft_loss.py:19


@dataclass
class MelRes:
    fs: float
    fft_size: int
    hop_size: int
    win_length: Optional[int]
    window: str = "hann_window"
    num_mels: int = 80
    fmin: Optional[float] = 80
    fmax: Optional[float] = 7600
    eps: float = 1e-10
    log_base: Optional[float] = 10.0

    def win(self) -> int:
        return self.fft_size if self.win_length is None else self.win_length


DEFAULT_STFT = [StftRes(1024, 120, 600), StftRes(2048, 240, 1200), StftRes(512, 50, 240)]


def mel_from_kwargs(**kw) -> List[MelRes]:
    """Expand MultiMelSpectrogramLoss(**kw) ctor kwargs (losses/mel_loss.py:100-138)."""
    d = dict(fs=22050, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
             window="hann_window", num_mels=80, fmin=80, fmax=7600, eps=1e-10, log_base=10.0)
    for k in ("center", "normalized", "onesided"):
        kw.pop(k, None)
    d.update(kw)
    assert len(d["fft_sizes"]) == len(d["hop_sizes"]) == len(d["win_lengths"])
    return [MelRes(d["fs"], n, h, w, d["window"], d["num_mels"], d["fmin"], d["fmax"], d["eps"], d["log_base"])
            for n, h, w in zip(d["fft_sizes"], d["hop_sizes"], d["win_lengths"])]


def stft_from_kwargs(**kw) -> List[StftRes]:
    d = dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
             window="hann_window")
    d.update(kw)
    assert len(d["fft_sizes"]) == len(d["hop_sizes"]) == len(d["win_lengths"])
    return [StftRes(n, h, w, d["window"]) for n, h, w in zip(d["fft_sizes"], d["hop_sizes"], d["win_lengths"])]


# --------------------------------------------------------------------------------------
# framing and spectra
# --------------------------------------------------------------------------------------
def padded_window(name: str, win_length: int, n_fft: int, dtype) -> torch.Tensor:
    """torch.<name>(win_length) centred in n_fft zeros, as torch.stft does internally.

    The reference registers the window as an fp32 buffer (stft_loss.py:97, mel_loss.py:49); its
    .double() runs therefore carry fp32-rounded window taps, and so does this restatement."""
    w = getattr(torch, name)(win_length, dtype=torch.float32).to(torch.float64)
    left = (n_fft - win_length) // 2
    full = torch.zeros(n_fft, dtype=torch.float64)
    full[left:left + win_length] = w
    return full.to(dtype)


def frames(x: torch.Tensor, n_fft: int, hop: int, win_length: int, window: str) -> torch.Tensor:
    """(B, T) -> (B, F, n_fft) windowed frames; reflect padding n_fft//2, F = 1 + T//hop."""
    half = n_fft // 2
    if x.shape[-1] <= half:
        raise RuntimeError("reflect padding needs T > n_fft/2 (torch.stft raises likewise)")
    xp = torch.nn.functional.pad(x.unsqueeze(1), (half, half), mode="reflect").squeeze(1)
    fr = xp.unfold(-1, n_fft, hop)
    return fr * padded_window(window, win_length, n_fft, x.dtype).to(x.device)


def spectrum(x: torch.Tensor, n_fft: int, hop: int, win_length: int, window: str,
             use_torch_stft: bool = False) -> torch.Tensor:
    """Complex one-sided STFT laid out (B, F, K)."""
    if use_torch_stft:
        w = getattr(torch, window)(win_length, dtype=torch.float32).to(device=x.device, dtype=x.dtype)
        return torch.stft(x, n_fft, hop, win_length, w, return_complex=True).transpose(1, 2)
    return torch.fft.rfft(frames(x, n_fft, hop, win_length, window), dim=-1)


def magnitude(x, n_fft, hop, win_length, window, eps, use_torch_stft=False) -> torch.Tensor:
    """sqrt(clamp(|X|^2, eps)), (B, F, K)  -- losses/stft_loss.py:33-35, losses/mel_loss.py:88-90."""
    s = spectrum(x, n_fft, hop, win_length, window, use_torch_stft)
    return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=eps))


def _flat(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-1]) if x.dim() == 3 else x


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def stft_terms(x, y, r: StftRes, use_torch_stft=False):
    """Per-resolution sums: S1 = sum (Ay-Ax)^2, S2 = sum Ay^2, S3 = sum |ln Ay - ln Ax|, n."""
    ax = magnitude(x, r.fft_size, r.hop_size, r.win_length, r.window, r.eps, use_torch_stft)
    ay = magnitude(y, r.fft_size, r.hop_size, r.win_length, r.window, r.eps, use_torch_stft)
    return ((ay - ax) ** 2).sum(), (ay ** 2).sum(), (torch.log(ay) - torch.log(ax)).abs().sum(), ax.numel()


def mr_stft_loss(x, y, resolutions: Sequence[StftRes] = None, use_torch_stft=False):
    """(sc, mag) as MultiResolutionSTFTLoss.forward, losses/stft_loss.py:146-170."""
    resolutions = DEFAULT_STFT if resolutions is None else resolutions
    x, y = _flat(x), _flat(y)
    sc = mag = 0.0
    for r in resolutions:
        ax = magnitude(x, r.fft_size, r.hop_size, r.win_length, r.window, r.eps, use_torch_stft)
        ay = magnitude(y, r.fft_size, r.hop_size, r.win_length, r.window, r.eps, use_torch_stft)
        # torch.norm(p="fro") as the reference calls it (stft_loss.py:56): in fp32 on CPU its
        # accumulation is ~2e-5 less accurate than sqrt(sum(.^2)) on 1e6-element inputs, and the
        # golden fp32 values carry that error, so the restatement keeps the same call.
        sc = sc + torch.norm(ay - ax, p="fro") / torch.norm(ay, p="fro")
        mag = mag + (torch.log(ay) - torch.log(ax)).abs().mean()
    return sc / len(resolutions), mag / len(resolutions)


def _logfn(base):
    if base is None:
        return torch.log
    if base == 2.0:
        return torch.log2
    if base == 10.0:
        return torch.log10
    raise ValueError(f"log_base: {base} is not supported.")


_MEL_CACHE = {}


def melmat(r: MelRes, dtype) -> torch.Tensor:
    fmin = 0 if r.fmin is None else r.fmin
    fmax = r.fs / 2 if r.fmax is None else r.fmax
    key = (r.fs, r.fft_size, r.num_mels, fmin, fmax)
    if key not in _MEL_CACHE:
        _MEL_CACHE[key] = torch.from_numpy(slaney_mel_filterbank(r.fs, r.fft_size, r.num_mels, fmin, fmax).T.copy())
    return _MEL_CACHE[key].to(dtype)     # (K, num_mels), fp32 values as the reference buffer


def log_mel(x, r: MelRes, use_torch_stft=False) -> torch.Tensor:
    """(B, num_mels, F) log-mel spectrogram, losses/mel_loss.py:74-94."""
    x = _flat(x)
    amp = magnitude(x, r.fft_size, r.hop_size, r.win(), r.window, r.eps, use_torch_stft)
    mel = torch.clamp(torch.matmul(amp, melmat(r, x.dtype).to(x.device)), min=r.eps)
    return _logfn(r.log_base)(mel).transpose(1, 2)


def multi_mel_loss(x, y, resolutions: Sequence[MelRes], use_torch_stft=False):
    tot = 0.0
    for r in resolutions:
        tot = tot + (log_mel(x, r, use_torch_stft) - log_mel(y, r, use_torch_stft)).abs().mean()
    return tot / len(resolutions)


def losses_and_grad(x, y, stft_res, mel_res, weights=(1.0, 1.0, 1.0), dtype=torch.float32,
                    use_torch_stft=False):
    """Autograd route: returns (sc, mag, mel) as python floats and d(w.sc+w.mag+w.mel)/dx."""
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    yy = y.detach().to(dtype)
    zero = torch.zeros((), dtype=dtype, device=xx.device)
    sc, mag = mr_stft_loss(xx, yy, stft_res, use_torch_stft) if stft_res else (zero, zero)
    mel = multi_mel_loss(xx, yy, mel_res, use_torch_stft) if mel_res else zero
    total = weights[0] * sc + weights[1] * mag + weights[2] * mel
    (g,) = torch.autograd.grad(total, xx)
    return (float(sc.detach()), float(mag.detach()), float(mel.detach())), g


# --------------------------------------------------------------------------------------
# analytic backward, independent of autograd; numpy, any float dtype
# --------------------------------------------------------------------------------------
def _np_frames(x: np.ndarray, n_fft, hop, win_length, window):
    half = n_fft // 2
    xp = np.pad(x, ((0, 0), (half, half)), mode="reflect")
    nfr = 1 + x.shape[1] // hop
    idx = (np.arange(nfr) * hop)[:, None] + np.arange(n_fft)[None, :]
    w = padded_window(window, win_length, n_fft, torch.float64).numpy().astype(x.dtype)
    return xp[:, idx] * w, w


def _adjoint_frames(gf: np.ndarray, w, n_fft, hop, t_len):
    """window, overlap-add into the padded signal, then fold the reflect margins (A.2 step 6)."""
    half = n_fft // 2
    b, nfr, _ = gf.shape
    gxp = np.zeros((b, t_len + 2 * half), dtype=gf.dtype)
    gfw = gf * w
    for t in range(nfr):
        gxp[:, t * hop:t * hop + n_fft] += gfw[:, t]
    gx = gxp[:, half:half + t_len].copy()
    for j in range(half):                       # left margin: padded j <- x[half - j]
        gx[:, half - j] += gxp[:, j]
    for m in range(1, half + 1):                # right margin: padded half+T-1+m <- x[T-1-m]
        gx[:, t_len - 1 - m] += gxp[:, half + t_len - 1 + m]
    return gx


def _adjoint_rfft(gre: np.ndarray, gim: np.ndarray, n_fft: int) -> np.ndarray:
    """gf[n] = sum_k gre[k] cos(2 pi k n/N) - gim[k] sin(2 pi k n/N)  (A.2 step 5)."""
    g = (gre + 1j * gim).astype(np.complex128 if gre.dtype == np.float64 else np.complex64)
    g = g.copy()
    g[..., 1:n_fft // 2] *= 0.5
    g[..., 0] = g[..., 0].real
    g[..., n_fft // 2] = g[..., n_fft // 2].real
    return (np.fft.irfft(g, n=n_fft, axis=-1) * n_fft).astype(gre.dtype)


def analytic(x, y, stft_res, mel_res, g=(1.0, 1.0, 1.0), dtype=np.float64):
    """Losses and dL/dx from the closed-form spec.  Returns ((sc, mag, mel), dx, partial_sums).

    partial_sums: list of per-transform tuples -- (S1, S2, S3, n) for STFT resolutions,
    (S4, n) for mel resolutions -- the quantities the CUDA path all-reduces.
    """
    x = np.asarray(_flat(torch.as_tensor(x)).numpy(), dtype=dtype)
    y = np.asarray(_flat(torch.as_tensor(y)).numpy(), dtype=dtype)
    t_len = x.shape[1]
    dx = np.zeros_like(x)
    sc = mag = mel = 0.0
    sums = []
    nres = len(stft_res)
    for r in stft_res:
        fx, w = _np_frames(x, r.fft_size, r.hop_size, r.win_length, r.window)
        fy, _ = _np_frames(y, r.fft_size, r.hop_size, r.win_length, r.window)
        sx, sy = np.fft.rfft(fx, axis=-1), np.fft.rfft(fy, axis=-1)
        px, py = sx.real ** 2 + sx.imag ** 2, sy.real ** 2 + sy.imag ** 2
        ax, ay = np.sqrt(np.maximum(px, r.eps)), np.sqrt(np.maximum(py, r.eps))
        s1, s2 = ((ay - ax) ** 2).sum(), (ay ** 2).sum()
        s3, n = np.abs(np.log(ay) - np.log(ax)).sum(), ax.size
        sums.append((float(s1), float(s2), float(s3), n))
        d, ny = math.sqrt(s1), math.sqrt(s2)
        sc += d / ny / nres
        mag += s3 / n / nres
        ga = np.zeros_like(ax)
        if d > 0:
            ga += (g[0] / nres) * (ax - ay) / (d * ny)
        ga += (g[1] / nres) * np.sign(np.log(ax) - np.log(ay)) / (n * ax)
        gate = (px >= r.eps) / ax
        dx += _adjoint_frames(_adjoint_rfft(ga * gate * sx.real, ga * gate * sx.imag, r.fft_size),
                              w, r.fft_size, r.hop_size, t_len)
    nmel = len(mel_res)
    for r in mel_res:
        wm = melmat(r, torch.float64).numpy().astype(dtype)            # (K, M)
        fx, w = _np_frames(x, r.fft_size, r.hop_size, r.win(), r.window)
        fy, _ = _np_frames(y, r.fft_size, r.hop_size, r.win(), r.window)
        sx, sy = np.fft.rfft(fx, axis=-1), np.fft.rfft(fy, axis=-1)
        px, py = sx.real ** 2 + sx.imag ** 2, sy.real ** 2 + sy.imag ** 2
        ax, ay = np.sqrt(np.maximum(px, r.eps)), np.sqrt(np.maximum(py, r.eps))
        mx, my = ax @ wm, ay @ wm
        lnb = 1.0 if r.log_base is None else math.log(r.log_base)
        lx, ly = np.log(np.maximum(mx, r.eps)) / lnb, np.log(np.maximum(my, r.eps)) / lnb
        s4, n = np.abs(lx - ly).sum(), lx.size
        sums.append((float(s4), n))
        mel += s4 / n / nmel
        gl = (g[2] / nmel) * np.sign(lx - ly) / n
        gm = gl / (np.maximum(mx, r.eps) * lnb) * (mx >= r.eps)
        ga = gm @ wm.T
        gate = (px >= r.eps) / ax
        dx += _adjoint_frames(_adjoint_rfft(ga * gate * sx.real, ga * gate * sx.imag, r.fft_size),
                              w, r.fft_size, r.hop_size, t_len)
    return (sc, mag, mel), dx, sums


# --------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md 8d recipe) -- shared by tests and bench so both see the same data
# --------------------------------------------------------------------------------------
def synth_pair(batch: int, t_len: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """y = 0.1 N(0,1), y_hat = y + 0.05 N(0,1); both (B, 1, T) fp32 on CPU."""
    g = torch.Generator().manual_seed(seed)
    y = 0.1 * torch.randn(batch, 1, t_len, generator=g)
    y_hat = y + 0.05 * torch.randn(batch, 1, t_len, generator=g)
    return y_hat, y


# --------------------------------------------------------------------------------------
# waveform shape loss: losses/waveform_loss.py:15-75 (WaveformShapeLoss / MultiWindowShapeLoss)
# --------------------------------------------------------------------------------------
def shape_loss_and_grad(y_hat, y, winlens: Sequence[int], dtype=np.float64):
    """numpy restatement of MultiWindowShapeLoss.forward and of what autograd derives for it.

    waveform_loss.py:36-38: ys = MaxPool1d(w)(|y|) (kernel = stride = w, no padding, floor mode: T // w disjoint
    windows), loss_w = L1Loss()(ys_hat, ys) = mean |ys_hat - ys|;  :70-73: mean over the window lengths.
    Gradient: sign(ys_hat - ys) / count routed to the first maximum of |y_hat| in the window (max_pool1d backward),
    times sign(y_hat) there (abs backward)."""
    x = np.asarray(y_hat, dtype=dtype)
    shape = x.shape
    x = x.reshape(-1, shape[-1])
    t = np.asarray(y, dtype=dtype).reshape(x.shape)
    rows, t_len = x.shape
    loss = 0.0
    grad = np.zeros_like(x)
    for w in winlens:
        n = t_len // w
        ax = np.abs(x[:, :n * w]).reshape(rows, n, w)
        ay = np.abs(t[:, :n * w]).reshape(rows, n, w)
        px, py = ax.max(-1), ay.max(-1)
        loss += np.abs(px - py).mean() / len(winlens)
        arg = ax.argmax(-1)                                       # first maximum
        idx = arg + np.arange(n)[None, :] * w
        r = np.repeat(np.arange(rows)[:, None], n, axis=1)
        np.add.at(grad, (r, idx), np.sign(px - py) * np.sign(x[r, idx]) / (rows * n * len(winlens)))
    return float(loss), grad.reshape(shape)


# --------------------------------------------------------------------------------------
# Mel_L1 evaluation metric: mel_spectrogram.py:36-44 (duplicate: sandbox.py:183-191)
#   mel_spectrogram = torchaudio.transforms.MelSpectrogram(48000);  Mel_L1 = nn.L1Loss()(mel(pred), mel(target))
# The arithmetic lives in torchaudio (third party; requirements.txt pins torchaudio==2.1.1, the container has 2.11):
# transforms.MelSpectrogram defaults = Spectrogram(n_fft 400, win 400 periodic Hann, hop 200, pad 0, power 2,
# normalized False, center True, reflect) -> MelScale(128 mels, f_min 0, f_max sr/2, norm None, mel_scale "htk").
# Pinned by tests/golden/mel_l1_*.npz, produced by torchaudio itself (tests/golden/make_golden_mel_l1.py).
# --------------------------------------------------------------------------------------
def htk_fbanks(sample_rate: int, n_fft: int = 400, n_mels: int = 128, f_min: float = 0.0, f_max=None) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk"):
    fp32 throughout, as torchaudio computes it; (n_freqs, n_mels)."""
    import math

    f_max = float(sample_rate // 2) if f_max is None else float(f_max)
    all_freqs = torch.linspace(0, sample_rate // 2, n_fft // 2 + 1)
    m_pts = torch.linspace(2595.0 * math.log10(1.0 + f_min / 700.0), 2595.0 * math.log10(1.0 + f_max / 700.0), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    return torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


def power_mel(x: torch.Tensor, sample_rate: int = 48000, n_fft: int = 400, hop: int = 200, n_mels: int = 128) -> torch.Tensor:
    """MelSpectrogram(sample_rate)(x): (..., T) -> (..., n_mels, 1 + T // hop), in x's dtype (the fp32 filterbank and window
    are cast, which is what module.double() does to torchaudio's buffers)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    spec = torch.fft.rfft(frames(x2, n_fft, hop, n_fft, "hann_window"), dim=-1)          # (B, F, K)
    power = spec.real ** 2 + spec.imag ** 2
    mel = torch.matmul(power, htk_fbanks(sample_rate, n_fft, n_mels).to(device=x.device, dtype=x.dtype))     # (B, F, n_mels)
    return mel.transpose(1, 2).reshape(lead + (n_mels, mel.shape[1]))


def mel_l1(pred: torch.Tensor, target: torch.Tensor, sample_rate: int = 48000) -> torch.Tensor:
    """Mel_L1(pred, target) = mean |MelSpectrogram(pred) - MelSpectrogram(target)|."""
    return (power_mel(pred, sample_rate) - power_mel(target, sample_rate)).abs().mean()
