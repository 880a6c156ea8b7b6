"""CPU-only: the kernels' device code, executed lane-accurately by the SIMT emulator through the
real C-ABI host logic and the real Python driver, against the golden vectors of the reference."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, plans_for, rel_l2, run_losses

SMALL = [n for n in golden_names() if n not in ("c1_clean1_noise1", "gauss_b2_t16000")]

LOSS_RTOL = 1e-4      # north star: loss values within 1e-4 relative
GRAD_RTOL = 1e-3      # north star: waveform gradients within 1e-3 relative (rel-L2)


@pytest.mark.parametrize("name", SMALL)
def test_emulated_kernels_match_reference(emu_engine, name):
    g = load_golden(name)
    vals, grad = run_losses(emu_engine, g)
    for i in range(3):
        assert abs(vals[i] - g["loss32"][i]) <= LOSS_RTOL * max(abs(g["loss32"][i]), 1e-12)
        assert abs(vals[i] - g["loss64"][i]) <= LOSS_RTOL * max(abs(g["loss64"][i]), 1e-12)
    grad = grad.reshape(g["grad32"].shape)
    assert rel_l2(grad, g["grad32"]) <= GRAD_RTOL
    assert rel_l2(grad, g["grad64"]) <= GRAD_RTOL


def test_emulated_weighted_backward(emu_engine):
    """Upstream gradients (lambda weights, trainerGAN.py:221,228-229) scale the three terms independently."""
    from oracle import spectral_oracle as so

    g = load_golden("gauss_b2_t4800")
    w = (45.0, 0.5, -2.0)
    _, grad = run_losses(emu_engine, g, weights=w)
    _, ref = so.losses_and_grad(g["y_hat"], g["y"], so.stft_from_kwargs(**g["stft_kwargs"]),
                                so.mel_from_kwargs(**g["mel_kwargs"]), weights=w, dtype=torch.float64)
    assert rel_l2(grad, ref.numpy().reshape(grad.shape)) <= GRAD_RTOL


@pytest.mark.parametrize("n_fft,hop,win,t_len", [(512, 50, 240, 777), (1024, 120, 600, 1500), (2048, 300, 2048, 2500),
                                                 (1024, 256, 1024, 1300), (512, 128, 512, 300)])
def test_emulated_spectrogram_matches_torch_stft(emu_engine, n_fft, hop, win, t_len):
    """stft() of the reference (stft_loss.py:19-35): sqrt(clamp(|torch.stft|^2, eps)) transposed to (B, F, K)."""
    from dl_speech_enhancement_b200.engine import twiddle_table

    g = torch.Generator().manual_seed(n_fft + t_len)
    x = 0.1 * torch.randn(3, t_len, generator=g)
    x[1, :200] = 0.0                                        # exact zeros: the clamp floor
    window = torch.hann_window(win)
    out = emu_engine.spectrogram(x, n_fft, hop, win, window, twiddle_table(n_fft), 1e-7)
    ref = torch.stft(x.double(), n_fft, hop, win, window.double(), return_complex=True)
    ref = torch.sqrt(torch.clamp(ref.real ** 2 + ref.imag ** 2, min=1e-7)).transpose(2, 1)
    assert out.shape == ref.shape
    assert rel_l2(out.numpy(), ref.numpy()) <= 2e-6
    assert float((out.double() - ref).abs().max()) <= 2e-5 * float(ref.max())


def test_emulated_spectrogram_tf32_split(emu_engine):
    """The GEMM operand form of the spectrogram: hi exactly TF32, hi + lo == the plain output, zero pad columns."""
    from dl_speech_enhancement_b200.engine import gemm_ld, twiddle_table

    x = 0.1 * torch.randn(2, 900, generator=torch.Generator().manual_seed(3))
    window, tw = torch.hann_window(512), twiddle_table(512)
    plain = emu_engine.spectrogram(x, 512, 128, 512, window, tw, 1e-10)
    hi, lo = emu_engine.spectrogram(x, 512, 128, 512, window, tw, 1e-10, ld=gemm_ld(512), split=True)
    assert hi.shape == (2, 8, 288) and lo.shape == hi.shape
    assert torch.equal(hi[:, :, :257] + lo[:, :, :257], plain)
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert float(hi[:, :, 257:].abs().max()) == 0.0 and float(lo[:, :, 257:].abs().max()) == 0.0
    assert float((lo.abs() / plain.abs().max()).max()) < 2.0 ** -10


def test_emulated_no_grad(emu_engine):
    from dl_speech_enhancement_b200.functional import spectral_losses

    g = load_golden("minlen_b1_t1025")
    with torch.no_grad():
        outs = spectral_losses(g["y_hat"], g["y"], plans_for(g), engine=emu_engine)
    np.testing.assert_allclose([float(o) for o in outs], g["loss64"], rtol=LOSS_RTOL)
    assert all(not o.requires_grad for o in outs)


def test_emulated_partial_sums_match_oracle(emu_engine):
    """The all-reduced quantities themselves, not only the finished losses."""
    from oracle import spectral_oracle as so

    g = load_golden("gauss_b2_t4800")
    plans = plans_for(g)
    st = emu_engine.forward(plans, g["y_hat"].reshape(-1, 4800), g["y"].reshape(-1, 4800), need_grad=False)
    _, _, sums = so.analytic(g["y_hat"], g["y"], so.stft_from_kwargs(**g["stft_kwargs"]),
                             so.mel_from_kwargs(**g["mel_kwargs"]))
    flat = [v for s in sums for v in s[:-1]]
    np.testing.assert_allclose(st.sums.numpy(), flat, rtol=2e-5)


def test_emulated_identical_signals_are_exactly_zero(emu_engine):
    """x == y: the reference returns 0 losses and a 0 gradient (norm backward at 0, sign(0)); the packed
    FFT must not leak rounding asymmetry into that case."""
    from dl_speech_enhancement_b200.functional import spectral_losses

    g = load_golden("gauss_b2_t4800")
    x = g["y"].clone().requires_grad_(True)
    outs = spectral_losses(x, g["y"], plans_for(g), engine=emu_engine)
    sum(outs).backward()
    assert [float(o.detach()) for o in outs] == [0.0, 0.0, 0.0]
    assert torch.count_nonzero(x.grad) == 0


def _ref_spectrogram(x, n_fft, hop, win, window, eps):
    s = torch.stft(x, n_fft, hop, win, window.to(x.dtype), return_complex=True)
    return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=eps)).transpose(2, 1)


@pytest.mark.parametrize("n_fft,hop,win,t_len", [(512, 50, 240, 777), (1024, 120, 600, 1500), (2048, 300, 2048, 2500),
                                                 (1024, 256, 1024, 1301), (512, 128, 512, 300)])
def test_emulated_spectrogram_backward_matches_autograd(emu_engine, n_fft, hop, win, t_len):
    """d/dx of stft() (stft_loss.py:19-35) for an arbitrary upstream gradient, against torch autograd in fp64;
    odd frame counts (last pair half empty), reflect margins and the clamp gate (exact-zero block) included."""
    from dl_speech_enhancement_b200._abi import SPL_KIND_STFT
    from dl_speech_enhancement_b200.engine import TransformPlan, twiddle_table
    from dl_speech_enhancement_b200.functional import spectrogram

    gen = torch.Generator().manual_seed(n_fft + t_len)
    x = 0.1 * torch.randn(3, t_len, generator=gen)
    x[1, :] = 0.0                                           # every bin clamped: zero gradient rows
    window = torch.hann_window(win)
    plan = TransformPlan(SPL_KIND_STFT, n_fft, hop, win, 1e-7, window, twiddle_table(n_fft))
    xg = x.clone().requires_grad_(True)
    out = spectrogram(xg, plan, engine=emu_engine)
    gout = torch.randn(out.shape, generator=gen)
    (out * gout).sum().backward()
    xr = x.double().requires_grad_(True)
    (_ref_spectrogram(xr, n_fft, hop, win, window, 1e-7) * gout.double()).sum().backward()
    assert torch.count_nonzero(xg.grad[1]) == 0
    assert rel_l2(xg.grad.numpy(), xr.grad.numpy()) <= 1e-5


@pytest.mark.parametrize("kw,t_len", [(dict(fs=48000, fft_size=2048, hop_size=300, win_length=None, num_mels=80, fmin=0,
                                            fmax=24000, log_base=None), 3100),
                                      (dict(fs=24000, fft_size=1024, hop_size=256, num_mels=80, fmin=80, fmax=7600,
                                            log_base=10.0), 1500),
                                      (dict(fs=24000, fft_size=512, hop_size=120, win_length=400, num_mels=40, fmin=0,
                                            fmax=24000, log_base=2.0), 900)])
def test_emulated_logmel_backward_matches_autograd(emu_engine, kw, t_len):
    """d/dx of MelSpectrogram.forward (mel_loss.py:74-94) for an arbitrary upstream gradient, fp64 autograd reference
    (includes empty filters above Nyquist -> clamped mel energies -> zero gradient through them)."""
    import math

    from dl_speech_enhancement_b200 import modules

    mod = modules.MelSpectrogram(**kw)
    gen = torch.Generator().manual_seed(t_len)
    x = 0.1 * torch.randn(2, t_len, generator=gen)
    frames = 1 + t_len // mod.hop_size
    g = torch.randn(2, mod.num_mels, frames, generator=gen)
    dx = emu_engine.spectrogram_backward(mod.plan(), x, g)
    xr = x.double().requires_grad_(True)
    amp = _ref_spectrogram(xr, mod.fft_size, mod.hop_size, mod.win_length, mod.window, mod.eps)
    mel = torch.clamp(torch.matmul(amp, mod.melmat.double()), min=mod.eps)
    logmel = torch.log(mel) / (1.0 if mod.log_base is None else math.log(mod.log_base))
    (logmel.transpose(1, 2) * g.double()).sum().backward()
    assert rel_l2(dx.numpy(), xr.grad.numpy()) <= 1e-5


def test_emulated_plain_magnitude_eps0(emu_engine):
    """eps = 0 (torchaudio power=1.0, the UnivNet discriminator front-end, discriminator.py:556-565): |X| exactly 0 and
    a zero gradient on silent input (torch.abs backward at 0), no NaN from rsqrt(0)."""
    from dl_speech_enhancement_b200._abi import SPL_KIND_STFT
    from dl_speech_enhancement_b200.engine import TransformPlan, twiddle_table
    from dl_speech_enhancement_b200.functional import spectrogram

    gen = torch.Generator().manual_seed(5)
    x = 0.1 * torch.randn(3, 1300, generator=gen)
    x[1] = 0.0
    window = torch.hann_window(600)
    plan = TransformPlan(SPL_KIND_STFT, 1024, 120, 600, 0.0, window, twiddle_table(1024))
    xg = x.clone().requires_grad_(True)
    out = spectrogram(xg, plan, engine=emu_engine)
    gout = torch.randn(out.shape, generator=gen)
    (out * gout).sum().backward()
    xr = x.double().requires_grad_(True)
    ref = torch.stft(xr, 1024, 120, 600, window.double(), return_complex=True).abs().transpose(2, 1)
    (ref * gout.double()).sum().backward()
    assert torch.count_nonzero(out[1]) == 0 and torch.count_nonzero(xg.grad[1]) == 0
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(xg.grad).all())
    assert rel_l2(out.detach().numpy(), ref.detach().numpy()) <= 2e-6
    assert rel_l2(xg.grad.numpy(), xr.grad.numpy()) <= 1e-5


def _ref_shape_loss(y_hat, y, winlens):
    """losses/waveform_loss.py:15-75 restated with the same torch ops (MaxPool1d of |.|, L1Loss, mean over windows)."""
    loss = 0.0
    for w in winlens:
        pool = torch.nn.MaxPool1d(w)
        loss = loss + torch.nn.functional.l1_loss(pool(torch.abs(y_hat)), pool(torch.abs(y)))
    return loss / len(winlens)


@pytest.mark.parametrize("winlens,shape", [([300, 200, 100], (3, 1, 5003)), ([64], (2, 1, 2048)), ([7, 2500], (2, 2, 4100)),
                                           ([1], (1, 1, 300)), ([300, 200, 100], (2, 1, 6000)), ([400, 200], (2, 1, 4800)),
                                           ([96, 36], (1, 2, 3001))])
def test_emulated_shape_loss_matches_reference_ops(emu_engine, winlens, shape):
    from dl_speech_enhancement_b200.functional import shape_loss

    gen = torch.Generator().manual_seed(sum(winlens))
    y = 0.1 * torch.randn(*shape, generator=gen)
    y_hat = y + 0.05 * torch.randn(*shape, generator=gen)
    y_hat[0, 0, :600] = y[0, 0, :600]                       # equal windows: sign(0) = 0
    y_hat[-1, -1, 700:1400] = 0.0                           # silent prediction: abs'(0) = 0, first index is the argmax
    xg = y_hat.clone().requires_grad_(True)
    loss = shape_loss(xg, y, winlens, engine=emu_engine)
    (3.0 * loss).backward()
    xr = y_hat.double().requires_grad_(True)
    ref = _ref_shape_loss(xr, y.double(), winlens)
    (3.0 * ref).backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-6 * abs(float(ref.detach()))
    assert xg.grad.shape == xg.shape
    np.testing.assert_allclose(xg.grad.numpy(), xr.grad.numpy(), rtol=1e-6, atol=1e-8)   # atol: fp32 cancellation between window lengths


@pytest.mark.parametrize("name", __import__("conftest").shape_golden_names())
def test_emulated_shape_loss_matches_golden(emu_engine, name):
    from conftest import load_shape_golden
    from dl_speech_enhancement_b200.functional import shape_loss

    g = load_shape_golden(name)
    x = g["y_hat"].clone().requires_grad_(True)
    loss = shape_loss(x, g["y"], g["winlens"], engine=emu_engine)
    loss.backward()
    assert abs(float(loss.detach()) - g["loss64"]) <= 1e-6 * abs(g["loss64"])
    np.testing.assert_allclose(x.grad.numpy(), g["grad32"], rtol=1e-6, atol=1e-9)


def test_emulated_explicit_magnitude_losses(emu_engine):
    """SpectralConvergenceLoss / LogSTFTMagnitudeLoss on explicit tensors (stft_loss.py:38-77): values and the gradients
    w.r.t. BOTH arguments against torch autograd in fp64."""
    from dl_speech_enhancement_b200.functional import magnitude_loss

    gen = torch.Generator().manual_seed(8)
    x = torch.rand(3, 37, 129, generator=gen) + 0.05
    y = torch.rand(3, 37, 129, generator=gen) + 0.05
    y[0, :5] = x[0, :5]                                      # equal entries: sign(0) = 0
    for which in (0, 1):
        xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        loss = magnitude_loss(xg, yg, which, engine=emu_engine)
        (2.5 * loss).backward()
        xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
        ref = (torch.norm(yr - xr, p="fro") / torch.norm(yr, p="fro")) if which == 0 else \
            torch.nn.functional.l1_loss(torch.log(yr), torch.log(xr))
        (2.5 * ref).backward()
        assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-6 * float(ref.detach())
        assert rel_l2(xg.grad.numpy(), xr.grad.numpy()) <= 1e-6 and rel_l2(yg.grad.numpy(), yr.grad.numpy()) <= 1e-6
    same = magnitude_loss(x.clone().requires_grad_(True), x, 0, engine=emu_engine)
    assert float(same.detach()) == 0.0


def test_eight_resolutions_in_one_call(emu_engine):
    """SPL_MAX_TRANSFORMS = 8 STFT resolutions = 24 sums in one call (the reduction and the exchange slot are sized for
    3 x SPL_MAX_TRANSFORMS); a ninth raises."""
    import torch

    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    res = [(512, 50 + 10 * i, 240) for i in range(8)]
    crit = modules.MultiResolutionSTFTLoss([r[0] for r in res], [r[1] for r in res], [r[2] for r in res])
    y_hat, y = so.synth_pair(1, 700, seed=5)
    x = y_hat.clone().requires_grad_(True)
    sc, mag = spectral_losses(x, y, crit.plans(), engine=emu_engine)
    (sc + mag).backward()
    ref, gref = so.losses_and_grad(y_hat, y, [so.StftRes(*r) for r in res], None, dtype=torch.float64)
    np.testing.assert_allclose([float(sc), float(mag)], ref[:2], rtol=1e-4)
    assert rel_l2(x.grad.numpy(), gref.numpy()) <= 1e-3
    nine = modules.MultiResolutionSTFTLoss([512] * 9, [50] * 9, [240] * 9)
    with pytest.raises(RuntimeError, match="resolutions"):
        spectral_losses(x, y, nine.plans(), engine=emu_engine)


def test_second_backward_needs_retain_graph(emu_engine):
    """The gradient workspace is a saved tensor: backward(retain_graph=True) allows a second backward with the same result
    (the reference's autograd graph behaves so), and a second backward without it raises torch's usual error."""
    import torch

    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    crit = modules.MultiResolutionSTFTLoss([512], [50], [240])
    y_hat, y = so.synth_pair(1, 700, seed=6)
    x = y_hat.clone().requires_grad_(True)
    sc, mag = spectral_losses(x, y, crit.plans(), engine=emu_engine)
    (sc + mag).backward(retain_graph=True)
    g1 = x.grad.clone()
    x.grad = None
    (sc + mag).backward()
    assert torch.equal(g1, x.grad)
    with pytest.raises(RuntimeError, match="second time|already been freed"):
        (sc + mag).backward()


def test_shared_memory_race_check_under_thread_sanitizer():
    """compute-sanitizer's racecheck is not available on this pool; the emulator's lanes are real threads and its barriers
    real barriers, so ThreadSanitizer over the emulated kernels is the shared-memory race detector (tests/emu/
    tsan_race_check.py: clean run must report nothing, a run with one __syncwarp() skipped must be caught)."""
    import os
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "tsan_race_check.py")
    res = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=1500)
    if res.returncode == 2:
        pytest.skip("libtsan / g++ -fsanitize=thread not available")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "data-race reports = 0" in res.stdout.splitlines()[0]


@pytest.mark.parametrize("win,hop,t_len", [(1000, 256, 3001), (1023, 200, 2500), (2048, 512, 2200), (1200, 240, 2047)])
def test_even_odd_2048_generic_windows_and_odd_lengths(emu_engine, win, hop, t_len, monkeypatch):
    """The 2048-point losses on the 32 x 32 geometry (csrc/transform_eo.cuh): window lengths without a compile-time
    kernel, an ODD window (the (even, odd) sample pairs then straddle the window's centring offset: scalar tap path), odd
    utterance lengths (rows not 8-byte aligned: scalar loads) -- STFT and mel, forward and gradient, against the fp64
    oracle; and the two kernel families (even/odd vs 64 points per lane) agree with each other."""
    import torch

    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    monkeypatch.setenv("SPECLOSS_EO_2048", "1")          # STFT too (the product default takes this route for mel only)
    y_hat, y = so.synth_pair(2, t_len, seed=win)
    stft = modules.MultiResolutionSTFTLoss([2048], [hop], [win])
    mel_kw = dict(fs=48000, fft_sizes=[2048], hop_sizes=[hop], win_lengths=[win], num_mels=80, fmin=0, fmax=24000, log_base=10.0)
    mel = modules.MultiMelSpectrogramLoss(**mel_kw)
    plans = stft.plans() + mel.plans()
    assert all(p.twiddle_eo is not None for p in plans)
    x = y_hat.clone().requires_grad_(True)
    outs = spectral_losses(x, y, plans, engine=emu_engine)
    sum(outs).backward()
    ref, gref = so.losses_and_grad(y_hat, y, [so.StftRes(2048, hop, win)], so.mel_from_kwargs(**mel_kw), dtype=torch.float64)
    got = [float(o.detach()) for o in outs]
    np.testing.assert_allclose(got, ref, rtol=1e-4)
    assert rel_l2(x.grad.numpy(), gref.numpy()) <= 1e-3
    # the 64-point-per-lane kernels on the same input (plans without the even/odd tables)
    for p in plans:
        p.twiddle_eo = None
    x2 = y_hat.clone().requires_grad_(True)
    outs2 = spectral_losses(x2, y, plans, engine=emu_engine)
    sum(outs2).backward()
    np.testing.assert_allclose([float(o.detach()) for o in outs2], got, rtol=2e-6)
    assert rel_l2(x2.grad.numpy(), x.grad.numpy()) <= 2e-4


def test_even_odd_2048_identical_inputs_exact_zero(emu_engine, monkeypatch):
    """loss(x, x) = 0 and a zero gradient EXACTLY: both signals run through the same instruction sequence."""
    import torch

    monkeypatch.setenv("SPECLOSS_EO_2048", "1")

    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses

    x = (0.2 * torch.randn(1, 2600, generator=torch.Generator().manual_seed(9))).requires_grad_(True)
    stft = modules.MultiResolutionSTFTLoss([2048], [240], [1200])
    mel = modules.MultiMelSpectrogramLoss(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], num_mels=80, fmin=0,
                                          fmax=24000, log_base=None)
    outs = spectral_losses(x, x.detach().clone(), stft.plans() + mel.plans(), engine=emu_engine)
    sum(outs).backward()
    assert all(float(o.detach()) == 0.0 for o in outs)
    assert float(x.grad.abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["0", "1"], ids=["64-points-per-lane", "even-odd"])
@pytest.mark.parametrize("name", ["gauss_b2_t4800", "silence_b3_t6000"])
def test_both_2048_routes_match_reference(emu_engine, monkeypatch, name, mode):
    """The product default sends the 2048-point mel loss down the even/odd route and the 2048-point STFT loss down the
    64-point-per-lane route; both kernels exist for both losses (SPECLOSS_EO_2048 = 0 | mel | 1) and must all meet the
    reference (the golden-vector tests above cover the default mix)."""
    monkeypatch.setenv("SPECLOSS_EO_2048", mode)
    g = load_golden(name)
    vals, grad = run_losses(emu_engine, g)
    for i in range(3):
        assert abs(vals[i] - g["loss64"][i]) <= LOSS_RTOL * max(abs(g["loss64"][i]), 1e-12)
    assert rel_l2(grad.reshape(g["grad64"].shape), g["grad64"]) <= GRAD_RTOL


@pytest.mark.parametrize("run_frames", [2, 3, 5, 16])
@pytest.mark.parametrize("name", ["gauss_b2_t4800", "ragged_b3_t5003_2d", "minlen_b1_t1025"])
def test_overlap_add_ring_runs(emu_engine, monkeypatch, name, run_frames):
    """STFT gradient through the shared-memory overlap-add ring (a warp takes a run of `run_frames` consecutive frames and
    writes (run_frames - 1) * hop + win taps per run instead of run_frames * win): same losses bit for bit, gradient equal to
    the one-slot-per-frame path up to fp32 summation order, and within the reference tolerances.  Run lengths that do not
    divide the frame count, runs longer than an utterance (minlen: 3 frames at n_fft 2048), two frames per warp (n_fft 512)."""
    g = load_golden(name)
    monkeypatch.setenv("SPECLOSS_RUN_FRAMES", "1")
    base_vals, base_grad = run_losses(emu_engine, g)
    monkeypatch.setenv("SPECLOSS_RUN_FRAMES", str(run_frames))
    vals, grad = run_losses(emu_engine, g)
    assert vals == base_vals
    assert rel_l2(grad, base_grad) <= 2e-6
    assert rel_l2(grad.reshape(g["grad64"].shape), g["grad64"]) <= GRAD_RTOL
    # the workspace really shrank: taps per run vs taps per frame
    from dl_speech_enhancement_b200 import _abi
    tr = plans_for(g)[0]
    st = _abi.SplTransform()
    st.kind, st.n_fft, st.hop, st.win = tr.kind, tr.n_fft, tr.hop, tr.win
    geo = _abi.SplGeometry()
    b, t_len = g["y_hat"].reshape(-1, g["y_hat"].shape[-1]).shape
    assert emu_engine.lib.spl_geometry_of(st, b, t_len, geo) == 0
    frames = 1 + t_len // tr.hop
    m = min(run_frames, frames) if run_frames <= frames else 1
    runs = -(-frames // m)
    assert geo.gframe_bytes == b * runs * ((m - 1) * tr.hop + tr.win) * 8


@pytest.mark.parametrize("case", ["weak_white", "weak_lowpass", "loud", "silent_prediction"])
def test_weak_prediction_keeps_reference_accuracy(emu_engine, case):
    """The pair-packed STFT kernels equalise the levels of prediction and target per frame by an exact power of two
    (equalise_pair): with the prediction 40 dB below (or above) the target, the MR-STFT + default-resolution mel gradient
    stays as close to fp64 as the reference's fp32 op sequence (without it: 5e-5 ... 2e-2 against 1e-6, measured); an
    all-zero prediction (level statistic 0: no scaling) still works."""
    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    g = torch.Generator().manual_seed(77)
    y = 0.1 * torch.randn(2, 4800, generator=g)
    k = torch.hann_window(65)
    x = {"weak_white": 1e-3 * torch.randn(2, 4800, generator=g),
         "weak_lowpass": 1e-3 * torch.nn.functional.conv1d(torch.randn(2, 1, 4864, generator=g), (k / k.sum()).view(1, 1, -1)).squeeze(1),
         "loud": 10.0 * torch.randn(2, 4800, generator=g),
         "silent_prediction": torch.zeros(2, 4800)}[case]
    stft = modules.MultiResolutionSTFTLoss()
    mel = modules.MultiMelSpectrogramLoss()          # reference defaults: 1024 / 2048 / 512, pair-packed mel kernels too
    xx = x.clone().requires_grad_(True)
    outs = spectral_losses(xx, y, stft.plans() + mel.plans(), engine=emu_engine)
    sum(outs).backward()
    mel_res = so.mel_from_kwargs()
    l64, g64 = so.losses_and_grad(x, y, so.DEFAULT_STFT, mel_res, dtype=torch.float64, use_torch_stft=True)
    l32, g32 = so.losses_and_grad(x, y, so.DEFAULT_STFT, mel_res, dtype=torch.float32, use_torch_stft=True)
    np.testing.assert_allclose([float(o.detach()) for o in outs], l64, rtol=1e-5)
    e_ours, e_ref = rel_l2(xx.grad.numpy(), g64.numpy()), rel_l2(g32.numpy(), g64.numpy())
    assert e_ours <= max(3.0 * e_ref, 1e-5), (case, e_ours, e_ref)


@pytest.mark.parametrize("x_level,y_level", [(1e-30, 1.0), (1e3, 1e-20), (1e-12, 1e-12)])
def test_level_equalisation_extremes_stay_finite(emu_engine, x_level, y_level):
    """Level disparities far beyond anything audio produces: the power-of-two shift is clamped (2^+-96), nothing overflows,
    the losses equal the fp64 oracle's and, where every bin of the prediction sits below the clamp floor of stft(), the
    gradient is exactly zero as in the reference."""
    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    g = torch.Generator().manual_seed(5)
    x = x_level * torch.randn(2, 3000, generator=g)
    y = y_level * torch.randn(2, 3000, generator=g)
    stft = modules.MultiResolutionSTFTLoss()
    xx = x.clone().requires_grad_(True)
    outs = spectral_losses(xx, y, stft.plans(), engine=emu_engine)
    sum(outs).backward()
    l64, g64 = so.losses_and_grad(x, y, so.DEFAULT_STFT, None, dtype=torch.float64, use_torch_stft=True)
    got = [float(o.detach()) for o in outs]
    assert all(np.isfinite(got)) and bool(torch.isfinite(xx.grad).all())
    np.testing.assert_allclose(got, l64[:2], rtol=1e-4)
    if float(g64.abs().max()) == 0.0:
        assert float(xx.grad.abs().max()) == 0.0
    else:
        assert rel_l2(xx.grad.numpy(), g64.numpy()) <= GRAD_RTOL


@pytest.mark.parametrize("fft,hop,win,t_len", [(512, 300, 240, 3001), (1024, 700, 600, 4000), (2048, 2100, 1200, 9000),
                                               (1024, 1024, 1024, 5000)])
def test_hop_larger_than_window(emu_engine, fft, hop, win, t_len):
    """hop >= win_length (legal in torch.stft, hence in the reference modules): frames with gaps between them, samples no
    frame covers get zero gradient.  STFT and mel losses, forward and gradient, against the fp64 oracle."""
    from dl_speech_enhancement_b200 import modules
    from dl_speech_enhancement_b200.functional import spectral_losses
    from oracle import spectral_oracle as so

    y_hat, y = so.synth_pair(2, t_len, seed=hop)
    stft_kw = dict(fft_sizes=[fft], hop_sizes=[hop], win_lengths=[win], window="hann_window")
    mel_kw = dict(fs=24000, fft_sizes=[fft], hop_sizes=[hop], win_lengths=[win], window="hann_window", num_mels=40,
                  fmin=0, fmax=12000, log_base=None)
    plans = modules.MultiResolutionSTFTLoss(**stft_kw).plans() + modules.MultiMelSpectrogramLoss(**mel_kw).plans()
    x = y_hat.clone().requires_grad_(True)
    outs = spectral_losses(x, y, plans, engine=emu_engine)
    sum(outs).backward()
    ref, gref = so.losses_and_grad(y_hat, y, so.stft_from_kwargs(**stft_kw), so.mel_from_kwargs(**mel_kw), dtype=torch.float64,
                                   use_torch_stft=True)
    np.testing.assert_allclose([float(o.detach()) for o in outs], ref, rtol=LOSS_RTOL)
    assert rel_l2(x.grad.numpy(), gref.numpy().reshape(x.grad.shape)) <= GRAD_RTOL
