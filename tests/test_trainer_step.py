"""BASELINE configs[2] / SURVEY 8(f1): the reference's own denoise trainer (trainer/denoise.py:52-84, the symAD generator of
config/denoise/symAD_vctk_48000_hop300.yaml, frozen quantizer + decoder) driven with the reference's criteria and with
this repo's drop-in criteria, same seed, same batches: the per-step losses the trainer records must track each other.
The reference code comes from /root/reference here and from its staged copy oracle/_ref/ on the GPU box
(oracle/make_ref.sh); without either the tests skip."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

needs_trainer = pytest.mark.skipif(not ref_loader.trainer_available(),
                                   reason="reference trainer neither mounted nor staged (oracle/make_ref.sh)")

TRACK_RTOL = 1e-4          # losses of the two runs track <= 1e-4 relative


@needs_trainer
def test_reference_trainer_runs_on_cpu_with_reference_criteria():
    """The harness itself (CPU, tiny): one _train_step of the unmodified trainer, mel + MR-STFT enabled."""
    from oracle import trainer_harness as th

    ns = th.load()
    tr = th.build_trainer(ns, ns.MultiMelSpectrogramLoss, ns.MultiResolutionSTFTLoss, torch.device("cpu"), seed=0)
    before = [p.detach().clone() for p in tr.model["generator"].encoder.parameters()]
    rows = th.run_steps(tr, th.synthetic_batches(1, 1, 2400, seed=3))
    assert rows[0]["mel_loss"] > 0 and rows[0]["spectral_convergence_loss"] > 0 and rows[0]["log_stft_magnitude_loss"] > 0
    assert tr.steps == 1
    after = list(tr.model["generator"].encoder.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, after))            # the encoder trains ...
    assert all(not p.requires_grad for p in tr.model["generator"].decoder.parameters())   # ... the decoder is frozen


@pytest.mark.gpu
@needs_trainer
@pytest.mark.parametrize("use_stft", [True, False], ids=["mel+stft", "mel-only-as-shipped"])
def test_trainer_step_losses_track_reference_criteria(use_stft):
    """configs[2]: batch 32 x 0.5 s @ 48 kHz through Trainer._train_step, 4 optimiser steps.  Run A: reference criteria on
    the GPU (torch.stft/cuFFT + ATen + autograd).  Run B: this repo's criteria (libspecloss.so).  Same initial weights,
    same batches: every recorded loss of every step agrees to 1e-4 relative."""
    import dl_speech_enhancement_b200 as pkg
    from oracle import trainer_harness as th

    dev = torch.device("cuda:0")
    ns = th.load()
    batches = th.synthetic_batches(4, 32, 24000, seed=11, device="cpu")
    tr_ref = th.build_trainer(ns, ns.MultiMelSpectrogramLoss, ns.MultiResolutionSTFTLoss, dev, seed=5, use_stft=use_stft)
    init = {k: v.detach().cpu().clone() for k, v in tr_ref.model["generator"].state_dict().items()}
    tr_our = th.build_trainer(ns, pkg.MultiMelSpectrogramLoss, pkg.MultiResolutionSTFTLoss, dev, seed=5, use_stft=use_stft,
                              init_state=init)
    assert type(tr_our.criterion["mel"]).__module__.startswith("dl_speech_enhancement_b200")
    rows_ref = th.run_steps(tr_ref, batches)
    rows_our = th.run_steps(tr_our, batches)
    keys = ["mel_loss", "generator_loss"] + (["spectral_convergence_loss", "log_stft_magnitude_loss"] if use_stft else [])
    worst = 0.0
    for step, (a, b) in enumerate(zip(rows_ref, rows_our)):
        for k in keys:
            rel = abs(a[k] - b[k]) / abs(a[k])
            worst = max(worst, rel)
            assert rel <= TRACK_RTOL, (step, k, a[k], b[k], rel)
    # step 0 runs on identical weights: there the criteria alone are compared (fp32 evaluation noise only)
    for k in keys:
        assert abs(rows_ref[0][k] - rows_our[0][k]) <= 1e-5 * abs(rows_ref[0][k]), (k, rows_ref[0][k], rows_our[0][k])
    # The weights the two runs end with: Adam divides by sqrt(v), so parameters whose gradient is at the fp32 noise level
    # (biases of the strided convolutions) take near-random +-lr steps in ANY two fp32 evaluations; the bound is therefore
    # on the scale of the weight movement itself (4 steps x lr 1e-4 on weights of magnitude ~5e-2), not on rounding.
    pa = torch.cat([p.detach().flatten() for p in tr_ref.model["generator"].encoder.parameters()])
    pb = torch.cat([p.detach().flatten() for p in tr_our.model["generator"].encoder.parameters()])
    p0 = torch.cat([init[k].flatten() for k, _ in tr_ref.model["generator"].encoder.named_parameters(prefix="encoder")]).to(dev)
    moved = float((pa - p0).norm() / p0.norm())
    apart = float((pa - pb).norm() / pa.norm())
    print(f"trainer step ({'mel+stft' if use_stft else 'mel'}): worst relative loss deviation over 4 steps {worst:.2e}; "
          f"encoder weights moved {moved:.2e} (relative), runs ended {apart:.2e} apart")
    assert apart <= 0.1 * moved, (apart, moved)
