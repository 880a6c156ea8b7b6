#!/usr/bin/env python3
"""BASELINE configs[2] / SURVEY 8(f1) measurement: trainer.denoise.Trainer._train_step (symAD_vctk_48000_hop300 encoder,
frozen decoder, batch 32 x 0.5 s @ 48 kHz) on one B200, with the reference's criteria and with this repo's criteria swapped
in; the spectral losses' share of the step ("dominates the training step" was never measured).
Prints one JSON object.  Uses oracle/ (harness + the reference's own code): measurement infrastructure, not product."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def crit_ms(mel, stft, dev, batch, length, steps=30, warmup=5):
    """fwd+bwd of the criteria alone on a (B,1,T) pair, as _metric_loss + backward use them; device time (CUDA events)."""
    g = torch.Generator(device=dev).manual_seed(3)
    y = 0.1 * torch.randn(batch, 1, length, device=dev, generator=g)
    x = (y + 0.05 * torch.randn(batch, 1, length, device=dev, generator=g)).requires_grad_(True)

    def step():
        x.grad = None
        loss = mel(x, y) * 45.0
        if stft is not None:
            sc, mag = stft(x, y)
            loss = loss + 45.0 * sc + 45.0 * mag
        loss.backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def gan_trainers(pkg, th, ns, dev):
    """SURVEY 8(f1), the other two trainers at their shipped batch (16 x 0.2 s): trainer.autoencoder (stage 1 metric-only and
    the adversarial stage) and trainer.vocoder (adversarial), mel + MR-STFT + shape criteria on, reference's vs this repo's."""
    out = {"config": "trainer.autoencoder / trainer.vocoder Trainer._train_step at the shipped batch 16 x 9600, mel + MR-STFT + "
                     "shape criteria enabled; wall clock per step incl. the trainer's .item() syncs",
           "gpu": torch.cuda.get_device_name(0)}
    n_steps, warm = 25, 5
    g = torch.Generator().manual_seed(2)
    batches = [(0.1 * torch.randn(16, 1, 9600, generator=g)).pin_memory() for _ in range(n_steps)]
    cases = (("autoencoder_metric_stage", th.build_autoencoder_trainer, "autoencoder/symAD_vctk_48000_hop300", False),
             ("autoencoder_adversarial_stage", th.build_autoencoder_trainer, "autoencoder/symAD_vctk_48000_hop300", True),
             ("vocoder_hifigan", th.build_vocoder_trainer, "vocoder/AudioDec_v1_symAD_vctk_48000_hop300_clean", True),
             ("vocoder_univnet", th.build_vocoder_trainer, "vocoder/AudioDec_v3_symADuniv_vctk_48000_hop300_clean", True))
    for tag, build, config, adv in cases:
        res = {"yaml": config}
        for name, classes in (("reference_criteria", (ns.MultiMelSpectrogramLoss, ns.MultiResolutionSTFTLoss, ns.MultiWindowShapeLoss)),
                              ("b200_criteria", (pkg.MultiMelSpectrogramLoss, pkg.MultiResolutionSTFTLoss, pkg.MultiWindowShapeLoss))):
            tr = build(ns, *classes, dev, config=config, seed=5, adversarial=adv)
            res[name] = {"step_ms": th.time_steps(tr, batches, warmup=warm)}
            del tr
            torch.cuda.empty_cache()
        res["step_speedup"] = res["reference_criteria"]["step_ms"] / res["b200_criteria"]["step_ms"]
        out[tag] = res
    return out


def main():
    import dl_speech_enhancement_b200 as pkg
    from oracle import trainer_harness as th

    dev = torch.device("cuda:0")
    ns = th.load()
    if os.environ.get("PROF_TRAINER", "denoise") != "denoise":
        print(json.dumps(gan_trainers(pkg, th, ns, dev), indent=1))
        return
    batch, length, n_steps, warm = 32, 24000, 25, 5
    batches = th.synthetic_batches(n_steps, batch, length, seed=11, device="cpu")
    pinned = [(a.pin_memory(), b.pin_memory()) for a, b in batches]
    out = {"config": "configs[2]: trainer.denoise.Trainer._train_step, symAD_vctk_48000_hop300 generator (encoder trains, "
                     "quantizer+decoder frozen), batch 32 x 0.5 s @ 48 kHz, Adam; host batches in pinned memory",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "steps": n_steps - warm, "warmup": warm}
    for use_stft in (False, True):
        tag = "mel+stft" if use_stft else "mel_only_as_shipped"
        res = {}
        for name, mel_cls, stft_cls in (("reference_criteria", ns.MultiMelSpectrogramLoss, ns.MultiResolutionSTFTLoss),
                                        ("b200_criteria", pkg.MultiMelSpectrogramLoss, pkg.MultiResolutionSTFTLoss)):
            tr = th.build_trainer(ns, mel_cls, stft_cls, dev, seed=5, use_stft=use_stft)
            step_ms = th.time_steps(tr, pinned, warmup=warm)
            c_ms = crit_ms(tr.criterion["mel"], tr.criterion.get("stft"), dev, batch, length)
            res[name] = {"step_ms": step_ms, "criteria_fwd_bwd_ms": c_ms, "criteria_share_of_step": c_ms / step_ms}
            del tr
            torch.cuda.empty_cache()
        res["step_speedup"] = res["reference_criteria"]["step_ms"] / res["b200_criteria"]["step_ms"]
        res["criteria_speedup"] = res["reference_criteria"]["criteria_fwd_bwd_ms"] / res["b200_criteria"]["criteria_fwd_bwd_ms"]
        out[tag] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
