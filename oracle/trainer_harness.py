"""BASELINE configs[2] / SURVEY 8(f1): the reference's OWN denoise trainer step with a chosen pair of criteria.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Everything that runs is the reference's unmodified code
(trainer.denoise.Trainer._train_step, trainer/denoise.py:52-84 -> TrainerGAN._metric_loss, trainerGAN.py:214-241 ->
_update_generator, :271-281; models.autoencoder.AudioDec.Generator with the symAD_vctk_48000_hop300 hyper-parameters),
imported through oracle/ref_loader.py; the only thing swapped is what sits in criterion["mel"] / criterion["stft"]:
the reference's losses.MultiMelSpectrogramLoss / MultiResolutionSTFTLoss, or this repo's drop-in modules.  What the
entry scripts of the reference do around the trainer (build the dicts, bin/train.py:66-103, train_denoise.py:100-135) is
restated here in a dozen lines, because the concrete entry script is not part of the reference repo.
"""
from __future__ import annotations

import copy
import tempfile
import time

import torch

from oracle import ref_loader


class _NoTqdm:
    def update(self, n=1):
        pass

    def close(self):
        pass


def synthetic_batches(n, batch, length, seed, device="cpu"):
    """(x_noisy, x_clean) pairs shaped like dataloader/collater.py:57-60 output: (B, 1, T) fp32."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        clean = 0.1 * torch.randn(batch, 1, length, generator=g)
        noisy = clean + 0.05 * torch.randn(batch, 1, length, generator=g)
        out.append((noisy.to(device), clean.to(device)))
    return out


def build_trainer(ns, mel_cls, stft_cls, device, seed=0, use_stft=True, init_state=None):
    """A trainer.denoise.Trainer over a freshly initialised (seeded) symAD generator.  mel_cls / stft_cls are the criterion
    classes (reference's or this repo's); both take the YAML blocks verbatim."""
    cfg = copy.deepcopy(ns.config)
    cfg["use_mel_loss"] = True
    cfg["use_stft_loss"] = bool(use_stft)         # shipped YAML: false; flipping it is the MR-STFT route
    cfg["outdir"] = tempfile.mkdtemp(prefix="specloss_trainer_")
    cfg["train_max_steps"] = 1 << 30
    torch.manual_seed(seed)
    gen = ns.Generator(**cfg["generator_params"])
    if init_state is not None:
        gen.load_state_dict(init_state)
    gen = gen.to(device)
    criterion = {"mel": mel_cls(**cfg["mel_loss_params"]).to(device)}
    if use_stft:
        criterion["stft"] = stft_cls(**cfg["stft_loss_params"]).to(device)
    opt = torch.optim.Adam(gen.parameters(), **cfg["generator_optimizer_params"])
    sch = torch.optim.lr_scheduler.StepLR(opt, **cfg["generator_scheduler_params"])
    tr = ns.Trainer(steps=0, epochs=0, data_loader={}, model={"generator": gen}, criterion=criterion,
                    optimizer={"generator": opt}, scheduler={"generator": sch}, config=cfg, device=device)
    tr.tqdm = _NoTqdm()
    return tr


def _gan_parts(ns, cfg, device, mel_cls, stft_cls, shape_cls, use_stft, use_shape):
    """What the AudioDec entry scripts put around a TrainerGAN (criterion / discriminator / optimizer / scheduler dicts named as
    trainerGAN.py:214-300 reads them); the concrete scripts (codecTrain.py) are not part of the reference repo."""
    cfg["use_mel_loss"] = True
    cfg["use_stft_loss"] = bool(use_stft)
    cfg["use_shape_loss"] = bool(use_shape)
    cfg["outdir"] = tempfile.mkdtemp(prefix="specloss_trainer_")
    cfg["train_max_steps"] = 1 << 30
    disc_cls = ns.UnivNetDiscriminator if cfg["model_type"] in ("symAudioDecUniv", "UnivNet") else ns.HiFiGANDiscriminator
    disc = disc_cls(**cfg["discriminator_params"]).to(device)
    criterion = {"mel": mel_cls(**cfg["mel_loss_params"]).to(device),
                 "gen_adv": ns.GeneratorAdversarialLoss(**cfg["generator_adv_loss_params"]).to(device),
                 "dis_adv": ns.DiscriminatorAdversarialLoss(**cfg["discriminator_adv_loss_params"]).to(device)}
    if cfg.get("use_feat_match_loss", False):
        criterion["feat_match"] = ns.FeatureMatchLoss(**cfg.get("feat_match_loss_params", {})).to(device)
    if use_stft:
        criterion["stft"] = stft_cls(**cfg["stft_loss_params"]).to(device)
    if use_shape:
        criterion["shape"] = shape_cls(**cfg["shape_loss_params"]).to(device)
    return disc, criterion


def _optim(cfg, who, params):
    opt = getattr(torch.optim, cfg[f"{who}_optimizer_type"])(params, **cfg[f"{who}_optimizer_params"])
    sch = getattr(torch.optim.lr_scheduler, cfg[f"{who}_scheduler_type"])(opt, **cfg[f"{who}_scheduler_params"])
    return opt, sch


def build_autoencoder_trainer(ns, mel_cls, stft_cls, shape_cls, device, config="autoencoder/symAD_vctk_48000_hop300", seed=0,
                              use_stft=True, use_shape=True, adversarial=False, init_state=None):
    """trainer.autoencoder.Trainer (trainer/autoencoder.py:19-129) over a seeded symAD generator and the YAML's discriminator.
    adversarial=False is stage 1 of the 'efficient' paradigm (steps < start_steps.discriminator: metric + VQ losses only,
    :96-98), adversarial=True stage 2 (encoder/quantizer frozen, discriminator + feature matching on, :63-75,100-108)."""
    cfg = copy.deepcopy(ns.configs[config])
    cfg.setdefault("start_steps", {})["discriminator"] = 0 if adversarial else 1 << 30
    torch.manual_seed(seed)
    gen = ns.Generator(**cfg["generator_params"])
    if init_state is not None:
        gen.load_state_dict(init_state)
    gen = gen.to(device)
    disc, criterion = _gan_parts(ns, cfg, device, mel_cls, stft_cls, shape_cls, use_stft, use_shape)
    og, sg = _optim(cfg, "generator", gen.parameters())
    od, sd = _optim(cfg, "discriminator", disc.parameters())
    tr = ns.trainers["autoencoder"](steps=0, epochs=0, data_loader={}, model={"generator": gen, "discriminator": disc},
                                    criterion=criterion, optimizer={"generator": og, "discriminator": od},
                                    scheduler={"generator": sg, "discriminator": sd}, config=cfg, device=device)
    tr.tqdm = _NoTqdm()
    return tr


def build_vocoder_trainer(ns, mel_cls, stft_cls, shape_cls, device, config="vocoder/AudioDec_v1_symAD_vctk_48000_hop300_clean",
                          seed=0, use_stft=True, use_shape=True, adversarial=True, init_state=None):
    """trainer.vocoder.Trainer (trainer/vocoder.py:19-111): a HiFiGAN generator trained on the codes of a frozen analyzer (the
    symAD generator of the matching autoencoder YAML, seeded instead of loaded from the checkpoint the YAML names; the
    generator's input normalisation statistics file is not part of the repo: stats=None).  Its step compares with `>`
    (:66,78,94), so the trainer starts at steps=2: generator on, discriminator on when adversarial."""
    cfg = copy.deepcopy(ns.configs[config])
    cfg["generator_train_start_steps"] = 0
    cfg["discriminator_train_start_steps"] = 1 if adversarial else 1 << 30
    ae_name = "autoencoder/" + ("symAD_libritts_24000_hop300" if "libritts" in config else
                                "symADuniv_vctk_48000_hop300" if "univ" in config else "symAD_vctk_48000_hop300")
    torch.manual_seed(seed)
    analyzer = ns.Generator(**ns.configs[ae_name]["generator_params"]).to(device)
    gp = dict(cfg["generator_params"])
    gp["stats"] = None
    gen = ns.HiFiGANGenerator(**gp)
    if init_state is not None:
        gen.load_state_dict(init_state)
    gen = gen.to(device)
    disc, criterion = _gan_parts(ns, cfg, device, mel_cls, stft_cls, shape_cls, use_stft, use_shape)
    og, sg = _optim(cfg, "generator", gen.parameters())
    od, sd = _optim(cfg, "discriminator", disc.parameters())
    tr = ns.trainers["vocoder"](steps=2, epochs=0, data_loader={}, model={"generator": gen, "discriminator": disc, "analyzer": analyzer},
                                criterion=criterion, optimizer={"generator": og, "discriminator": od},
                                scheduler={"generator": sg, "discriminator": sd}, config=cfg, device=device)
    tr.tqdm = _NoTqdm()
    return tr


TRAIN_KEYS = ("train/mel_loss", "train/spectral_convergence_loss", "train/log_stft_magnitude_loss", "train/shape_loss",
              "train/generator_loss", "train/adversarial_loss", "train/feature_matching_loss", "train/discriminator_loss")


def run_steps(trainer, batches):
    """Runs Trainer._train_step on every batch; returns the per-step records the trainer itself keeps
    (total_train_loss is a running sum: the per-step value is the difference)."""
    keys = TRAIN_KEYS
    prev = {k: 0.0 for k in keys}
    rows = []
    for batch in batches:
        trainer._train_step(batch)
        row = {}
        for k in keys:
            cur = trainer.total_train_loss.get(k, 0.0)
            row[k.split("/", 1)[1]] = cur - prev[k]
            prev[k] = cur
        rows.append(row)
    return rows


def time_steps(trainer, batches, warmup=3):
    """Wall-clock ms per _train_step (the step ends with .item() syncs, trainerGAN.py:299-300) after `warmup` steps."""
    for b in batches[:warmup]:
        trainer._train_step(b)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        trainer._train_step(b)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / max(1, len(batches) - warmup)


class PortMel(torch.nn.Module):
    """The oracle's restatement of MultiMelSpectrogramLoss on the explicit-rFFT route (oracle/spectral_oracle.py): a second,
    independently rounded fp32 implementation of the reference's math -- the CONTROL for free-running trainer comparisons."""

    def __init__(self, **kw):
        super().__init__()
        from oracle import spectral_oracle as so
        self.res = so.mel_from_kwargs(**kw)

    def forward(self, y_hat, y):
        from oracle import spectral_oracle as so
        return so.multi_mel_loss(y_hat, y, self.res, use_torch_stft=False)


class PortStft(torch.nn.Module):
    def __init__(self, **kw):
        super().__init__()
        from oracle import spectral_oracle as so
        self.res = so.stft_from_kwargs(**kw)

    def forward(self, x, y):
        from oracle import spectral_oracle as so
        return so.mr_stft_loss(x, y, self.res, use_torch_stft=False)


class InDouble(torch.nn.Module):
    """A criterion evaluated in fp64 on fp32 tensors (inputs upcast, losses and hence gradients rounded back to fp32 once):
    the 'as exact as fp32 storage allows' run that two fp32 implementations are both measured against."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner.double()

    def forward(self, pred, target):
        out = self.inner(pred.double(), target.double())
        return tuple(o.float() for o in out) if isinstance(out, tuple) else out.float()


def in_double(cls):
    """Class-like factory for build_trainer(): cls(**kw) wrapped in InDouble."""
    return lambda **kw: InDouble(cls(**kw))


class Tee(torch.nn.Module):
    """Wraps a criterion of the running trainer: forwards to it unchanged (its losses drive the optimiser) and records, per
    call, the (prediction, target) pair it was given and -- through a tensor hook -- the gradient autograd delivers to the
    prediction after backward.  Lets a second implementation be evaluated on EXACTLY the tensors the trainer produced."""

    def __init__(self, inner, log, hook):
        super().__init__()
        self.inner, self.log, self.hook = inner, log, hook

    def forward(self, pred, target):
        if self.hook and pred.requires_grad:
            rec = {"pred": pred.detach().clone(), "target": target.detach().clone()}
            pred.register_hook(lambda g, rec=rec: rec.__setitem__("grad", g.detach().clone()))
            self.log.append(rec)
        return self.inner(pred, target)


def load():
    return ref_loader.load_reference_trainer()
