"""torch.autograd.Function over the C ABI, and the functional entry point `spectral_losses`.

Forward launches the fused kernels (losses AND the un-scaled waveform-gradient pieces, while the
spectra are still on chip); backward is one gather/scale kernel.  CUDA tensors only: CPU tensors,
non-fp32 dtypes and missing libspecloss.so raise -- there is no fallback path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .engine import Engine, TransformPlan, cuda_engine


def _as_2d(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dim() == 3:
        # the reference does x.view(-1, T) (stft_loss.py:158-160), which already requires contiguity
        x = x.reshape(-1, x.size(2))
    elif x.dim() != 2:
        raise RuntimeError(f"{name}: expected (B, T) or (B, C, T), got shape {tuple(x.shape)}")
    return x.contiguous()


def _check_inputs(x: torch.Tensor, y: torch.Tensor):
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("dl_speech_enhancement_b200 losses are CUDA-only (sm_100a kernels); got a CPU tensor. "
                           "There is no CPU fallback: use the reference modules on CPU.")
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise RuntimeError(f"fp32-only implementation; got {x.dtype} / {y.dtype}")
    if x.device != y.device:
        raise RuntimeError("prediction and target are on different devices")
    if x.shape != y.shape:
        raise RuntimeError(f"shape mismatch: {tuple(x.shape)} vs {tuple(y.shape)}")


class _SpectralLossFn(torch.autograd.Function):
    """outputs: (sc, mag, mel) with None for the absent loss family."""

    @staticmethod
    def forward(ctx, x, y, plans, engine, group, global_batch):
        x2, y2 = _as_2d(x, "prediction"), _as_2d(y, "target")
        need_grad = ctx.needs_input_grad[0]
        st = engine.forward(plans, x2, y2, need_grad, group, global_batch)      # only shapes and addresses are read
        if need_grad:
            # The gradient workspace travels as a SAVED TENSOR: autograd frees it right after backward() unless the
            # caller asked for retain_graph=True (then a second backward works, as with the reference's graph), and a
            # second backward without it raises torch's usual "backward through the graph a second time" error.
            ctx.save_for_backward(st.detach_workspace())
        ctx.state = st
        ctx.engine = engine
        ctx.x_shape = x.shape
        ctx.set_materialize_grads(False)
        outs = tuple(t for t in (st.sc, st.mag, st.mel) if t is not None)
        ctx.layout = (st.sc is not None, st.mel is not None)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        has_stft, has_mel = ctx.layout
        g = list(grads)
        g_sc = g_mag = g_mel = None
        if has_stft:
            g_sc, g_mag = g[0], g[1]
            g = g[2:]
        if has_mel:
            g_mel = g[0]
        if ctx.needs_input_grad[1]:
            raise NotImplementedError(
                "gradient w.r.t. the target is not implemented (no caller in the reference needs it: "
                "trainer/denoise.py:75, autoencoder.py:98, vocoder.py:76 pass a constant target)")
        (ws,) = ctx.saved_tensors          # keeps the workspace alive across the launch; raises if already released
        dx = ctx.engine.backward(ctx.state, g_sc, g_mag, g_mel)
        del ws
        return dx.view(ctx.x_shape), None, None, None, None, None


def spectral_losses(x: torch.Tensor, y: torch.Tensor, plans: Sequence[TransformPlan],
                    group=None, global_batch: Optional[int] = None, engine: Optional[Engine] = None):
    """Returns the tuple of 0-dim losses the plans define: (sc, mag) and/or (mel,), in that order.

    group: a torch.distributed process group whose ranks hold disjoint utterances of one logical
    batch; the partial sums are all-reduced so every rank returns the full-batch losses and its
    own rows' gradients."""
    if engine is None:
        _check_inputs(x, y)
        engine = cuda_engine()
    return _SpectralLossFn.apply(x, y, tuple(plans), engine, group, global_batch)


class _SpectrogramFn(torch.autograd.Function):
    """Explicit magnitude spectrogram stft() (stft_loss.py:19-35) with its backward (ABI spl_spectrogram /
    spl_spectrogram_backward).  plan: an STFT TransformPlan carrying window, twiddle and eps."""

    @staticmethod
    def forward(ctx, x, plan, engine):
        ctx.plan, ctx.engine = plan, engine
        xd = x.detach()
        ctx.save_for_backward(xd)
        return engine.spectrogram(xd, plan.n_fft, plan.hop, plan.win, plan.window, plan.twiddle, plan.eps)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.engine.spectrogram_backward(ctx.plan, x, g), None, None


class _LogMelFn(torch.autograd.Function):
    """MelSpectrogram.forward (mel_loss.py:74-94): spectrogram kernel + tcgen05 mel projection GEMM forward;
    backward recomputes the spectra and applies the banded transposed projection inside the FFT kernel."""

    @staticmethod
    def forward(ctx, x, plan, w_hi, w_lo, ld, log_scale, engine):
        ctx.plan, ctx.engine = plan, engine
        xd = x.detach()
        ctx.save_for_backward(xd)
        hi, lo = engine.spectrogram(xd, plan.n_fft, plan.hop, plan.win, plan.window, plan.twiddle, plan.eps, ld=ld,
                                    split=True)
        return engine.mel_project(hi, lo, w_hi, w_lo, plan.n_mels, plan.eps, log_scale)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.engine.spectrogram_backward(ctx.plan, x, g), None, None, None, None, None, None


def spectrogram(x: torch.Tensor, plan: TransformPlan, engine: Optional[Engine] = None) -> torch.Tensor:
    return _SpectrogramFn.apply(x, plan, engine or cuda_engine())


def log_mel_spectrogram(x: torch.Tensor, plan: TransformPlan, w_hi, w_lo, ld: int, log_scale: float,
                        engine: Optional[Engine] = None) -> torch.Tensor:
    return _LogMelFn.apply(x, plan, w_hi, w_lo, ld, log_scale, engine or cuda_engine())


class _ShapeLossFn(torch.autograd.Function):
    """MultiWindowShapeLoss.forward (waveform_loss.py:59-75) over the C ABI (spl_shape_*)."""

    @staticmethod
    def forward(ctx, x, y, winlens, engine, group):
        x2, y2 = _as_2d(x, "prediction"), _as_2d(y, "target")
        loss, records, rows_global = engine.shape_forward(x2.detach(), y2.detach(), winlens, group)
        ctx.engine, ctx.winlens, ctx.x_shape = engine, winlens, x.shape
        ctx.rows, ctx.t_len, ctx.rows_global = x2.shape[0], x2.shape[1], rows_global
        ctx.records = records if ctx.needs_input_grad[0] else None
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("gradient w.r.t. the target is not implemented (no reference caller needs it)")
        dx = ctx.engine.shape_backward(ctx.records, ctx.rows, ctx.rows_global, ctx.t_len, ctx.winlens, g)
        ctx.records = None
        return dx.view(ctx.x_shape), None, None, None, None


def shape_loss(x: torch.Tensor, y: torch.Tensor, winlens: Sequence[int], group=None,
               engine: Optional[Engine] = None) -> torch.Tensor:
    if engine is None:
        _check_inputs(x, y)
        engine = cuda_engine()
    return _ShapeLossFn.apply(x, y, tuple(int(w) for w in winlens), engine, group)


class _MagLossFn(torch.autograd.Function):
    """SpectralConvergenceLoss (which = 0) / LogSTFTMagnitudeLoss (which = 1) on explicit magnitude tensors
    (stft_loss.py:38-77), differentiable w.r.t. both arguments like the reference."""

    @staticmethod
    def forward(ctx, x_mag, y_mag, which, engine):
        xd, yd = x_mag.detach().contiguous(), y_mag.detach().contiguous()
        sc, mag, sums = engine.mag_loss_forward(xd, yd, which == 0, which == 1)
        ctx.engine, ctx.which, ctx.sums = engine, which, sums
        ctx.save_for_backward(xd, yd)
        return sc if which == 0 else mag

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        xd, yd = ctx.saved_tensors
        gx, gy = ctx.engine.mag_loss_backward(xd, yd, ctx.sums, g if ctx.which == 0 else None, g if ctx.which == 1 else None,
                                              ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gx, gy, None, None


def magnitude_loss(x_mag: torch.Tensor, y_mag: torch.Tensor, which: int, engine: Optional[Engine] = None) -> torch.Tensor:
    if engine is None:
        _check_inputs(x_mag, y_mag)
        engine = cuda_engine()
    return _MagLossFn.apply(x_mag, y_mag, which, engine)
