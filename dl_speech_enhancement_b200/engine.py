"""Host-side driver of the C ABI: builds the per-resolution constant tables, sizes the per-call
workspace and issues spl_forward / spl_reduce_finalize (one GPU) or spl_reduce_exchange_finalize (sharded: the
partial sums cross NVLink inside that kernel; NCCL all-reduce as fallback) / spl_backward.

All buffers are torch tensors (device memory, caching allocator, current stream); the library
itself never allocates or synchronises.  Nothing here computes losses or gradients in Python.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi
from ._abi import SPL_KIND_MEL, SPL_KIND_STFT, SplGeometry, SplTransform

# SPECLOSS_NVTX=1: every engine entry point opens an NVTX range ("specloss.forward", ...) around its launches, so a
# timeline tool (nsys, torch.profiler) shows where the criteria sit inside a trainer step.  Off by default: two extra
# Python calls per entry point are measurable on the host-bound small-batch path.
_NVTX = os.environ.get("SPECLOSS_NVTX", "0") not in ("", "0")


def _nvtx(name):
    def wrap(fn):
        if not _NVTX:
            return fn

        def inner(*a, **k):
            torch.cuda.nvtx.range_push("specloss." + name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        inner.__name__, inner.__doc__ = fn.__name__, fn.__doc__
        return inner
    return wrap

SUPPORTED_NFFT = (512, 1024, 2048)
MAX_MELS = 512           # check_transform() in csrc/specloss_host.inl enforces the same bound
COUNTER_SLOTS = 1024     # reduce tickets per device, handed out by recipe serial (see Engine._counter)


class _DeviceGuard:
    """Makes the tensors' device current around a C-ABI call: the library sizes grids, picks side streams and opts in to
    shared memory for cudaGetDevice()'s device, and the stream handle belongs to the tensors' device.  (The reference
    modules work on cuda:1 tensors while cuda:0 is current; so must these.)  One integer compare in the common case."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index if dev.type == "cuda" else None

    def __enter__(self):
        if self.idx is not None:
            self.prev = torch.cuda.current_device()
            if self.prev != self.idx:
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.idx is not None and self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _on_tensor_device(arg_index: int):
    """Runs an Engine method with the device of its arg_index-th positional argument (a tensor) current."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            t = args[arg_index]
            if t.device.type != "cuda" or t.device.index == torch.cuda.current_device():
                return fn(self, *args, **kwargs)
            with _DeviceGuard(t.device):
                return fn(self, *args, **kwargs)
        return wrapper
    return deco


def exchange_timeout_ns() -> int:
    """How long the fused exchange waits for its peers: SPECLOSS_EXCHANGE_TIMEOUT_S seconds, default 0 = as long as it
    takes (what a collective does; rank skew of many seconds is routine around checkpoints and validation)."""
    return int(float(os.environ.get("SPECLOSS_EXCHANGE_TIMEOUT_S", "0")) * 1e9)


def fft_geometry(n_fft: int):
    """(L, R): lanes per frame and points per lane; mirrors spl::FftGeom in specloss_kernels.cuh."""
    if n_fft not in SUPPORTED_NFFT:
        raise NotImplementedError(
            f"fft_size {n_fft} is outside the sm_100a kernels' envelope {SUPPORTED_NFFT} "
            "(the reference accepts any size torch.stft does; this implementation raises instead of falling back)")
    lanes = 16 if n_fft == 512 else 32
    return lanes, n_fft // lanes


def twiddle_table(n_fft: int) -> torch.Tensor:
    """W_N^(n1*k2) = exp(-2 pi i n1 k2 / N), fp32 roundings of fp64 values, layout [k2][n1] (re, im)."""
    lanes, rows = fft_geometry(n_fft)
    e = (np.arange(rows)[:, None] * np.arange(lanes)[None, :]) % n_fft
    th = 2.0 * np.pi * e.astype(np.float64) / n_fft
    tab = np.stack([np.cos(th), -np.sin(th)], axis=-1).astype(np.float32)
    return torch.from_numpy(tab.reshape(-1).copy())


def twiddle_eo_table() -> torch.Tensor:
    """Tables of the even/odd 2048-point kernels (csrc/transform_eo.cuh; mirrors spl_fill_twiddle_eo): W_1024^(n1*k2),
    layout [k2 < 32][n1 < 32], then W_2048^k for k = 0..512 zero-padded to 516 entries; fp32 roundings of fp64."""
    e = (np.arange(32)[:, None] * np.arange(32)[None, :]) % 1024
    th = 2.0 * np.pi * e.astype(np.float64) / 1024
    a = np.stack([np.cos(th), -np.sin(th)], axis=-1).reshape(-1)
    k = np.arange(516)
    th = 2.0 * np.pi * k.astype(np.float64) / 2048
    b = np.stack([np.where(k <= 512, np.cos(th), 0.0), np.where(k <= 512, -np.sin(th), 0.0)], axis=-1).reshape(-1)
    return torch.from_numpy(np.concatenate([a, b]).astype(np.float32))


def slot_offset_eo(k: int) -> int:
    """float2 index of the amplitude pair of bin k <= 1024 in the prediction's slot of the even/odd 2048-point kernel:
    the natural position of k in the 32 x 32 layout (row k % 32, pitch 33, column k // 32); bin 1024 = pad word of row 0."""
    return 32 if k == 1024 else (k % 32) * 33 + k // 32


# bins summed per lane per round of the projection schedule: 0 (default) picks, per filterbank, the target that minimises the
# modelled cost of the schedule (sum over rounds of its longest lane + a fixed per-round cost); a positive value forces it
MEL_ITER_TARGET = int(os.environ.get("SPECLOSS_MEL_ITER", "0"))
MEL_ROUND_COST = 5            # task load + shuffle reduction + store of one round, in units of one table entry


def slot_offset(n_fft: int, k: int) -> int:
    """float2 index of spectrum element k inside a frame slot: row k % R (pitch L + 1), column k // R."""
    lanes, rows = fft_geometry(n_fft)
    return (k % rows) * (lanes + 1) + k // rows


def mel_tables(melmat: np.ndarray, n_fft: int):
    """Band structure of the (K, n_mels) filterbank, as the kernels consume it.

    Forward projection: every mel row is a run of consecutive non-zero bins; a group of 1..L lanes
    (power of two, sized so each lane sums about `target` bins, see schedule()) strides over the run and the
    group is reduced with shuffles.  Groups are packed into rounds of L lanes, largest first, which
    keeps them aligned to their size.  Each lane walks a dense list of (amplitude slot, weight)
    entries, so the inner loop has no index arithmetic; padding entries carry weight 0.
    The amplitudes of bin k <= n_fft/2 are parked by the kernel at the bin's own position in the frame slot
    (those columns are free: the spectrum lives in registers), see specloss_kernels.cuh.
    Backward projection: every bin feeds at most two adjacent rows m0, m0+1.
    Slaney/HTK triangles always satisfy both; anything else raises (there is no dense fallback)."""
    lanes, _ = fft_geometry(n_fft)
    k_bins, n_mels = melmat.shape
    assert k_bins == n_fft // 2 + 1
    if not (2 <= n_mels <= MAX_MELS):
        raise NotImplementedError(f"num_mels must be in [2, {MAX_MELS}] for the sm_100a kernels")
    rows = []
    for m in range(n_mels):
        nz = np.flatnonzero(melmat[:, m])
        rows.append((m, int(nz[0]), int(nz[-1] - nz[0] + 1)) if nz.size else (m, 0, 0))

    def schedule(target):
        """Groups (lanes, mel row, first bin, bins) packed into rounds of `lanes` lanes: largest groups first (keeps every group
        aligned to its size), longest lanes first within a size (rounds then hold lanes of similar length: a round costs its
        longest lane)."""
        def group_size(length):
            g = 1
            while g < lanes and g * target < length:
                g *= 2
            return g

        groups = sorted(((group_size(ln), m, st, ln) for m, st, ln in rows), key=lambda t: (-t[0], -(-(-t[3] // t[0])), t[1]))
        out, cur, used = [], [], 0
        for g in groups:
            if used + g[0] > lanes:
                out.append(cur)
                cur, used = [], 0
            cur.append(g)
            used += g[0]
        if cur:
            out.append(cur)
        cost = sum(max((-(-ln // g) for g, _, _, ln in r), default=0) + MEL_ROUND_COST for r in out)
        return cost, out

    if MEL_ITER_TARGET > 0:
        rounds = schedule(MEL_ITER_TARGET)[1]
    else:
        rounds = min((schedule(t) for t in (8, 10, 12, 14, 16, 20, 24, 28, 32)), key=lambda c: c[0])[1]
    tasks = np.zeros((len(rounds), lanes, 4), np.int32)
    entries, entries_eo = [], []
    amp_slot = lambda k: slot_offset(n_fft, k)      # noqa: E731
    for r, grp_list in enumerate(rounds):
        iters = max((-(-ln // g) for g, _, _, ln in grp_list), default=0)
        ent = np.zeros((iters, lanes, 2), np.int32)
        ent_eo = np.zeros((iters, lanes, 2), np.int32)
        tasks[r, :, 0] = 0xfff | (1 << 12) | (iters << 20)             # idle lanes
        tasks[r, :, 1] = len(entries)
        lane = 0
        for g, m, st, ln in grp_list:
            for j in range(g):
                tasks[r, lane + j, 0] = m | (g << 12) | (iters << 20)
                for s_i, k in enumerate(range(st + j, st + ln, g)):
                    ent[s_i, lane + j, 0] = amp_slot(k)
                    ent[s_i, lane + j, 1] = np.float32(melmat[k, m]).view(np.int32)
                    if n_fft == 2048:
                        ent_eo[s_i, lane + j] = (slot_offset_eo(k), ent[s_i, lane + j, 1])
            lane += g
        entries.extend(ent)
        entries_eo.extend(ent_eo)
    entries = np.stack(entries) if entries else np.zeros((1, lanes, 2), np.int32)
    entries_eo = np.stack(entries_eo) if entries_eo else np.zeros((1, lanes, 2), np.int32)

    bin_tab = np.zeros((k_bins, 4), np.int32)
    f2i = lambda v: np.float32(v).view(np.int32)                                 # noqa: E731
    for k in range(k_bins):
        nz = np.flatnonzero(melmat[k])
        if nz.size == 0:
            continue
        if nz.size > 2 or (nz.size == 2 and nz[1] != nz[0] + 1):
            raise NotImplementedError(
                f"mel filterbank row for bin {k} feeds mels {nz.tolist()}: only banded filterbanks "
                "(<= 2 adjacent mels per bin, e.g. librosa/Slaney triangles) are supported")
        if nz.size == 2:
            bin_tab[k, :3] = (nz[0], f2i(melmat[k, nz[0]]), f2i(melmat[k, nz[1]]))
        elif nz[0] < n_mels - 1:
            bin_tab[k, :3] = (nz[0], f2i(melmat[k, nz[0]]), 0)
        else:
            bin_tab[k, :3] = (n_mels - 2, 0, f2i(melmat[k, nz[0]]))
    t = torch.from_numpy
    out = dict(mel_tasks=t(tasks.reshape(-1).copy()), mel_entries=t(entries.reshape(-1).copy()),
               bin_tab=t(bin_tab.reshape(-1).copy()))
    if n_fft == 2048:
        out["mel_entries_eo"] = t(entries_eo.reshape(-1).copy())
    return out


@dataclass
class TransformPlan:
    """One resolution of one loss, with its device-resident constant tables."""
    kind: int
    n_fft: int
    hop: int
    win: int
    eps: float
    window: torch.Tensor                    # (win,) fp32 -- the module's registered buffer
    twiddle: torch.Tensor                   # (2 * n_fft,) fp32
    n_mels: int = 0
    inv_ln_base: float = 1.0
    tables: dict = field(default_factory=dict)   # mel only: tensors named as the spl_transform fields
    twiddle_eo: Optional[torch.Tensor] = None    # n_fft == 2048: tables of the even/odd kernels (None: 64-point-per-lane kernels)

    def validate(self):
        fft_geometry(self.n_fft)
        if not (1 <= self.win <= self.n_fft):
            raise RuntimeError(f"win_length {self.win} must be in [1, fft_size={self.n_fft}] (torch.stft raises likewise)")
        if self.hop < 1:
            raise RuntimeError(f"hop_size {self.hop} < 1")          # hop > win_length is legal, as in torch.stft
        if self.window.dtype != torch.float32 or self.window.numel() != self.win:
            raise RuntimeError("window buffer must be float32 with win_length taps (fp32-only implementation)")

    def validate_explicit(self):
        """Envelope of the explicit spectrogram kernels (any hop >= 1)."""
        fft_geometry(self.n_fft)
        if not (1 <= self.win <= self.n_fft):
            raise RuntimeError(f"win_length {self.win} must be in [1, fft_size={self.n_fft}] (torch.stft raises likewise)")
        if self.hop < 1:
            raise RuntimeError(f"hop_size {self.hop} < 1")
        if self.window.dtype != torch.float32 or self.window.numel() != self.win:
            raise RuntimeError("window buffer must be float32 with win_length taps (fp32-only implementation)")


def melpow_twiddle_table() -> torch.Tensor:
    """W_400^(n1*k2) for the 25 x 16 transform of the power-mel metric: fp32 roundings of fp64, layout [k2 < 16][n1 < 25]."""
    e = (np.arange(16)[:, None] * np.arange(25)[None, :]) % 400
    th = 2.0 * np.pi * e.astype(np.float64) / 400
    tab = np.stack([np.cos(th), -np.sin(th)], axis=-1).astype(np.float32)
    return torch.from_numpy(tab.reshape(-1).copy())


def melpow_csr(fb: np.ndarray):
    """(n_freqs, n_mels) filterbank -> CSR by mel row: (ptr int32 [n_mels + 1], entries int32 [nnz][2] = {bin, weight bits})."""
    n_freqs, n_mels = fb.shape
    ptr, ent = [0], []
    for m in range(n_mels):
        for k in np.flatnonzero(fb[:, m]):
            ent.append((int(k), int(np.float32(fb[k, m]).view(np.int32))))
        ptr.append(len(ent))
    if not ent:
        raise RuntimeError("the mel filterbank is all zeros")
    return (torch.from_numpy(np.asarray(ptr, dtype=np.int32)),
            torch.from_numpy(np.asarray(ent, dtype=np.int32).reshape(-1).copy()))


def gemm_ld(n_fft: int) -> int:
    """Row pitch (floats) of the amplitude operand of the mel GEMM: n_fft/2+1 bins padded to whole 32-float k-blocks."""
    return ((n_fft // 2 + 1) + 31) // 32 * 32


def split_tf32(w: np.ndarray):
    """w = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared) -- the 3xTF32 operand split."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    hi = (w.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    return hi, (w - hi).astype(np.float32)


def mel_gemm_weights(melmat: np.ndarray, n_fft: int):
    """(K, n_mels) filterbank -> (w_hi, w_lo): melmat^T zero-padded to (n_pad, ld), K-major like the amplitudes."""
    k_bins, n_mels = melmat.shape
    n_pad = (n_mels + 15) // 16 * 16
    if n_pad > 128:
        raise NotImplementedError("the tensor-core mel projection supports num_mels <= 128")
    w = np.zeros((n_pad, gemm_ld(n_fft)), np.float32)
    w[:n_mels, :k_bins] = melmat.T
    hi, lo = split_tf32(w)
    return torch.from_numpy(hi.copy()), torch.from_numpy(lo.copy())


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class ForwardState:
    """What backward needs: the recipe (template transforms + workspace offsets) and the workspace address."""
    __slots__ = ("rec", "n", "keep", "batch", "t_len", "has_grad", "sc", "mag", "mel", "n_launches",
                 "ws", "ws_ptr", "off_sums", "off_coefs", "n_sums", "coefs_ptr", "device", "_transforms")

    def __init__(self):
        self.rec = None
        self._transforms = None
        self.n = 0
        self.keep = None
        self.ws = None
        self.ws_ptr = 0
        self.batch = self.t_len = 0
        self.has_grad = False
        self.sc = self.mag = self.mel = None
        self.n_launches = 0

    @property
    def transforms(self):
        """The spl_transform array of THIS call (template + pointers into this call's workspace), for the multi-step entry
        points (spl_reduce / spl_finalize / spl_reduce_exchange_finalize / spl_backward).  Built on first use: the
        unsharded path never needs it (spl_loss_forward / spl_loss_backward place the workspace themselves)."""
        if self._transforms is None:
            rec, n = self.rec, self.n
            arr = (SplTransform * n)()
            ctypes.memmove(arr, rec.template, rec.nbytes)
            for i in range(n):
                arr[i].partials = self.ws_ptr + rec.off_partials[i]
                arr[i].gframes = self.ws_ptr + rec.off_gframes[i] if self.has_grad else None
            self._transforms = arr
        return self._transforms

    def detach_workspace(self) -> torch.Tensor:
        """Hands the workspace tensor over to the caller (the autograd function saves it with save_for_backward, which
        ties its lifetime to the autograd graph); this object keeps only raw addresses into it."""
        ws = self.ws
        self.ws = None
        self.keep = self.keep[1:]
        return ws

    # views into the workspace, made on demand (the hot path only needs their addresses)
    @property
    def sums(self) -> torch.Tensor:
        return self.ws[self.off_sums:self.off_sums + 8 * self.n_sums].view(torch.float64)

    @property
    def coefs(self) -> torch.Tensor:
        return self.ws[self.off_coefs:self.off_coefs + 8 * self.n].view(torch.float32)


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _Recipe:
    """Everything about a (plan list, batch shape, grad mode) that does not change from call to call:
    the filled ctypes transform array and the carve-up of the single per-call workspace buffer."""
    __slots__ = ("template", "nbytes", "n", "off_partials", "off_gframes", "off_sums", "off_coefs", "ws_bytes",
                 "n_sums", "has_stft", "has_mel", "keep", "off_lsums", "serial", "plan_refs", "c_off_partials",
                 "c_off_gframes", "counter_ptr")


class Engine:
    def __init__(self, lib: ctypes.CDLL):
        self.lib = lib
        self._recipes = {}
        self._recipes_by_id = {}
        self._counters = {}
        self._exchanges = {}
        self._recipe_serial = 0
        self.launches = 0          # kernels launched so far (bench.py reports the per-step count)

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _stream(ref: torch.Tensor):
        """The caller's current CUDA stream as a raw handle (None for the CPU emulator of the test-suite)."""
        if not ref.is_cuda:
            return None
        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # no Stream object: ~10x cheaper per call
        if raw is not None:
            return ctypes.c_void_p(raw(ref.device.index if ref.device.index is not None else torch.cuda.current_device()))
        return ctypes.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream)

    def geometry(self, tr: SplTransform, batch: int, t_len: int) -> SplGeometry:
        g = SplGeometry()
        _abi.check(self.lib, self.lib.spl_geometry_of(ctypes.byref(tr), batch, t_len, ctypes.byref(g)))
        return g

    def _counter_ptr(self, dev, serial: int) -> int:
        """Address of the zero-initialised, self-resetting ticket of spl_reduce_finalize for recipe `serial`: one per
        recipe, so that two criteria evaluated concurrently on different streams never share one.  The tickets of a
        device live in ONE buffer that is never freed or moved (captured CUDA graphs keep raw pointers into it); a slot is
        reused only by the recipe created COUNTER_SLOTS recipes later."""
        pool = self._counters.get(dev)
        if pool is None:
            pool = self._counters[dev] = torch.zeros(COUNTER_SLOTS, dtype=torch.int32, device=dev)
        return pool.data_ptr() + 4 * (serial % COUNTER_SLOTS)

    def _exchange(self, group, dev, owner):
        """Peer-mapped exchange buffers of one recipe for `group` (torch symmetric memory over NVLink), or None when the
        group cannot use them (not NCCL / more than 8 ranks / SPECLOSS_NCCL_ALLREDUCE=1): the caller then all-reduces the
        sums with NCCL.  The first call per recipe is collective (every rank builds its recipes in the same order)."""
        key = (id(group), str(dev), owner)
        ex = self._exchanges.get(key)
        if ex is not None or key in self._exchanges:
            return ex
        ex = None
        import torch.distributed as dist
        world = dist.get_world_size(group)
        usable = (dev.type == "cuda" and os.environ.get("SPECLOSS_NCCL_ALLREDUCE", "0") != "1" and 1 < world <= 8
                  and dist.get_backend(group) == "nccl")
        if usable:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                nbytes = int(self.lib.spl_exchange_buffer_bytes())
                buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, group)
                torch.cuda.synchronize(dev)
                dist.barrier(group)                    # every buffer is zero before any peer writes into it
                ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
                # error word in pinned (mapped) host memory: the kernel stores to it, the host reads it without a sync
                err = torch.zeros(1, dtype=torch.int32).pin_memory()
                ex = dict(buf=buf, hdl=hdl, ptrs=ptrs, rank=dist.get_rank(group), world=world,
                          state=torch.zeros(2, dtype=torch.int32, device=dev), err=err, timeout_ns=exchange_timeout_ns())
            except Exception as exc:          # peer mapping unavailable (no NVLink/IPC): NCCL does the exchange instead
                import warnings
                warnings.warn(f"specloss: symmetric-memory exchange unavailable ({exc!r}); using NCCL all-reduce")
                ex = None
        self._exchanges[key] = ex
        return ex

    def _evict(self):
        """Drop every cached recipe together with its peer-exchange buffers (symmetric memory is returned only here, so
        the caches cannot grow without bound when batch shapes keep changing).  Ranks evict at the same call, because
        they build the same recipes in the same order."""
        self._recipes.clear()
        self._recipes_by_id.clear()
        self._exchanges.clear()

    def check_exchange_errors(self):
        """Raises if a fused exchange gave up waiting for a peer (finite SPECLOSS_EXCHANGE_TIMEOUT_S only).  Reads mapped
        host memory: no synchronisation.  Called at the start of every sharded forward."""
        for key, ex in self._exchanges.items():
            if ex is not None and int(ex["err"][0]) != 0:
                epoch = int(ex["err"][0])
                ex["err"][0] = 0
                raise _abi.SpecLossError(
                    f"specloss: peer exchange call #{epoch} of recipe {key[2]} timed out waiting for another rank "
                    f"(SPECLOSS_EXCHANGE_TIMEOUT_S={os.environ.get('SPECLOSS_EXCHANGE_TIMEOUT_S')}); the losses of that "
                    "step are NaN.  Every rank must call the criterion collectively.")

    def peer_exchange_active(self) -> bool:
        """True when every sharded recipe used so far exchanges its sums over peer memory (no NCCL call per step)."""
        return bool(self._exchanges) and all(v is not None for v in self._exchanges.values())

    def _recipe(self, plans: Sequence[TransformPlan], batch: int, t_len: int, need_grad: bool, dev) -> _Recipe:
        # fast path: the same plan OBJECTS as last time (modules cache their plans until a buffer moves; the recipe keeps
        # them alive, so an id cannot be recycled while its entry exists)
        fast = (tuple(map(id, plans)), batch, t_len, need_grad, dev)
        rec = self._recipes_by_id.get(fast)
        if rec is not None:
            return rec
        key = (tuple((p.kind, p.n_fft, p.hop, p.win, p.eps, p.n_mels, p.inv_ln_base, p.window.data_ptr(),
                      p.twiddle.data_ptr(), _ptr(p.twiddle_eo)) + tuple(t.data_ptr() for t in p.tables.values()) for p in plans),
               batch, t_len, need_grad, str(dev))
        rec = self._recipes.get(key)
        if rec is not None:
            if len(self._recipes_by_id) > 256:
                self._recipes_by_id.clear()
            self._recipes_by_id[fast] = rec
            if len(rec.plan_refs) < 8:           # keeps the plan objects (hence their ids) alive; bounded
                rec.plan_refs.append(tuple(plans))
            else:
                self._recipes_by_id.pop(fast)    # callers that build fresh plans per call take the keyed lookup
            return rec
        if len(plans) < 1 or len(plans) > _abi.SPL_MAX_TRANSFORMS:
            raise RuntimeError(f"{len(plans)} resolutions: supported range is 1..{_abi.SPL_MAX_TRANSFORMS}")
        n = len(plans)
        arr = (SplTransform * n)()
        rec = _Recipe()
        rec.n, rec.off_partials, rec.off_gframes, rec.keep = n, [], [], []
        off = 0
        n_sums = 0
        for i, pl in enumerate(plans):
            pl.validate()
            if pl.window.device != dev or pl.twiddle.device != dev:
                raise RuntimeError(f"loss module buffers are on {pl.window.device}, inputs on {dev}: call .to(device)")
            if t_len <= pl.n_fft // 2:
                raise RuntimeError(f"reflect padding needs T > fft_size/2 (T={t_len}, fft_size={pl.n_fft}); "
                                   "torch.stft raises for the same input")
            tr = arr[i]
            tr.kind, tr.n_fft, tr.hop, tr.win = pl.kind, pl.n_fft, pl.hop, pl.win
            tr.eps = pl.eps
            tr.window, tr.twiddle = _ptr(pl.window), _ptr(pl.twiddle)
            tr.n_mels, tr.inv_ln_base = pl.n_mels, pl.inv_ln_base
            if pl.twiddle_eo is not None and pl.n_fft == 2048:
                if pl.twiddle_eo.device != dev:
                    raise RuntimeError("loss module buffers are not on the input device: call .to(device)")
                tr.twiddle_eo = _ptr(pl.twiddle_eo)
            if pl.kind == SPL_KIND_MEL:
                for name, t in pl.tables.items():
                    if t.device != dev:
                        raise RuntimeError("mel tables are not on the input device: call .to(device)")
                    setattr(tr, name, _ptr(t))
                lanes = fft_geometry(pl.n_fft)[0]
                tr.mel_rounds = pl.tables["mel_tasks"].numel() // (4 * lanes)
                tr.mel_entry_rows = pl.tables["mel_entries"].numel() // (2 * lanes)
            g = self.geometry(tr, batch, t_len)
            rec.off_partials.append(off)
            off = _align(off + 8 * g.partial_count)
            if need_grad:
                rec.off_gframes.append(off)
                off = _align(off + g.gframe_bytes)
            n_sums += g.n_sums
            rec.keep.extend([pl.window, pl.twiddle, pl.twiddle_eo] + list(pl.tables.values()))
        rec.off_sums = off
        off = _align(off + 8 * n_sums)
        rec.off_lsums = off                    # this rank's sums when the global ones come from the peer exchange
        off = _align(off + 8 * n_sums)
        rec.off_coefs = off
        off = _align(off + 4 * 2 * n)
        rec.ws_bytes, rec.n_sums = off, n_sums
        rec.has_stft = any(p.kind == SPL_KIND_STFT for p in plans)
        rec.has_mel = any(p.kind == SPL_KIND_MEL for p in plans)
        rec.template, rec.nbytes = arr, ctypes.sizeof(arr)
        rec.c_off_partials = (ctypes.c_int64 * n)(*rec.off_partials)
        rec.c_off_gframes = (ctypes.c_int64 * n)(*rec.off_gframes) if need_grad else None
        self._recipe_serial += 1
        rec.serial = self._recipe_serial       # same on every rank (SPMD): names the recipe's peer-exchange buffers
        rec.counter_ptr = self._counter_ptr(dev, rec.serial)
        if len(self._recipes) > 64:
            self._evict()
        self._recipes[key] = rec
        rec.plan_refs = [tuple(plans)]
        self._recipes_by_id[fast] = rec
        return rec

    # -- forward ---------------------------------------------------------------------------------
    @_on_tensor_device(1)
    @_nvtx("forward")
    def forward(self, plans: Sequence[TransformPlan], x: torch.Tensor, y: torch.Tensor, need_grad: bool,
                group=None, global_batch: Optional[int] = None) -> ForwardState:
        """x, y: (B, T) fp32 contiguous on one device.  Launches on the current stream; one workspace
        allocation per call, no host synchronisation."""
        batch, t_len = x.shape
        dev = x.device
        rec = self._recipe(plans, batch, t_len, need_grad, dev)
        n = rec.n
        ws = torch.empty(rec.ws_bytes, dtype=torch.uint8, device=dev)
        base = ws.data_ptr()
        st = ForwardState()
        st.rec, st.ws_ptr = rec, base
        st.batch, st.t_len, st.has_grad, st.n = batch, t_len, need_grad, n
        st.keep = (ws, rec.keep)
        # separate 0-dim outputs: callers scale them in place (trainer/trainerGAN.py:221,228-229)
        if rec.has_stft:
            st.sc = torch.empty((), dtype=torch.float32, device=dev)
            st.mag = torch.empty((), dtype=torch.float32, device=dev)
        if rec.has_mel:
            st.mel = torch.empty((), dtype=torch.float32, device=dev)
        st.ws, st.off_sums, st.off_coefs, st.n_sums, st.device = ws, rec.off_sums, rec.off_coefs, rec.n_sums, dev
        st.coefs_ptr = base + rec.off_coefs
        stream = self._stream(x)
        lib = self.lib
        if group is None and (global_batch is None or global_batch == batch):
            # unsharded: transforms + reduce + finalize behind ONE C-ABI call
            counter = rec.counter_ptr
            rc = lib.spl_loss_forward(rec.template, n, x.data_ptr(), y.data_ptr(), batch, t_len, base, rec.c_off_partials,
                                      rec.c_off_gframes, rec.off_sums, rec.off_coefs, _ptr(st.sc), _ptr(st.mag), _ptr(st.mel),
                                      counter, stream)
            if rc:
                _abi.check(lib, rc)
            st.n_launches = n + 1
            self.launches += st.n_launches
            return st
        arr = st.transforms
        _abi.check(lib, lib.spl_forward(arr, n, x.data_ptr(), y.data_ptr(), batch, t_len, stream))
        sums_ptr, coefs_ptr = base + rec.off_sums, base + rec.off_coefs
        ex = self._exchange(group, dev, rec.serial) if group is not None else None
        if ex is not None:
            if ex["timeout_ns"] > 0:
                self.check_exchange_errors()
            # reduce + NVLink peer-memory exchange + finalize in one launch (spl_reduce_exchange_finalize)
            if global_batch is None:
                global_batch = batch * ex["world"]
            lsums_ptr = base + rec.off_lsums
            _abi.check(lib, lib.spl_reduce_exchange_finalize(
                arr, n, batch, t_len, int(global_batch), lsums_ptr, sums_ptr, ex["rank"], ex["world"],
                ex["ptrs"], ex["state"].data_ptr(), ex["timeout_ns"], ex["err"].data_ptr(),
                _ptr(st.sc), _ptr(st.mag), _ptr(st.mel), coefs_ptr, stream))
            st.keep = st.keep + (ex,)
            st.n_launches = n + 1
        else:
            _abi.check(lib, lib.spl_reduce(arr, n, batch, t_len, sums_ptr, stream))
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(st.sums, op=dist.ReduceOp.SUM, group=group)   # the single exchange step
                if global_batch is None:
                    global_batch = batch * dist.get_world_size(group)
            _abi.check(lib, lib.spl_finalize(arr, n, sums_ptr, int(global_batch), t_len, _ptr(st.sc), _ptr(st.mag),
                                             _ptr(st.mel), coefs_ptr, stream))
            st.n_launches = n + 2
        self.launches += st.n_launches
        return st

    # -- explicit spectrogram / log-mel spectrogram -----------------------------------------------------
    @_on_tensor_device(0)
    @_nvtx("spectrogram")
    def spectrogram(self, x: torch.Tensor, n_fft: int, hop: int, win: int, window: torch.Tensor,
                    twiddle: torch.Tensor, eps: float, ld: Optional[int] = None, split: bool = False):
        """(B, T) fp32 -> magnitude spectrogram (B, 1 + T // hop, n_fft // 2 + 1), the tensor the reference's
        stft() returns (stft_loss.py:19-35).  With `ld` the rows are padded to ld floats (zeros) and a view is
        returned.  split=True returns (hi, lo) with hi = tf32(A) and lo = A - hi, full (B, F, ld) buffers: the
        operands of mel_project()."""
        fft_geometry(n_fft)
        batch, t_len = x.shape
        n_bins = n_fft // 2 + 1
        ld = n_bins if ld is None else int(ld)
        out = torch.empty(batch, 1 + t_len // hop, ld, dtype=torch.float32, device=x.device)
        lo = torch.empty_like(out) if split else None
        _abi.check(self.lib, self.lib.spl_spectrogram(x.data_ptr(), batch, t_len, n_fft, hop, win, window.data_ptr(),
                                                      twiddle.data_ptr(), eps, out.data_ptr(), _ptr(lo), ld,
                                                      self._stream(x)))
        self.launches += 1
        if split:
            return out, lo
        return out if ld == n_bins else out[:, :, :n_bins]

    @_on_tensor_device(0)
    def mel_project(self, amp_hi: torch.Tensor, amp_lo: torch.Tensor, w_hi: torch.Tensor, w_lo: torch.Tensor,
                    n_mels: int, eps: float, log_scale: float) -> torch.Tensor:
        """(B, F, ld) split amplitudes x (n_pad, ld) split melmat^T -> log-mel (B, n_mels, F) on the tensor cores:
        log_b(clamp(matmul(x_amp, melmat), eps)).transpose(1, 2) of MelSpectrogram.forward (mel_loss.py:91-94)."""
        batch, frames, ld = amp_hi.shape
        out = torch.empty(batch, n_mels, frames, dtype=torch.float32, device=amp_hi.device)
        _abi.check(self.lib, self.lib.spl_mel_project(amp_hi.data_ptr(), amp_lo.data_ptr(), batch * frames, ld,
                                                      w_hi.data_ptr(), w_lo.data_ptr(), n_mels, w_hi.shape[0], frames,
                                                      eps, log_scale, out.data_ptr(), self._stream(amp_hi)))
        self.launches += 1
        return out

    @_on_tensor_device(1)
    @_nvtx("spectrogram_backward")
    def spectrogram_backward(self, plan: TransformPlan, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        """dL/dx (B, T) from the gradient of an explicit spectrogram of x: g = dL/d stft(x) (B, F, K) for an STFT plan
        (stft_loss.py:19-35), g = dL/d MelSpectrogram(x) (B, n_mels, F) for a mel plan (mel_loss.py:74-94).  Two
        launches (recompute + adjoint transform, overlap-add gather) on the current stream."""
        plan.validate_explicit()
        batch, t_len = x.shape
        frames = 1 + t_len // plan.hop
        dev = x.device
        tr = SplTransform()
        tr.kind, tr.n_fft, tr.hop, tr.win, tr.eps = plan.kind, plan.n_fft, plan.hop, plan.win, plan.eps
        tr.window, tr.twiddle = _ptr(plan.window), _ptr(plan.twiddle)
        tr.n_mels, tr.inv_ln_base = plan.n_mels, plan.inv_ln_base
        if plan.kind == SPL_KIND_MEL:
            want = (batch, plan.n_mels, frames)
            for name, t in plan.tables.items():
                if name != "mel_entries_eo":          # the explicit path runs the 64-point-per-lane kernels
                    setattr(tr, name, _ptr(t))
            lanes = fft_geometry(plan.n_fft)[0]
            tr.mel_rounds = plan.tables["mel_tasks"].numel() // (4 * lanes)
            tr.mel_entry_rows = plan.tables["mel_entries"].numel() // (2 * lanes)
            ld = frames
        else:
            want = (batch, frames, plan.n_fft // 2 + 1)
            ld = want[2]
        if tuple(g.shape) != want or g.dtype != torch.float32 or g.device != dev:
            raise RuntimeError(f"spectrogram gradient must be fp32 {want} on {dev}, got {g.dtype} {tuple(g.shape)} on {g.device}")
        g = g.contiguous()
        gframes = torch.empty(batch * frames * plan.win, dtype=torch.float32, device=dev)
        tr.gframes = gframes.data_ptr()
        dx = torch.empty(batch, t_len, dtype=torch.float32, device=dev)
        _abi.check(self.lib, self.lib.spl_spectrogram_backward(ctypes.byref(tr), x.data_ptr(), batch, t_len, g.data_ptr(),
                                                               ld, dx.data_ptr(), self._stream(x)))
        self.launches += 2
        return dx

    # -- waveform shape loss ---------------------------------------------------------------------
    @_on_tensor_device(0)
    @_nvtx("shape_forward")
    def shape_forward(self, x: torch.Tensor, y: torch.Tensor, winlens: Sequence[int], group=None,
                      global_rows: Optional[int] = None):
        """x, y: (rows, T) fp32 contiguous.  Returns (loss 0-dim, records, rows_global): MultiWindowShapeLoss.forward
        (waveform_loss.py:59-75).  Launches: forward + reduce [+ all-reduce] + finalize."""
        rows, t_len = x.shape
        dev = x.device
        n = len(winlens)
        wl = (ctypes.c_int32 * n)(*[int(w) for w in winlens])
        n_rec, n_part = ctypes.c_int64(), ctypes.c_int64()
        _abi.check(self.lib, self.lib.spl_shape_geometry(rows, t_len, wl, n, ctypes.byref(n_rec), ctypes.byref(n_part)))
        records = torch.empty(max(1, n_rec.value), dtype=torch.int32, device=dev)
        partials = torch.empty(n_part.value, dtype=torch.float64, device=dev)
        sums = torch.empty(n, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stream = self._stream(x)
        _abi.check(self.lib, self.lib.spl_shape_forward(x.data_ptr(), y.data_ptr(), rows, t_len, wl, n, records.data_ptr(),
                                                        partials.data_ptr(), sums.data_ptr(), stream))
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            if global_rows is None:
                global_rows = rows * dist.get_world_size(group)
        rows_global = int(global_rows) if global_rows is not None else rows
        _abi.check(self.lib, self.lib.spl_shape_finalize(sums.data_ptr(), rows_global, t_len, wl, n, loss.data_ptr(), stream))
        self.launches += 3
        return loss, records, rows_global

    @_on_tensor_device(0)
    @_nvtx("shape_backward")
    def shape_backward(self, records: torch.Tensor, rows: int, rows_global: int, t_len: int, winlens: Sequence[int],
                       g: torch.Tensor) -> torch.Tensor:
        dev = records.device
        n = len(winlens)
        wl = (ctypes.c_int32 * n)(*[int(w) for w in winlens])
        if g.dtype != torch.float32 or g.device != dev or g.dim() != 0:
            g = g.detach().to(device=dev, dtype=torch.float32).reshape(())
        dx = torch.empty(rows, t_len, dtype=torch.float32, device=dev)
        _abi.check(self.lib, self.lib.spl_shape_backward(records.data_ptr(), rows, rows_global, t_len, wl, n, g.data_ptr(),
                                                         dx.data_ptr(), self._stream(dx)))
        self.launches += 1
        return dx

    # -- power-mel L1 metric (Mel_L1 of the reference's evaluation scripts) ------------------------------------
    @_on_tensor_device(0)
    @_nvtx("melpow_l1")
    def melpow_l1(self, x: torch.Tensor, y: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, twiddle: torch.Tensor,
                  n_mels: int, mel_ptr: torch.Tensor, mel_ent: torch.Tensor, want_mels: bool = False):
        """x, y: (rows, T) fp32 contiguous -> (loss 0-dim, mel_x, mel_y): nn.L1Loss()(M(x), M(y)) with M the power-mel
        spectrogram (mel_spectrogram.py:36-44); mel_x / mel_y (rows, n_mels, 1 + T // hop) only when want_mels."""
        rows, t_len = x.shape
        dev = x.device
        n_part = ctypes.c_int64()
        _abi.check(self.lib, self.lib.spl_melpow_geometry(rows, t_len, n_fft, hop, ctypes.byref(n_part)))
        partials = torch.empty(n_part.value, dtype=torch.float64, device=dev)
        total = torch.empty(1, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        frames = 1 + t_len // hop
        mel_x = torch.empty(rows, n_mels, frames, dtype=torch.float32, device=dev) if want_mels else None
        mel_y = torch.empty(rows, n_mels, frames, dtype=torch.float32, device=dev) if want_mels else None
        _abi.check(self.lib, self.lib.spl_melpow_l1(
            x.data_ptr(), y.data_ptr(), rows, t_len, n_fft, hop, window.data_ptr(), twiddle.data_ptr(), n_mels,
            mel_ent.numel() // 2, mel_ptr.data_ptr(), mel_ent.data_ptr(), partials.data_ptr(), total.data_ptr(),
            loss.data_ptr(), _ptr(mel_x), _ptr(mel_y), self._stream(x)))
        self.launches += 3
        return loss, mel_x, mel_y

    # -- losses on explicit magnitude tensors ----------------------------------------------------
    @_on_tensor_device(0)
    @_nvtx("mag_loss_forward")
    def mag_loss_forward(self, x_mag: torch.Tensor, y_mag: torch.Tensor, want_sc: bool, want_mag: bool):
        """x_mag, y_mag: fp32 contiguous, same shape.  Returns (sc or None, mag or None, sums): SpectralConvergenceLoss /
        LogSTFTMagnitudeLoss.forward (stft_loss.py:38-77) on the streaming kernels."""
        n = x_mag.numel()
        dev = x_mag.device
        n_part = ctypes.c_int64()
        _abi.check(self.lib, self.lib.spl_mag_loss_geometry(n, ctypes.byref(n_part)))
        partials = torch.empty(n_part.value, dtype=torch.float64, device=dev)
        sums = torch.empty(6, dtype=torch.float64, device=dev)
        sc = torch.empty((), dtype=torch.float32, device=dev) if want_sc else None
        mag = torch.empty((), dtype=torch.float32, device=dev) if want_mag else None
        _abi.check(self.lib, self.lib.spl_mag_loss_forward(x_mag.data_ptr(), y_mag.data_ptr(), n, partials.data_ptr(),
                                                           sums.data_ptr(), _ptr(sc), _ptr(mag), self._stream(x_mag)))
        self.launches += 3
        return sc, mag, sums

    @_on_tensor_device(0)
    @_nvtx("mag_loss_backward")
    def mag_loss_backward(self, x_mag, y_mag, sums, g_sc, g_mag, need_x: bool, need_y: bool):
        dev = x_mag.device

        def scalar(g):
            if g is None:
                return None
            if g.dtype != torch.float32 or g.device != dev or g.dim() != 0:
                g = g.detach().to(device=dev, dtype=torch.float32).reshape(())
            return g

        g_sc, g_mag = scalar(g_sc), scalar(g_mag)
        gx = torch.empty_like(x_mag) if need_x else None
        gy = torch.empty_like(y_mag) if need_y else None
        _abi.check(self.lib, self.lib.spl_mag_loss_backward(x_mag.data_ptr(), y_mag.data_ptr(), x_mag.numel(), sums.data_ptr(),
                                                            _ptr(g_sc), _ptr(g_mag), _ptr(gx), _ptr(gy), self._stream(x_mag)))
        self.launches += 1
        return gx, gy

    # -- backward --------------------------------------------------------------------------------
    @_nvtx("backward")
    def backward(self, st: ForwardState, g_sc: Optional[torch.Tensor], g_mag: Optional[torch.Tensor],
                 g_mel: Optional[torch.Tensor]) -> torch.Tensor:
        if not st.has_grad:
            raise RuntimeError("backward requested but forward ran without gradient workspace")
        dev = st.device
        if dev.type == "cuda" and dev.index != torch.cuda.current_device():
            with _DeviceGuard(dev):
                return self.backward(st, g_sc, g_mag, g_mel)
        dx = torch.empty(st.batch, st.t_len, dtype=torch.float32, device=dev)

        def scalar(g):
            if g is None:
                return None
            if g.dtype != torch.float32 or g.device != dev or g.dim() != 0:
                g = g.detach().to(device=dev, dtype=torch.float32).reshape(())
            return g

        gs = (scalar(g_sc), scalar(g_mag), scalar(g_mel))
        rec = st.rec
        rc = self.lib.spl_loss_backward(rec.template, st.n, st.batch, st.t_len, st.ws_ptr, rec.c_off_gframes, rec.off_coefs,
                                        _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), dx.data_ptr(), self._stream(dx))
        if rc:
            _abi.check(self.lib, rc)
        self.launches += 1
        return dx


_ENGINE: Optional[Engine] = None


def cuda_engine() -> Engine:
    """The process-wide engine over libspecloss.so.  Raises when the library is not built."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(_abi.load_library())
    return _ENGINE
