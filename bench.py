#!/usr/bin/env python3
"""Benchmark of the spectral-loss hot path: MultiResolutionSTFTLoss (3 resolutions) +
MultiMelSpectrogramLoss (2048/300, 80 mels) forward+backward, audio-seconds per second.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own modules on the host cores

Workload (BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is quoted on): a GLOBAL batch of
256 x 4 s synthetic 48 kHz audio, batch-sharded over the N ranks (256/N utterances per GPU, N=1 runs all 256 on one
GPU) -- strong scaling.  The loss partial sums are exchanged between the ranks so every rank returns the losses of the
global batch.  A step = one fwd+bwd of both criteria through the drop-in nn.Modules, launched eagerly, exactly
as trainer/trainerGAN.py:214-241,271-281 drives them (no CUDA graphs in the headline).  The timed region is repeated
blocks of exactly K steps (each bracketed by barrier + synchronize, device time by CUDA events, max over ranks); the
median block is reported.  `config.secondary` carries BASELINE configs[1] (16 x 1 s per GPU, the round-1 headline) so
rounds stay comparable.  At N > 1 a `sharded_parity` block (outside the timed region) checks that all ranks hold
bit-identical losses equal to rank 0's unsharded evaluation of the gathered global batch, and that every rank's gradient
rows equal that run's; a failure makes the process exit non-zero.  Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000
GLOBAL_BATCH = 256              # configs[3]
T_LEN = 4 * FS
SEC_BATCH, SEC_T = 16, FS       # configs[1] (per GPU), reported as config.secondary
MEL_KW = dict(fs=FS, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
STFT_RES = [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)]
METRIC = "spectral_loss_fwd_bwd_audio_seconds_per_second"
UNIT = "audio-s/s"
POOL_BYTES = 192 << 20          # > 126 MB L2
MIN_TIMED_S = 0.25              # repeat K-step blocks until the timed region is at least this long
MAX_BLOCKS = 40


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def nominal_flops(batch, t_len):
    """5 passes x sum_res B*F*2.5*N*log2(N)."""
    tot = 0.0
    for n, hop, _ in STFT_RES + [(2048, 300, 2048)]:
        tot += batch * (1 + t_len // hop) * 2.5 * n * math.log2(n)
    return 5.0 * tot


def workload_name(batch, t_len):
    tag = "configs[3]: " if (batch, t_len) == (256, 4 * FS) else ("configs[1]: " if (batch, t_len) == (16, FS) else "non-default size: ")
    return tag + f"global batch {batch} x {t_len / FS:g} s synthetic 48 kHz, MR-STFT(1024/2048/512) + 80-mel hop-300, fwd+bwd"


def csrc_hash():
    """Identifies the kernel sources a committed ncu capture belongs to."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "dl_speech_enhancement_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inl")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.3)           # nvidia-smi start-up: the first sample must not post-date the timed region
            self.rows.clear()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.08)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's OWN modules (losses/stft_loss.py, losses/mel_loss.py) on the host cores.
# oracle/_ref holds their unmodified copies (oracle/make_ref.sh) -- kind "reference"; if that staging is missing the
# oracle's restatement of the same op sequence is timed instead -- kind "port".
# --------------------------------------------------------------------------------------------
def cpu_reference_step_time(batch, t_len, steps, warmup):
    import torch
    from oracle import ref_loader, spectral_oracle as so     # checker/baseline only; never on the product path

    torch.set_num_threads(os.cpu_count())
    y_hat, y = so.synth_pair(batch, t_len, seed=0)
    kind = "port"
    if ref_loader.available():
        stft_mod, mel_mod = ref_loader.load_reference_losses()
        stft = stft_mod.MultiResolutionSTFTLoss()
        mel = mel_mod.MultiMelSpectrogramLoss(**MEL_KW)
        kind = "reference"

        def step():
            xx = y_hat.detach().clone().requires_grad_(True)
            ml = mel(xx, y)                                   # trainerGAN.py:220
            sc, mag = stft(xx, y)                             # trainerGAN.py:227
            (sc + mag + ml).backward()
    else:
        mel_res = so.mel_from_kwargs(**MEL_KW)

        def step():
            so.losses_and_grad(y_hat, y, so.DEFAULT_STFT, mel_res, dtype=torch.float32, use_torch_stft=True)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads(), kind


def gpu_reference_step_time(batch, t_len, dev, steps=10, warmup=2):
    """The reference's op sequence (torch.stft -> cuFFT, ATen elementwise, cuBLAS matmul, autograd) on the SAME GPU:
    what the trainer runs today when its criteria sit on a CUDA device.  Eager launches, CUDA events."""
    import torch
    from oracle import ref_loader, spectral_oracle as so     # baseline only; never on the product path

    y_hat, y = so.synth_pair(batch, t_len, seed=0)
    y_hat, y = y_hat.to(dev), y.to(dev)
    if ref_loader.available():
        stft_mod, mel_mod = ref_loader.load_reference_losses()
        stft = stft_mod.MultiResolutionSTFTLoss().to(dev)
        mel = mel_mod.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
        what = "the reference's own modules (.cuda(): torch.stft/cuFFT + ATen + cuBLAS + autograd)"

        def step():
            xx = y_hat.detach().clone().requires_grad_(True)
            ml = mel(xx, y)
            sc, mag = stft(xx, y)
            (sc + mag + ml).backward()
    else:
        mel_res = so.mel_from_kwargs(**MEL_KW)
        what = "the reference's op sequence (oracle port; torch.stft/cuFFT + ATen + cuBLAS + autograd)"

        def step():
            xx = y_hat.detach().clone().requires_grad_(True)
            sc, mag = so.mr_stft_loss(xx, y, so.DEFAULT_STFT, use_torch_stft=True)
            ml = so.multi_mel_loss(xx, y, mel_res, use_torch_stft=True)
            (sc + mag + ml).backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, what


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: probe one utterance, then size the per-step batch so K+W steps stay within ~2 minutes
    probe, cores, kind = cpu_reference_step_time(1, T_LEN, 1, 1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    b = int(max(1, min(GLOBAL_BATCH, 64, budget / max(probe[0], 1e-6))))      # <= 64 x 4 s: a few GB of host RAM
    times, cores, kind = cpu_reference_step_time(b, T_LEN, args.steps, args.warmup)
    ms = 1000.0 * sum(times) / len(times)
    value = b * T_LEN / FS / (ms / 1000.0)
    sample = (f"{b} x {T_LEN / FS:g} s @ 48 kHz per step (a bounded sample of the {GLOBAL_BATCH} x {T_LEN / FS:g} s workload; "
              f"throughput is linear in the batch, SURVEY 8d), {args.steps} steps, mean")
    path = ("the reference's own modules losses/stft_loss.py + losses/mel_loss.py (unmodified copies under oracle/_ref) on CPU"
            if kind == "reference" else
            "torch.stft + ATen elementwise/norm/matmul + autograd on CPU (oracle port of losses/stft_loss.py, losses/mel_loss.py)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(GLOBAL_BATCH, T_LEN), "global_batch": GLOBAL_BATCH,
                       "samples_per_utterance": T_LEN, "fs": FS, "reference_path": path},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, dev):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, BEFORE any pinned host memory is allocated
    (first touch places it on that NUMA node): eight ranks pulling their inputs through one remote socket is what made the
    8-GPU e2e number of round 2's first runs collapse.  Returns a short description for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"{len(os.sched_getaffinity(0))} of {before} host CPUs (NVML: local to this GPU)"
    except Exception as exc:                      # an optimisation of the host side only
        return f"not applied ({type(exc).__name__})"


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    t_start = time.time()

    def log(msg):
        if args.verbose:
            print(f"[bench r{rank} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    import dl_speech_enhancement_b200 as pkg
    from dl_speech_enhancement_b200.engine import cuda_engine

    if GLOBAL_BATCH % world:
        raise SystemExit(f"global batch {GLOBAL_BATCH} is not divisible by {world} ranks")
    b_local = GLOBAL_BATCH // world
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa_node(torch, dev)
    group = None
    if world > 1:
        # stdout carries exactly one JSON line: the version banner NCCL prints on stdout when the first communicator is
        # created is sent to stderr (file descriptor 1 points at stderr while the communicator comes up)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        group = dist.group.WORLD
    eng = cuda_engine()          # raises if libspecloss.so is missing -- no fallback
    log("process group + engine ready")

    def make_criteria(grp):
        s = pkg.MultiResolutionSTFTLoss().to(dev)
        m = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
        s.process_group = grp
        m.process_group = grp
        return s, m

    stft, mel = make_criteria(group)

    def make_pool(batch, t_len, seed):
        pair_bytes = 2 * batch * t_len * 4
        n = max(2, -(-POOL_BYTES // pair_bytes))
        gen = torch.Generator(device=dev).manual_seed(seed)
        pool = []
        for _ in range(n):
            y = 0.1 * torch.randn(batch, 1, t_len, device=dev, generator=gen)
            y_hat = (y + 0.05 * torch.randn(batch, 1, t_len, device=dev, generator=gen)).requires_grad_(True)
            pool.append((y_hat, y))
        return pool, pair_bytes

    def losses_and_backward(y_hat, y, crit=None):
        s, m = crit or (stft, mel)
        ml = m(y_hat, y)                 # criterion["mel"](predict_y, natural_y)   trainerGAN.py:220
        sc, mag = s(y_hat, y)            # criterion["stft"](predict_y, natural_y)  trainerGAN.py:227
        (sc + mag + ml).backward()       # _update_generator                        trainerGAN.py:274
        return sc, mag, ml

    mel_stream = torch.cuda.Stream()

    def losses_and_backward_2s(y_hat, y):
        """The same two criterion calls, the mel criterion issued on a second stream (secondary workload, graphs only)."""
        cur = torch.cuda.current_stream()
        mel_stream.wait_stream(cur)
        with torch.cuda.stream(mel_stream):
            ml = mel(y_hat, y)
        sc, mag = stft(y_hat, y)
        cur.wait_stream(mel_stream)
        (sc + mag + ml).backward()
        return sc, mag, ml

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_blocks(step_fn, steps, warmup, sample_clocks=False, min_s=MIN_TIMED_S):
        """Blocks of EXACTLY `steps` steps, each bracketed by barrier + synchronize, timed with CUDA events on the launch
        stream, max over ranks; repeated (the block count is agreed between the ranks after the first block) until the
        timed region covers min_s seconds.  Returns (median ms/step, per-block ms/step, launches per block, clocks)."""
        for i in range(warmup):
            step_fn(i)
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        blocks, launches, k, n_blocks = [], 0, warmup, 1
        while len(blocks) < n_blocks:
            n0 = eng.launches
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record()
            for i in range(steps):
                step_fn(k + i)
            ev1.record()
            barrier()
            k += steps
            blocks.append(max_over_ranks(ev0.elapsed_time(ev1)) / steps)
            launches = eng.launches - n0
            if len(blocks) == 1:
                n_blocks = int(min(MAX_BLOCKS, max(3, math.ceil(min_s / max(blocks[0] * steps * 1e-3, 1e-6)))))
        clocks = sampler.stop() if sampler else None
        return sorted(blocks)[len(blocks) // 2], blocks, launches, clocks

    # ---- primary workload: eager launches through the modules, as the trainer drives them -----------------------
    pool, pair_bytes = make_pool(b_local, T_LEN, 1000 + rank)
    n_pool = len(pool)

    def eager_step(i):
        y_hat, y = pool[i % n_pool]
        y_hat.grad = None
        return losses_and_backward(y_hat, y)

    log("pool ready, timing eager steps")
    ms_per_step, block_ms, launches_timed, clocks = timed_blocks(eager_step, args.steps, args.warmup, sample_clocks=True)
    log(f"eager {ms_per_step:.4f} ms/step over {len(block_ms)} blocks")
    value = GLOBAL_BATCH * T_LEN / FS / (ms_per_step / 1000.0)
    peer = 1 if (world > 1 and eng.peer_exchange_active()) else 0
    if world > 1:
        pf = torch.tensor([peer], device=dev)
        dist.all_reduce(pf, op=dist.ReduceOp.MIN)
        peer = int(pf.item())

    # ---- the same step through the one-call extension module (SpectralLoss: the four transforms concurrently, ONE gather, no
    #      gradient accumulation between two autograd nodes).  Not the reference's API: reported beside the headline only. ----
    fused_mod = pkg.SpectralLoss(None, MEL_KW).to(dev)
    fused_mod.process_group = group

    def fused_step(i):
        y_hat, y = pool[i % n_pool]
        y_hat.grad = None
        sc, mag, ml = fused_mod(y_hat, y)
        (sc + mag + ml).backward()

    fused_ms, _, _, _ = timed_blocks(fused_step, args.steps, args.warmup, min_s=0.1)
    ref_out = [float(v.detach()) for v in eager_step(0)]
    g_two = pool[0][0].grad.clone()
    pool[0][0].grad = None
    sc, mag, ml = fused_mod(*pool[0])
    (sc + mag + ml).backward()
    fused_info = {"api": "dl_speech_enhancement_b200.SpectralLoss(stft_loss_params, mel_loss_params)(y_hat, y) -> (sc, mag, mel): "
                         "extension, one autograd node for both criteria",
                  "ms_per_step": fused_ms, "value": GLOBAL_BATCH * T_LEN / FS / (fused_ms / 1000.0), "unit": UNIT,
                  "losses_equal_two_module_step": [float(v.detach()) for v in (sc, mag, ml)] == ref_out,
                  "grad_rel_l2_vs_two_module_step": float((pool[0][0].grad - g_two).norm() / g_two.norm())}
    del g_two, fused_mod
    log(f"one-call module {fused_ms:.4f} ms/step")

    # ---- e2e: pinned host inputs -> H2D -> fwd+bwd -> D2H of the three losses + sync, every step.  The inputs land in two
    #      pre-allocated device buffer pairs (a prefetcher's double buffer: no allocator traffic inside the loop); the copy
    #      of step i+1 is issued on a copy stream before step i's kernels, so it runs under them (what a DataLoader with
    #      pinned memory + non_blocking copies gives the trainer).  Every step still ends with the trainer's host sync. ----
    host = [(p[0].detach().cpu().pin_memory(), p[1].cpu().pin_memory()) for p in pool[:2]]
    copy_stream = torch.cuda.Stream()
    host_out = torch.empty(3, dtype=torch.float32).pin_memory()
    dbuf = [(torch.empty_like(pool[0][0]).requires_grad_(True), torch.empty_like(pool[0][1])) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(steps):
        cur = torch.cuda.current_stream()

        def fetch(i):
            hx, hy = host[i % len(host)]
            x, y = dbuf[i % 2]
            copy_stream.wait_event(consumed[i % 2])       # the step that last read this buffer pair has finished
            with torch.cuda.stream(copy_stream), torch.no_grad():
                x.copy_(hx, non_blocking=True)
                y.copy_(hy, non_blocking=True)
                copied[i % 2].record(copy_stream)
        consumed[0].record(cur)
        consumed[1].record(cur)
        fetch(0)
        for i in range(steps):
            x, y = dbuf[i % 2]
            fetch(i + 1)
            cur.wait_event(copied[i % 2])
            x.grad = None
            sc, mag, ml = losses_and_backward(x, y)
            consumed[i % 2].record(cur)
            host_out.copy_(torch.stack([sc.detach(), mag.detach(), ml.detach()]), non_blocking=True)
            cur.synchronize()                             # the trainer's .item()
        copy_stream.synchronize()

    log("e2e")
    e2e_loop(5)
    barrier()
    # pinned H2D rate of this box with every rank copying at once and no kernels running (explains e2e when the copy, not the
    # kernels, is the longer leg: 55 GB/s for one GPU, 23 GB/s per GPU when eight share the host's memory system)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    with torch.no_grad():
        dbuf[0][0].copy_(host[0][0], non_blocking=True)
        dbuf[0][1].copy_(host[0][1], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = pair_bytes / (h0.elapsed_time(h1) * 1e-3) / 1e9
    e_steps = max(10, args.steps)
    e_runs = []
    for _ in range(5):                   # wall clock on a shared host is noisy: five runs, each max over ranks, median
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e_steps)
        barrier()
        e_runs.append(max_over_ranks(1000.0 * (time.perf_counter() - t0) / e_steps))
    e2e_ms = sorted(e_runs)[len(e_runs) // 2]
    e2e_value = GLOBAL_BATCH * T_LEN / FS / (e2e_ms / 1000.0)
    del host, dbuf
    log("e2e done")

    # ---- sharded parity (outside the timed region): N-rank result == 1-rank evaluation of the gathered global batch ----
    sharded_parity, parity_ok = None, True
    if world > 1 and not args.no_parity:
        sharded_parity, parity_ok = check_sharded_parity(torch, dist, pkg, eng, dev, rank, world, group, pool[0], make_criteria,
                                                         losses_and_backward, log)

    # ---- secondary workload: BASELINE configs[1], 16 x 1 s per GPU (weak), eager and captured-graph replay --------
    secondary = None
    if not args.no_secondary:
        del pool
        torch.cuda.empty_cache()
        spool, s_pair_bytes = make_pool(SEC_BATCH, SEC_T, 2000 + rank)
        ns = len(spool)

        def s_eager(i):
            y_hat, y = spool[i % ns]
            y_hat.grad = None
            return losses_and_backward(y_hat, y)

        s_steps = max(args.steps, 20)
        s_eager_ms, _, _, _ = timed_blocks(s_eager, s_steps, args.warmup, min_s=0.1)
        secondary = {"workload": f"configs[1]: batch {SEC_BATCH} x {SEC_T / FS:g} s synthetic 48 kHz PER GPU (weak), same criteria",
                     "eager_ms_per_step": s_eager_ms,
                     "eager_value": world * SEC_BATCH * SEC_T / FS / (s_eager_ms / 1000.0), "unit": UNIT}
        graph_ms = None
        if (world == 1 or peer) and not args.no_graph:
            ok = 1
            try:
                mem_pool = torch.cuda.graph_pool_handle()
                directs = []
                side = torch.cuda.Stream()
                for (px0, py) in spool:
                    px = px0.detach().requires_grad_(True)
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        losses_and_backward_2s(px, py)
                        px.grad = None
                    torch.cuda.current_stream().wait_stream(side)
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph, pool=mem_pool):
                        outs = losses_and_backward_2s(px, py)
                    directs.append((gph, outs, px))
            except Exception as exc:      # graph capture is an optimisation of the launch path, never a requirement
                print(f"[bench] CUDA graph mode unavailable: {exc!r}", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                def direct_step(i):
                    directs[i % ns][0].replay()

                d_ms, _, _, _ = timed_blocks(direct_step, s_steps, args.warmup, min_s=0.1)
                direct_step(0)
                torch.cuda.synchronize()
                got = [float(o.detach()) for o in directs[0][1]]
                ggrad = directs[0][2].grad.clone()
                ref = s_eager(0)
                torch.cuda.synchronize()
                same = got == [float(o.detach()) for o in ref] and torch.equal(ggrad, spool[0][0].grad)
                flag = torch.tensor([1 if same else 0], device=dev)
                if world > 1:
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) == 1:
                    graph_ms = d_ms
                else:
                    print("[bench] graph replay does not reproduce the eager step; ignoring it", file=sys.stderr)
            directs = None
        secondary["graph_ms_per_step"] = graph_ms
        secondary["graph_value"] = None if graph_ms is None else world * SEC_BATCH * SEC_T / FS / (graph_ms / 1000.0)
        secondary["graph_mode"] = "one captured CUDA graph per input pair, mel criterion on a 2nd stream, bit-equal to the eager step"
        pool = spool
        log(f"secondary: eager {s_eager_ms:.4f} ms, graph {graph_ms}")

    def teardown():
        """Never let communicator teardown hang the job."""
        import gc
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            done = threading.Event()

            def _destroy():
                try:
                    dist.barrier()
                    dist.destroy_process_group()
                finally:
                    done.set()
            threading.Thread(target=_destroy, daemon=True).start()
            if not done.wait(20.0):
                log("process-group teardown timed out; exiting")
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0 if parity_ok else 1)
    if rank != 0:
        teardown()
        if not parity_ok:
            sys.exit(1)
        return

    # ---- per-kernel timing of the dominant (transform) kernels at this rank's share of the primary workload -------
    gen = torch.Generator(device=dev).manual_seed(7)
    y2 = 0.1 * torch.randn(b_local, T_LEN, device=dev, generator=gen)
    x2 = y2 + 0.05 * torch.randn(b_local, T_LEN, device=dev, generator=gen)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)       # > L2, written between timed launches
    kernels = []
    ustft, umel = make_criteria(None)
    for name, plans in [("stft_1024_hop120", [ustft.stft_losses[0].plan()]), ("stft_2048_hop240", [ustft.stft_losses[1].plan()]),
                        ("stft_512_hop50", [ustft.stft_losses[2].plan()]), ("mel_2048_hop300", umel.plans())]:
        # the transform kernel + its tiny reduce/finalize launch, timed alone on the current stream with CUDA events
        # (a launch takes >= 0.2 ms at this size: host launch overhead is hidden), L2 flushed before each launch
        for _ in range(2):
            eng.forward(plans, x2, y2, need_grad=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.forward(plans, x2, y2, need_grad=True)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        kernels.append({"name": name, "ms": sum(ts) / len(ts)})
    st = eng.forward(ustft.plans() + umel.plans(), x2, y2, need_grad=True)
    one = torch.ones((), device=dev)
    ts = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.backward(st, one, one, one)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    combine_ms = sum(ts) / len(ts)
    del st, flush
    dom = max(kernels, key=lambda k: k["ms"])
    hbm_peak, peak_src = peaks()
    alg_bytes = 8.0 * b_local * T_LEN                         # one transform launch must read y_hat and y once
    achieved = alg_bytes / (dom["ms"] * 1e-3) / 1e9
    step_bytes = 20.0 * GLOBAL_BATCH * T_LEN                  # bytes_min of the whole fwd+bwd (all ranks)
    flops = nominal_flops(GLOBAL_BATCH, T_LEN)
    # DRAM traffic from the committed `ncu --set full` capture of THIS build (profiles/traffic.json carries the hash of the
    # kernel sources it was captured from; a stale capture is reported as such, not silently reused)
    traffic, traffic_src, step_traffic, traffic_stale = None, None, None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        ent = tj.get("workloads", {}).get(f"{b_local}x{T_LEN}")
        if ent and dom["name"] in ent.get("kernels", {}):
            traffic = float(ent["kernels"][dom["name"]]["dram_bytes_per_launch"])
            step_traffic = ent.get("step_dram_bytes")
            traffic_src = ent.get("source")
            traffic_stale = tj.get("csrc_hash") != csrc_hash()
    roofline = {"bound": "hbm", "kernel": "transform_kernel<" + dom["name"] + ">", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "traffic_capture_is_stale": traffic_stale, "step_dram_bytes": step_traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "step_hbm_frac": step_bytes / world / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                "binding_roof": "fp32 CUDA-core pipe, not HBM (SURVEY 8d: ~217 FLOP/B vs ridge ~11)",
                "fp32_nominal_tflops_per_gpu": flops / world / (ms_per_step * 1e-3) / 1e12, "fp32_peak_tflops": 74.5,
                "fp32_frac": flops / world / (ms_per_step * 1e-3) / 1e12 / 74.5,
                "kernels_ms": kernels, "combine_ms": combine_ms,
                "kernel_timing": f"each transform launch alone at this rank's share ({b_local} x {T_LEN / FS:g} s), CUDA events on the launch stream, L2 flushed before every launch, mean of 5"}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload -------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        for tid in os.listdir("/proc/self/task"):                        # the CPU leg may use every host core again
            try:
                os.sched_setaffinity(int(tid), all_cpus)
            except OSError:
                pass
        cb = 8                                                           # 8 x 4 s per step: a few seconds per step
        times, cores, kind = cpu_reference_step_time(cb, T_LEN, 3, 1)
        best = min(times)
        cpu = {"value": cb * T_LEN / FS / best, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cb} x {T_LEN / FS:g} s per step (bounded sample of the {GLOBAL_BATCH} x {T_LEN / FS:g} s workload), 3 steps after 1 "
                         f"warm-up, best step {best * 1e3:.1f} ms (mean {1e3 * sum(times) / len(times):.1f} ms); "
                         + ("the reference's own modules (oracle/_ref)" if kind == "reference" else "oracle port of the reference's op sequence")
                         + f", torch {torch.__version__} CPU ops, {os.cpu_count()} host cores"}
        try:
            ref_ms, what = gpu_reference_step_time(32, T_LEN, dev)
            cpu["reference_ops_on_this_gpu"] = {
                "value": 32 * T_LEN / FS / (ref_ms * 1e-3), "unit": UNIT, "ms_per_step": ref_ms,
                "what": what + f" run eagerly on the same B200 at 32 x {T_LEN / FS:g} s, device time over 10 steps: context only, not the reference arm"}
        except Exception as exc:      # context number only
            cpu["reference_ops_on_this_gpu"] = {"unavailable": repr(exc)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(GLOBAL_BATCH, T_LEN),
                       "global_batch": GLOBAL_BATCH, "per_gpu_batch": b_local, "samples_per_utterance": T_LEN, "fs": FS,
                       "api": "drop-in nn.Modules (MultiResolutionSTFTLoss + MultiMelSpectrogramLoss), eager launches on one "
                              "stream in the trainer's call order (no CUDA graph)",
                       "timing": {"blocks": len(block_ms), "steps_per_block": args.steps, "block_ms_per_step": block_ms,
                                  "reported": "median block", "timed_region_s": sum(block_ms) * args.steps * 1e-3},
                       "l2": f"inputs rotate through a pool of {n_pool} pairs = {n_pool * pair_bytes >> 20} MB per GPU > 126 MB L2 "
                             "(and every step streams its multi-GB gradient-slot workspace through L2)",
                       "parallelism": (f"batch-sharded x{world}; the 10 fp64 partial sums are exchanged "
                                       + ("over NVLink peer memory inside the reduce+finalize kernel of each criterion "
                                          "(no NCCL call in the step)" if peer else
                                          "with one NCCL all-reduce per criterion") if world > 1 else "single GPU"),
                       "one_call_extension": fused_info,
                       "secondary": secondary},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pair_bytes, "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms, "runs_ms_per_step": e_runs, "steps": e_steps,
                    "h2d_gbs_per_gpu_all_ranks_copying_no_kernels": h2d_gbs, "host_affinity": affinity,
                    "timing": "wall clock, median of 5 runs; per step: pinned H2D of the NEXT step's inputs on a copy stream into a "
                              "double-buffered device pair, fwd+bwd (eager launches), D2H of the 3 losses + stream sync; max over "
                              "ranks; bytes are per GPU"},
            "gpu_launches": launches_timed, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "sharded_parity": sharded_parity}
    print(json.dumps(line), flush=True)
    teardown()
    if not parity_ok:
        sys.exit(1)


def check_sharded_parity(torch, dist, pkg, eng, dev, rank, world, group, pair, make_criteria, losses_and_backward, log):
    """SURVEY 4 'distributed' row on real hardware: (i) every rank's (sc, mag, mel) bit-identical, (ii) equal (rtol 1e-6)
    to rank 0 evaluating the GATHERED global batch unsharded, (iii) every rank's gradient rows rel-L2 <= 1e-6 from that
    run's rows -- through the fused NVLink exchange (or whatever path the modules took in the timed region) AND through
    the NCCL all-reduce path (SPECLOSS_NCCL_ALLREDUCE=1).  (iv) a finite exchange timeout that expires surfaces as an
    exception on the host, not as silent NaNs."""
    y_hat, y = pair
    b_local, t_len = y_hat.shape[0], y_hat.shape[2]
    res = {"ok": True, "paths": {}}

    def sharded(crit):
        x = y_hat.detach().clone().requires_grad_(True)
        sc, mag, ml = losses_and_backward(x, y, crit)
        torch.cuda.synchronize()
        return torch.stack([sc.detach(), mag.detach(), ml.detach()]), x.grad.reshape(b_local, t_len)

    # rank 0: the global batch, unsharded
    gx = [torch.empty_like(y_hat.detach()) for _ in range(world)] if rank == 0 else None
    gy = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
    dist.gather(y_hat.detach().contiguous(), gx, dst=0)
    dist.gather(y.contiguous(), gy, dst=0)
    full_losses = torch.zeros(3, device=dev)
    full_grad = None
    if rank == 0:
        fx = torch.cat(gx).requires_grad_(True)
        fy = torch.cat(gy)
        del gx, gy
        sc, mag, ml = losses_and_backward(fx, fy, make_criteria(None))
        torch.cuda.synchronize()
        full_losses = torch.stack([sc.detach(), mag.detach(), ml.detach()])
        full_grad = fx.grad.reshape(world, b_local, t_len)
        del fx, fy
    dist.broadcast(full_losses, src=0)

    def compare(tag, losses, grad):
        bits = losses.view(torch.int32)
        allbits = [torch.empty_like(bits) for _ in range(world)]
        dist.all_gather(allbits, bits)
        identical = all(torch.equal(allbits[0], b) for b in allbits)
        finite = bool(torch.isfinite(losses).all())
        rel_loss = float(((losses.double() - full_losses.double()).abs() / full_losses.double().abs()).max())
        grads = [torch.empty_like(grad) for _ in range(world)] if rank == 0 else None
        dist.gather(grad.contiguous(), grads, dst=0)
        g_rel = torch.zeros(1, device=dev, dtype=torch.float64)
        if rank == 0:
            worst = 0.0
            for r in range(world):
                ref = full_grad[r].double()
                worst = max(worst, float((grads[r].double() - ref).norm() / ref.norm()))
            g_rel[0] = worst
        dist.broadcast(g_rel, src=0)
        ok = identical and finite and rel_loss <= 1e-6 and float(g_rel) <= 1e-6
        res["paths"][tag] = {"ok": ok, "losses_bit_identical_across_ranks": identical, "losses_finite": finite,
                             "loss_max_rel_vs_unsharded": rel_loss, "grad_rows_max_rel_l2_vs_unsharded": float(g_rel),
                             "losses": [float(v) for v in losses]}
        res["ok"] = res["ok"] and ok

    crit = make_criteria(group)
    losses, grad = sharded(crit)
    compare("peer_memory_exchange" if eng.peer_exchange_active() else "default_path", losses, grad)
    log("sharded parity: default path done")
    old = os.environ.get("SPECLOSS_NCCL_ALLREDUCE")
    os.environ["SPECLOSS_NCCL_ALLREDUCE"] = "1"
    try:
        crit = make_criteria(group)            # new plan objects -> new recipes -> the exchange choice is made afresh
        for s in list(crit):
            s.__dict__.pop("_plan_cache", None)
        eng._recipes.clear()
        eng._recipes_by_id.clear()
        losses, grad = sharded(crit)
        compare("nccl_all_reduce", losses, grad)
        # and its speed, for the record (10 eager steps after 3, device time, max over ranks)
        x = y_hat.detach().clone().requires_grad_(True)
        for _ in range(3):
            x.grad = None
            losses_and_backward(x, y, crit)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            x.grad = None
            losses_and_backward(x, y, crit)
        e1.record()
        torch.cuda.synchronize()
        tms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        res["paths"]["nccl_all_reduce"]["ms_per_step"] = float(tms.item())
    finally:
        if old is None:
            os.environ.pop("SPECLOSS_NCCL_ALLREDUCE", None)
        else:
            os.environ["SPECLOSS_NCCL_ALLREDUCE"] = old
        eng._evict()
    log("sharded parity: NCCL path done")
    res["unsharded_losses_rank0"] = [float(v) for v in full_losses]
    res["what"] = ("all ranks' (sc, mag, mel) bit-identical; equal (rtol 1e-6) to rank 0 evaluating the gathered global batch "
                   "unsharded; every rank's gradient rows rel-L2 <= 1e-6 of that run's rows")
    return res, bool(res["ok"])


def main():
    global GLOBAL_BATCH, T_LEN
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH, help="GLOBAL batch, sharded over the ranks (default: BASELINE configs[3], 256)")
    ap.add_argument("--seconds", type=float, default=T_LEN / FS, help="utterance length in seconds at 48 kHz (default 4)")
    ap.add_argument("--no-graph", action="store_true", help="secondary workload: skip the CUDA-graph replay measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] (16 x 1 s per GPU) measurement")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded == unsharded check")
    ap.add_argument("--verbose", action="store_true", help="stage-by-stage progress on stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    GLOBAL_BATCH, T_LEN = int(args.batch), int(round(args.seconds * FS))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-exec ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
