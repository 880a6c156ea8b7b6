#!/usr/bin/env python3
"""Benchmark of the spectral-loss hot path: MultiResolutionSTFTLoss (3 resolutions) +
MultiMelSpectrogramLoss (2048/300, 80 mels) forward+backward, audio-seconds per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU op sequence (oracle port)

Workload (BASELINE.json configs[1]): batch 16 x 1 s synthetic 48 kHz audio per GPU (weak scaling:
every rank holds its own 16 utterances; the loss partial sums are all-reduced over NCCL so each rank
returns the losses of the global batch).  A step = one fwd+bwd of both criteria through the drop-in
nn.Modules; the input pair of each step is taken round-robin from a pool larger than L2.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000
BATCH = 16
T_LEN = 48000
MEL_KW = dict(fs=FS, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
              num_mels=80, fmin=0, fmax=24000, log_base=None)
STFT_RES = [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)]
METRIC = "spectral_loss_fwd_bwd_audio_seconds_per_second"
UNIT = "audio-s/s"
POOL_BYTES = 192 << 20          # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def nominal_flops(batch, t_len):
    """5 passes x sum_res B*F*2.5*N*log2(N)."""
    import math
    tot = 0.0
    for n, hop, _ in STFT_RES + [(2048, 300, 2048)]:
        tot += batch * (1 + t_len // hop) * 2.5 * n * math.log2(n)
    return 5.0 * tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
# reference arm / CPU baseline: the reference's ATen op sequence on the host cores (oracle port)
# --------------------------------------------------------------------------------------------
def cpu_reference_step_time(batch, t_len, steps, warmup):
    import torch
    from oracle import spectral_oracle as so     # checker/baseline only; never on the product path

    torch.set_num_threads(os.cpu_count())
    mel = so.mel_from_kwargs(**MEL_KW)
    y_hat, y = so.synth_pair(batch, t_len, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        so.losses_and_grad(y_hat, y, so.DEFAULT_STFT, mel, dtype=torch.float32, use_torch_stft=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def gpu_reference_step_time(batch, t_len, dev, steps=20, warmup=3):
    """The reference's op sequence (torch.stft -> cuFFT, ATen elementwise, cuBLAS matmul, autograd) on the SAME GPU:
    what the trainer runs today when its criteria sit on a CUDA device.  Eager launches, CUDA events."""
    import torch
    from oracle import spectral_oracle as so     # baseline only; never on the product path

    mel = so.mel_from_kwargs(**MEL_KW)
    y_hat, y = so.synth_pair(batch, t_len, seed=0)
    y_hat, y = y_hat.to(dev), y.to(dev)

    def step():
        xx = y_hat.detach().clone().requires_grad_(True)
        sc, mag = so.mr_stft_loss(xx, y, so.DEFAULT_STFT, use_torch_stft=True)
        ml = so.multi_mel_loss(xx, y, mel, use_torch_stft=True)
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
    return a.elapsed_time(b) / steps


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: probe one utterance, then size the per-step batch so K+W steps stay within ~2 minutes
    probe, cores = cpu_reference_step_time(1, T_LEN, 1, 1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    b = int(max(1, min(BATCH, budget / max(probe[0], 1e-6))))
    times, cores = cpu_reference_step_time(b, T_LEN, args.steps, args.warmup)
    ms = 1000.0 * sum(times) / len(times)
    value = b * T_LEN / FS / (ms / 1000.0)
    sample = f"{b} x {T_LEN / FS:g} s @ 48 kHz per step (of the {BATCH} x {T_LEN / FS:g} s workload), {args.steps} steps, mean"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("configs[1]: " if (BATCH, T_LEN) == (16, 48000) else "non-default size: ")
                                   + f"batch {BATCH} x {T_LEN / FS:g} s synthetic 48 kHz, MR-STFT(1024/2048/512) + 80-mel hop-300, fwd+bwd",
                       "reference_path": "torch.stft + ATen elementwise/norm/matmul + autograd on CPU (oracle port of losses/stft_loss.py, losses/mel_loss.py)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    t_start = time.time()

    def log(msg):
        if args.verbose:
            print(f"[bench r{rank} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    import dl_speech_enhancement_b200 as pkg
    from dl_speech_enhancement_b200.engine import cuda_engine

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
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
    cuda_engine()          # raises if libspecloss.so is missing -- no fallback
    log("process group + engine ready")

    stft = pkg.MultiResolutionSTFTLoss().to(dev)
    mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
    stft.process_group = group
    mel.process_group = group

    pair_bytes = 2 * BATCH * T_LEN * 4
    n_pool = max(2, POOL_BYTES // pair_bytes)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = []
    for _ in range(n_pool):
        y = 0.1 * torch.randn(BATCH, 1, T_LEN, device=dev, generator=gen)
        y_hat = (y + 0.05 * torch.randn(BATCH, 1, T_LEN, device=dev, generator=gen)).requires_grad_(True)
        pool.append((y_hat, y))

    eng = cuda_engine()

    def losses_and_backward(y_hat, y):
        ml = mel(y_hat, y)                 # criterion["mel"](predict_y, natural_y)   trainerGAN.py:220
        sc, mag = stft(y_hat, y)           # criterion["stft"](predict_y, natural_y)  trainerGAN.py:227
        (sc + mag + ml).backward()
        return sc, mag, ml

    mel_stream = torch.cuda.Stream()

    def losses_and_backward_2s(y_hat, y):
        """The same two criterion calls, the mel criterion issued on a second stream: its kernels (forward, and the
        backward node autograd runs on the forward's stream) overlap the STFT criterion's.  Module API unchanged."""
        cur = torch.cuda.current_stream()
        mel_stream.wait_stream(cur)
        with torch.cuda.stream(mel_stream):
            ml = mel(y_hat, y)
        sc, mag = stft(y_hat, y)
        cur.wait_stream(mel_stream)
        (sc + mag + ml).backward()
        return sc, mag, ml

    globals_losses_and_backward = losses_and_backward

    def eager_step(i):
        y_hat, y = pool[i % n_pool]
        y_hat.grad = None
        return losses_and_backward(y_hat, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, sample_clocks=False):
        for i in range(warmup):
            step_fn(i)
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        n0 = eng.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(steps):
            step_fn(warmup + i)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, eng.launches - n0, clocks

    # ---- eager: one Python-driven launch sequence per step -----------------------------------------
    log("pool ready, timing eager steps")
    eager_ms, eager_launches, clocks = timed(eager_step, args.steps, args.warmup, sample_clocks=True)
    log(f"eager {eager_ms:.4f} ms/step")
    mode, ms_per_step, launches_timed = "eager launches", eager_ms, eager_launches

    # ---- CUDA graph: the same step captured once and replayed (inputs are copied into the graph's static
    #      buffers inside the timed region: device-to-device from the rotating pool for `value`, from pinned
    #      host memory for `e2e`).  Two buffer sets so that e2e can copy step i+1 while step i runs. ---------
    def capture_set(step_fn=None):
        losses_and_backward = step_fn or globals_losses_and_backward
        sx = pool[0][0].detach().clone().requires_grad_(True)
        sy = pool[0][1].clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                sx.grad = None
                losses_and_backward(sx, sy)
        torch.cuda.current_stream().wait_stream(side)
        sx.grad = None
        n_before = eng.launches
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sc, mag, ml = losses_and_backward(sx, sy)
            vec = torch.stack([sc.detach(), mag.detach(), ml.detach()])
        return dict(sx=sx, sy=sy, graph=graph, losses=(sc, mag, ml), vec=vec, launches=eng.launches - n_before)

    # Variants of the captured step: the trainer's literal call sequence on one stream, and the same two criterion calls
    # with the mel criterion on a second stream (no module/API change).  Sharded runs take the second variant only when
    # the exchange step is the in-kernel peer-memory one: NCCL collectives of one communicator stay on one stream.
    # sharded: is the exchange step the in-kernel NVLink peer-memory one on EVERY rank (then the step holds no NCCL call)?
    peer = 1 if (world > 1 and eng.peer_exchange_active()) else 0
    if world > 1:
        pf = torch.tensor([peer], device=dev)
        dist.all_reduce(pf, op=dist.ReduceOp.MIN)
        peer = int(pf.item())
    variants = [("1 stream", losses_and_backward)]
    if (world == 1 or peer) and not args.one_stream:
        variants.append(("mel criterion on a 2nd stream", losses_and_backward_2s))
    graph_ms, sets, graph_variant, graph_times = None, None, None, {}
    for vname, vfn in ([] if args.no_graph else variants):
        ok = 1
        log(f"capturing CUDA graphs ({vname})")
        try:
            vsets = [capture_set(vfn), capture_set(vfn)]
        except Exception as exc:      # graph capture is an optimisation of the launch path, never a requirement
            print(f"[bench] CUDA graph mode ({vname}) unavailable: {exc!r}", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            continue
        g0 = vsets[0]

        def graph_step(i, g0=g0):
            y_hat, y = pool[i % n_pool]
            g0["sx"].data.copy_(y_hat.data)
            g0["sy"].copy_(y)
            g0["graph"].replay()

        log("timing graph replays")
        v_ms, _, clocks_g = timed(graph_step, args.steps, args.warmup, sample_clocks=True)
        log(f"graph ({vname}) {v_ms:.4f} ms/step")
        # the replayed step must reproduce the eager result bit for bit
        graph_step(0)
        ref = eager_step(0)
        torch.cuda.synchronize()
        same = all(float(a.detach()) == float(b.detach()) for a, b in zip(g0["losses"], ref))
        same = same and torch.equal(g0["sx"].grad, pool[0][0].grad)
        flag = torch.tensor([1 if same else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # every rank must take the same branch below
        same = bool(int(flag.item()))
        if not same:
            print(f"[bench] CUDA graph replay ({vname}) does not reproduce the eager step; ignoring it", file=sys.stderr)
            continue
        graph_times[vname] = v_ms
        if graph_ms is None or v_ms < graph_ms:
            graph_ms, sets, graph_variant = v_ms, vsets, vname
            if v_ms < ms_per_step:
                mode, ms_per_step, clocks = f"CUDA graph replay, {vname}", v_ms, clocks_g
                launches_timed = g0["launches"] * args.steps
    # ---- one captured graph PER pool entry (shared memory pool): the replayed step reads its inputs where they already
    #      are in HBM -- no staging copy into static buffers inside the timed region -- and the rotation over the whole
    #      pool (> L2) still makes every step find its inputs outside the cache ---------------------------------------
    graph_direct_ms = None
    if sets is not None and not args.no_direct_graphs:
        ok = 1
        try:
            vfn = dict(variants)[graph_variant]
            mem_pool = torch.cuda.graph_pool_handle()
            directs = []
            side = torch.cuda.Stream()
            for (px0, py) in pool:
                # a fresh leaf over the same storage: its AccumulateGrad node is first used on a side stream (the pool's
                # own leaves were used on the default stream by the eager steps, which a capture must not touch)
                px = px0.detach().requires_grad_(True)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    vfn(px, py)
                    px.grad = None
                torch.cuda.current_stream().wait_stream(side)
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, pool=mem_pool):
                    outs = vfn(px, py)
                directs.append((gph, outs, px))
        except Exception as exc:
            print(f"[bench] per-entry graphs unavailable: {exc!r}", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            def direct_step(i):
                directs[i % n_pool][0].replay()

            d_ms, _, clocks_d = timed(direct_step, args.steps, args.warmup, sample_clocks=True)
            log(f"graph per pool entry ({graph_variant}) {d_ms:.4f} ms/step")
            # entry 0 replayed must reproduce the eager step on entry 0 bit for bit (graphs share one memory pool: the
            # outputs of a graph are only valid until the next replay)
            direct_step(0)
            torch.cuda.synchronize()
            got = [float(o.detach()) for o in directs[0][1]]
            ggrad = directs[0][2].grad.clone()
            ref = eager_step(0)
            torch.cuda.synchronize()
            same = got == [float(o.detach()) for o in ref] and torch.equal(ggrad, pool[0][0].grad)
            flag = torch.tensor([1 if same else 0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                graph_direct_ms = d_ms
                if d_ms < ms_per_step:
                    mode, ms_per_step, clocks = f"CUDA graph replay (one graph per input pair), {graph_variant}", d_ms, clocks_d
            else:
                print("[bench] per-entry graph replay does not reproduce the eager step; ignoring it", file=sys.stderr)
            directs.clear()
    value = world * BATCH * T_LEN / FS / (ms_per_step / 1000.0)

    # ---- e2e: pinned host inputs -> H2D -> fwd+bwd -> D2H of the three losses + sync, every step.  The copy
    #      of step i+1 is issued on a copy stream before step i's result is awaited (what a DataLoader with
    #      pinned memory + non_blocking copies gives the trainer). -------------------------------------
    host = [(p[0].detach().cpu().pin_memory(), p[1].cpu().pin_memory()) for p in pool[:4]]
    copy_stream = torch.cuda.Stream()
    host_out = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_loop_eager(steps):
        def fetch(i):
            hx, hy = host[i % len(host)]
            with torch.cuda.stream(copy_stream):
                x = hx.to(dev, non_blocking=True)
                y = hy.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return x, y, ev
        nxt = fetch(0)
        for i in range(steps):
            x, y, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            nxt = fetch(i + 1)
            x.requires_grad_(True)
            sc, mag, ml = losses_and_backward(x, y)
            x.record_stream(torch.cuda.current_stream())
            y.record_stream(torch.cuda.current_stream())
            host_out.copy_(torch.stack([sc.detach(), mag.detach(), ml.detach()]), non_blocking=True)
            torch.cuda.current_stream().synchronize()            # the trainer's .item()

    def e2e_loop_graph(steps):
        evs = [None, None]

        def issue_copy(i):
            hx, hy = host[i % len(host)]
            st = sets[i % 2]
            with torch.cuda.stream(copy_stream):
                st["sx"].data.copy_(hx, non_blocking=True)
                st["sy"].copy_(hy, non_blocking=True)
                evs[i % 2] = torch.cuda.Event()
                evs[i % 2].record(copy_stream)
        issue_copy(0)
        for i in range(steps):
            st = sets[i % 2]
            torch.cuda.current_stream().wait_event(evs[i % 2])
            issue_copy(i + 1)          # other buffer set: its last consumer (step i-1) was synchronised on
            st["graph"].replay()
            host_out.copy_(st["vec"], non_blocking=True)
            torch.cuda.current_stream().synchronize()            # the trainer's .item()

    e2e_loop = e2e_loop_graph if sets is not None else e2e_loop_eager
    e2e_mode = f"CUDA graph replay, {graph_variant}" if sets is not None else "eager launches"
    log("e2e (" + e2e_mode + ")")
    e2e_loop(max(3, args.warmup // 4))
    barrier()
    # wall-clock numbers on a shared host are noisy (PCIe, wake-up latency of the per-step synchronize): three runs,
    # each the max over ranks, the median reported
    e_steps = max(10, args.steps // 2)
    e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e_steps)
        barrier()
        e_ms = 1000.0 * (time.perf_counter() - t0) / e_steps
        t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_runs.append(float(t.item()))
    e2e_ms = sorted(e_runs)[1]
    e2e_value = world * BATCH * T_LEN / FS / (e2e_ms / 1000.0)
    log("e2e done")

    def teardown():
        """Release the captured graphs before the communicator; never let teardown hang the job."""
        nonlocal sets
        for st in (sets or []):
            st.clear()               # drops the CUDAGraph objects (they hold NCCL work when world > 1)
        sets = None
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
                os._exit(0)
    if rank != 0:
        teardown()
        return

    # ---- per-kernel timing of the dominant (transform) kernels, CUDA events on the launch stream ---
    x2 = pool[0][0].detach().reshape(BATCH, T_LEN)
    y2 = pool[0][1].reshape(BATCH, T_LEN)
    kernels = []
    for name, plans in [("stft_1024_hop120", [stft.stft_losses[0].plan()]), ("stft_2048_hop240", [stft.stft_losses[1].plan()]),
                        ("stft_512_hop50", [stft.stft_losses[2].plan()]), ("mel_2048_hop300", mel.plans())]:
        # `inner` launches captured in one CUDA graph and replayed: device time per launch without host overhead
        # (the transform kernel + the tiny reduce/finalize launch; inputs 6 MB, L2-warm -- the kernel is compute bound)
        inner, reps = 10, 10
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                eng.forward(plans, x2, y2, need_grad=True)
        torch.cuda.current_stream().wait_stream(side)
        kg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(kg):
            for _ in range(inner):
                eng.forward(plans, x2, y2, need_grad=True)
        kg.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            kg.replay()
        b.record()
        torch.cuda.synchronize()
        kernels.append({"name": name, "ms": a.elapsed_time(b) / (reps * inner)})
        del kg
    dom = max(kernels, key=lambda k: k["ms"])
    hbm_peak, peak_src = peaks()
    alg_bytes = 8.0 * BATCH * T_LEN                           # one transform launch must read y_hat and y once
    achieved = alg_bytes / (dom["ms"] * 1e-3) / 1e9
    step_bytes = 20.0 * BATCH * T_LEN                         # bytes_min of the whole fwd+bwd
    flops = nominal_flops(BATCH, T_LEN)
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/traffic.json), per launch
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and dom["name"].startswith("mel_2048") and (BATCH, T_LEN) == (16, 48000):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj["source"].split(":")[0]
    roofline = {"bound": "hbm", "kernel": "transform_kernel<" + dom["name"] + ">", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "step_hbm_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                "binding_roof": "fp32 CUDA-core pipe, not HBM (SURVEY 8d: ~217 FLOP/B vs ridge ~11)",
                "fp32_nominal_tflops": flops / (ms_per_step * 1e-3) / 1e12, "fp32_peak_tflops": 74.5,
                "fp32_frac": flops / (ms_per_step * 1e-3) / 1e12 / 74.5,
                "kernels_ms": kernels}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload -------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_steps = 10 if BATCH * T_LEN <= 16 * 48000 else 3        # bounded: ~1.3 s at configs[1], a few seconds beyond
        times, cores = cpu_reference_step_time(BATCH, T_LEN, cpu_steps, 1)
        best = min(times)
        cpu = {"value": BATCH * T_LEN / FS / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the full {BATCH} x {T_LEN / FS:g} s workload, {cpu_steps} steps after 1 warm-up, best step {best * 1e3:.1f} ms "
                         f"(mean {1e3 * sum(times) / len(times):.1f} ms); torch {torch.__version__} CPU ops, {os.cpu_count()} host cores"}

        try:
            ref_ms = gpu_reference_step_time(BATCH, T_LEN, dev)
            cpu["reference_ops_on_this_gpu"] = {
                "value": BATCH * T_LEN / FS / (ref_ms * 1e-3), "unit": UNIT, "ms_per_step": ref_ms,
                "what": "the reference's op sequence (torch.stft/cuFFT + ATen + cuBLAS + autograd, oracle port) run eagerly "
                        "on the same B200, device time over 20 steps: context only, not the reference arm"}
        except Exception as exc:      # context number only
            cpu["reference_ops_on_this_gpu"] = {"unavailable": repr(exc)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("configs[1]: " if (BATCH, T_LEN) == (16, 48000) else "non-default size: ")
                                   + f"batch {BATCH} x {T_LEN / FS:g} s synthetic 48 kHz per GPU, MR-STFT(1024/2048/512) + 80-mel hop-300, fwd+bwd",
                       "global_batch": world * BATCH, "samples_per_utterance": T_LEN, "fs": FS,
                       "api": "drop-in nn.Modules (MultiResolutionSTFTLoss + MultiMelSpectrogramLoss), " + mode,
                       "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
                       "graph_variants_ms_per_step": graph_times, "graph_per_input_pair_ms_per_step": graph_direct_ms,
                       "l2": f"inputs rotate through a pool of {n_pool} pairs = {n_pool * pair_bytes >> 20} MB > 126 MB L2",
                       "parallelism": (f"batch-sharded x{world}; the 10 fp64 partial sums are exchanged "
                                       + ("over NVLink peer memory inside the reduce+finalize kernel of each criterion "
                                          "(no NCCL call in the step)" if peer else
                                          "with one NCCL all-reduce per criterion") if world > 1 else "single GPU")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pair_bytes, "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms, "runs_ms_per_step": e_runs,
                    "steps": e_steps, "timing": "wall clock, median of 3 runs; per step: pinned H2D of the NEXT step's inputs on a copy stream, fwd+bwd ("
                                                + e2e_mode + "), D2H of the 3 losses + stream sync; max over ranks"},
            "gpu_launches": launches_timed, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    teardown()


def main():
    global BATCH, T_LEN
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=BATCH, help="utterances per GPU (default: BASELINE configs[1], 16)")
    ap.add_argument("--seconds", type=float, default=T_LEN / FS, help="utterance length in seconds at 48 kHz (default 1)")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay measurement")
    ap.add_argument("--no-direct-graphs", action="store_true",
                    help="skip the variant with one captured graph per input pair (no staging copy in the timed region)")
    ap.add_argument("--one-stream", action="store_true",
                    help="do not try the captured step with the mel criterion on a second stream")
    ap.add_argument("--verbose", action="store_true", help="stage-by-stage progress on stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    BATCH, T_LEN = int(args.batch), int(round(args.seconds * FS))
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
