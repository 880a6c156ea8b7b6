#!/usr/bin/env python3
"""Shared-memory race check of the kernels WITHOUT a GPU tool (compute-sanitizer is closed on this pool, see
profiles/r3d_compute_sanitizer_closed.log): the SIMT emulator plays every lane of a warp with a real host thread and
every __syncwarp()/shuffle with a real barrier, so a missing barrier between a lane's shared-memory write and another
lane's read IS a host data race -- which ThreadSanitizer detects.  This script

  1. builds tests/emu/specloss_emu.cpp with -fsanitize=thread,
  2. runs every kernel body once through the C ABI (losses fwd+grad / forward-only at all three FFT sizes, mel, explicit
     spectrogram forward + backward, shape loss, magnitude losses, the 400-point power-mel metric) under TSan,
  3. runs the positive control: the same workload with ONE __syncwarp() skipped (SPECLOSS_EMU_SKIP_SYNC) must make TSan
     report a race -- proving the detector sees this class of bug.

    python tests/emu/tsan_race_check.py            # prints a summary, exit code 0 = clean and control caught
Test infrastructure only."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

WORKLOAD = r'''
import ctypes, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
from dl_speech_enhancement_b200 import _abi, modules
from dl_speech_enhancement_b200.engine import Engine, TransformPlan, twiddle_table
from dl_speech_enhancement_b200._abi import SPL_KIND_STFT
from dl_speech_enhancement_b200.functional import spectral_losses, spectrogram, shape_loss, magnitude_loss
eng = Engine(_abi.bind(ctypes.CDLL(%(so)r)))
g = torch.Generator().manual_seed(0)
y = 0.1 * torch.randn(2, 1300, generator=g)
x = (y + 0.05 * torch.randn(2, 1300, generator=g)).requires_grad_(True)
stft = modules.MultiResolutionSTFTLoss()
mel = modules.MultiMelSpectrogramLoss(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], num_mels=80, fmin=0, fmax=24000, log_base=None)
gen = modules.MultiMelSpectrogramLoss(fs=24000, fft_sizes=[1024, 512], hop_sizes=[256, 128], win_lengths=[None, 400])
outs = spectral_losses(x, y, stft.plans() + mel.plans() + gen.plans(), engine=eng)
sum(outs).backward()
with torch.no_grad():
    spectral_losses(x, y, stft.plans() + mel.plans(), engine=eng)
x2 = x.detach().clone().requires_grad_(True)
for n_fft, hop, win in ((512, 50, 240), (1024, 120, 600), (2048, 240, 1200)):
    plan = TransformPlan(SPL_KIND_STFT, n_fft, hop, win, 1e-7, torch.hann_window(win), twiddle_table(n_fft))
    a = spectrogram(x2, plan, eng)
    b = spectrogram(y, plan, eng)
    (magnitude_loss(a, b, 0, eng) + magnitude_loss(a, b, 1, eng)).backward()
mplan = mel.mel_transfers[0].plan()
eng.spectrogram_backward(mplan, x.detach(), torch.ones(2, 80, 1 + 1300 // 300))
shape_loss(x, y, (300, 200, 100), engine=eng).backward()
shape_loss(x, y, (300, 77), engine=eng).backward()
m = modules.MelL1(48000)
eng.melpow_l1(x.detach(), y, 400, 200, m.window, m._twiddle, m.n_mels, m._mel_ptr, m._mel_ent, True)
print("workload done", [round(float(o.detach()), 6) for o in outs])
'''


def main():
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    tsan = os.path.realpath(tsan)
    if not os.path.exists(tsan):
        print("libtsan not found; cannot run the race check")
        return 2
    tmp = tempfile.mkdtemp(prefix="specloss_tsan_")
    so = os.path.join(tmp, "libspecloss_emu_tsan.so")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-fPIC", "-shared", "-pthread", "-I" + HERE, "-o", so,
           os.path.join(HERE, "specloss_emu.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stderr)
        return 2
    script = os.path.join(tmp, "workload.py")
    with open(script, "w") as f:
        f.write(WORKLOAD % {"root": ROOT, "so": so})

    def run(skip):
        env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="halt_on_error=0 exitcode=0 report_signal_unsafe=0")
        env.pop("SPECLOSS_EMU_SKIP_SYNC", None)
        if skip:
            env["SPECLOSS_EMU_SKIP_SYNC"] = str(skip)
        r = subprocess.run([sys.executable, script], capture_output=True, text=True, env=env, timeout=3000)
        out = r.stdout + r.stderr
        races = out.count("WARNING: ThreadSanitizer: data race")
        done = "workload done" in out
        first = ""
        if races:
            i = out.index("WARNING: ThreadSanitizer: data race")
            first = "\n".join(l for l in out[i:].splitlines()[:40] if "spl::" in l or "data race" in l)[:1500]
        return races, done, first

    races, done, first = run(0)
    print(f"clean run: workload completed = {done}, ThreadSanitizer data-race reports = {races}")
    if races:
        print(first)
    ok = done and races == 0
    caught = 0
    tried = 0
    all_controls = "--all-controls" in sys.argv
    for skip in (3, 9, 40):                        # three different barriers of the first kernels
        r, d, f = run(skip)
        tried += 1
        print(f"positive control, __syncwarp() #{skip} of every lane skipped: data-race reports = {r}")
        if r:
            caught += 1
            print("  " + f.replace("\n", "\n  ")[:600])
            if not all_controls:                   # one caught control proves the detector; --all-controls runs the rest
                break
    print(f"controls caught: {caught} of {tried}")
    ok = ok and caught >= 1
    print("RACE CHECK", "PASSED" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
