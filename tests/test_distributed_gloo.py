"""N>1 path on CPU: world_size-2 gloo, batch sharded by utterance, one all-reduce of the fp64 partial
sums.  The kernels run through the SIMT emulator; what is under test is the host logic
the NCCL path shares: sharding, the exchange step, global counts, per-rank gradients."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_so, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes

    from conftest import load_golden, plans_for
    from dl_speech_enhancement_b200 import _abi
    from dl_speech_enhancement_b200.engine import Engine
    from dl_speech_enhancement_b200.functional import spectral_losses

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = Engine(_abi.bind(ctypes.CDLL(emu_so)))
    g = load_golden("uniform_b2_c2_t6000")           # (2, 2, 6000) -> 4 rows, 2 per rank
    yh = g["y_hat"].reshape(4, 6000)
    y = g["y"].reshape(4, 6000)
    x = yh[2 * rank:2 * rank + 2].clone().requires_grad_(True)
    outs = spectral_losses(x, y[2 * rank:2 * rank + 2], plans_for(g), group=dist.group.WORLD, engine=eng)
    sum(outs).backward()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array([float(o.detach()) for o in outs]),
             grad=x.grad.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_equals_full_batch(emu_engine, tmp_path):
    from conftest import EMU_DIR, load_golden, rel_l2, run_losses

    emu_so = os.path.join(EMU_DIR, "libspecloss_emu.so")
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, emu_so, str(tmp_path)), nprocs=2, join=True)
    g = load_golden("uniform_b2_c2_t6000")
    full_losses, full_grad = run_losses(emu_engine, g)
    full_grad = full_grad.reshape(4, 6000)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # every rank reports the losses of the GLOBAL batch ...
    np.testing.assert_allclose(r0["losses"], full_losses, rtol=1e-6)
    np.testing.assert_array_equal(r0["losses"], r1["losses"])
    # ... and the gradients of its own rows, scaled with the global norms / counts
    assert rel_l2(np.concatenate([r0["grad"], r1["grad"]]), full_grad) <= 1e-6
    # and both agree with the reference on the full batch
    np.testing.assert_allclose(r0["losses"], g["loss64"], rtol=1e-4)
    assert rel_l2(np.concatenate([r0["grad"], r1["grad"]]), g["grad64"].reshape(4, 6000)) <= 1e-3
