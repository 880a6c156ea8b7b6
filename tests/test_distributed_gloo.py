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


def test_peer_exchange_protocol_three_ranks(emu_engine):
    """spl_reduce_exchange_finalize (reduce + peer-memory exchange + finalize in one launch, the NVLink path of the
    sharded loss): three "ranks" = three host threads of this process that call into the emulated kernels concurrently,
    their symmetric buffers being plain host arrays mapped into each other's pointer tables.  Three calls in a row
    exercise the epoch / parity alternation.  Every rank must end with the full-batch losses, bit-identical."""
    import ctypes
    import threading

    from conftest import load_golden, plans_for, run_losses

    eng, lib = emu_engine, emu_engine.lib
    g = load_golden("ragged_b3_t5003_2d")                     # (3, 5003): one utterance per rank
    yh, y = g["y_hat"].reshape(3, -1), g["y"].reshape(3, -1)
    t_len, world = yh.shape[1], 3
    plans = plans_for(g)
    nbytes = int(lib.spl_exchange_buffer_bytes())
    bufs = [np.zeros(nbytes, dtype=np.uint8) for _ in range(world)]
    results = [[None] * 3 for _ in range(world)]
    errors = []

    def rank_main(rank):
        try:
            ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in bufs])
            state = np.zeros(2, dtype=np.uint32)
            st = eng.forward(plans, yh[rank:rank + 1].contiguous(), y[rank:rank + 1].contiguous(), need_grad=False)
            for call in range(3):                                # the partial sums stay valid: three exchanges of them
                n_sums = st.sums.numel()
                lsums, gsums = np.zeros(n_sums), np.zeros(n_sums)
                sc, mag, mel = (np.zeros(1, dtype=np.float32) for _ in range(3))
                coefs = np.zeros(2 * st.n, dtype=np.float32)
                rc = lib.spl_reduce_exchange_finalize(st.transforms, st.n, 1, t_len, world, lsums.ctypes.data,
                                                      gsums.ctypes.data, rank, world, ptrs, state.ctypes.data, 0, None,
                                                      sc.ctypes.data, mag.ctypes.data, mel.ctypes.data, coefs.ctypes.data, None)
                assert rc == 0, lib.spl_last_error()
                assert int(state[1]) == call + 1 and int(state[0]) == 0
                results[rank][call] = (float(sc[0]), float(mag[0]), float(mel[0]), gsums.copy(), coefs.copy())
        except Exception as exc:   # surfaced in the main thread
            errors.append(exc)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert all(not t.is_alive() for t in threads)
    full, _ = run_losses(eng, g)
    for call in range(3):
        for rank in range(world):
            sc, mag, mel, gsums, coefs = results[rank][call]
            np.testing.assert_allclose([sc, mag, mel], full, rtol=1e-6)
            assert (sc, mag, mel) == results[0][call][:3]
            np.testing.assert_array_equal(gsums, results[0][call][3])
            np.testing.assert_array_equal(coefs, results[0][call][4])


def test_peer_exchange_timeout_sets_host_visible_error(emu_engine):
    """A finite timeout with a peer that never arrives: the call ends, the error word receives the call's epoch (the host
    reads it without synchronising and raises on its next sharded call) and the losses are NaN -- never silent."""
    import ctypes

    from conftest import load_golden, plans_for

    eng, lib = emu_engine, emu_engine.lib
    g = load_golden("ragged_b3_t5003_2d")
    yh, y = g["y_hat"].reshape(3, -1), g["y"].reshape(3, -1)
    world = 2
    nbytes = int(lib.spl_exchange_buffer_bytes())
    bufs = [np.zeros(nbytes, dtype=np.uint8) for _ in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in bufs])
    state = np.zeros(2, dtype=np.uint32)
    err = np.zeros(1, dtype=np.uint32)
    st = eng.forward(plans_for(g), yh[:1].contiguous(), y[:1].contiguous(), need_grad=False)
    n_sums = st.sums.numel()
    lsums, gsums = np.zeros(n_sums), np.zeros(n_sums)
    sc, mag, mel = (np.zeros(1, dtype=np.float32) for _ in range(3))
    coefs = np.zeros(2 * st.n, dtype=np.float32)
    rc = lib.spl_reduce_exchange_finalize(st.transforms, st.n, 1, yh.shape[1], world, lsums.ctypes.data, gsums.ctypes.data,
                                          0, world, ptrs, state.ctypes.data, 100000, err.ctypes.data,
                                          sc.ctypes.data, mag.ctypes.data, mel.ctypes.data, coefs.ctypes.data, None)
    assert rc == 0, lib.spl_last_error()
    assert int(err[0]) == 1                       # epoch of the failed call
    assert np.isnan(sc[0]) and np.isnan(mag[0]) and np.isnan(mel[0])
    assert np.all(lsums != 0) and not np.any(np.isnan(lsums))     # the local sums are untouched


def test_engine_raises_on_exchange_error_flag(emu_engine):
    """Engine.check_exchange_errors(): a non-zero error word raises SpecLossError and is cleared."""
    from dl_speech_enhancement_b200 import _abi

    err = torch.zeros(1, dtype=torch.int32)
    emu_engine._exchanges[("g", "cpu", 7)] = dict(err=err, timeout_ns=1)
    try:
        emu_engine.check_exchange_errors()        # clean: no raise
        err[0] = 3
        with pytest.raises(_abi.SpecLossError, match="timed out"):
            emu_engine.check_exchange_errors()
        assert int(err[0]) == 0
    finally:
        emu_engine._exchanges.pop(("g", "cpu", 7))
