"""Multi-GPU parity on real hardware (SURVEY 4 'distributed' row, 8e): two ranks, one process per GPU (spawned here),
NCCL process group, the batch sharded by utterance.  The losses every rank returns must be bit-identical across ranks and
equal to the single-GPU evaluation of the full batch; each rank's gradient rows must equal that run's rows.  Exercised
through BOTH exchange paths: the fused reduce + NVLink peer-memory exchange + finalize kernel
(spl_reduce_exchange_finalize) and the NCCL all-reduce path (SPECLOSS_NCCL_ALLREDUCE=1).  Skips on a 1-GPU box; the
same check runs inside `bench.py --gpus N` (`sharded_parity`)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

MEL48 = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window",
             num_mels=80, fmin=0, fmax=24000, log_base=None)


def _worker(rank, world, port, out_dir, nccl_path, timeout_s):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if nccl_path:
        os.environ["SPECLOSS_NCCL_ALLREDUCE"] = "1"
    if timeout_s:
        os.environ["SPECLOSS_EXCHANGE_TIMEOUT_S"] = str(timeout_s)
    import torch.distributed as dist

    import dl_speech_enhancement_b200 as pkg
    from dl_speech_enhancement_b200.engine import cuda_engine
    from oracle import spectral_oracle as so

    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    y_hat, y = so.synth_pair(8, 24000, seed=14)
    per = 8 // world
    stft = pkg.MultiResolutionSTFTLoss().to(dev)
    mel = pkg.MultiMelSpectrogramLoss(**MEL48).to(dev)
    stft.process_group = mel.process_group = dist.group.WORLD
    t = y[per * rank:per * (rank + 1)].to(dev)
    rows = []
    for call in range(3):                                # three calls: epoch / parity alternation of the exchange slots
        x = y_hat[per * rank:per * (rank + 1)].to(dev).requires_grad_(True)
        ml = mel(x, t)
        sc, mag = stft(x, t)
        (sc + mag + ml).backward()
        torch.cuda.synchronize()
        rows.append((np.array([float(sc), float(mag), float(ml)], dtype=np.float32), x.grad.cpu().numpy()))
    peer = cuda_engine().peer_exchange_active()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.stack([r[0] for r in rows]),
             grad=np.stack([r[1] for r in rows]), peer=np.array(int(peer)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nccl_path", [False, True], ids=["peer_memory_exchange", "nccl_all_reduce"])
def test_two_rank_sharded_equals_single_gpu(tmp_path, nccl_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import dl_speech_enhancement_b200 as pkg
    from oracle import spectral_oracle as so

    port = 29100 + (os.getpid() % 800) + (50 if nccl_path else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), nccl_path, 0), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert bool(r0["peer"]) == (not nccl_path)           # the path under test is the one that ran
    # single-GPU evaluation of the full batch
    dev = torch.device("cuda:0")
    y_hat, y = so.synth_pair(8, 24000, seed=14)
    stft = pkg.MultiResolutionSTFTLoss().to(dev)
    mel = pkg.MultiMelSpectrogramLoss(**MEL48).to(dev)
    x = y_hat.to(dev).requires_grad_(True)
    ml = mel(x, y.to(dev))
    sc, mag = stft(x, y.to(dev))
    (sc + mag + ml).backward()
    full = np.array([float(sc), float(mag), float(ml)], dtype=np.float32)
    fgrad = x.grad.cpu().numpy()
    for call in range(3):
        assert np.array_equal(r0["losses"][call].view(np.int32), r1["losses"][call].view(np.int32))     # bit-identical
        np.testing.assert_allclose(r0["losses"][call], full, rtol=1e-6)
        got = np.concatenate([r0["grad"][call], r1["grad"][call]])
        rel = float(np.linalg.norm(got.astype(np.float64) - fgrad) / np.linalg.norm(fgrad.astype(np.float64)))
        assert rel <= 1e-6, rel
    # and against the fp64 oracle on the full batch
    ref, gref = so.losses_and_grad(y_hat, y, so.DEFAULT_STFT, so.mel_from_kwargs(**MEL48), dtype=torch.float64)
    np.testing.assert_allclose(r0["losses"][0], ref, rtol=1e-4)
    got = np.concatenate([r0["grad"][0], r1["grad"][0]]).astype(np.float64)
    assert float(np.linalg.norm(got - gref.numpy()) / np.linalg.norm(gref.numpy())) <= 1e-3


def test_tensors_on_a_non_current_device():
    """Inputs on cuda:1 while cuda:0 is current (plain .to('cuda:1') without set_device -- the reference modules handle it):
    the library must launch on cuda:1's stream with cuda:1's SM count.  ADVICE r1 (engine.py device guard)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import dl_speech_enhancement_b200 as pkg
    from oracle import spectral_oracle as so

    torch.cuda.set_device(0)
    y_hat, y = so.synth_pair(2, 9600, seed=3)
    outs = []
    for d in ("cuda:0", "cuda:1"):
        dev = torch.device(d)
        stft = pkg.MultiResolutionSTFTLoss().to(dev)
        mel = pkg.MultiMelSpectrogramLoss(**MEL48).to(dev)
        shape = pkg.MultiWindowShapeLoss().to(dev)
        x = y_hat.to(dev).requires_grad_(True)
        sc, mag = stft(x, y.to(dev))
        ml = mel(x, y.to(dev))
        sh = shape(x, y.to(dev))
        (sc + mag + ml + sh).backward()
        assert torch.cuda.current_device() == 0
        outs.append(([float(sc), float(mag), float(ml), float(sh)], x.grad.cpu().numpy()))
        assert x.grad.device == dev
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
