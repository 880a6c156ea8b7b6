"""Product mel filterbank (restating librosa 0.8.1 filters.mel) vs the oracle's independent
restatement and vs torchaudio's Slaney filterbank (third implementation)."""
import numpy as np
import pytest

from dl_speech_enhancement_b200 import melfb
from dl_speech_enhancement_b200.engine import mel_tables
from oracle import spectral_oracle as so

CASES = [(48000, 2048, 80, 0, 24000), (24000, 2048, 80, 0, 12000), (24000, 2048, 80, 0, 24000),
         (22050, 1024, 80, 80, 7600), (22050, 512, 80, 80, 7600), (22050, 2048, 80, 80, 7600)]


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", CASES)
def test_matches_oracle_and_torchaudio(sr, n_fft, n_mels, fmin, fmax):
    a = melfb.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    b = so.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    assert a.shape == (n_mels, n_fft // 2 + 1) and a.dtype == np.float32
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    import torchaudio
    c = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr,
                                              norm="slaney", mel_scale="slaney").numpy().T
    assert np.abs(a - c).max() <= 3e-6 * np.abs(a).max() + 1e-7


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", CASES)
def test_band_tables_reproduce_matrix(sr, n_fft, n_mels, fmin, fmax):
    from dl_speech_enhancement_b200.engine import fft_geometry, slot_offset
    lanes, _ = fft_geometry(n_fft)
    w = melfb.mel_filterbank(sr, n_fft, n_mels, fmin, fmax).T        # (K, M) as the melmat buffer
    t = {k: v.numpy() for k, v in mel_tables(w, n_fft).items()}
    tasks = t["mel_tasks"].reshape(-1, lanes, 4)
    entries = t["mel_entries"].reshape(-1, lanes, 2)
    slot_to_bin = {slot_offset(n_fft, k): k for k in range(w.shape[0])}    # natural position of bin k in the frame slot
    assert len(slot_to_bin) == w.shape[0]           # amplitude slots are distinct
    rebuilt = np.zeros_like(w)
    seen = set()
    for r in range(tasks.shape[0]):
        for lane in range(lanes):
            head, base = tasks[r, lane, 0], tasks[r, lane, 1]
            row, grp, iters = head & 0xfff, (head >> 12) & 0xff, head >> 20
            for s in range(iters):
                off, wbits = entries[base + s, lane]
                wt = np.int32(wbits).view(np.float32)
                if row == 0xfff:
                    assert wt == 0
                elif wt != 0:
                    rebuilt[slot_to_bin[off], row] += wt
            if row != 0xfff:
                seen.add(row)
                assert lane % grp == lane - (lane // grp) * grp
    assert seen == set(range(n_mels))               # empty filters still get a task
    np.testing.assert_array_equal(rebuilt, w)
    bt = t["bin_tab"].reshape(-1, 4)
    rebuilt2 = np.zeros_like(w)
    k = np.arange(w.shape[0])
    rebuilt2[k, bt[:, 0]] += bt[:, 1].view(np.float32)
    rebuilt2[k, bt[:, 0] + 1] += bt[:, 2].view(np.float32)
    np.testing.assert_array_equal(rebuilt2, w)


def test_dense_filterbank_rejected():
    with pytest.raises(NotImplementedError):
        mel_tables(np.ones((513, 8), np.float32), 1024)
