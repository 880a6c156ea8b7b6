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
    w = melfb.mel_filterbank(sr, n_fft, n_mels, fmin, fmax).T        # (K, M) as the melmat buffer
    t = {k: v.numpy() for k, v in mel_tables(w).items()}
    rebuilt = np.zeros_like(w)
    for m in range(n_mels):
        s, n, p = t["mel_row_start"][m], t["mel_row_len"][m], t["mel_row_ptr"][m]
        rebuilt[s:s + n, m] = t["mel_row_val"][p:p + n]
    np.testing.assert_array_equal(rebuilt, w)
    rebuilt2 = np.zeros_like(w)
    k = np.arange(w.shape[0])
    rebuilt2[k, t["bin_m0"]] += t["bin_w0"]
    rebuilt2[k, t["bin_m0"] + 1] += t["bin_w1"]
    np.testing.assert_array_equal(rebuilt2, w)


def test_dense_filterbank_rejected():
    with pytest.raises(NotImplementedError):
        mel_tables(np.ones((513, 8), np.float32))
