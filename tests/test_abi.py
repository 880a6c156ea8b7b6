"""The C-ABI library builds for sm_100a, loads, exports every symbol include/specloss.h declares,
and rejects bad arguments with messages.  No kernel is launched here (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from dl_speech_enhancement_b200 import _abi
from dl_speech_enhancement_b200.engine import twiddle_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _abi.build_library()
    return _abi.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "specloss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_abi.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.spl_abi_version() == _abi.ABI_VERSION


def test_library_is_sm100a_and_native():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_twiddle_table_matches_native(lib):
    for n in (512, 1024, 2048):
        buf = np.zeros(2 * n, np.float32)
        assert lib.spl_fill_twiddle(n, buf.ctypes.data) == 0
        np.testing.assert_array_equal(buf, twiddle_table(n).numpy())
    assert lib.spl_fill_twiddle(300, buf.ctypes.data) == -1


def _tr(**kw):
    t = _abi.SplTransform()
    t.kind, t.n_fft, t.hop, t.win, t.eps = 0, 1024, 120, 600, 1e-7
    for k, v in kw.items():
        setattr(t, k, v)
    return t


def test_geometry(lib):
    g = _abi.SplGeometry()
    assert lib.spl_geometry_of(ctypes.byref(_tr()), 16, 48000, ctypes.byref(g)) == 0
    assert (g.n_frames, g.n_bins, g.n_sums) == (401, 513, 3)
    assert g.partial_count == (16 * 401 + 32) * 3 and g.gframe_bytes == 16 * 401 * 600 * 8
    assert g.smem_table_bytes == (2 * 32 * 33 + 600) * 4 and 0 < g.smem_warp_bytes <= 32 * 1024
    assert lib.spl_geometry_of(ctypes.byref(_tr()), 256, 192000, ctypes.byref(g)) == 0
    assert g.partial_count == (32768 + 32) * 3     # capped: one row per warp of the launch


@pytest.mark.parametrize("kw,frag", [
    (dict(n_fft=768), "n_fft"), (dict(win=2000), "win"), (dict(hop=0), "hop"),
    (dict(kind=7), "kind"),
])
def test_invalid_arguments_are_reported(lib, kw, frag):
    g = _abi.SplGeometry()
    assert lib.spl_geometry_of(ctypes.byref(_tr(**kw)), 2, 4800, ctypes.byref(g)) == -1
    assert frag in lib.spl_last_error().decode()


def test_hop_larger_than_window_is_legal(lib):
    """torch.stft accepts hop > win_length (frames with gaps between them); so does the library."""
    g = _abi.SplGeometry()
    assert lib.spl_geometry_of(ctypes.byref(_tr(hop=700)), 2, 4800, ctypes.byref(g)) == 0
    assert g.n_frames == 1 + 4800 // 700


def test_too_short_signal_is_an_error(lib):
    g = _abi.SplGeometry()
    assert lib.spl_geometry_of(ctypes.byref(_tr()), 1, 512, ctypes.byref(g)) == -1   # T <= n_fft/2
    assert "reflect" in lib.spl_last_error().decode()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_abi, "_LIB", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_abi.SpecLossError):
        _abi.load_library()


def test_new_entry_points_reject_bad_arguments(lib):
    """Argument checks of the widened ABI (shape loss, spectrogram backward, peer exchange) run before any launch."""
    wl = (ctypes.c_int32 * 3)(300, 200, 100)
    n_rec, n_part = ctypes.c_int64(), ctypes.c_int64()
    assert lib.spl_shape_geometry(2, 250, wl, 3, ctypes.byref(n_rec), ctypes.byref(n_part)) == -1      # window > T
    assert "window length" in lib.spl_last_error().decode()
    assert lib.spl_shape_geometry(2, 9000, wl, 9, ctypes.byref(n_rec), ctypes.byref(n_part)) == -1     # > 8 windows
    assert lib.spl_shape_forward(None, None, 2, 9000, wl, 3, None, None, None, None) == -1
    assert lib.spl_shape_backward(None, 2, 2, 9000, wl, 3, None, None, None) == -1
    assert lib.spl_spectrogram_backward(ctypes.byref(_tr()), None, 2, 9000, None, 513, None, None) == -1
    assert "null" in lib.spl_last_error().decode()
    assert lib.spl_exchange_buffer_bytes() == 2 * 8 * 24 * 8 + 2 * 8 * 4     # 24 = 3 sums x SPL_MAX_TRANSFORMS
    dummy = (ctypes.c_double * 24)()
    state = (ctypes.c_uint32 * 2)()
    ptrs = (ctypes.c_void_p * 2)(ctypes.addressof(dummy), ctypes.addressof(dummy))
    rc = lib.spl_reduce_exchange_finalize(ctypes.byref(_tr()), 1, 2, 9000, 4, ctypes.addressof(dummy), ctypes.addressof(dummy),
                                          5, 2, ptrs, ctypes.addressof(state), 0, None, None, None, None, ctypes.addressof(dummy), None)
    assert rc == -1 and "rank" in lib.spl_last_error().decode()


def test_even_odd_tables_match_the_python_builder(lib):
    """spl_fill_twiddle_eo (C) and engine.twiddle_eo_table (numpy) produce the same fp32 tables."""
    from dl_speech_enhancement_b200.engine import twiddle_eo_table
    n = 2 * 1024 + 2 * 516
    buf = (ctypes.c_float * n)()
    assert lib.spl_fill_twiddle_eo(buf) == 0
    np.testing.assert_array_equal(np.frombuffer(buf, dtype=np.float32), twiddle_eo_table().numpy())
    assert lib.spl_fill_twiddle_eo(None) == -1


def test_one_call_entry_points_reject_bad_arguments(lib):
    """spl_loss_forward / spl_loss_backward check their template and offsets before any launch."""
    t = _tr()
    off = (ctypes.c_int64 * 1)(0)
    dummy = (ctypes.c_double * 64)()
    a = ctypes.addressof(dummy)
    assert lib.spl_loss_forward(ctypes.byref(t), 0, a, a, 2, 9000, a, off, None, 0, 0, None, None, None, a, None) == -1
    assert "transforms" in lib.spl_last_error().decode()
    assert lib.spl_loss_forward(ctypes.byref(t), 1, a, a, 2, 9000, None, off, None, 0, 0, None, None, None, a, None) == -1
    assert lib.spl_loss_forward(ctypes.byref(t), 1, a, a, 2, 9000, a, None, None, 0, 0, None, None, None, a, None) == -1
    assert "offsets" in lib.spl_last_error().decode()
    assert lib.spl_loss_backward(ctypes.byref(t), 1, 2, 9000, a, None, 0, None, None, None, a, None) == -1
    assert "gradient workspace" in lib.spl_last_error().decode()


def test_geometry_with_runs(lib, monkeypatch):
    """Workspace of the overlap-add ring (SPECLOSS_RUN_FRAMES is read at the first use of the CUDA library, so this
    only checks the default here: one gradient slot per frame)."""
    g = _abi.SplGeometry()
    assert lib.spl_geometry_of(ctypes.byref(_tr()), 256, 192000, ctypes.byref(g)) == 0
    assert g.gframe_bytes == 256 * 1601 * 600 * 8
