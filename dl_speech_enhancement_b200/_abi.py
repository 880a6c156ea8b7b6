"""ctypes binding of libspecloss.so (C ABI declared in include/specloss.h).

The library is built in-tree by `build_library()` (nvcc, sm_100a) -- `__graft_entry__.build()`
calls it.  There is no CPU fallback: if the shared object is missing or does not load,
`load_library()` raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspecloss.so")
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 15

SPL_KIND_STFT = 0
SPL_KIND_MEL = 1
SPL_MAX_TRANSFORMS = 8

# No -ftz: the packed fp32 instructions of sm_100 (FADD2/FMUL2/FFMA2) have no flush-to-zero form, and with
# -ftz=true ptxas emulates it with extra scalar FADDs and MOVs (fft32: 216 -> 471 instructions).  The two
# hot transcendental calls use explicit .ftz PTX instead (specloss_kernels.cuh: spl_fast_rsqrt / spl_fast_log2).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class SplTransform(ctypes.Structure):
    _fields_ = [
        ("kind", c_int32), ("n_fft", c_int32), ("hop", c_int32), ("win", c_int32),
        ("eps", c_float),
        ("window", c_void_p), ("twiddle", c_void_p),
        ("n_mels", c_int32), ("inv_ln_base", c_float),
        ("mel_tasks", c_void_p), ("mel_entries", c_void_p), ("mel_rounds", c_int32), ("mel_entry_rows", c_int32), ("bin_tab", c_void_p),
        ("twiddle_eo", c_void_p), ("mel_entries_eo", c_void_p),
        ("partials", c_void_p), ("gframes", c_void_p),
    ]


class SplGeometry(ctypes.Structure):
    _fields_ = [
        ("n_frames", c_int32), ("n_bins", c_int32), ("n_sums", c_int32), ("reserved", c_int32),
        ("partial_count", c_int64), ("gframe_bytes", c_int64), ("smem_table_bytes", c_int64), ("smem_warp_bytes", c_int64),
    ]


EXPORTS = ("spl_abi_version", "spl_last_error", "spl_fill_twiddle", "spl_fill_twiddle_eo", "spl_geometry_of", "spl_forward",
           "spl_reduce", "spl_finalize", "spl_reduce_finalize", "spl_exchange_buffer_bytes", "spl_reduce_exchange_finalize", "spl_backward", "spl_spectrogram", "spl_spectrogram_backward", "spl_mel_project",
           "spl_shape_geometry", "spl_shape_forward", "spl_shape_finalize", "spl_shape_backward",
           "spl_mag_loss_geometry", "spl_mag_loss_forward", "spl_mag_loss_backward",
           "spl_melpow_geometry", "spl_melpow_l1", "spl_loss_forward", "spl_loss_backward")


class SpecLossError(RuntimeError):
    pass


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Attach argument/return types to every exported entry point."""
    lib.spl_abi_version.restype = c_int32
    lib.spl_abi_version.argtypes = []
    lib.spl_last_error.restype = c_char_p
    lib.spl_last_error.argtypes = []
    lib.spl_fill_twiddle.restype = c_int32
    lib.spl_fill_twiddle.argtypes = [c_int32, c_void_p]
    lib.spl_fill_twiddle_eo.restype = c_int32
    lib.spl_fill_twiddle_eo.argtypes = [c_void_p]
    lib.spl_geometry_of.restype = c_int32
    lib.spl_geometry_of.argtypes = [POINTER(SplTransform), c_int32, c_int32, POINTER(SplGeometry)]
    lib.spl_forward.restype = c_int32
    lib.spl_forward.argtypes = [POINTER(SplTransform), c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.spl_reduce.restype = c_int32
    lib.spl_reduce.argtypes = [POINTER(SplTransform), c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.spl_finalize.restype = c_int32
    lib.spl_finalize.argtypes = [POINTER(SplTransform), c_int32, c_void_p, c_int64, c_int32,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_reduce_finalize.restype = c_int32
    lib.spl_reduce_finalize.argtypes = [POINTER(SplTransform), c_int32, c_int32, c_int32, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_exchange_buffer_bytes.restype = c_int64
    lib.spl_exchange_buffer_bytes.argtypes = []
    lib.spl_reduce_exchange_finalize.restype = c_int32
    lib.spl_reduce_exchange_finalize.argtypes = [POINTER(SplTransform), c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p,
                                                 c_int32, c_int32, POINTER(c_void_p), c_void_p, c_int64, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_backward.restype = c_int32
    lib.spl_backward.argtypes = [POINTER(SplTransform), c_int32, c_int32, c_int32, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_spectrogram.restype = c_int32
    lib.spl_spectrogram.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                    c_float, c_void_p, c_void_p, c_int32, c_void_p]
    lib.spl_spectrogram_backward.restype = c_int32
    lib.spl_spectrogram_backward.argtypes = [POINTER(SplTransform), c_void_p, c_int32, c_int32, c_void_p, c_int32,
                                             c_void_p, c_void_p]
    lib.spl_mel_project.restype = c_int32
    lib.spl_mel_project.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_float, c_float, c_void_p, c_void_p]
    lib.spl_shape_geometry.restype = c_int32
    lib.spl_shape_geometry.argtypes = [c_int32, c_int32, POINTER(c_int32), c_int32, POINTER(c_int64), POINTER(c_int64)]
    lib.spl_shape_forward.restype = c_int32
    lib.spl_shape_forward.argtypes = [c_void_p, c_void_p, c_int32, c_int32, POINTER(c_int32), c_int32,
                                      c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_shape_finalize.restype = c_int32
    lib.spl_shape_finalize.argtypes = [c_void_p, c_int64, c_int32, POINTER(c_int32), c_int32, c_void_p, c_void_p]
    lib.spl_shape_backward.restype = c_int32
    lib.spl_shape_backward.argtypes = [c_void_p, c_int32, c_int64, c_int32, POINTER(c_int32), c_int32,
                                       c_void_p, c_void_p, c_void_p]
    lib.spl_mag_loss_geometry.restype = c_int32
    lib.spl_mag_loss_geometry.argtypes = [c_int64, POINTER(c_int64)]
    lib.spl_mag_loss_forward.restype = c_int32
    lib.spl_mag_loss_forward.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_mag_loss_backward.restype = c_int32
    lib.spl_mag_loss_backward.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]
    lib.spl_melpow_geometry.restype = c_int32
    lib.spl_melpow_geometry.argtypes = [c_int32, c_int32, c_int32, c_int32, POINTER(c_int64)]
    lib.spl_melpow_l1.restype = c_int32
    lib.spl_melpow_l1.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.spl_loss_forward.restype = c_int32
    lib.spl_loss_forward.argtypes = [POINTER(SplTransform), c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                     POINTER(c_int64), POINTER(c_int64), c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]
    lib.spl_loss_backward.restype = c_int32
    lib.spl_loss_backward.argtypes = [POINTER(SplTransform), c_int32, c_int32, c_int32, c_void_p, POINTER(c_int64), c_int64,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    ver = lib.spl_abi_version()
    if ver != ABI_VERSION:
        raise SpecLossError(f"libspecloss ABI version {ver}, expected {ABI_VERSION}")
    return lib


def check(lib: ctypes.CDLL, rc: int) -> None:
    if rc != 0:
        msg = lib.spl_last_error()
        text = msg.decode() if msg else "unknown error"
        if rc == -1:
            # -1 = SPL_E_INVALID: the reference raises RuntimeError from torch.stft for the same inputs
            raise RuntimeError(f"specloss: {text}")
        raise SpecLossError(f"specloss (code {rc}): {text}")


def build_library(verbose: bool = False) -> str:
    """Compile csrc/specloss.cu into libspecloss.so for sm_100a (cross-compiles without a GPU)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise SpecLossError("nvcc not found; libspecloss.so cannot be built")
    srcs = [os.path.join(CSRC, f) for f in ("specloss.cu", "specloss_kernels.cuh", "specloss_host.inl",
                                            "fft_codelets.cuh", "melgemm.cuh", "melpower.cuh", "transform_eo.cuh")]
    hdr = os.path.join(os.path.dirname(_HERE), "include", "specloss.h")
    newest = max(os.path.getmtime(p) for p in srcs + [hdr])
    if os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, srcs[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SpecLossError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_LIB = None


def load_library() -> ctypes.CDLL:
    """Load the CUDA library.  Raises if it has not been built -- there is no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SpecLossError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
        _LIB = bind(ctypes.CDLL(LIB_PATH))
    return _LIB
