"""ctypes binding of librealise_b200.so (the C ABI declared in include/realise_b200.h).

There is no fallback: if the shared object cannot be built or loaded the import raises.
"""
import ctypes
import os

from . import build as _build

_LIB = None


class GemmDesc(ctypes.Structure):
    """Mirror of `rl_gemm_desc` (include/realise_b200.h)."""
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
        ("M", ctypes.c_int64), ("N", ctypes.c_int64), ("K", ctypes.c_int64),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64),
        ("a_mode", ctypes.c_int32),
        ("conv_C", ctypes.c_int32), ("conv_W", ctypes.c_int32), ("conv_H", ctypes.c_int32),
        ("conv_P", ctypes.c_int32), ("conv_NIMG", ctypes.c_int32),
        ("ntaps", ctypes.c_int32),
        ("tap_dw", ctypes.c_int8 * 12), ("tap_dh", ctypes.c_int8 * 12), ("tap_plane", ctypes.c_int8 * 12),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64), ("out_dtype", ctypes.c_int32),
        ("out2", ctypes.c_void_p), ("ldo2", ctypes.c_int64),
        ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("res", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("res_dtype", ctypes.c_int32),
        ("act", ctypes.c_int32), ("out_remap", ctypes.c_int32),
        ("remap_plane", ctypes.c_int32), ("conv_Cuse", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("a_major", ctypes.c_int32),
        ("drop_p", ctypes.c_float), ("drop_site", ctypes.c_uint32), ("drop_seed", ctypes.c_uint64),
        ("b_major", ctypes.c_int32),
        ("a_dtype", ctypes.c_int32), ("b_dtype", ctypes.c_int32),
        ("drop_counter", ctypes.c_void_p),
        ("tune_tile_n", ctypes.c_int32), ("tune_no_pair", ctypes.c_int32),
        ("colsum", ctypes.c_void_p), ("colsumsq", ctypes.c_void_p),
        ("sm_reserve", ctypes.c_int32),
        ("b_mode", ctypes.c_int32),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = _build.ensure_built()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: the CUDA extension is required (no CPU fallback)")
        _LIB = ctypes.CDLL(path)
        _LIB.rl_last_error.restype = ctypes.c_char_p
        _LIB.rl_version.restype = ctypes.c_int
        _LIB.rl_workspace_bytes.restype = ctypes.c_int64
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().rl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"librealise_b200 {what} failed (rc={rc}): {msg}")
