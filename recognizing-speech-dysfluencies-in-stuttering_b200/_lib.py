"""ctypes binding of libdysb200.so (C ABI in include/dysfluency_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no CUDA device is
visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdysb200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE = 0, 1, 2, 3
FEATURE_LEN = 149
AUDIO_FEATURE_LEN = 144
SAMPLE_RATE = 16000
CMVN_ACC_LEN = 299
CMVN_PARTIALS = 128 * 298
STATUS_SHORT, STATUS_NONFINITE, STATUS_CLEAN_FALLBACK, STATUS_BAD_LENGTH = 1, 2, 4, 8

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

_SIGNATURES = {
    "dys_version": (C.c_int, []),
    "dys_last_error": (C.c_char_p, []),
    "dys_init": (C.c_int, []),
    "dys_set_overlap": (C.c_int, [_i32]),
    "dys_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "dys_workspace_min_bytes": (_i64, [_i32, _i32, _i32]),
    "dys_features_raw": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "dys_features_raw_clean": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dys_features_raw_pcm16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "dys_features_raw_clean_pcm16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dys_resampled_length": (_i64, [_i64, _i32]),
    "dys_resample_to_16k": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "dys_resample_table": (_i64, [_i32, _vp, _i64, _vp]),
    "dys_cmvn_accumulate": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "dys_cmvn_finalize": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "dys_cmvn_apply": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "dys_qc_workspace_bytes": (_i64, [_i32, _i32]),
    "dys_qc_metrics": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _vp]),
    "dys_get_table": (_i64, [_i32, _i32, _vp, _i64]),
    "dys_debug_feature_stages": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dys_debug_denoise": (C.c_int, [_vp, _i32, _f32, _vp, _vp, _vp]),
    "dys_kernel_count": (C.c_int, []),
    "dys_kernel_name": (C.c_char_p, [_i32]),
    "dys_profile_enable": (C.c_int, [_i32]),
    "dys_profile_read": (C.c_int, [_vp, _vp, _i32]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class DysError(RuntimeError):
    pass


def load():
    """Loads the shared library (once). Raises DysError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DysError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                       "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != OK:
        msg = load().dys_last_error().decode("utf-8", "replace")
        raise DysError(f"{what} failed (code {rc}): {msg}")


def kernel_names():
    lib = load()
    return [lib.dys_kernel_name(i).decode() for i in range(lib.dys_kernel_count())]


def profile_enable(on: bool):
    check(load().dys_profile_enable(1 if on else 0), "dys_profile_enable")


def profile_read(reset: bool = True):
    """-> {kernel name: (total ms while profiling was on, launches)} since the last reset."""
    lib = load()
    n = lib.dys_kernel_count()
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib.dys_profile_read(C.cast(ms, C.c_void_p), C.cast(cnt, C.c_void_p), 1 if reset else 0), "dys_profile_read")
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(kernel_names())}
