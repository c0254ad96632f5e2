"""ctypes loader for libb200mel.so (C ABI declared in include/b200mel.h).

There is no CPU path and no fallback: if the shared library is missing or a call fails, this
module raises.  Build the library with ``python -m audio_transformers_b200.build`` (or
``__graft_entry__.build()``); it is compiled in-tree so it travels with the repo snapshot.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MEL_LIB") or os.path.join(_HERE, "libb200mel.so")   # env override: kernel experiments only

PRESET_WHISPER = 0
PRESET_URBAN = 1
TABLE_WINDOW = 0
TABLE_FILTERBANK = 1

# every symbol include/b200mel.h declares: name -> (restype, argtypes)
_c = ctypes
SYMBOLS = {
    "b200mel_version": (_c.c_int, []),
    "b200mel_last_error": (_c.c_char_p, []),
    "b200mel_create": (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_c.c_void_p)]),
    "b200mel_destroy": (_c.c_int, [_c.c_void_p]),
    "b200mel_workspace_bytes": (_c.c_size_t, [_c.c_void_p, _c.c_int32]),
    "b200mel_whisper_logmel_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int32,
                                              _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "b200mel_whisper_frame_mask": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p]),
    "b200mel_mel_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_float,
                                   _c.c_void_p, _c.c_void_p]),
    "b200mel_urban_prep_workspace_bytes": (_c.c_size_t, [_c.c_void_p, _c.c_int32]),
    "b200mel_urban_prep_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int32, _c.c_int32,
                                          _c.c_int32, _c.c_int32, _c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_int64,
                                          _c.c_int32, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "b200mel_host_pack": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int64, _c.c_void_p,
                                     _c.c_int64, _c.c_void_p, _c.c_int32]),
    "b200mel_whisper_logmel_host": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_void_p,
                                               _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_size_t, _c.c_int32, _c.c_void_p]),
    "b200mel_encoder_stem_workspace_bytes": (_c.c_size_t, [_c.c_void_p, _c.c_int32]),
    "b200mel_encoder_stem_bf16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                             _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "b200mel_profile_begin": (_c.c_int, [_c.c_void_p, _c.c_int32]),
    "b200mel_profile_end": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_double), _c.POINTER(_c.c_int32)]),
    "b200mel_get_table": (_c.c_int64, [_c.c_int, _c.c_int, _c.c_void_p, _c.c_int64]),
}

_lib = None
_lock = threading.Lock()


class B200MelError(RuntimeError):
    """A libb200mel.so call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise B200MelError(
                    f"{LIB_PATH} is missing: the sm_100a library has not been built and there is no CPU "
                    "fallback.  Run `python -m audio_transformers_b200.build` (needs nvcc).")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SYMBOLS.items():
                fn = getattr(lib, name)          # AttributeError if the .so does not export it
                fn.restype, fn.argtypes = restype, argtypes
            _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().b200mel_last_error().decode("utf-8", "replace")
        raise B200MelError(f"{what} failed with status {status}: {msg}")


def get_table(preset: int, table: int):
    """Host copy of one constant table as a numpy array (no device needed)."""
    import numpy as np
    n_fft, n_mel = (400, 80) if preset == PRESET_WHISPER else (1024, 64)
    shape = (n_fft,) if table == TABLE_WINDOW else (n_fft // 2 + 1, n_mel)
    buf = np.empty(shape, dtype=np.float32)
    n = load().b200mel_get_table(preset, table, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    if n != buf.size:
        check(int(n) if n < 0 else -1, "b200mel_get_table")
    return buf
