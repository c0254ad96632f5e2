"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/b200mel.h declares, serves its constant tables without a device, and refuses to compute
without one (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from audio_transformers_b200 import build, _lib
    build.build()                      # no-op when up to date; nvcc cross-compiles without a GPU
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200mel.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mel_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(lib):
    from audio_transformers_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 11
    assert sorted(_lib.SYMBOLS) == declared          # the ctypes table binds exactly what the header declares
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in b200mel.h but not exported by libb200mel.so"
    assert lib.b200mel_version() == 100


def test_tables_match_reference_libraries(lib, golden_dir):
    from audio_transformers_b200 import _lib
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    assert np.array_equal(_lib.get_table(_lib.PRESET_WHISPER, _lib.TABLE_WINDOW), g["whisper_window"])
    assert np.array_equal(_lib.get_table(_lib.PRESET_WHISPER, _lib.TABLE_FILTERBANK), g["whisper_mel_filters"].astype(np.float32))
    assert np.array_equal(_lib.get_table(_lib.PRESET_URBAN, _lib.TABLE_WINDOW), g["urban_window"])
    assert np.array_equal(_lib.get_table(_lib.PRESET_URBAN, _lib.TABLE_FILTERBANK), g["urban_fb"])
    # REF:whisper_finetune/experiments.ipynb:563-569, straight from the library's own table
    fb = _lib.get_table(_lib.PRESET_WHISPER, _lib.TABLE_FILTERBANK)
    assert abs(float(fb[1, 0]) - 0.02486259) < 5e-9 and abs(float(fb[199, 79]) - 0.00044876) < 5e-9


def test_error_reporting(lib):
    buf = (ctypes.c_float * 4)()
    n = lib.b200mel_get_table(0, 0, buf, 4)            # capacity too small
    assert n == -1 and b"capacity" in lib.b200mel_last_error()
    assert lib.b200mel_get_table(7, 0, buf, 4) == -1 and b"preset" in lib.b200mel_last_error()
    out = ctypes.c_void_p()
    assert lib.b200mel_create(0, 9, ctypes.byref(out)) == -1
    assert lib.b200mel_workspace_bytes(None, 64) == 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = ctypes.c_void_p()
    st = lib.b200mel_create(0, 0, ctypes.byref(out))
    assert st in (-3, -4) and not out.value             # CUDA error / unsupported arch, never a CPU handle
    from audio_transformers_b200 import B200MelSpectrogram, B200WhisperFeatureExtractor, ops
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.whisper_logmel(torch.zeros(1, 480000), None)
    with pytest.raises(RuntimeError):
        B200WhisperFeatureExtractor()(np.zeros(16000, np.float32), sampling_rate=16000, return_tensors="pt")
    with pytest.raises(RuntimeError):
        B200MelSpectrogram()(torch.zeros(1, 88200))


def test_meta_shapes():
    import torch
    from audio_transformers_b200 import ops  # noqa: F401  (registers the ops)
    w = torch.empty(5, 480000, device="meta")
    assert torch.ops.b200mel.whisper_logmel(w, None).shape == (5, 80, 3000)
    assert torch.ops.b200mel.mel_power(torch.empty(3, 88200, device="meta"), 1e-9).shape == (3, 64, 173)
    assert torch.ops.b200mel.whisper_frame_mask(torch.empty(4, dtype=torch.int32, device="meta")).shape == (4, 3000)


def test_host_pack_matches_numpy_cast(lib):
    """b200mel_host_pack: ragged float64 / float32 clips -> one float32 staging buffer, truncated at max_samples,
    rows beyond a clip's length untouched; the cast is numpy's (HF:models/whisper/feature_extraction_whisper.py:285-286)."""
    rng = np.random.default_rng(0)
    clips = [rng.standard_normal(n) for n in (480000, 5, 0, 123457, 600000)]
    for dt in (np.float64, np.float32):
        arrs = [np.ascontiguousarray(c.astype(dt)) for c in clips]
        n, stride = len(arrs), 480000
        ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.int64)
        for threads in (1, 5):
            dst = np.full((n, stride), 7.0, dtype=np.float32)
            out = np.zeros(n, dtype=np.int32)
            st = lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), n, int(dt == np.float64), 480000,
                                       dst.ctypes.data_as(ctypes.c_void_p), stride, out.ctypes.data_as(ctypes.c_void_p), threads)
            assert st == 0
            for i, a in enumerate(arrs):
                L = min(len(a), 480000)
                assert out[i] == L
                assert np.array_equal(dst[i, :L], a[:L].astype(np.float32)) and (dst[i, L:] == 7.0).all()
    # a clip longer than the row is an error, not an overrun
    bad = lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), n, 0, 480000, dst.ctypes.data_as(ctypes.c_void_p),
                                1000, None, 2)
    assert bad == -1 and b"longer than dst_stride" in lib.b200mel_last_error()


def _pack_once(lib, clips, threads):
    n = len(clips)
    ptrs = (ctypes.c_void_p * n)(*[c.ctypes.data for c in clips])
    lens = np.array([len(c) for c in clips], dtype=np.int64)
    dst = np.full((n, 480000), 7.0, dtype=np.float32)
    out = np.zeros(n, dtype=np.int32)
    st = lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), n, 1, 480000, dst.ctypes.data_as(ctypes.c_void_p),
                               480000, out.ctypes.data_as(ctypes.c_void_p), threads)
    assert st == 0
    for i, c in enumerate(clips):
        L = min(len(c), 480000)
        assert out[i] == L and np.array_equal(dst[i, :L], c[:L].astype(np.float32)) and (dst[i, L:] == 7.0).all()


def test_host_pack_worker_pool(lib):
    """The pack runs on a persistent worker pool: many calls with changing thread counts, concurrent callers from
    several Python threads (ctypes drops the GIL), and a forked child (which must not wait on the parent's workers)."""
    import threading
    rng = np.random.default_rng(1)
    clips = [rng.standard_normal(n) for n in (1, 159, 70000, 480000, 600000, 33333)]
    for k in range(40):
        _pack_once(lib, clips, 1 + (7 * k) % 16)
    errors = []

    def worker(seed):
        try:
            r = np.random.default_rng(seed)
            for _ in range(20):
                _pack_once(lib, clips, int(r.integers(1, 17)))
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors
    pid = os.fork()
    if pid == 0:                                  # child: the pool of the parent has no threads here
        code = 1
        try:
            _pack_once(lib, clips, 8)
            code = 0
        finally:
            os._exit(code)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
