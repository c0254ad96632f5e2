"""Pin the CPU oracle (oracle/logmel_oracle.py) before anything is compared against it.

Sources of truth, in order:
  1. the reference's own notebook: extractor config + six filter coefficients
     (REF:whisper_finetune/experiments.ipynb:558-573), urban geometry
     (REF:urban_sounds/experiments.ipynb:30-32 + REF:urban_sounds/dataset.py:9);
  2. committed golden vectors produced by the installed HF / torchaudio implementations, i.e.
     the code the reference executes (tests/golden/make_golden.py);
  3. the live libraries, when importable (same image on the GPU box).
"""
import os

import numpy as np
import pytest

from audio_transformers_b200 import signals
from oracle import logmel_oracle as O

TOL = 1e-4   # max-abs on normalised log-mel / log(mel+1e-9)


@pytest.fixture(scope="module")
def tables(golden_dir):
    return np.load(os.path.join(golden_dir, "tables.npz"))


@pytest.fixture(scope="module")
def wgold(golden_dir):
    return np.load(os.path.join(golden_dir, "whisper_golden.npz"))


@pytest.fixture(scope="module")
def ugold(golden_dir):
    return np.load(os.path.join(golden_dir, "urban_golden.npz"))


def test_reference_notebook_config():
    # REF:whisper_finetune/experiments.ipynb:558-573,576-578
    p = O.WHISPER
    assert (p["chunk_length"], p["n_mels"], p["hop_length"], p["n_fft"]) == (30, 80, 160, 400)
    assert (p["n_samples"], p["nb_max_frames"], p["sampling_rate"]) == (480000, 3000, 16000)
    # REF:urban_sounds/dataset.py:9 -> 88 200 samples -> 1 + 88200//512 = 173 frames
    assert 1 + O.URBAN["n_samples"] // O.URBAN["hop_length"] == O.URBAN["n_frames"] == 173


def test_reference_notebook_filter_coefficients():
    # REF:whisper_finetune/experiments.ipynb:563-569 (printed float32 values of mel_filters)
    fb = O.whisper_mel_filters().astype(np.float32)
    known = {(1, 0): 0.02486259, (1, 1): 0.00199082, (2, 1): 0.02287177, (2, 2): 0.00398164,
             (198, 79): 0.00089752, (199, 79): 0.00044876}
    for (r, c), v in known.items():
        assert abs(float(fb[r, c]) - v) < 5e-9, (r, c, fb[r, c], v)
    assert not fb[0].any() and not fb[200].any()


def test_whisper_filterbank_matches_library(tables):
    fb = O.whisper_mel_filters()
    ref = tables["whisper_mel_filters"]
    assert fb.shape == ref.shape == (201, 80)
    assert np.array_equal(fb != 0, ref != 0)
    assert int((fb != 0).sum()) == 391
    assert np.abs(fb - ref).max() < 1e-15
    assert np.array_equal(fb.astype(np.float32), ref.astype(np.float32))


def test_urban_filterbank_matches_library(tables):
    fb = O.urban_mel_filters(dtype=np.float32)
    ref = tables["urban_fb"]
    assert fb.shape == ref.shape == (513, 64)
    assert int((ref != 0).sum()) == 998
    assert not ref[0].any() and not ref[512].any()
    # torchaudio builds the table with FP32 torch ops; the restatement agrees to FP32 round-off
    assert np.abs(fb.astype(np.float64) - ref.astype(np.float64)).max() < 2e-5
    fb64 = O.urban_mel_filters(dtype=np.float64)
    assert np.abs(fb64 - ref.astype(np.float64)).max() < 5e-5


def test_windows_match_library(tables):
    for n, key in ((400, "whisper_window"), (1024, "urban_window")):
        ref = tables[key]
        assert np.abs(O.hann_periodic(n, np.float32) - ref).max() < 3e-7
        assert np.abs(O.hann_periodic(n, np.float64) - ref).max() < 3e-7


def _whisper_case(wgold, name):
    index, length = (int(v) for v in wgold[f"{name}/meta"])
    kind = str(wgold[f"{name}/kind"])
    wav = signals.whisper_clip(index, seed=0, n_samples=length, kind=kind)
    return wav, wgold[f"{name}/frames"], wgold[f"{name}/values"], wgold[f"{name}/stats"]


def test_whisper_oracle_vs_golden(wgold):
    names = [str(n) for n in wgold["names"]]
    assert len(names) >= 20
    worst = {}
    for name in names:
        wav, frames, values, stats = _whisper_case(wgold, name)
        for dt in (np.float32, np.float64):
            out = O.whisper_logmel(wav, dtype=dt)
            assert out.shape == (1, 80, 3000) and out.dtype == np.float32
            err = float(np.abs(out[0][:, frames] - values).max())
            worst[(name, dt.__name__)] = err
            assert err <= TOL, (name, dt.__name__, err)
            assert abs(float(out.max()) - stats[1]) <= TOL and abs(float(out.min()) - stats[2]) <= TOL
    # broadband inputs agree far better than the contract (SURVEY.md section 8c self-consistency)
    assert worst[("noise_full", "float32")] < 1e-5
    assert worst[("zeros_full", "float32")] < 1e-6


def test_whisper_oracle_batched_ragged(wgold):
    lengths = [int(v) for v in wgold["ragged3/lengths"]]
    clips = [signals.whisper_clip(40 + i, seed=0, n_samples=L) for i, L in enumerate(lengths)]
    out = O.whisper_logmel(clips, dtype=np.float32)
    assert out.shape == (3, 80, 3000)
    assert np.abs(out[:, :, ::25] - wgold["ragged3/values"]).max() <= TOL


def test_whisper_floor_and_truncation():
    z = O.whisper_logmel(np.zeros(1000, dtype=np.float32))
    assert np.abs(z + 1.5).max() < 1e-6                    # (-10+4)/4
    long = signals.whisper_clip(3, n_samples=600000, kind="noise")
    assert np.array_equal(O.whisper_logmel(long), O.whisper_logmel(long[:480000]))


def test_attention_mask_rescale():
    m = O.whisper_attention_mask([1, 160, 161, 480000, 600000])
    assert m.shape == (5, 3000) and m.dtype == np.int32
    assert m.sum(axis=1).tolist() == [1, 1, 2, 3000, 3000]


def test_attention_mask_vs_live_hf():
    """HF:models/whisper/feature_extraction_whisper.py:328-337 with return_attention_mask=True, element by element."""
    tr = pytest.importorskip("transformers")
    lens = [1, 159, 160, 161, 319, 320, 321, 4000, 479999, 480000, 480001, 600000]
    clips = [np.full(n, 0.01, dtype=np.float32) for n in lens]
    ref = tr.WhisperFeatureExtractor()(clips, sampling_rate=16000, return_tensors="np", return_attention_mask=True)
    assert np.array_equal(O.whisper_attention_mask(lens), ref["attention_mask"])


def test_urban_oracle_vs_golden(ugold):
    wave = signals.urban_batch(int(ugold["batch"]), seed=int(ugold["seed"]))
    for dt in (np.float32, np.float64):
        out = O.urban_melspec(wave, log_eps=1e-9, dtype=dt)
        assert out.shape == (4, 1, 64, 173)
        assert np.abs(out - ugold["logmel"]).max() <= TOL
    lin = O.urban_melspec(wave, log_eps=None)
    assert np.abs(lin - ugold["mel"]).max() <= 1e-4 * np.abs(ugold["mel"]).max()
    z = O.urban_melspec(np.zeros((1, 1, 88200), np.float32))
    assert np.abs(z - ugold["zeros_logmel"]).max() < 1e-5
    assert abs(float(z.min()) - (-20.7233)) < 1e-3


def _strong(logmel_ref):
    """Bins within 40 dB of the clip's loudest mel value (and well above the 1e-9 epsilon).  Resampled clips have
    an almost empty band above the source's Nyquist (e.g. above 4 kHz for an 8 kHz clip): the mel energy there is
    FP32 round-off of the FFT, where no two FP32 implementations agree in the log.  Those bins are checked on the
    linear mel instead."""
    return (logmel_ref > logmel_ref.max() - 9.2) & (logmel_ref > np.log(1e-6))


def test_urban_prep_oracle_vs_golden(golden_dir):
    """REF:urban_sounds/dataset.py:26-58 process_audio: mono mean, T.Resample, pad/trim, peak normalisation, mel,
    log -- oracle vs outputs of the same torch/torchaudio calls (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "urban_prep_golden.npz"))
    for name in [str(n) for n in g["names"]]:
        rate, channels, n_in = (int(v) for v in g[f"{name}/meta"])
        audio = signals.urban_raw_clip(name, rate, channels, n_in)
        w = O.urban_preprocess(audio, orig_sr=rate)
        assert w.shape == (1, 88200) and w.dtype == np.float32
        assert np.abs(w[0, ::37] - g[f"{name}/wave_sub"]).max() <= 2e-6, name
        assert abs(float(w.astype(np.float64).sum()) - g[f"{name}/wave_stats"][0]) <= 2e-2, name
        assert abs(float(np.abs(w).max()) - g[f"{name}/wave_stats"][1]) <= 1e-6, name
        lm = O.urban_melspec(w)[0][:, ::9]
        ref = g[f"{name}/logmel_sub"]
        ok = _strong(ref)
        assert np.abs(lm[ok] - ref[ok]).max(initial=0.0) <= TOL, name
        assert np.abs(np.exp(lm) - np.exp(ref)).max() <= 1e-5 * np.exp(ref).max() + 1e-9, name


def test_sinc_resample_taps_match_torchaudio():
    ta = pytest.importorskip("torchaudio")
    import math
    from torchaudio.functional.functional import _get_sinc_resample_kernel
    from audio_transformers_b200.urban import sinc_resample_kernel as shim_kernel
    for rate in (44100, 48000, 16000, 8000, 96000, 11025, 32000):
        gcd = math.gcd(rate, 22050)
        ref, width = _get_sinc_resample_kernel(rate, 22050, gcd)
        k, w, orig, new = O.sinc_resample_kernel(rate, 22050)
        assert (w, orig, new) == (width, rate // gcd, 22050 // gcd)
        assert k.shape == tuple(ref.shape[::2]) and np.abs(k - ref.numpy()[:, 0]).max() <= 1e-7
        k2, w2, o2, n2 = shim_kernel(rate, 22050)
        assert (w2, o2, n2) == (w, orig, new) and np.array_equal(k2, k)


def test_oracle_vs_live_libraries():
    """Same image on the GPU box: compare against the live HF / torchaudio implementations."""
    tr = pytest.importorskip("transformers")
    ta = pytest.importorskip("torchaudio")
    import torch
    fe = tr.WhisperFeatureExtractor()
    clips = [signals.whisper_clip(i, seed=3, n_samples=n) for i, n in enumerate((480000, 31337, 250000, 480000))]
    ref = fe([c.astype(np.float64) for c in clips], sampling_rate=16000, return_tensors="pt").input_features.numpy()
    out = O.whisper_logmel(clips)
    assert np.abs(out - ref).max() <= TOL
    wave = signals.urban_batch(3, seed=5)
    tf = ta.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
    ref = torch.log(tf(torch.from_numpy(wave)) + 1e-9).numpy()
    assert np.abs(O.urban_melspec(wave) - ref).max() <= TOL
