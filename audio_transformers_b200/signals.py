"""Seeded synthetic waveforms shared by the tests, the golden-vector generator and bench.py.

The classes follow SURVEY.md section 8(d) "Config 2": they are deterministic functions of
``(seed, index)`` so that every process (CPU container, GPU box, any rank) regenerates the
same bytes without shipping audio around.
"""
from __future__ import annotations

import numpy as np

WHISPER_SR = 16000
WHISPER_SAMPLES = 480000
URBAN_SR = 22050
URBAN_SAMPLES = 88200

CLASSES = ("noise", "tone_noise", "chirp", "am_noise")


def whisper_clip(index: int, seed: int = 0, n_samples: int = WHISPER_SAMPLES, sr: int = WHISPER_SR,
                 kind: str | None = None) -> np.ndarray:
    """One float32 clip.  ``kind`` defaults to ``CLASSES[index % 4]`` (equal shares by index).

    * ``noise``      0.1 N(0,1)
    * ``tone_noise`` the reference's own dummy recipe, 0.5 sin(2 pi 440 t) + 0.01 N(0,1)
                     (REF:whisper_finetune/inference.py:246-255)
    * ``chirp``      linear chirp 50 -> 7900 Hz over the clip, amplitude 0.5
    * ``am_noise``   0.3 N(0,1) (0.5 + 0.5 sin(2 pi 3 t))^4
    plus, for edge-case tests: ``zeros``, ``click`` (unit impulse at n/3), ``tone1k``.
    """
    kind = kind or CLASSES[index % len(CLASSES)]
    rng = np.random.default_rng([seed, index])
    t = np.arange(n_samples, dtype=np.float64) / sr
    if kind == "noise":
        x = 0.1 * rng.standard_normal(n_samples)
    elif kind == "tone_noise":
        x = 0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.01 * rng.standard_normal(n_samples)
    elif kind == "chirp":
        dur = max(n_samples / sr, 1e-9)
        f0, f1 = 50.0, 7900.0
        x = 0.5 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t))
    elif kind == "am_noise":
        x = 0.3 * rng.standard_normal(n_samples) * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t)) ** 4
    elif kind == "zeros":
        x = np.zeros(n_samples)
    elif kind == "click":
        x = np.zeros(n_samples)
        if n_samples:
            x[n_samples // 3] = 1.0
    elif kind == "tone1k":
        x = 0.5 * np.sin(2 * np.pi * 1000.0 * t)
    else:
        raise ValueError(f"unknown signal kind {kind!r}")
    return x.astype(np.float32)


def whisper_batch(batch: int, seed: int = 0, start: int = 0, n_samples: int = WHISPER_SAMPLES) -> np.ndarray:
    """(batch, n_samples) float32, clip ``i`` = ``whisper_clip(start + i, seed)``."""
    out = np.empty((batch, n_samples), dtype=np.float32)
    for i in range(batch):
        out[i] = whisper_clip(start + i, seed, n_samples)
    return out


def ragged_lengths(batch: int, seed: int = 1, lo: int = 16000, hi: int = 480000) -> np.ndarray:
    """SURVEY.md section 8(d) "Config 4": L_i ~ U{lo..hi}."""
    return np.random.default_rng(seed).integers(lo, hi + 1, size=batch).astype(np.int64)


EDGE_LENGTHS = (1, 159, 160, 161, 399, 400, 401, 479999, 480000, 480001, 600000)


def urban_batch(batch: int = 32, seed: int = 0, n_samples: int = URBAN_SAMPLES) -> np.ndarray:
    """(batch, 1, n_samples) float32 white noise, each clip peak-normalised
    (mirrors REF:urban_sounds/dataset.py:51-52).  SURVEY.md section 8(d) "Config 1"."""
    rng = np.random.default_rng([seed, 7])
    x = rng.standard_normal((batch, 1, n_samples)).astype(np.float32)
    x /= np.abs(x).max(axis=-1, keepdims=True)
    return x


# Raw decoded clips for the urban pre-steps (REF:urban_sounds/dataset.py:64-68 hands `audio['array']` and
# `audio['sampling_rate']` to process_audio).  (name, sampling rate, channels, samples per channel):
# UrbanSound8K mixes rates and channel counts, clips are <= 4 s.
URBAN_PREP_CASES = (
    ("stereo_44k1_4s", 44100, 2, 176400),
    ("mono_44k1_short", 44100, 1, 61234),
    ("stereo_48k_4s", 48000, 2, 192000),
    ("mono_48k_long", 48000, 1, 250000),       # longer than 4 s after resampling: trimmed
    ("mono_22k05_full", 22050, 1, 88200),      # no resampling
    ("stereo_22k05_short", 22050, 2, 40000),
    ("mono_16k", 16000, 1, 50000),             # upsampling
    ("mono_8k_tiny", 8000, 1, 2000),
    ("stereo_96k", 96000, 2, 300000),
    ("mono_11k025", 11025, 1, 44100),
    ("silence_44k1", 44100, 2, 30000),         # all zeros: no normalisation
)


def urban_raw_clip(name: str, rate: int, channels: int, n_in: int) -> np.ndarray:
    """float64 array shaped like `datasets` decodes it: (n,) for mono, (channels, n) otherwise.  Band-limited
    tones plus noise at an arbitrary scale, so that resampling, the mono mean and the peak normalisation all
    have something to do."""
    seed = sum(ord(c) for c in name)
    rng = np.random.default_rng([seed, rate, channels])
    if name.startswith("silence"):
        x = np.zeros((channels, n_in), dtype=np.float64)
    else:
        t = np.arange(n_in, dtype=np.float64) / rate
        x = np.empty((channels, n_in), dtype=np.float64)
        for c in range(channels):
            f1, f2 = 220.0 * (c + 1), min(0.18 * rate, 3100.0 + 400.0 * c)
            x[c] = 0.31 * np.sin(2 * np.pi * f1 * t + c) + 0.12 * np.sin(2 * np.pi * f2 * t) + 0.05 * rng.standard_normal(n_in)
        x *= 0.37
    return x[0] if channels == 1 else x
