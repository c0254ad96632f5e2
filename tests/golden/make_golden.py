"""Generate the committed golden vectors by running the code the reference actually executes.

The reference (k0r1g/audio-transformers) reaches the hot path through two third-party calls:
``WhisperProcessor(...)(audio, sampling_rate=16000, return_tensors="pt").input_features``
(REF:whisper_finetune/dataset.py:58-62, inference.py:154,200) and
``torchaudio.transforms.MelSpectrogram(22050, n_fft=1024, hop_length=512, n_mels=64)`` followed
by ``torch.log(mel + 1e-9)`` (REF:urban_sounds/dataset.py:19-24,55-56).  This script imports the
installed ``transformers`` / ``torchaudio`` (the only runnable oracle), runs
them on the seeded synthetic waveforms of ``audio_transformers_b200.signals`` and stores
inputs-by-seed + outputs.  Run from the repo root:

    python tests/golden/make_golden.py

Outputs: tests/golden/whisper_golden.npz, tests/golden/urban_golden.npz, tests/golden/tables.npz
Versions used are recorded inside each file.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torchaudio  # noqa: E402
import transformers  # noqa: E402
from transformers import WhisperFeatureExtractor  # noqa: E402

from audio_transformers_b200 import signals  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, kind, index, length in samples)  -- length > 480000 exercises truncation
WHISPER_CASES = [
    ("noise_full", "noise", 0, 480000),
    ("tone_noise_full", "tone_noise", 1, 480000),
    ("chirp_full", "chirp", 2, 480000),
    ("am_noise_full", "am_noise", 3, 480000),
    ("zeros_full", "zeros", 4, 480000),
    ("click_full", "click", 5, 480000),
    ("tone1k_full", "tone1k", 6, 480000),
    ("ref_dummy_12s", "tone_noise", 7, 192000),      # REF:whisper_finetune/inference.py:246-255
    ("tone1k_1s", "tone1k", 8, 16000),
    ("noise_5s", "noise", 9, 80000),                 # one 5 s emotion segment, inference.py:187-200
] + [(f"noise_len{L}", "noise", 20 + i, L) for i, L in enumerate(signals.EDGE_LENGTHS)]


def frame_subset(length: int) -> np.ndarray:
    e = int(np.clip(min(length, 480000) // 160, 40, 2960))
    idx = np.concatenate([np.arange(0, 64), np.arange(e - 40, e + 40), np.arange(2984, 3000)])
    return np.unique(idx)


def main() -> None:
    torch.manual_seed(0)
    versions = dict(transformers=transformers.__version__, torch=torch.__version__,
                    torchaudio=torchaudio.__version__, numpy=np.__version__)
    print("versions:", versions)

    fe = WhisperFeatureExtractor()
    # ---- constant tables straight from the libraries ---------------------------------
    mel_tf = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
    np.savez_compressed(
        os.path.join(OUT, "tables.npz"),
        whisper_mel_filters=np.asarray(fe.mel_filters, dtype=np.float64),          # (201, 80) f64
        whisper_window=torch.hann_window(400).numpy(),                               # f32
        urban_fb=mel_tf.mel_scale.fb.numpy(),                                        # (513, 64) f32
        urban_window=mel_tf.spectrogram.window.numpy(),                              # f32
        versions=np.array(repr(versions)),
    )

    # ---- Whisper -------------------------------------------------------------------------
    store = {"versions": np.array(repr(versions)), "names": np.array([c[0] for c in WHISPER_CASES])}
    for name, kind, index, length in WHISPER_CASES:
        wav = signals.whisper_clip(index, seed=0, n_samples=length, kind=kind)
        feats = fe(wav.astype(np.float64), sampling_rate=16000, return_tensors="pt").input_features
        feats = feats.squeeze(0).numpy()                                              # (80, 3000)
        assert feats.shape == (80, 3000) and feats.dtype == np.float32
        idx = frame_subset(length)
        store[f"{name}/meta"] = np.array([index, length], dtype=np.int64)
        store[f"{name}/kind"] = np.array(kind)
        store[f"{name}/frames"] = idx.astype(np.int32)
        store[f"{name}/values"] = feats[:, idx].copy()
        store[f"{name}/stats"] = np.array([feats.astype(np.float64).sum(), feats.max(), feats.min()], dtype=np.float64)
        print(f"whisper {name:18s} max={feats.max():+.6f} min={feats.min():+.6f}")
    # a batched call (list of ragged arrays), which is how a collated batch reaches the extractor
    ragged = [signals.whisper_clip(40 + i, seed=0, n_samples=L) for i, L in enumerate((48000, 160000, 480000))]
    fb = fe(ragged, sampling_rate=16000, return_tensors="pt").input_features.numpy()
    store["ragged3/lengths"] = np.array([48000, 160000, 480000], dtype=np.int64)
    store["ragged3/frames"] = np.arange(0, 3000, 25, dtype=np.int32)
    store["ragged3/values"] = fb[:, :, ::25].copy()
    np.savez_compressed(os.path.join(OUT, "whisper_golden.npz"), **store)

    # ---- Urban ---------------------------------------------------------------------------
    wave = torch.from_numpy(signals.urban_batch(4, seed=0))
    with torch.no_grad():
        mel = mel_tf(wave)
        logmel = torch.log(mel + 1e-9)
    zeros = torch.zeros(1, 1, signals.URBAN_SAMPLES)
    with torch.no_grad():
        z = torch.log(mel_tf(zeros) + 1e-9)
    np.savez_compressed(
        os.path.join(OUT, "urban_golden.npz"),
        versions=np.array(repr(versions)),
        batch=np.array(4), seed=np.array(0),
        mel=mel.numpy(), logmel=logmel.numpy(), zeros_logmel=z.numpy(),
    )
    print("urban", tuple(logmel.shape), float(logmel.max()), float(logmel.min()), float(z.min()))

    # ---- Urban pre-steps: REF:urban_sounds/dataset.py:26-58 process_audio, restated with the same torch /
    # torchaudio calls (the class itself cannot be built offline: its __init__ downloads the dataset) ----
    import torchaudio.transforms as T

    def process_audio(audio_array, orig_sr, sr=22050, target_length=88200):
        waveform = torch.from_numpy(audio_array).float()
        waveform = waveform.mean(dim=0, keepdim=True) if len(waveform.shape) > 1 else waveform.unsqueeze(0)
        if orig_sr != sr:
            waveform = T.Resample(orig_sr, sr)(waveform)
        if waveform.shape[1] < target_length:
            waveform = torch.nn.functional.pad(waveform, (0, target_length - waveform.shape[1]))
        else:
            waveform = waveform[:, :target_length]
        if torch.max(torch.abs(waveform)) > 0:
            waveform = waveform / torch.max(torch.abs(waveform))
        return waveform, torch.log(mel_tf(waveform) + 1e-9)

    prep = {"versions": np.array(repr(versions)), "names": np.array([c[0] for c in signals.URBAN_PREP_CASES])}
    for name, rate, channels, n_in in signals.URBAN_PREP_CASES:
        audio = signals.urban_raw_clip(name, rate, channels, n_in)
        with torch.no_grad():
            wav, lm = process_audio(audio, rate)
        prep[f"{name}/meta"] = np.array([rate, channels, n_in], dtype=np.int64)
        prep[f"{name}/wave_sub"] = wav.numpy()[0, ::37].copy()
        prep[f"{name}/wave_stats"] = np.array([wav.double().sum().item(), wav.abs().max().item()], dtype=np.float64)
        prep[f"{name}/logmel_sub"] = lm.numpy()[0, :, ::9].copy()
        print(f"urban prep {name:16s} rate={rate} ch={channels} n={n_in} wave sum={prep[f'{name}/wave_stats'][0]:+.4f}")
    np.savez_compressed(os.path.join(OUT, "urban_prep_golden.npz"), **prep)


if __name__ == "__main__":
    main()
