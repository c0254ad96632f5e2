"""Waveform collate: the batched replacement for the reference's per-sample feature extraction.

The reference extracts features one clip at a time inside ``Dataset.__getitem__``
(REF:whisper_finetune/dataset.py:53-82) and then ``collate_fn`` allocates a CPU ``(B, 80, 3000)``
tensor and copies them in (REF:whisper_finetune/dataset.py:84-110), after which the training loop
moves the features to the device (REF:whisper_finetune/train.py:188).  With a GPU front end that
order is backwards: it would call the kernel batch-1 from a serial loop and drag the result back to
the host.  Here ``__getitem__`` keeps the raw waveform, the collate stacks the ragged waveforms into
ONE pinned host buffer (+ lengths), and :func:`features_on_device` issues one H2D copy and one
kernel launch for the whole batch.  Label handling mirrors the reference's collate exactly.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

N_SAMPLES = 480000


def stack_waveforms(clips: Sequence[Any], max_samples: int = N_SAMPLES, pin: Optional[bool] = None):
    """Ragged clips -> ((B, T4) float32 host tensor, (B,) int32 lengths).  T4 = longest clip (capped at
    ``max_samples``) rounded up to a multiple of 4; the tail of shorter rows is zero.  Pinned when a
    CUDA device is present (or ``pin=True``)."""
    arrs = [np.asarray(c.detach().cpu() if isinstance(c, torch.Tensor) else c, dtype=np.float32).reshape(-1) for c in clips]
    lens = np.fromiter((min(a.shape[0], max_samples) for a in arrs), dtype=np.int32, count=len(arrs))
    width = (max(int(lens.max()) if len(arrs) else 0, 4) + 3) // 4 * 4
    pin = torch.cuda.is_available() if pin is None else pin
    wave = torch.zeros((len(arrs), width), dtype=torch.float32, pin_memory=bool(pin))
    wnp = wave.numpy()
    for i, a in enumerate(arrs):
        wnp[i, :lens[i]] = a[:lens[i]]
    return wave, torch.from_numpy(lens)


class WaveformCollator:
    """``collate_fn`` for samples shaped like the reference's dataset items, but carrying
    ``"waveform"`` instead of ``"input_features"``.

    Output keys: ``waveform`` (B, T4) pinned float32, ``lengths`` (B,) int32, ``labels`` (B, Lmax)
    long padded with ``pad_token_id`` and ``emotion_labels`` (B,) long -- the last two exactly as
    REF:whisper_finetune/dataset.py:85-110 builds them.
    """

    def __init__(self, pad_token_id: int = 50257, max_samples: int = N_SAMPLES):
        self.pad_token_id = pad_token_id
        self.max_samples = max_samples

    def __call__(self, batch: List[Dict[str, Any]]) -> Dict[str, torch.Tensor]:
        wave, lengths = stack_waveforms([b["waveform"] for b in batch], self.max_samples)
        out: Dict[str, torch.Tensor] = {"waveform": wave, "lengths": lengths}
        if "labels" in batch[0]:
            max_len = max(int(b["labels"].shape[0]) for b in batch)
            labels = torch.ones(len(batch), max_len, dtype=torch.long) * self.pad_token_id
            for i, b in enumerate(batch):
                labels[i, : b["labels"].shape[0]] = b["labels"]
            out["labels"] = labels
        if "emotion_label" in batch[0]:
            out["emotion_labels"] = torch.stack([torch.as_tensor(b["emotion_label"], dtype=torch.long) for b in batch])
        return out


def features_on_device(batch: Dict[str, torch.Tensor], extractor, device=None) -> Dict[str, torch.Tensor]:
    """One H2D + one fused launch for the whole collated batch; adds ``input_features`` (B, 80, 3000) on
    the GPU, so that ``batch["input_features"].to(device)`` in the training loop is a no-op."""
    feats = extractor(batch["waveform"], sampling_rate=16000, return_tensors="pt", device=device,
                      lengths=batch["lengths"]).input_features
    out = dict(batch)
    out["input_features"] = feats
    return out


def segment_chunks(audio: np.ndarray, sampling_rate: int = 16000, segment_duration: int = 5) -> List[np.ndarray]:
    """The segment slicing of REF:whisper_finetune/inference.py:176-190 (ceil(N / (sr*dur)) chunks, empty
    ones skipped), so that all chunks of an upload go through the front end in one batched call
    instead of 1 + 6 serial batch-1 calls."""
    n = len(audio)
    per = sampling_rate * segment_duration
    k = int(np.ceil(n / per))
    if k == 0 and n > 0:
        k = 1
    chunks = [audio[i * per: min((i + 1) * per, n)] for i in range(k)]
    return [c for c in chunks if len(c) > 0]
