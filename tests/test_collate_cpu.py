"""Waveform collate: labels/emotion labels are padded exactly like the
reference's collate_fn (REF:whisper_finetune/dataset.py:84-110); waveforms are stacked once."""
import numpy as np
import torch

from audio_transformers_b200.collate import WaveformCollator, segment_chunks, stack_waveforms


def _reference_collate(batch, pad_token_id):
    # restatement of REF:whisper_finetune/dataset.py:85-110 for the label fields
    max_label_length = max(x["labels"].size(0) for x in batch)
    labels = torch.ones(len(batch), max_label_length, dtype=torch.long) * pad_token_id
    emotion_labels = torch.zeros(len(batch), dtype=torch.long)
    for i, item in enumerate(batch):
        labels[i, :item["labels"].size(0)] = item["labels"]
        emotion_labels[i] = item["emotion_label"]
    return labels, emotion_labels


def test_collate_matches_reference_label_handling():
    rng = np.random.default_rng(0)
    batch = []
    for i, n in enumerate((16000, 123457, 480000, 600000, 1)):
        batch.append({"waveform": rng.standard_normal(n).astype(np.float64),      # datasets yields float64
                      "labels": torch.arange(3 + 2 * i, dtype=torch.long),
                      "emotion_label": torch.tensor(i % 3, dtype=torch.long)})
    out = WaveformCollator(pad_token_id=50257)(batch)
    labels, emo = _reference_collate(batch, 50257)
    assert torch.equal(out["labels"], labels) and torch.equal(out["emotion_labels"], emo)
    assert out["waveform"].shape == (5, 480000) and out["waveform"].dtype == torch.float32
    assert out["lengths"].tolist() == [16000, 123457, 480000, 480000, 1]
    for i, b in enumerate(batch):
        n = int(out["lengths"][i])
        assert np.array_equal(out["waveform"][i, :n].numpy(), b["waveform"][:n].astype(np.float32))
        assert not out["waveform"][i, n:].any()


def test_stack_rounds_width_to_four():
    wave, lens = stack_waveforms([np.ones(5), np.ones(2)])
    assert wave.shape == (2, 8) and lens.tolist() == [5, 2]


def test_segment_chunks_like_inference():
    audio = np.arange(16000 * 12, dtype=np.float32)                 # the reference's 12 s dummy clip
    chunks = segment_chunks(audio, 16000, 5)                        # REF:whisper_finetune/inference.py:176-190
    assert [len(c) for c in chunks] == [80000, 80000, 32000]
    assert np.array_equal(np.concatenate(chunks), audio)
    assert len(segment_chunks(np.zeros(3, np.float32))) == 1
