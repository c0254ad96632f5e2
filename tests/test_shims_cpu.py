"""Host-side behaviour of the drop-in shims that needs no GPU: argument checking and error text
mirror the reference-facing interfaces (HF:models/whisper/feature_extraction_whisper.py:261-276)."""
import numpy as np
import pytest
import torch

from audio_transformers_b200 import B200MelSpectrogram, B200WhisperFeatureExtractor, B200WhisperProcessor


def test_extractor_config_attributes():
    fe = B200WhisperFeatureExtractor()
    # REF:whisper_finetune/experiments.ipynb:558-573
    assert (fe.feature_size, fe.hop_length, fe.chunk_length, fe.n_fft) == (80, 160, 30, 400)
    assert (fe.n_samples, fe.nb_max_frames, fe.sampling_rate) == (480000, 3000, 16000)
    assert fe.mel_filters.shape == (201, 80) and fe.mel_filters.dtype == np.float64
    assert abs(fe.mel_filters[1, 0] - 0.02486259) < 5e-9


def test_wrong_sampling_rate_raises_like_hf():
    fe = B200WhisperFeatureExtractor()
    with pytest.raises(ValueError, match="was trained using a sampling rate of 16000"):
        fe(np.zeros(100, np.float32), sampling_rate=22050)
    tr = pytest.importorskip("transformers")
    with pytest.raises(ValueError) as hf_err:
        tr.WhisperFeatureExtractor()(np.zeros(100, np.float32), sampling_rate=22050)
    with pytest.raises(ValueError) as my_err:
        fe(np.zeros(100, np.float32), sampling_rate=22050)
    assert str(hf_err.value) == str(my_err.value)


def test_unsupported_options_fail_loudly():
    fe = B200WhisperFeatureExtractor()
    x = np.zeros(100, np.float32)
    for kw in (dict(do_normalize=True), dict(padding="longest"), dict(truncation=False), dict(max_length=1000),
               dict(pad_to_multiple_of=8), dict(return_tensors="tf")):
        with pytest.raises(NotImplementedError):
            fe(x, sampling_rate=16000, **kw)
    with pytest.raises(NotImplementedError):
        B200WhisperFeatureExtractor(n_fft=512)
    with pytest.raises(NotImplementedError):
        B200WhisperFeatureExtractor(dither=0.1)
    with pytest.raises(NotImplementedError):
        B200MelSpectrogram(sample_rate=16000)


def test_processor_passthrough():
    class Tok:
        pad_token_id, eos_token_id = 50257, 50256

        def __call__(self, text=None, text_target=None, **kw):
            return {"input_ids": [1, 2, 3]}

        def decode(self, ids, **kw):
            return "decoded"

        def batch_decode(self, ids, **kw):
            return ["decoded"]

    proc = B200WhisperProcessor(tokenizer=Tok())
    assert proc.tokenizer.pad_token_id == 50257                       # REF:whisper_finetune/dataset.py:23
    assert proc.tokenizer(text_target="hi")["input_ids"] == [1, 2, 3]  # REF:whisper_finetune/dataset.py:66
    assert proc.decode([1, 2]) == "decoded"                            # REF:whisper_finetune/inference.py:170
    assert proc(text="hello")["input_ids"] == [1, 2, 3]
    with pytest.raises(ValueError, match="either an `audio` or `text`"):
        proc()
    assert isinstance(proc.feature_extractor, B200WhisperFeatureExtractor)


def test_urban_module_buffers():
    m = B200MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
    sd = m.state_dict()
    assert sd["spectrogram.window"].shape == (1024,) and sd["mel_scale.fb"].shape == (513, 64)
    assert (m.n_fft, m.hop_length, m.n_mels, m.sample_rate) == (1024, 512, 64, 22050)


def test_urban_front_end_host_side():
    """B200UrbanFrontEnd mirrors UrbanSoundDataset's constructor defaults (REF:urban_sounds/dataset.py:8-24); the tap
    table has torchaudio's shape; segment mode validates its arguments before touching the device."""
    from audio_transformers_b200.urban import B200UrbanFrontEnd, sinc_resample_kernel
    k, width, orig, new = sinc_resample_kernel(44100, 22050)
    assert (orig, new, width) == (2, 1, 13) and k.shape == (1, 28) and k.dtype == np.float32
    assert abs(float(k.sum()) - 1.0) < 2e-3                                                         # unit DC gain
    k, width, orig, new = sinc_resample_kernel(48000, 22050)
    assert (orig, new) == (320, 147) and k.shape == (147, 2 * width + 320)
    if not torch.cuda.is_available():
        fe = B200UrbanFrontEnd(device="cpu")
        assert fe.target_length == 88200
        with pytest.raises(RuntimeError):
            fe.process_audio(np.zeros(1000), 44100)
    ext = B200WhisperFeatureExtractor()
    with pytest.raises(ValueError):
        ext.segment_features(np.zeros(10, np.float32), 80001)
    with pytest.raises(ValueError):
        ext.segment_features(np.zeros(10, np.float32), 80000, sampling_rate=8000)
