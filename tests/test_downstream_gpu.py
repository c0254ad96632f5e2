"""Downstream parity (BASELINE north_star; SURVEY.md section 8c "Downstream parity", 8d config 3): the features of
the CUDA path must drive the reference's model to the SAME greedy token ids and the SAME emotion-head argmax as
the features of the reference's own extractor.

Real whisper-tiny weights and the tokenizer are not available offline, so -- as SURVEY.md prescribes -- the model
is a seeded random-init ``EmotionWhisperModel(WhisperConfig(), num_emotions_classes=10)``: the reference's own class
(REF:whisper_finetune/model.py) when the reference tree is present, else its restatement below (composition :6-18,
sequence-level branch of forward :57-97) -- the tree does not travel to the GPU box.  32 clips, eight of each signal
class, generation settings and token budget of REF:whisper_finetune/evaluate_simple.py:125-135.
"""
import importlib.util
import os
import numpy as np
import pytest
import torch

from audio_transformers_b200 import signals

pytestmark = pytest.mark.gpu


REF_MODEL = "/root/reference/whisper_finetune/model.py"


def _reference_class():
    """The reference's own EmotionWhisperModel, or None when the tree is not there (the GPU box)."""
    if not os.path.exists(REF_MODEL):
        return None
    import sys
    spec = importlib.util.spec_from_file_location("ref_whisper_model", REF_MODEL)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.EmotionWhisperModel


def _restated_class():
    from transformers import WhisperForConditionalGeneration, WhisperPreTrainedModel

    class EmotionWhisperModel(WhisperPreTrainedModel):           # REF:whisper_finetune/model.py:6-18
        def __init__(self, config, num_emotions_classes=10):
            super().__init__(config)
            self.whisper = WhisperForConditionalGeneration(config)
            self.emotion_classifier = torch.nn.Linear(config.d_model, num_emotions_classes)
            self.post_init()

        def forward(self, input_features, decoder_input_ids=None):   # REF:whisper_finetune/model.py:57-97
            outputs = self.whisper(input_features, decoder_input_ids=decoder_input_ids, return_dict=True,
                                   output_hidden_states=True)
            hidden_states = outputs.decoder_hidden_states[-1]
            return {"logits": outputs.logits,
                    "emotion_logits": self.emotion_classifier(torch.mean(hidden_states, dim=1))}

    return EmotionWhisperModel


def _build_model():
    tr = pytest.importorskip("transformers")
    from transformers import WhisperConfig
    cls = _reference_class()
    print("model class:", "the reference's own EmotionWhisperModel" if cls is not None else "restatement (reference tree absent)")
    if cls is None:
        cls = _restated_class()
    torch.manual_seed(1234)
    return tr, cls(WhisperConfig(), num_emotions_classes=10).eval().cuda()


def test_greedy_ids_and_emotion_argmax_identical():
    tr, model = _build_model()
    from audio_transformers_b200 import B200WhisperFeatureExtractor
    lens = [480000] * 24 + [200000, 64000, 333333, 16000, 479999, 160000, 96000, 31999]
    clips = [signals.whisper_clip(i, seed=31, n_samples=n) for i, n in enumerate(lens)]       # classes cycle with i
    ref_fe = tr.WhisperFeatureExtractor()
    ref = ref_fe([c.astype(np.float64) for c in clips], sampling_rate=16000, return_tensors="pt").input_features.cuda()
    ours = B200WhisperFeatureExtractor(device="cuda")(clips, sampling_rate=16000, return_tensors="pt").input_features
    assert ours.is_cuda and ours.shape == ref.shape == (len(clips), 80, 3000)
    print("max-abs feature difference:", float((ours - ref).abs().max()))
    assert float((ours - ref).abs().max()) <= 1e-4

    eos = model.config.eos_token_id

    def run(feats):
        ids_all, emo_all = [], []
        with torch.no_grad():
            for b0 in range(0, feats.shape[0], 4):                     # evaluate_simple.py runs batch 4
                f = feats[b0:b0 + 4]
                ids = model.whisper.generate(f, max_new_tokens=100, eos_token_id=eos, pad_token_id=eos, do_sample=False,
                                             no_repeat_ngram_size=3, repetition_penalty=1.15, length_penalty=-0.5,
                                             forced_decoder_ids=None)
                out = model(input_features=f, decoder_input_ids=ids)
                ids_all.append(torch.nn.functional.pad(ids, (0, 128 - ids.shape[1]), value=-1))
                emo_all.append(out["emotion_logits"])
        return torch.cat(ids_all), torch.cat(emo_all)

    ids_ref, emo_ref = run(ref)
    ids_ours, emo_ours = run(ours)
    assert ids_ref.shape == ids_ours.shape and torch.equal(ids_ref, ids_ours), "greedy token ids differ"
    assert torch.equal(emo_ref.argmax(-1), emo_ours.argmax(-1)), "emotion argmax differs"
    print("tokens per clip:", ids_ref.shape[1], " emotion argmax:", emo_ref.argmax(-1).tolist(),
          " max |d emotion logit|:", float((emo_ref - emo_ours).abs().max()))
