"""The GPU downstream-parity test restates ``EmotionWhisperModel`` because the reference tree does not travel to the GPU
box.  Here, where the tree IS present, the restatement is checked against the reference's own class
(REF:whisper_finetune/model.py:6-18, :57-107): same parameters under the same seed, same logits and emotion logits on the
same inputs.  Skipped when /root/reference is absent."""
import os

import pytest
import torch

import test_downstream_gpu as G


@pytest.mark.skipif(not os.path.exists(G.REF_MODEL), reason="reference tree not present")
def test_restated_model_equals_reference_class():
    pytest.importorskip("transformers")
    from transformers import WhisperConfig
    cfg = WhisperConfig(encoder_layers=1, decoder_layers=1, d_model=64, encoder_attention_heads=2, decoder_attention_heads=2,
                        encoder_ffn_dim=128, decoder_ffn_dim=128)
    torch.manual_seed(7)
    ref = G._reference_class()(cfg, num_emotions_classes=10).eval()
    torch.manual_seed(7)
    mine = G._restated_class()(cfg, num_emotions_classes=10).eval()
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    for (k, a), (_, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert torch.equal(a, b), k
    feats = torch.randn(2, 80, 3000)
    ids = torch.tensor([[50258, 11, 12, 13], [50258, 21, 22, 23]])
    with torch.no_grad():
        a = ref(input_features=feats, decoder_input_ids=ids)
        b = mine(input_features=feats, decoder_input_ids=ids)
    assert torch.equal(a["logits"], b["logits"]) and torch.equal(a["emotion_logits"], b["emotion_logits"])
