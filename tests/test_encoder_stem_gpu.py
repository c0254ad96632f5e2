"""GPU parity of the tensor-core encoder stem against plain PyTorch FP32 conv1d / gelu, i.e.
what HF:models/whisper/modeling_whisper.py:619-625 executes.

Two references:
* the FP32 module as it is (weights and features in FP32): the kernel rounds its operands to BF16 (2^-9 relative), so
  the tolerance is the BF16 one -- max-abs <= 2e-2 and relative Frobenius error <= 4e-3 on hidden states of order 1;
* the same arithmetic with the operands the kernel actually multiplies (features, weights and the activations
  between the convolutions rounded to BF16, products and sums in FP32): this isolates the kernel from the rounding of its
  inputs, tolerance 2e-3 max-abs (FP32 accumulation order, and BF16 re-rounding of activations that sit on a rounding
  boundary) and 2e-4 relative Frobenius.
"""
import numpy as np
import pytest
import torch

from audio_transformers_b200 import signals

pytestmark = pytest.mark.gpu


def _encoder():
    tr = pytest.importorskip("transformers")
    torch.manual_seed(99)
    model = tr.WhisperModel(tr.WhisperConfig())            # whisper-tiny geometry by default, random init
    return model.encoder.eval().cuda()


def _reference(enc, feats, round_operands: bool):
    F = torch.nn.functional
    bf = (lambda t: t.to(torch.bfloat16).float()) if round_operands else (lambda t: t)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        x = F.gelu(F.conv1d(bf(feats), bf(enc.conv1.weight), enc.conv1.bias, padding=1))
        x = F.gelu(F.conv1d(bf(x), bf(enc.conv2.weight), enc.conv2.bias, stride=2, padding=1))
        return x.permute(0, 2, 1) + enc.embed_positions.weight


def _features(n, seed):
    from audio_transformers_b200 import ops
    lens = [480000] * (n - 2) + [200000, 16001]
    host = np.zeros((n, 480000), np.float32)
    for i, L in enumerate(lens):
        host[i, :L] = signals.whisper_clip(i, seed=seed, n_samples=L)
    return ops.whisper_logmel(torch.from_numpy(host).cuda(), torch.tensor(lens, dtype=torch.int32).cuda())


@pytest.mark.parametrize("batch", [1, 5])
def test_stem_matches_torch_conv(batch):
    from audio_transformers_b200 import B200WhisperEncoderStem
    enc = _encoder()
    stem = B200WhisperEncoderStem.from_encoder(enc).cuda()
    feats = _features(max(batch, 2), seed=5)[:batch].contiguous()
    out = stem(feats)
    torch.cuda.synchronize()
    assert out.shape == (batch, 1500, 384) and out.dtype == torch.float32 and out.is_contiguous()
    assert torch.isfinite(out).all()
    exact = _reference(enc, feats, round_operands=False)
    same_operands = _reference(enc, feats, round_operands=True)
    e_fp32 = float((out - exact).abs().max())
    r_fp32 = float((out - exact).norm() / exact.norm())
    e_ops = float((out - same_operands).abs().max())
    r_ops = float((out - same_operands).norm() / same_operands.norm())
    print(f"batch {batch}: vs FP32 module max-abs {e_fp32:.2e} rel-fro {r_fp32:.2e}; vs BF16-operand reference max-abs {e_ops:.2e} rel-fro {r_ops:.2e}")
    assert e_ops <= 2e-3 and r_ops <= 2e-4
    assert e_fp32 <= 2e-2 and r_fp32 <= 4e-3


def test_stem_feeds_the_encoder_layers():
    """Hidden states from the stem drive the encoder's transformer layers to the same output as the module's own stem,
    within the BF16 tolerance (the downstream consumer of this row)."""
    from audio_transformers_b200 import B200WhisperEncoderStem
    enc = _encoder()
    stem = B200WhisperEncoderStem.from_encoder(enc).cuda()
    feats = _features(2, seed=8)
    with torch.no_grad():
        ref = enc(feats).last_hidden_state
        hs = stem(feats)
        for layer in enc.layers:
            out = layer(hs, None)
            hs = out[0] if isinstance(out, tuple) else out
        ours = enc.layer_norm(hs)
    rel = float((ours - ref).norm() / ref.norm())
    print("encoder output rel-fro error:", rel)
    assert rel <= 1e-2


def test_stem_rejects_cpu_and_wrong_shapes():
    from audio_transformers_b200 import B200WhisperEncoderStem, ops
    enc = _encoder()
    stem = B200WhisperEncoderStem.from_encoder(enc).cuda()
    with pytest.raises(RuntimeError):
        stem(torch.zeros(1, 80, 3000))
    with pytest.raises(RuntimeError):
        ops.encoder_stem(torch.zeros(1, 80, 2999, device="cuda"), stem.w1, stem.bias1, stem.w2, stem.bias2, stem.positions)
