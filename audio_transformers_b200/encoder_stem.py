"""The Whisper encoder stem on the tensor cores.

Drop-in for the first lines of ``WhisperEncoder.forward`` (HF:models/whisper/modeling_whisper.py:619-625)::

    inputs_embeds = nn.functional.gelu(self.conv1(input_features))
    inputs_embeds = nn.functional.gelu(self.conv2(inputs_embeds))
    inputs_embeds = inputs_embeds.permute(0, 2, 1)
    hidden_states = inputs_embeds + self.embed_positions(all_positions)

for the whisper-tiny geometry (80 mels, d_model 384, 1500 positions; ``conv1``/``conv2``/``embed_positions`` as built
at :567-570).  BF16 operands, FP32 accumulation and FP32 output; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops

N_MEL, D_MODEL, N_POS = 80, 384, 1500


def pack_weights(conv1_weight: torch.Tensor, conv2_weight: torch.Tensor):
    """``Conv1d`` weights (out, in, tap) -> the K-major BF16 matrices of include/b200mel.h: column ``tap * in + ci``;
    conv1's 240 columns are padded with zeros to 256."""
    if tuple(conv1_weight.shape) != (D_MODEL, N_MEL, 3) or tuple(conv2_weight.shape) != (D_MODEL, D_MODEL, 3):
        raise ValueError(f"encoder stem is built for whisper-tiny: conv1 (384, 80, 3), conv2 (384, 384, 3); got "
                         f"{tuple(conv1_weight.shape)}, {tuple(conv2_weight.shape)}")
    w1 = torch.zeros((D_MODEL, 256), dtype=torch.bfloat16, device=conv1_weight.device)
    w1[:, :240] = conv1_weight.detach().permute(0, 2, 1).reshape(D_MODEL, 240).to(torch.bfloat16)
    w2 = conv2_weight.detach().permute(0, 2, 1).reshape(D_MODEL, 3 * D_MODEL).to(torch.bfloat16).contiguous()
    return w1, w2


class B200WhisperEncoderStem(torch.nn.Module):
    """``stem(input_features) -> hidden_states`` (B, 1500, 384) float32, ready for ``encoder.layers``."""

    def __init__(self, conv1: torch.nn.Conv1d, conv2: torch.nn.Conv1d, embed_positions: torch.nn.Embedding):
        super().__init__()
        if conv1.stride != (1,) or conv1.padding != (1,) or conv2.stride != (2,) or conv2.padding != (1,):
            raise ValueError("encoder stem: conv1 must be (stride 1, padding 1) and conv2 (stride 2, padding 1)")
        if tuple(embed_positions.weight.shape) != (N_POS, D_MODEL):
            raise ValueError(f"encoder stem: embed_positions must be (1500, 384), got {tuple(embed_positions.weight.shape)}")
        w1, w2 = pack_weights(conv1.weight, conv2.weight)
        self.register_buffer("w1", w1)
        self.register_buffer("w2", w2)
        self.register_buffer("bias1", conv1.bias.detach().float().clone())
        self.register_buffer("bias2", conv2.bias.detach().float().clone())
        self.register_buffer("positions", embed_positions.weight.detach().float().clone())

    @classmethod
    def from_encoder(cls, encoder) -> "B200WhisperEncoderStem":
        """``encoder``: a ``transformers`` ``WhisperEncoder`` (``model.whisper.model.encoder`` in the reference's
        ``EmotionWhisperModel``, REF:whisper_finetune/model.py:6-18)."""
        return cls(encoder.conv1, encoder.conv2, encoder.embed_positions)

    def forward(self, input_features: torch.Tensor) -> torch.Tensor:
        if not input_features.is_cuda:
            raise RuntimeError("B200WhisperEncoderStem: input_features must be a CUDA tensor (there is no CPU path)")
        return ops.encoder_stem(input_features.float(), self.w1, self.bias1, self.w2, self.bias2, self.positions)


def use_b200_stem(encoder) -> B200WhisperEncoderStem:
    """Make a ``transformers`` ``WhisperEncoder`` run its first lines (HF:models/whisper/modeling_whisper.py:619-625:
    conv1, GELU, conv2, GELU, permute, positions) on the tensor-core stem; the transformer layers and the final layer
    norm (:628-645) stay the module's own.  For inference call sites of the reference (REF:whisper_finetune/
    evaluate_simple.py:115-143, inference.py:154-170: ``model.whisper.generate`` / ``model(...)`` under
    ``torch.no_grad()``): the stem has no backward, so a call that would need gradients raises instead of silently
    training through a constant.  The weights are packed once, here; call again after loading new weights.
    ``encoder.forward`` is replaced on the instance; ``encoder._b200_original_forward`` keeps the module's own.
    """
    from transformers.modeling_outputs import BaseModelOutput

    dev = encoder.conv1.weight.device
    if dev.type != "cuda":
        raise RuntimeError("use_b200_stem: move the encoder to a CUDA device first (there is no CPU path)")
    stem = B200WhisperEncoderStem.from_encoder(encoder).to(dev)
    original = getattr(encoder, "_b200_original_forward", None) or encoder.forward
    expected = encoder.config.max_source_positions * encoder.conv1.stride[0] * encoder.conv2.stride[0]

    def forward(input_features, attention_mask=None, **kwargs):
        if torch.is_grad_enabled() and (encoder.training or input_features.requires_grad):
            raise NotImplementedError("the B200 encoder stem is inference only (no backward): wrap the call in "
                                      "torch.no_grad() / model.eval(), or restore encoder._b200_original_forward")
        if input_features.shape[-1] != expected:
            raise ValueError(f"Whisper expects the mel input features to be of length {expected}, but found "
                             f"{input_features.shape[-1]}. Make sure to pad the input mel features to {expected}.")
        hidden = stem(input_features).to(encoder.layer_norm.weight.dtype)
        for layer in encoder.layers:
            out = layer(hidden, None)
            hidden = out[0] if isinstance(out, tuple) else out
        return BaseModelOutput(last_hidden_state=encoder.layer_norm(hidden))

    encoder._b200_original_forward = original
    encoder.forward = forward
    return stem
