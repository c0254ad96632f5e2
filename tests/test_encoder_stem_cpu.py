"""CPU checks of the encoder stem's host side (no kernel runs here): the packed weight layout of include/b200mel.h
reproduces conv1 / conv2 as plain matrix products over the im2col rows the kernels build, and the module refuses CPU
input loudly."""
import pytest
import torch

from audio_transformers_b200.encoder_stem import B200WhisperEncoderStem, pack_weights

F = torch.nn.functional


def test_packed_weights_are_the_convolutions_as_gemms():
    torch.manual_seed(3)
    c1 = torch.nn.Conv1d(80, 384, 3, padding=1)
    c2 = torch.nn.Conv1d(384, 384, 3, stride=2, padding=1)
    w1, w2 = pack_weights(c1.weight, c2.weight)
    assert w1.shape == (384, 256) and w1.dtype == torch.bfloat16 and float(w1[:, 240:].abs().max()) == 0.0
    assert w2.shape == (384, 1152) and w2.dtype == torch.bfloat16
    x = torch.randn(2, 80, 40)
    # conv1: row t of the im2col image is [x[:, t-1], x[:, t], x[:, t+1]] (es_im2col_kernel), zero outside the clip
    xp = F.pad(x, (1, 1))
    a1 = torch.cat([xp[:, :, k:k + 40] for k in range(3)], dim=1).permute(0, 2, 1)          # (B, T, 240)
    with torch.no_grad():
        w1f = w1[:, :240].float()
        got1 = a1 @ w1f.T + c1.bias
        ref1 = F.conv1d(x, w1f.reshape(384, 3, 80).permute(0, 2, 1), c1.bias, padding=1).permute(0, 2, 1)
        assert torch.allclose(got1, ref1, atol=1e-5)
        # conv2: output row t' reads rows 2t'-1, 2t', 2t'+1 of h = rows 2t', 2t'+1, 2t'+2 of h with one zero row in front
        # and one behind (the 4-D tensor map of the kernel: pair t' + tap // 2, parity tap % 2)
        h = torch.randn(2, 40, 384)
        hp = F.pad(h, (0, 0, 1, 1))
        a2 = torch.cat([hp[:, k:k + 40:2] for k in range(3)], dim=2)                         # (B, T/2, 1152)
        w2f = w2.float()
        got2 = a2 @ w2f.T + c2.bias
        ref2 = F.conv1d(h.permute(0, 2, 1), w2f.reshape(384, 3, 384).permute(0, 2, 1), c2.bias, stride=2, padding=1).permute(0, 2, 1)
        assert torch.allclose(got2, ref2, atol=1e-4)


def test_wrong_geometry_and_cpu_input_are_rejected():
    with pytest.raises(ValueError):
        pack_weights(torch.zeros(384, 128, 3), torch.zeros(384, 384, 3))
    c1, c2 = torch.nn.Conv1d(80, 384, 3, padding=1), torch.nn.Conv1d(384, 384, 3, stride=2, padding=1)
    with pytest.raises(ValueError):
        B200WhisperEncoderStem(c1, torch.nn.Conv1d(384, 384, 3, stride=1, padding=1), torch.nn.Embedding(1500, 384))
    stem = B200WhisperEncoderStem(c1, c2, torch.nn.Embedding(1500, 384))
    assert set(stem.state_dict()) == {"w1", "w2", "bias1", "bias2", "positions"}
    with pytest.raises(RuntimeError, match="no CPU path"):
        stem(torch.zeros(1, 80, 3000))
