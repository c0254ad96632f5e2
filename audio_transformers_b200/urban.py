"""Drop-in for the reference's urban-sounds mel transform.

The reference builds ``T.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)``
(REF:urban_sounds/dataset.py:19-24), calls it on a ``(1, 88200)`` waveform and takes
``torch.log(mel + 1e-9)`` (REF:urban_sounds/dataset.py:55-56).  :class:`B200MelSpectrogram` keeps the
constructor of ``torchaudio.transforms.MelSpectrogram`` (TA:transforms/_transforms.py:566-585), the
``spectrogram.window`` / ``mel_scale.fb`` buffers (so a ``state_dict`` round-trips), and the
``(..., T) -> (..., 64, 1 + T // 512)`` forward, computed by the fused sm_100a kernel on the GPU.
``log_eps`` optionally fuses the reference's ``log(. + 1e-9)`` into the same launch.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops
from ._lib import PRESET_URBAN, TABLE_FILTERBANK, TABLE_WINDOW, get_table


class _Buffers(torch.nn.Module):
    def __init__(self, name: str, value: torch.Tensor):
        super().__init__()
        self.register_buffer(name, value)


class B200MelSpectrogram(torch.nn.Module):
    """CUDA-only ``MelSpectrogram``; only the reference's configuration is compiled in."""

    def __init__(self, sample_rate: int = 22050, n_fft: int = 1024, win_length: Optional[int] = None,
                 hop_length: Optional[int] = 512, f_min: float = 0.0, f_max: Optional[float] = None, pad: int = 0,
                 n_mels: int = 64, window_fn: Callable[..., torch.Tensor] = torch.hann_window, power: float = 2.0,
                 normalized: bool = False, wkwargs: Optional[dict] = None, center: bool = True,
                 pad_mode: str = "reflect", onesided: Optional[bool] = None, norm: Optional[str] = None,
                 mel_scale: str = "htk", log_eps: Optional[float] = None) -> None:
        super().__init__()
        win_length = win_length if win_length is not None else n_fft
        hop_length = hop_length if hop_length is not None else win_length // 2
        f_max_eff = float(f_max) if f_max is not None else float(sample_rate // 2)
        supported = (sample_rate == 22050 and n_fft == 1024 and win_length == 1024 and hop_length == 512
                     and f_min == 0.0 and f_max_eff == 11025.0 and pad == 0 and n_mels == 64
                     and window_fn is torch.hann_window and power == 2.0 and not normalized and wkwargs is None
                     and center and pad_mode == "reflect" and norm is None and mel_scale == "htk")
        if not supported:
            raise NotImplementedError(
                "B200MelSpectrogram is compiled for MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, "
                "n_mels=64) with torchaudio defaults otherwise (REF:urban_sounds/dataset.py:19-24); no CPU fallback")
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = sample_rate, n_fft, win_length, hop_length
        self.pad, self.power, self.normalized, self.n_mels = pad, power, normalized, n_mels
        self.f_min, self.f_max = f_min, f_max
        self.log_eps = log_eps
        # same buffer names as torchaudio: spectrogram.window, mel_scale.fb
        self.spectrogram = _Buffers("window", torch.from_numpy(get_table(PRESET_URBAN, TABLE_WINDOW)))
        self.mel_scale = _Buffers("fb", torch.from_numpy(get_table(PRESET_URBAN, TABLE_FILTERBANK)))

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if not waveform.is_cuda:
            raise RuntimeError("B200MelSpectrogram computes on CUDA only: move the (collated) waveform batch to the "
                               "GPU in the main process before calling it (no CPU fallback)")
        shape = waveform.shape
        x = waveform.reshape(-1, shape[-1])
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        out = ops.mel_power(x.contiguous(), -1.0 if self.log_eps is None else float(self.log_eps))
        return out.reshape(shape[:-1] + out.shape[-2:])
