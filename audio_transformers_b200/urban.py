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

import math
from typing import Callable, Optional, Sequence

import numpy as np
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


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The tap table ``T.Resample(orig_freq, new_freq)`` convolves with (sinc_interp_hann, torchaudio defaults),
    restated from TA:functional/functional.py ``_get_sinc_resample_kernel`` in FP64 and rounded once to FP32.

    Returns ``(taps[new][2*width + orig] float32, width, orig, new)`` with the rates divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torch.arange(0, -new, -1) is int64 there and `/ new_freq` yields the default dtype: the phase term is FP32
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = np.clip((phase + idx) * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * window * (base / orig)
    return kern.astype(np.float32), width, orig, new


class B200UrbanFrontEnd:
    """GPU version of ``UrbanSoundDataset.process_audio`` (REF:urban_sounds/dataset.py:26-58) for whole batches:
    mono mean, ``T.Resample(orig_sr, 22050)``, pad/trim to 4 s, peak normalisation, 64-mel spectrogram and
    ``log(. + 1e-9)``.  ``process_batch`` takes the raw decoded arrays and their sampling rates (what
    ``__getitem__`` reads from the dataset, REF:urban_sounds/dataset.py:64-68), groups them by (rate, channels),
    stages each group through one pinned host buffer and returns ``(B, 1, 64, 173)`` on the GPU."""

    def __init__(self, sr: int = 22050, duration: float = 4.0, n_fft: int = 1024, hop_length: int = 512,
                 n_mels: int = 64, device="cuda") -> None:
        self.sr, self.duration = sr, duration
        self.target_length = int(sr * duration)
        self.device = torch.device(device)
        self.mel_transform = B200MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                                                log_eps=1e-9)
        self._taps: dict = {}

    def _taps_for(self, orig_sr: int):
        key = int(orig_sr)
        if key not in self._taps:
            k, width, orig, new = sinc_resample_kernel(key, self.sr)
            self._taps[key] = (torch.from_numpy(k).to(self.device), width, orig, new)
        return self._taps[key]

    def waveforms(self, arrays: Sequence[np.ndarray], rates: Sequence[int]) -> torch.Tensor:
        """The normalised ``(B, target_length)`` waveforms the mel transform is applied to (dataset.py:28-52)."""
        if len(arrays) != len(rates):
            raise ValueError("one sampling rate per clip")
        out = torch.empty((len(arrays), self.target_length), dtype=torch.float32, device=self.device)
        groups: dict = {}
        for i, (a, r) in enumerate(zip(arrays, rates)):
            a = np.asarray(a)
            if a.ndim > 2:
                raise ValueError("audio arrays must be (samples,) or (channels, samples)")
            ch = a.shape[0] if a.ndim == 2 else 1
            groups.setdefault((int(r), ch), []).append(i)
        for (rate, ch), idxs in groups.items():
            tmax = max(np.asarray(arrays[i]).shape[-1] for i in idxs)
            host = torch.zeros((len(idxs), ch, max(tmax, 1)), dtype=torch.float32).pin_memory()
            lens = torch.empty((len(idxs),), dtype=torch.int32)
            for j, i in enumerate(idxs):
                a = np.asarray(arrays[i], dtype=np.float32).reshape(ch, -1)        # dataset.py:28 .float()
                host[j, :, :a.shape[1]] = torch.from_numpy(a)
                lens[j] = a.shape[1]
            dev = host.to(self.device, non_blocking=True)
            if rate == self.sr:
                taps, width, orig, new = None, 0, 1, 1
            else:
                taps, width, orig, new = self._taps_for(rate)
            w = ops.urban_prep(dev, lens.to(self.device), orig, new, taps, width, self.target_length)
            out[torch.as_tensor(idxs, device=self.device)] = w
        return out

    def process_batch(self, arrays: Sequence[np.ndarray], rates: Sequence[int]) -> torch.Tensor:
        return self.mel_transform(self.waveforms(arrays, rates)).unsqueeze(1)

    def process_audio(self, audio_array: np.ndarray, orig_sr: int) -> torch.Tensor:
        """Same signature and output shape as the reference method: ``(1, 64, 173)``."""
        return self.process_batch([audio_array], [orig_sr])[0]
