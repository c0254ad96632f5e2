"""torch custom ops over the C ABI (CUDA only, no autograd -- the reference never differentiates
through its features: the extractor returns numpy, HF:models/whisper/feature_extraction_whisper.py:164).

    torch.ops.b200mel.whisper_logmel(wave, lengths) -> (B, 80, 3000) float32
    torch.ops.b200mel.whisper_frame_mask(lengths)   -> (B, 3000) int32
    torch.ops.b200mel.mel_power(wave, log_eps)       -> (B, 64, 1 + T // 512) float32
    torch.ops.b200mel.urban_prep(audio, lengths, orig, new, taps, width, out_samples) -> (B, out_samples) float32
    torch.ops.b200mel.encoder_stem(features, w1, bias1, w2, bias2, positions) -> (B, 1500, 384) float32

All inputs and outputs live on the same CUDA device; the kernels are enqueued on the current
stream of that device without any host synchronisation.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Optional

import torch

from . import _lib

W_NMEL, W_NFRAME, W_NSAMP = 80, 3000, 480000
U_NMEL, U_HOP, U_NFFT = 64, 512, 1024

_handles: dict = {}
_hlock = threading.Lock()


def _handle(device: torch.device, preset: int) -> ctypes.c_void_p:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, preset)
    h = _handles.get(key)
    if h is None:
        with _hlock:
            h = _handles.get(key)
            if h is None:
                lib = _lib.load()
                out = ctypes.c_void_p()
                _lib.check(lib.b200mel_create(idx, preset, ctypes.byref(out)), "b200mel_create")
                _handles[key] = h = out
    return h


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"b200mel: `{name}` must be a CUDA tensor (this library has no CPU path)")


_LIBDEF = torch.library.Library("b200mel", "DEF")
_LIBDEF.define("whisper_logmel(Tensor wave, Tensor? lengths) -> Tensor")
_LIBDEF.define("whisper_frame_mask(Tensor lengths) -> Tensor")
_LIBDEF.define("mel_power(Tensor wave, float log_eps) -> Tensor")
_LIBDEF.define("urban_prep(Tensor audio, Tensor? lengths, int orig_freq, int new_freq, Tensor? taps, int width, "
               "int out_samples) -> Tensor")


_LIBDEF.define("encoder_stem(Tensor features, Tensor w1, Tensor bias1, Tensor w2, Tensor bias2, Tensor positions) -> Tensor")


_workspaces: dict = {}        # (device index, stream, batch) -> zero-initialised workspace tensor


def _whisper_workspace(lib, h, dev: torch.device, stream: int, batch: int) -> torch.Tensor:
    """The per-(clip, tile, warp) maxima the floor pass reduces (include/b200mel.h: no initialisation needed).  One
    buffer per (device, stream, batch) is kept and reused, so a call allocates nothing; calls on different streams
    never share one."""
    key = (dev.index, stream, batch)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = max(int(lib.b200mel_workspace_bytes(h, batch)), 16)
        ws = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
        if not torch.cuda.is_current_stream_capturing():
            if len(_workspaces) > 256:
                _workspaces.clear()
            _workspaces[key] = ws
    return ws


def _whisper_logmel_cuda(wave: torch.Tensor, lengths: Optional[torch.Tensor]) -> torch.Tensor:
    _require_cuda(wave, "wave")
    if wave.dim() != 2 or wave.dtype != torch.float32:
        raise ValueError("b200mel.whisper_logmel: wave must be a (B, T) float32 tensor")
    batch, t = wave.shape
    if lengths is not None:
        _require_cuda(lengths, "lengths")
        if lengths.device != wave.device:
            raise ValueError("b200mel.whisper_logmel: lengths must live on the same device as wave")
        if lengths.numel() != batch:
            raise ValueError("b200mel.whisper_logmel: lengths must have one entry per clip")
        if lengths.dtype != torch.int32:
            lengths = lengths.to(torch.int32)
        # a clip never extends past its row of the (B, T) view (stream-ordered, no host sync)
        lengths = lengths.clamp(max=t).contiguous()
    if t % 4 != 0 or wave.stride(1) != 1 or wave.stride(0) % 4 != 0 or wave.data_ptr() % 16 != 0 or \
            (batch > 1 and wave.stride(0) < t):
        # keep the kernel's alignment contract: repack into a 16-byte aligned, stride%4==0 buffer
        tp = (t + 3) // 4 * 4
        if lengths is None:
            lengths = torch.full((batch,), t, dtype=torch.int32, device=wave.device)
        buf = torch.zeros((batch, tp), dtype=torch.float32, device=wave.device)
        buf[:, :t] = wave
        wave = buf
    stride = wave.stride(0) if batch > 1 else wave.shape[1]
    if lengths is None and stride != t:
        # a column-sliced view: without lengths the C ABI takes the whole row stride as the clip
        lengths = torch.full((batch,), t, dtype=torch.int32, device=wave.device)
    out = torch.empty((batch, W_NMEL, W_NFRAME), dtype=torch.float32, device=wave.device)
    if batch == 0:
        return out
    lib = _lib.load()
    dev = wave.device
    h = _handle(dev, _lib.PRESET_WHISPER)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _whisper_workspace(lib, h, dev, stream, batch)
    # the host side of a call is a few tens of microseconds of Python against ~60 us on the GPU for 64 clips, so keep
    # this path lean: no device context manager unless the tensor lives on another device than the current one
    args = (h, wave.data_ptr(), stride, lengths.data_ptr() if lengths is not None else None, batch, out.data_ptr(),
            ws.data_ptr(), ws.numel(), stream)
    if dev.index == torch.cuda.current_device():
        st = lib.b200mel_whisper_logmel_f32(*args)
    else:
        with torch.cuda.device(dev):
            st = lib.b200mel_whisper_logmel_f32(*args)
    if st != 0:
        _lib.check(st, "b200mel_whisper_logmel_f32")
    return out


def _whisper_frame_mask_cuda(lengths: torch.Tensor) -> torch.Tensor:
    _require_cuda(lengths, "lengths")
    lengths = lengths.to(torch.int32).contiguous()
    batch = lengths.numel()
    out = torch.empty((batch, W_NFRAME), dtype=torch.int32, device=lengths.device)
    if batch == 0:
        return out
    lib = _lib.load()
    h = _handle(lengths.device, _lib.PRESET_WHISPER)
    with torch.cuda.device(lengths.device):
        st = lib.b200mel_whisper_frame_mask(h, ctypes.c_void_p(lengths.data_ptr()), batch,
                                            ctypes.c_void_p(out.data_ptr()), _stream_ptr(lengths.device))
    _lib.check(st, "b200mel_whisper_frame_mask")
    return out


def _mel_power_cuda(wave: torch.Tensor, log_eps: float) -> torch.Tensor:
    _require_cuda(wave, "wave")
    if wave.dim() != 2 or wave.dtype != torch.float32:
        raise ValueError("b200mel.mel_power: wave must be a (B, T) float32 tensor")
    batch, t = wave.shape
    if t <= U_NFFT // 2:
        raise ValueError(f"b200mel.mel_power: reflect padding needs more than {U_NFFT // 2} samples, got {t}")
    if t % 4 != 0 or wave.stride(1) != 1 or wave.stride(0) % 4 != 0 or wave.data_ptr() % 16 != 0:
        tp = (t + 3) // 4 * 4
        buf = torch.zeros((batch, tp), dtype=torch.float32, device=wave.device)
        buf[:, :t] = wave
        wave = buf
    stride = wave.stride(0) if batch > 1 else wave.shape[1]
    out = torch.empty((batch, U_NMEL, 1 + t // U_HOP), dtype=torch.float32, device=wave.device)
    if batch == 0:
        return out
    lib = _lib.load()
    h = _handle(wave.device, _lib.PRESET_URBAN)
    with torch.cuda.device(wave.device):
        st = lib.b200mel_mel_f32(h, ctypes.c_void_p(wave.data_ptr()), stride, t, batch, float(log_eps),
                                 ctypes.c_void_p(out.data_ptr()), _stream_ptr(wave.device))
    _lib.check(st, "b200mel_mel_f32")
    return out


def _urban_prep_cuda(audio: torch.Tensor, lengths: Optional[torch.Tensor], orig_freq: int, new_freq: int,
                     taps: Optional[torch.Tensor], width: int, out_samples: int) -> torch.Tensor:
    """(B, C, T) planar float32 -> (B, out_samples): mono mean, sinc resampling, pad/trim, peak normalisation."""
    _require_cuda(audio, "audio")
    if audio.dim() != 3 or audio.dtype != torch.float32:
        raise ValueError("b200mel.urban_prep: audio must be a (B, C, T) float32 tensor")
    audio = audio.contiguous()
    batch, channels, t = audio.shape
    if lengths is not None:
        _require_cuda(lengths, "lengths")
        if lengths.dtype != torch.int32 or lengths.shape != (batch,):
            raise ValueError("b200mel.urban_prep: lengths must be an int32 tensor of shape (B,)")
        lengths = lengths.contiguous()
    if orig_freq != new_freq:
        if taps is None:
            raise ValueError("b200mel.urban_prep: resampling needs the tap table")
        _require_cuda(taps, "taps")
        if taps.dtype != torch.float32 or tuple(taps.shape) != (new_freq, 2 * width + orig_freq):
            raise ValueError("b200mel.urban_prep: taps must be float32 of shape (new_freq, 2*width + orig_freq)")
        taps = taps.contiguous()
    out_stride = (out_samples + 3) // 4 * 4
    out = torch.empty((batch, out_stride), dtype=torch.float32, device=audio.device)
    if batch == 0:
        return out[:, :out_samples]
    lib = _lib.load()
    h = _handle(audio.device, _lib.PRESET_URBAN)
    ws_bytes = lib.b200mel_urban_prep_workspace_bytes(h, batch)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=audio.device)
    with torch.cuda.device(audio.device):
        st = lib.b200mel_urban_prep_f32(
            h, ctypes.c_void_p(audio.data_ptr()), t, ctypes.c_void_p(lengths.data_ptr()) if lengths is not None else None,
            channels, batch, int(orig_freq), int(new_freq),
            ctypes.c_void_p(taps.data_ptr()) if (taps is not None and orig_freq != new_freq) else None, int(width),
            ctypes.c_void_p(out.data_ptr()), out_stride, int(out_samples),
            ctypes.c_void_p(ws.data_ptr()), ws_bytes, _stream_ptr(audio.device))
    _lib.check(st, "b200mel_urban_prep_f32")
    return out[:, :out_samples] if out_stride == out_samples else out[:, :out_samples]


_LIBDEF.impl("whisper_logmel", _whisper_logmel_cuda, "CUDA")
_LIBDEF.impl("urban_prep", _urban_prep_cuda, "CUDA")
_LIBDEF.impl("whisper_frame_mask", _whisper_frame_mask_cuda, "CUDA")
_LIBDEF.impl("mel_power", _mel_power_cuda, "CUDA")


def _encoder_stem_cuda(features: torch.Tensor, w1: torch.Tensor, bias1: torch.Tensor, w2: torch.Tensor,
                       bias2: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py:619-625 (whisper-tiny geometry) on the tensor cores; weights in the
    layout include/b200mel.h describes (encoder_stem.pack_weights builds them)."""
    for t, name in ((features, "features"), (w1, "w1"), (bias1, "bias1"), (w2, "w2"), (bias2, "bias2"), (positions, "positions")):
        _require_cuda(t, name)
        if t.device != features.device:
            raise RuntimeError(f"b200mel: encoder_stem: `{name}` is on {t.device}, features on {features.device}")
    if features.dtype != torch.float32 or features.dim() != 3 or tuple(features.shape[1:]) != (W_NMEL, W_NFRAME):
        raise RuntimeError(f"b200mel: encoder_stem: features must be float32 (B, 80, 3000), got {features.dtype} {tuple(features.shape)}")
    if w1.dtype != torch.bfloat16 or tuple(w1.shape) != (384, 256) or w2.dtype != torch.bfloat16 or tuple(w2.shape) != (384, 1152):
        raise RuntimeError("b200mel: encoder_stem: w1 must be bfloat16 (384, 256) and w2 bfloat16 (384, 1152)")
    for t, shape, name in ((bias1, (384,), "bias1"), (bias2, (384,), "bias2"), (positions, (1500, 384), "positions")):
        if t.dtype != torch.float32 or tuple(t.shape) != shape:
            raise RuntimeError(f"b200mel: encoder_stem: `{name}` must be float32 {shape}")
    features, w1, bias1, w2, bias2, positions = (t.contiguous() for t in (features, w1, bias1, w2, bias2, positions))
    lib = _lib.load()
    dev = features.device
    h = _handle(dev, _lib.PRESET_WHISPER)
    batch = features.shape[0]
    out = torch.empty((batch, 1500, 384), dtype=torch.float32, device=dev)
    if batch == 0:
        return out
    nbytes = int(lib.b200mel_encoder_stem_workspace_bytes(h, batch))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200mel_encoder_stem_bf16(h, features.data_ptr(), batch, w1.data_ptr(), bias1.data_ptr(), w2.data_ptr(),
                                                 bias2.data_ptr(), positions.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes,
                                                 _stream_ptr(dev)), "b200mel_encoder_stem_bf16")
    return out


_LIBDEF.impl("encoder_stem", _encoder_stem_cuda, "CUDA")


@torch.library.register_fake("b200mel::encoder_stem")
def _encoder_stem_fake(features, w1, bias1, w2, bias2, positions):
    return features.new_empty((features.shape[0], 1500, 384))


@torch.library.register_fake("b200mel::whisper_logmel")
def _whisper_logmel_fake(wave, lengths):
    return wave.new_empty((wave.shape[0], W_NMEL, W_NFRAME), dtype=torch.float32)


@torch.library.register_fake("b200mel::whisper_frame_mask")
def _whisper_frame_mask_fake(lengths):
    return lengths.new_empty((lengths.shape[0], W_NFRAME), dtype=torch.int32)


@torch.library.register_fake("b200mel::mel_power")
def _mel_power_fake(wave, log_eps):
    return wave.new_empty((wave.shape[0], U_NMEL, 1 + wave.shape[1] // U_HOP), dtype=torch.float32)


@torch.library.register_fake("b200mel::urban_prep")
def _urban_prep_fake(audio, lengths, orig_freq, new_freq, taps, width, out_samples):
    return audio.new_empty((audio.shape[0], out_samples), dtype=torch.float32)


def urban_prep(audio: torch.Tensor, lengths: Optional[torch.Tensor], orig_freq: int, new_freq: int,
               taps: Optional[torch.Tensor], width: int, out_samples: int) -> torch.Tensor:
    return torch.ops.b200mel.urban_prep(audio, lengths, orig_freq, new_freq, taps, width, out_samples)


def whisper_logmel(wave: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    return torch.ops.b200mel.whisper_logmel(wave, lengths)


def encoder_stem(features: torch.Tensor, w1: torch.Tensor, bias1: torch.Tensor, w2: torch.Tensor, bias2: torch.Tensor,
                 positions: torch.Tensor) -> torch.Tensor:
    return torch.ops.b200mel.encoder_stem(features, w1, bias1, w2, bias2, positions)


def whisper_frame_mask(lengths: torch.Tensor) -> torch.Tensor:
    return torch.ops.b200mel.whisper_frame_mask(lengths)


def mel_power(wave: torch.Tensor, log_eps: float = -1.0) -> torch.Tensor:
    return torch.ops.b200mel.mel_power(wave, log_eps)


def profile_begin(device, preset: int = _lib.PRESET_WHISPER, max_launches: int = 4096) -> None:
    """Benchmark hook: bracket the dominant kernel of every following call with CUDA events."""
    dev = torch.device(device)
    _lib.check(_lib.load().b200mel_profile_begin(_handle(dev, preset), max_launches), "b200mel_profile_begin")


def profile_end(device, preset: int = _lib.PRESET_WHISPER):
    """Returns (summed dominant-kernel milliseconds, launches covered)."""
    dev = torch.device(device)
    ms, n = ctypes.c_double(0.0), ctypes.c_int32(0)
    _lib.check(_lib.load().b200mel_profile_end(_handle(dev, preset), ctypes.byref(ms), ctypes.byref(n)), "b200mel_profile_end")
    return ms.value, n.value
