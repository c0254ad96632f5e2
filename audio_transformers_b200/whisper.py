"""Drop-in for the reference's Whisper feature call site.

The reference calls ``processor(audio_array, sampling_rate=16000, return_tensors="pt").input_features``
(REF:whisper_finetune/dataset.py:58-62, REF:whisper_finetune/inference.py:154,200).  The classes here
keep that call shape -- same arguments, same error behaviour, same key/attribute, same
``(B, 80, 3000)`` float32 result -- but the features are computed by the fused sm_100a kernel and
are born on the GPU.

* :class:`B200WhisperFeatureExtractor` mirrors ``WhisperFeatureExtractor.__call__``
  (HF:models/whisper/feature_extraction_whisper.py:189-343).
* :class:`B200WhisperProcessor` mirrors ``WhisperProcessor`` (HF:models/whisper/processing_whisper.py:23-57)
  and passes everything that is not feature extraction through to the wrapped tokenizer/processor.

There is no CPU fallback: without a CUDA device (or without the built library) a call raises.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch

from . import ops
from . import _lib
from ._lib import PRESET_WHISPER, TABLE_FILTERBANK, get_table

logger = logging.getLogger(__name__)

_CLASS_NAME = "WhisperFeatureExtractor"      # used in the error text the reference's users see


class FeatureBatch(dict):
    """Minimal ``BatchFeature`` stand-in (key *and* attribute access, ``.to()``) used when
    ``transformers`` is not importable."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(item) from e

    def to(self, *args, **kwargs):
        return FeatureBatch({k: (v.to(*args, **kwargs) if isinstance(v, torch.Tensor) else v) for k, v in self.items()})


_BATCH_FEATURE = None          # resolved on first use: transformers' BatchFeature, or the stand-in above


def _batch_feature(data: dict):
    global _BATCH_FEATURE
    if _BATCH_FEATURE is None:
        try:
            from transformers.feature_extraction_utils import BatchFeature
            _BATCH_FEATURE = BatchFeature
        except Exception:  # transformers absent or broken: keep the same access pattern
            _BATCH_FEATURE = FeatureBatch
    return _BATCH_FEATURE(data)


def default_pack_threads() -> int:
    """Host threads one process may use for the cast into pinned memory: the cores it may run on, shared fairly with
    the other ranks of the node (torchrun exports LOCAL_WORLD_SIZE), 16 at most -- the cast is bound by memory
    bandwidth long before that, and eight ranks with 16 threads each would only fight over the same cores."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = os.cpu_count() or 1
    try:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        ranks = 1
    return max(1, min(16, cores // ranks))


class B200WhisperFeatureExtractor:
    """GPU log-mel extractor with ``WhisperFeatureExtractor``'s interface.

    Only the configuration the reference uses is supported (whisper-tiny's
    ``preprocessor_config.json``, printed at REF:whisper_finetune/experiments.ipynb:558-573);
    anything else raises ``NotImplementedError`` instead of silently computing something different.
    """

    model_input_names = ["input_features"]

    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 chunk_length: int = 30, n_fft: int = 400, padding_value: float = 0.0, dither: float = 0.0,
                 return_attention_mask: bool = False, device: Union[str, torch.device, None] = None, **kwargs):
        if (feature_size, sampling_rate, hop_length, chunk_length, n_fft) != (80, 16000, 160, 30, 400):
            raise NotImplementedError(
                "B200WhisperFeatureExtractor is compiled for feature_size=80, sampling_rate=16000, hop_length=160, "
                "chunk_length=30, n_fft=400 (the configuration the reference uses)")
        if padding_value != 0.0 or dither != 0.0:
            raise NotImplementedError("padding_value != 0 and dither != 0 are not supported (no CPU fallback)")
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_fft = n_fft
        self.padding_value = padding_value
        self.padding_side = "right"
        self.dither = dither
        self.return_attention_mask = return_attention_mask
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.device = torch.device(device) if device is not None else None
        self._mel_filters = None
        # two pinned staging buffers used in turn, so that packing batch k+1 overlaps the H2D copy of batch k
        self._stage = [None, None]
        self._stage_turn = 0
        self._pack_threads = default_pack_threads()

    # -- attributes the reference's notebook prints (experiments.ipynb:558-573) --------------------
    @property
    def mel_filters(self) -> np.ndarray:
        """(201, 80) filterbank the kernel is compiled against (float32 values, float64 array like HF's)."""
        if self._mel_filters is None:
            self._mel_filters = get_table(PRESET_WHISPER, TABLE_FILTERBANK).astype(np.float64)
        return self._mel_filters

    def to_dict(self) -> dict:
        return dict(feature_extractor_type=_CLASS_NAME, feature_size=self.feature_size, sampling_rate=self.sampling_rate,
                    hop_length=self.hop_length, chunk_length=self.chunk_length, n_fft=self.n_fft,
                    n_samples=self.n_samples, nb_max_frames=self.nb_max_frames, padding_value=self.padding_value,
                    padding_side=self.padding_side, dither=self.dither, return_attention_mask=self.return_attention_mask)

    def __repr__(self) -> str:
        return f"B200WhisperFeatureExtractor {self.to_dict()}"

    # -- helpers ------------------------------------------------------------------------------------
    def _target_device(self, device) -> torch.device:
        if device is not None and str(device) != "cpu":
            dev = torch.device(device)
        elif self.device is not None:
            dev = self.device
        else:
            dev = torch.device("cuda")
        if dev.type != "cuda":
            raise RuntimeError("B200WhisperFeatureExtractor computes on CUDA only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("B200WhisperFeatureExtractor needs a CUDA device (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    @staticmethod
    def _canonicalise(raw_speech) -> list:
        """HF:models/whisper/feature_extraction_whisper.py:274-292: always a batch, mono.  float32 / float64 arrays are
        kept as they are (the float64 -> float32 cast of :285-286 happens inside the native packer, in parallel);
        anything else is converted to float32 here."""
        is_batched_numpy = isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1
        if is_batched_numpy and raw_speech.ndim > 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {_CLASS_NAME}")
        is_batched = is_batched_numpy or (
            isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
            and isinstance(raw_speech[0], (np.ndarray, tuple, list, torch.Tensor)))

        def one(c):
            if isinstance(c, torch.Tensor):
                c = c.detach().cpu().numpy()
            c = np.asarray(c)
            if c.dtype not in (np.float32, np.float64):
                c = c.astype(np.float32)
            return np.ascontiguousarray(c)

        clips = [one(c) for c in raw_speech] if is_batched else [one(raw_speech)]
        for c in clips:
            if c.ndim != 1:
                raise ValueError(f"Only mono-channel audio is supported for input to {_CLASS_NAME}")
        return clips

    def _slot(self, dev: torch.device, n: int, width: int) -> dict:
        """One of two staging slots used in turn (pinned host rows + lengths, their device twins, an event that marks
        the last use), grown on demand; the previous use of the slot is awaited before it is handed out again."""
        turn = self._stage_turn
        self._stage_turn ^= 1
        slot = self._stage[turn]
        if slot is None or slot["dev"].device != dev or slot["cap"] < n * width or slot["ncap"] < n:
            cap, ncap = max(n * width, 4), max(n, 64)
            slot = {"host": torch.empty((cap,), dtype=torch.float32, pin_memory=True),
                    "lens": torch.empty((ncap,), dtype=torch.int32, pin_memory=True),
                    "dev": torch.empty((cap,), dtype=torch.float32, device=dev),
                    "dev_lens": torch.empty((ncap,), dtype=torch.int32, device=dev),
                    "event": torch.cuda.Event(), "pending": False, "cap": cap, "ncap": ncap}
            self._stage[turn] = slot
        elif slot["pending"]:
            slot["event"].synchronize()
        return slot

    def _features_from_host(self, clips: Sequence[np.ndarray], dev: torch.device):
        """Ragged host clips (float32 / float64) -> features, through ONE native call: the worker pool casts the clips
        into the pinned slot, every worker copies its piece to the device as soon as it is converted, then the kernels
        are enqueued (b200mel_whisper_logmel_host).  Only ``min(len, 480000)`` samples per clip cross PCIe; padding is
        never materialised.  Returns (features, device lengths view)."""
        n = len(clips)
        lens64 = np.fromiter((c.shape[0] for c in clips), dtype=np.int64, count=n)
        ptrs = np.fromiter((c.__array_interface__["data"][0] for c in clips), dtype=np.uint64, count=n)
        is64 = np.fromiter((c.dtype == np.float64 for c in clips), dtype=np.uint8, count=n)
        width = (max(int(min(lens64.max(), self.n_samples)) if n else 0, 4) + 3) // 4 * 4
        slot = self._slot(dev, n, width)
        feats = torch.empty((n, self.feature_size, self.nb_max_frames), dtype=torch.float32, device=dev)
        if n == 0:
            return feats, slot["dev_lens"][:0]
        lib = _lib.load()
        h = ops._handle(dev, PRESET_WHISPER)
        stream = torch.cuda.current_stream(dev)
        ws = ops._whisper_workspace(lib, h, dev, stream.cuda_stream, n)
        args = (h, ptrs.ctypes.data, lens64.ctypes.data, is64.ctypes.data, n, slot["host"].data_ptr(), width,
                slot["lens"].data_ptr(), slot["dev"].data_ptr(), slot["dev_lens"].data_ptr(), feats.data_ptr(),
                ws.data_ptr(), ws.numel(), self._pack_threads, stream.cuda_stream)
        if dev.index == torch.cuda.current_device():
            st = lib.b200mel_whisper_logmel_host(*args)
        else:
            with torch.cuda.device(dev):
                st = lib.b200mel_whisper_logmel_host(*args)
        if st != 0:
            _lib.check(st, "b200mel_whisper_logmel_host")
        slot["event"].record(stream)
        slot["pending"] = True
        return feats, slot["dev_lens"][:n]

    # -- the call -----------------------------------------------------------------------------------
    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of: Optional[int] = None,
                 return_tensors: Optional[str] = None, return_attention_mask: Optional[bool] = None,
                 padding: Optional[str] = "max_length", max_length: Optional[int] = None,
                 sampling_rate: Optional[int] = None, do_normalize: Optional[bool] = None,
                 device: Union[str, torch.device, None] = None, lengths: Optional[torch.Tensor] = None, **kwargs):
        # HF:models/whisper/feature_extraction_whisper.py:261-272
        if sampling_rate is not None:
            if sampling_rate != self.sampling_rate:
                raise ValueError(
                    f"The model corresponding to this feature extractor: {_CLASS_NAME} was trained using a"
                    f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                    f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        else:
            logger.warning(
                f"It is strongly recommended to pass the `sampling_rate` argument to `{_CLASS_NAME}()`. "
                "Failing to do so can result in silent errors that might be hard to debug.")
        if do_normalize:
            raise NotImplementedError("do_normalize=True is not supported by the fused kernel (no CPU fallback)")
        if str(getattr(padding, "value", padding)) != "max_length" or not truncation or pad_to_multiple_of is not None:
            raise NotImplementedError("only padding='max_length', truncation=True, pad_to_multiple_of=None are supported")
        if max_length is not None and max_length != self.n_samples:
            raise NotImplementedError(f"max_length must be {self.n_samples} (30 s); the encoder requires 3000 frames")
        rt = None if return_tensors is None else str(getattr(return_tensors, "value", return_tensors))
        if rt not in (None, "pt", "np"):
            raise NotImplementedError(f"return_tensors={return_tensors!r} is not supported")

        dev = self._target_device(device if not isinstance(raw_speech, torch.Tensor) or not raw_speech.is_cuda
                                  else raw_speech.device)
        if isinstance(raw_speech, torch.Tensor) and raw_speech.is_cuda:
            # new: device-resident waveforms (collated batch) + optional per-clip lengths
            wave = raw_speech
            if wave.dim() == 1:
                wave = wave.unsqueeze(0)
            if wave.dim() != 2:
                raise ValueError(f"Only mono-channel audio is supported for input to {_CLASS_NAME}")
            if wave.dtype != torch.float32:
                wave = wave.to(torch.float32)
            dev_lengths = lengths.to(dev) if lengths is not None else None
        elif (isinstance(raw_speech, torch.Tensor) and raw_speech.dim() == 2 and raw_speech.dtype == torch.float32
              and raw_speech.is_contiguous()):
            # new: an already collated host batch (ideally pinned): one asynchronous H2D, no restaging
            wave = raw_speech.to(dev, non_blocking=True)
            dev_lengths = lengths.to(dev, non_blocking=True) if lengths is not None else None
        else:
            if isinstance(raw_speech, torch.Tensor):
                raw_speech = raw_speech.numpy()
            clips = self._canonicalise(raw_speech)
            feats, dev_lengths = self._features_from_host(clips, dev)
            wave = None
        if wave is not None:
            feats = ops.whisper_logmel(wave, dev_lengths)

        data = {"input_features": feats}
        want_mask = self.return_attention_mask if return_attention_mask is None else return_attention_mask
        if want_mask:
            if dev_lengths is None:
                dev_lengths = torch.full((wave.shape[0],), min(wave.shape[1], self.n_samples), dtype=torch.int32, device=dev)
            elif dev_lengths.dtype != torch.int32:
                dev_lengths = dev_lengths.to(torch.int32)
            data["attention_mask"] = ops.whisper_frame_mask(dev_lengths)
        if rt != "pt":
            # HF hands back numpy unless return_tensors="pt"; that means a device->host copy here
            data = {k: v.cpu().numpy() for k, v in data.items()}
        return _batch_feature(data)


    # -- segment mode --------------------------------------------------------------------------------
    def segment_features(self, raw_speech, segment_samples: int, sampling_rate: Optional[int] = None,
                         device: Union[str, torch.device, None] = None) -> torch.Tensor:
        """Features of consecutive segments of ONE clip, each padded to 30 s, in a single launch.

        REF:whisper_finetune/inference.py:176-200 cuts the upload into ``segment_duration``-second pieces and calls
        the processor once per piece (so each piece is right-padded with zeros to 30 s).  Here the clip crosses
        PCIe once and the pieces are rows of a strided view of that one device buffer (row stride =
        ``segment_samples``, per-row lengths), so no padded copy is ever materialised.  Returns
        ``(ceil(N / segment_samples), 80, 3000)`` float32 on the GPU; an empty clip gives ``(0, 80, 3000)``."""
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor: {_CLASS_NAME} was trained using a"
                f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        if segment_samples <= 0 or segment_samples % 4 != 0 or segment_samples > self.n_samples:
            raise ValueError(f"segment_samples must be a positive multiple of 4 and at most {self.n_samples}")
        if isinstance(raw_speech, torch.Tensor):
            clip = raw_speech.detach().to(torch.float32).reshape(-1)
            dev = self._target_device(clip.device if clip.is_cuda else device)
        else:
            a = np.asarray(raw_speech, dtype=np.float32)
            if a.ndim != 1:
                raise ValueError(f"Only mono-channel audio is supported for input to {_CLASS_NAME}")
            clip = torch.from_numpy(np.ascontiguousarray(a))
            dev = self._target_device(device)
        n = int(clip.shape[0])
        k = (n + segment_samples - 1) // segment_samples
        if k == 0:
            return torch.empty((0, self.feature_size, self.nb_max_frames), dtype=torch.float32, device=dev)
        buf = torch.empty((k * segment_samples + 4,), dtype=torch.float32, device=dev)   # +4: TMA boxes are 16-byte granular
        buf[:n].copy_(clip if clip.is_cuda else clip.pin_memory(), non_blocking=True)
        buf[n:].zero_()
        lens = torch.full((k,), segment_samples, dtype=torch.int32)
        lens[-1] = n - (k - 1) * segment_samples
        rows = buf[:k * segment_samples].view(k, segment_samples)
        return ops.whisper_logmel(rows, lens.to(dev, non_blocking=True))


class B200WhisperProcessor:
    """``WhisperProcessor`` look-alike: audio goes to the B200 extractor, everything else
    (``tokenizer``, ``decode``, ``batch_decode``, ``save_pretrained``, ...) to the wrapped objects.

    REF call sites that rely on the passthrough: whisper_finetune/dataset.py:23,66 (tokenizer),
    inference.py:162,170 (eos id, decode), train.py:135,336 (save_pretrained).
    """

    def __init__(self, feature_extractor: Optional[B200WhisperFeatureExtractor] = None, tokenizer: Any = None,
                 base_processor: Any = None, device: Union[str, torch.device, None] = None):
        self.feature_extractor = feature_extractor or B200WhisperFeatureExtractor(device=device)
        self._base = base_processor
        self.tokenizer = tokenizer if tokenizer is not None else getattr(base_processor, "tokenizer", None)

    @classmethod
    def from_pretrained(cls, name_or_path, device: Union[str, torch.device, None] = None, **kwargs):
        """Load tokenizer/config through ``transformers`` (hub or local path) and swap in the GPU extractor."""
        from transformers import WhisperProcessor
        base = WhisperProcessor.from_pretrained(name_or_path, **kwargs)
        fe = base.feature_extractor
        ext = B200WhisperFeatureExtractor(feature_size=fe.feature_size, sampling_rate=fe.sampling_rate,
                                          hop_length=fe.hop_length, chunk_length=fe.chunk_length, n_fft=fe.n_fft,
                                          padding_value=fe.padding_value, dither=getattr(fe, "dither", 0.0),
                                          return_attention_mask=fe.return_attention_mask, device=device)
        return cls(ext, base.tokenizer, base, device)

    def __call__(self, *args, **kwargs):
        # HF:models/whisper/processing_whisper.py:31-54
        audio = kwargs.pop("audio", None)
        sampling_rate = kwargs.pop("sampling_rate", None)
        text = kwargs.pop("text", None)
        if len(args) > 0:
            audio = args[0]
            args = args[1:]
        if audio is None and text is None:
            raise ValueError("You need to specify either an `audio` or `text` input to process.")
        if audio is not None:
            inputs = self.feature_extractor(audio, *args, sampling_rate=sampling_rate, **kwargs)
        if text is not None:
            if self.tokenizer is None:
                raise ValueError("this processor was built without a tokenizer")
            encodings = self.tokenizer(text, **kwargs)
        if text is None:
            return inputs
        if audio is None:
            return encodings
        inputs["labels"] = encodings["input_ids"]
        return inputs

    def get_decoder_prompt_ids(self, task=None, language=None, no_timestamps=True):
        return self.tokenizer.get_decoder_prompt_ids(task=task, language=language, no_timestamps=no_timestamps)

    def get_prompt_ids(self, text: str, return_tensors="np"):
        return self.tokenizer.get_prompt_ids(text, return_tensors=return_tensors)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def save_pretrained(self, *args, **kwargs):
        if self._base is None:
            raise RuntimeError("save_pretrained needs the wrapped transformers processor (use from_pretrained)")
        return self._base.save_pretrained(*args, **kwargs)

    def __getattr__(self, item):
        base = self.__dict__.get("_base")
        if base is not None:
            return getattr(base, item)
        raise AttributeError(item)
