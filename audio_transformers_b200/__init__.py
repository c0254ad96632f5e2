"""B200-native log-mel spectrogram front end: a drop-in for the two feature call sites of
k0r1g/audio-transformers (Whisper ``input_features`` and the urban-sounds ``MelSpectrogram``).

Importing the package never touches CUDA; the compiled sm_100a library is loaded on first use
and every entry point fails loudly when it (or a CUDA device) is missing.  See DESIGN.md.
"""
__version__ = "0.1.0"

__all__ = ["B200WhisperFeatureExtractor", "B200WhisperProcessor", "B200MelSpectrogram", "B200UrbanFrontEnd", "B200WhisperEncoderStem",
           "ops", "signals"]


def __getattr__(name):
    if name in ("B200WhisperFeatureExtractor", "B200WhisperProcessor"):
        from . import whisper
        return getattr(whisper, name)
    if name in ("B200MelSpectrogram", "B200UrbanFrontEnd"):
        from . import urban
        return getattr(urban, name)
    if name == "B200WhisperEncoderStem":
        from . import encoder_stem
        return encoder_stem.B200WhisperEncoderStem
    if name in ("ops", "signals", "whisper", "urban", "collate", "encoder_stem"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
