"""B200-native log-mel spectrogram front end (drop-in for the two call sites of
k0r1g/audio-transformers).  See DESIGN.md."""
__version__ = "0.1.0"
