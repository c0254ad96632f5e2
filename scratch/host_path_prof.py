import sys, time, torch, numpy as np, cProfile, pstats
sys.path.insert(0, '.')
from audio_transformers_b200 import B200WhisperFeatureExtractor, signals
fe = B200WhisperFeatureExtractor(device="cuda")
clips64 = [signals.whisper_clip(i, seed=1).astype(np.float64) for i in range(64)]
for _ in range(3): fe(clips64, sampling_rate=16000, return_tensors="pt"); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): fe(clips64, sampling_rate=16000, return_tensors="pt"); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
