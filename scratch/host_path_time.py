"""Drop-in call with the reference's own argument shape: a list of float64 numpy arrays (what `datasets` yields)."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import B200WhisperFeatureExtractor, signals
fe = B200WhisperFeatureExtractor(device="cuda")
B = 64
clips64 = [signals.whisper_clip(i, seed=1).astype(np.float64) for i in range(B)]
clips32 = [c.astype(np.float32) for c in clips64]
for name, clips in (("float64 list", clips64), ("float32 list", clips32)):
    for _ in range(6): fe(clips, sampling_rate=16000, return_tensors="pt"); torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 20
    for _ in range(n): out = fe(clips, sampling_rate=16000, return_tensors="pt").input_features
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{name}: {dt*1e3:.1f} ms per {B}-clip call = {B/dt:.0f} clips/s")
# per-clip calls, exactly like REF:whisper_finetune/dataset.py:58-62
t0 = time.perf_counter()
for c in clips64[:16]: fe(c, sampling_rate=16000, return_tensors="pt").input_features
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 16
print(f"one float64 clip per call: {dt*1e3:.2f} ms per call = {1/dt:.0f} clips/s")
