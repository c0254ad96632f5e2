"""Where the time of a one-clip-per-call drop-in call goes (REF:whisper_finetune/dataset.py:58-62 pattern)."""
import sys, time, torch, numpy as np, cProfile, pstats, ctypes
sys.path.insert(0, '.')
from audio_transformers_b200 import B200WhisperFeatureExtractor, signals, _lib
fe = B200WhisperFeatureExtractor(device="cuda")
clips = [signals.whisper_clip(i, seed=1).astype(np.float64) for i in range(32)]
for c in clips[:8]: fe(c, sampling_rate=16000, return_tensors="pt")
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for c in clips: fe(c, sampling_rate=16000, return_tensors="pt").input_features
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / len(clips)
    print(f"one float64 clip per call: {dt*1e3:.3f} ms per call = {1/dt:.0f} clips/s")
lib = _lib.load()
dst = torch.empty(480000, dtype=torch.float32, pin_memory=True); lens = np.array([480000], dtype=np.int64); out = np.zeros(1, dtype=np.int32)
for nt in (1, 2, 4, 8, 16):
    ptrs = (ctypes.c_void_p * 1)(clips[0].ctypes.data)
    t0 = time.perf_counter()
    for _ in range(200):
        lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), 1, 1, 480000, ctypes.c_void_p(dst.data_ptr()), 480000, out.ctypes.data_as(ctypes.c_void_p), nt)
    print(f"pack one clip, {nt} threads: {(time.perf_counter()-t0)/200*1e6:.0f} us")
pr = cProfile.Profile(); pr.enable()
for c in clips: fe(c, sampling_rate=16000, return_tensors="pt")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
