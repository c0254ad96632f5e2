"""Where the time of the float64-list call goes at N = 1 (run on the GPU box)."""
import sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from audio_transformers_b200 import B200WhisperFeatureExtractor, _lib, signals, ops
B = 64
dev = torch.device("cuda", 0)
clips = [[signals.whisper_clip(i, seed=p).astype(np.float64) for i in range(B)] for p in range(2)]
lib = _lib.load()
host_in = torch.empty((B, 480000), dtype=torch.float32).pin_memory()
dev_in = torch.empty((B, 480000), dtype=torch.float32, device=dev)
host_out = [torch.empty((B, 80, 3000), dtype=torch.float32).pin_memory() for _ in range(2)]
ptrs = (ctypes.c_void_p * B)(*[c.ctypes.data for c in clips[0]])
lens = np.full(B, 480000, np.int64)

def wall(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

for th in (8, 16, 24, 32):
    ms = wall(lambda: lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), B, 1, 480000, ctypes.c_void_p(host_in.data_ptr()), 480000, None, th))
    print(f"cast only, {th:2d} threads: {ms:.2f} ms ({B / ms:.1f} k clips/s)")
ms = wall(lambda: dev_in.copy_(host_in, non_blocking=True))
print(f"H2D one copy of 123 MB: {ms:.2f} ms ({host_in.numel() * 4 / ms / 1e6:.1f} GB/s)")
flat_h, flat_d = host_in.view(-1), dev_in.view(-1)
for piece in (1 << 17, 1 << 19, 1 << 21):
    def pieces():
        for o in range(0, flat_h.numel(), piece):
            flat_d[o:o + piece].copy_(flat_h[o:o + piece], non_blocking=True)
    ms = wall(pieces, 5)
    print(f"H2D in pieces of {piece * 4 >> 10} KB (torch copies, python loop): {ms:.2f} ms")
feats = ops.whisper_logmel(dev_in, None)
ms = wall(lambda: host_out[0].copy_(feats, non_blocking=True))
print(f"D2H 61 MB: {ms:.2f} ms")

fe = B200WhisperFeatureExtractor(device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
def run(nsteps, d2h, two_streams=True):
    t_calls = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(nsteps):
        s = streams[k % 2 if two_streams else 0]
        with torch.cuda.stream(s):
            ta = time.perf_counter()
            f = fe(clips[k % 2], sampling_rate=16000, return_tensors="pt").input_features
            t_calls.append(time.perf_counter() - ta)
            if d2h: host_out[k % 2].copy_(f, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t2 - t0) / nsteps * 1e3, (t1 - t0) / nsteps * 1e3, float(np.median(t_calls)) * 1e3
for th in (8, 16, 24, 32):
    fe._pack_threads = th
    run(4, True)
    for d2h, two in ((False, True), (True, True), (True, False)):
        tot, host, call = run(30, d2h, two)
        print(f"threads {th:2d} d2h={d2h} two_streams={two}: {tot:.2f} ms per step ({B / tot:.1f} k clips/s); host loop {host:.2f} ms per step; the call itself {call:.2f} ms")
