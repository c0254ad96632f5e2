"""Where does the host time of one extractor call on one float64 clip go?"""
import sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from audio_transformers_b200 import B200WhisperFeatureExtractor, ops, signals, _lib
from audio_transformers_b200._lib import PRESET_WHISPER

dev = torch.device("cuda", 0)
fe = B200WhisperFeatureExtractor(device=dev)
clips = [signals.whisper_clip(i, seed=3).astype(np.float64) for i in range(8)]
for c in clips: fe(c, sampling_rate=16000, return_tensors="pt")
torch.cuda.synchronize()

def bench(fn, n=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6

print("full call            host %.1f us  incl. drain %.1f us" % bench(lambda i: fe(clips[i % 8], sampling_rate=16000, return_tensors="pt").input_features.squeeze(0)))
print("canonicalise         %.1f us" % bench(lambda i: fe._canonicalise(clips[i % 8]))[0])
print("_features_from_host  host %.1f us  incl. drain %.1f us" % bench(lambda i: fe._features_from_host([clips[i % 8]], dev)))
lib = _lib.load(); h = ops._handle(dev, PRESET_WHISPER)
stream = torch.cuda.current_stream(dev)
ws = ops._whisper_workspace(lib, h, dev, stream.cuda_stream, 1)
slot = fe._slot(dev, 1, 480000)
feats = torch.empty((1, 80, 3000), dtype=torch.float32, device=dev)
for threads in (1, 2, 4, 8, 16):
    def native(i, threads=threads):
        c = clips[i % 8]
        ptrs = np.array([c.__array_interface__["data"][0]], dtype=np.uint64); lens = np.array([480000], dtype=np.int64); f = np.array([1], dtype=np.uint8)
        lib.b200mel_whisper_logmel_host(h, ptrs.ctypes.data, lens.ctypes.data, f.ctypes.data, 1, slot["host"].data_ptr(), 480000,
                                        slot["lens"].data_ptr(), slot["dev"].data_ptr(), slot["dev_lens"].data_ptr(), feats.data_ptr(),
                                        ws.data_ptr(), ws.numel(), threads, stream.cuda_stream)
    print("native call, %2d threads  host %.1f us  incl. drain %.1f us" % ((threads,) + bench(native)))
def kernel_only(i):
    lib.b200mel_whisper_logmel_f32(h, slot["dev"].data_ptr(), 480000, slot["dev_lens"].data_ptr(), 1, feats.data_ptr(), ws.data_ptr(), ws.numel(), stream.cuda_stream)
print("kernel launch only   host %.1f us  incl. drain %.1f us" % bench(kernel_only))
print("torch.empty feats    %.1f us" % bench(lambda i: torch.empty((1, 80, 3000), dtype=torch.float32, device=dev))[0])
ev = torch.cuda.Event()
print("event record         %.1f us" % bench(lambda i: ev.record(stream))[0])
from audio_transformers_b200.whisper import _batch_feature
print("BatchFeature         %.1f us" % bench(lambda i: _batch_feature({"input_features": feats}))[0])
# list of 64
clips64 = [signals.whisper_clip(i, seed=4).astype(np.float64) for i in range(64)]
print("list of 64 f64       host %.1f us  incl. drain %.1f us" % bench(lambda i: fe(clips64, sampling_rate=16000, return_tensors="pt"), 20))
for threads in (4, 8, 16, 32):
    fe._pack_threads = threads
    print("  pack threads %2d     host %.1f us  incl. drain %.1f us" % ((threads,) + bench(lambda i: fe(clips64, sampling_rate=16000, return_tensors="pt"), 20)))
import os; print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
