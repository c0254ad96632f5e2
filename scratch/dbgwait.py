import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
x = base.repeat(B // 8, 1).contiguous()
for _ in range(3): ops.whisper_logmel(x, None)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 4)()
lib.b200mel_debug_read(buf); a = list(buf)
ops.whisper_logmel(x, None); torch.cuda.synchronize()
lib.b200mel_debug_read(buf); b = list(buf)
n = b[1] - a[1]
print(f"tiles {n}  avg TMA wait (thread 0) {(b[0]-a[0])/n:.0f} cyc  avg top-barrier wait {(b[2]-a[2])/n:.0f} cyc")
