import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
pools = [base.repeat(B // 8, 1).contiguous() * (1.0 + 0.01 * i) for i in range(3)]
for p in pools: ops.whisper_logmel(p, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 30
e0.record()
for i in range(iters): out = ops.whisper_logmel(pools[i % 3], None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
clips_s = B / (ms * 1e-3)
print(f"B={B} ms/step={ms:.4f} clips/s={clips_s:.0f} frac_hbm={clips_s*2.88e6/6538.9e9:.3f}")
