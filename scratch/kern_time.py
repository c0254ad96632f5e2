import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
pools = [base.repeat(B // 8, 1).contiguous() * (1.0 + 0.01 * i) for i in range(3)]
for p in pools: ops.whisper_logmel(p, None)
torch.cuda.synchronize()
iters = 30
ops.profile_begin(pools[0].device, max_launches=iters)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters): out = ops.whisper_logmel(pools[i % 3], None)
e1.record(); torch.cuda.synchronize()
kms, n = ops.profile_end(pools[0].device)
ms = e0.elapsed_time(e1) / iters
print(f"B={B} step={ms:.4f} ms  main kernel={kms/n:.4f} ms  rest(memset+clamp+gaps)={ms-kms/n:.4f} ms  kernel-only clips/s={B/(kms/n*1e-3):.0f} frac={B/(kms/n*1e-3)*2.88e6/6538.9e9:.3f}")
