import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = 64
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
pools = [base.repeat(B // 8, 1).contiguous() * (1.0 + 0.01 * i) for i in range(4)]
for p in pools: ops.whisper_logmel(p, None)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(2000): out = ops.whisper_logmel(pools[i % 4], None)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e6*(t1-t0)/2000:.1f} us/call (CPU), total {1e6*(t2-t0)/2000:.1f} us/call")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(2000): out = ops.whisper_logmel(pools[i % 4], None)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(8)
