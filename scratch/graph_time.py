import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
pools = [base.repeat(B // 8, 1).contiguous() * (1.0 + 0.01 * i) for i in range(4)]
for p in pools: ops.whisper_logmel(p, None)
torch.cuda.synchronize()
graphs = []
for p in pools:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o = ops.whisper_logmel(p, None)
    graphs.append((g, o))
for g, _ in graphs: g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 400
e0.record()
for i in range(iters): graphs[i % 4][0].replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"graph replay B={B} ms/step={ms:.4f} clips/s={B/(ms*1e-3):.0f}")
e0.record()
for i in range(iters): ops.whisper_logmel(pools[i % 4], None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"eager        B={B} ms/step={ms:.4f} clips/s={B/(ms*1e-3):.0f}")
