"""Graph-replay step time of the Whisper op like bench.py measures it (4 input pools > L2, one graph each).
usage: [B200MEL_LIB=scratch/variants/x.so] python scratch/graph_time.py [batch] [steps]"""
import sys, os, statistics, torch
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
KIND = sys.argv[3] if len(sys.argv) > 3 else None
import numpy as np
if KIND:
    base = [torch.from_numpy(np.stack([signals.whisper_clip(i, seed=p, kind=KIND) for i in range(8)])).cuda() for p in range(4)]
else:
    base = [torch.from_numpy(signals.whisper_batch(min(B, 64), seed=p)).cuda() for p in range(4)]
pools = [b.repeat(B // b.shape[0], 1).contiguous() for b in base]
s = torch.cuda.Stream()
graphs = []
with torch.cuda.stream(s):
    for p in pools:
        for _ in range(2): out = ops.whisper_logmel(p, None)
    torch.cuda.synchronize()
    for p in pools:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            o = ops.whisper_logmel(p, None)
        graphs.append((g, o))
    for i in range(8): graphs[i % 4][0].replay()
    torch.cuda.synchronize()
    ops.profile_begin(pools[0].device, max_launches=64)
    for i in range(64): ops.whisper_logmel(pools[i % 4], None)
    torch.cuda.synchronize()
    kms, n = ops.profile_end(pools[0].device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for blk in range(20):
        for i in range(K // 20): graphs[i % 4][0].replay()
        ev[blk + 1].record()
    torch.cuda.synchronize()
per = [ev[i].elapsed_time(ev[i + 1]) / (K // 20) * 1e3 for i in range(20)]
print(f"{os.environ.get('B200MEL_LIB', 'default'):40s} B={B} {KIND or 'mixed'} step median {statistics.median(per):.2f} us best {min(per):.2f} us  main kernel {kms / n * 1e3:.2f} us  rest {statistics.median(per) - kms / n * 1e3:.2f} us  checksum {float(graphs[0][1].sum()):.4f}")
