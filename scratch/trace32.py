import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
x = base.repeat(B // 8, 1).contiguous()
for _ in range(3): ops.whisper_logmel(x, None)
tr = torch.zeros(32 * 4 * 16, dtype=torch.int64, device='cuda')
lib = _lib.load()
lib.b200mel_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.b200mel_debug_set_trace(ctypes.c_void_p(tr.data_ptr())) == 0
ops.whisper_logmel(x, None); torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(32, 4, 16)[:, :, :8]
t0 = t[0, 0].min()
names = ['top(after barrier)', 'phase A end', 'after mid barrier', 'phase B end']
for it in range(4, 9):
    base_t = t[it, 0].min()
    print(f"iter {it}: start +{base_t - t[it-1,0].min()} cyc since previous start")
    for k in range(4):
        row = t[it, k] - base_t
        print(f"   {names[k]:20s} min {row.min():6d} max {row.max():6d}  per-warp: " + ' '.join(f"{v:5d}" for v in row))
