import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals, _lib
B = 512
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
x = base.repeat(B // 8, 1).contiguous()
for _ in range(3): ops.whisper_logmel(x, None)
tr = torch.zeros(32 * 4 * 16, dtype=torch.int64, device='cuda')
lib = _lib.load()
lib.b200mel_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.b200mel_debug_set_trace(ctypes.c_void_p(tr.data_ptr())) == 0
ops.profile_begin(x.device, max_launches=4)
ops.whisper_logmel(x, None); torch.cuda.synchronize()
ms, n = ops.profile_end(x.device)
t = tr.cpu().numpy()
cyc = t[1024:1024 + 296].astype(np.float64); sm = t[512:512 + 296]
print(f"kernel {ms/n*1e3:.1f} us; per-CTA cycles: min {cyc.min():.0f} median {np.median(cyc):.0f} max {cyc.max():.0f}")
first, second = cyc[:148], cyc[148:]
print(f"first-wave CTAs (0..147): mean {first.mean():.0f}  second (148..295): mean {second.mean():.0f}")
# pair CTAs by SM
by_sm = {}
for b in range(296): by_sm.setdefault(int(sm[b]), []).append((b, cyc[b]))
pairs = [v for v in by_sm.values()]
print("CTAs per SM histogram:", np.bincount([len(v) for v in pairs]))
slow = sorted(((max(c for _, c in v), k) for k, v in by_sm.items()), reverse=True)[:5]
fast = sorted(((max(c for _, c in v), k) for k, v in by_sm.items()))[:5]
print("slowest SMs:", [(k, int(c), [b for b, _ in by_sm[k]]) for c, k in slow])
print("fastest SMs:", [(k, int(c), [b for b, _ in by_sm[k]]) for c, k in fast])
