import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
from oracle import logmel_oracle as O
B = 1500
rng = np.random.default_rng(5)
base = signals.whisper_batch(12, seed=9)
idx = rng.integers(0, 12, size=B)
lens = rng.integers(1, 480001, size=B).astype(np.int32)
lens[::7] = 480000
wave = torch.from_numpy(base).cuda()[torch.from_numpy(idx).cuda()].contiguous()
out = ops.whisper_logmel(wave, torch.from_numpy(lens).cuda())
torch.cuda.synchronize()
worst = 0.0
for b in (0, 1, 7, 733, 1499, int(np.argmin(lens)), int(np.argmax(lens))):
    ref = O.whisper_logmel([base[idx[b]][:lens[b]]])[0]
    worst = max(worst, float(np.abs(out[b].cpu().numpy() - ref).max()))
print("B=1500 ragged: max-abs vs oracle on 7 clips:", worst, " finite:", bool(torch.isfinite(out).all()))
assert worst <= 1e-4
