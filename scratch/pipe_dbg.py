import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
wave = base.repeat(B // 8, 1).contiguous()
for i in range(3):
    out = ops.whisper_logmel(wave, None)
torch.cuda.synchronize()
ws = list(ops._workspaces.values())[0].view(torch.int64).cpu().numpy()
d = ws[B + 2: B + 2 + 8 * 2 * 148].reshape(296, 8)
print("per floor warp: total cycles, gate, poll, spins, items, wait, clamp")
print("mean", d.mean(0)[:7].astype(int).tolist())
print("max ", d.max(0)[:7].tolist())
print("min ", d.min(0)[:7].tolist())
