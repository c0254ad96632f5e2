import sys, torch
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
wave = base.repeat(B // 8, 1).contiguous()
for i in range(4):
    out = ops.whisper_logmel(wave, None)
torch.cuda.synchronize()
print("ok", float(out.sum()))
