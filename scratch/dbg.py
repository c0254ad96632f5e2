import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
sys.path.insert(0, 'oracle')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = torch.from_numpy(signals.whisper_batch(B, seed=0)).cuda()
try:
    y = ops.whisper_logmel(x, None); torch.cuda.synchronize()
    print("ran ok", y.shape, float(y.abs().max()))
    from oracle import logmel_oracle as o
    ref = o.whisper_logmel([r for r in x.cpu().numpy()])
    if ref is not None: print("max abs err", float(np.abs(ref - y.cpu().numpy()).max()))
except Exception as e:
    print("FAILED:", str(e).splitlines()[0])
