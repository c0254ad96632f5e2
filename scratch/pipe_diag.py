import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
from oracle import logmel_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
clips = [signals.whisper_clip(i, seed=5) for i in range(B)]
wave = torch.from_numpy(np.stack(clips)).cuda()
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter()
    out = ops.whisper_logmel(wave, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ws = list(ops._workspaces.values())[0].view(torch.int64).cpu().numpy()
    print(f"rep {rep}: {dt*1e3:.3f} ms; workspace nonzero words: {np.nonzero(ws)[0].tolist()[:10]} tail={ws[B]:#x}")
ref = O.whisper_logmel(clips[:4], dtype=np.float32)
o = out[:4].cpu().numpy()
print("max-abs on 4 clips:", [float(np.abs(o[i]-ref[i]).max()) for i in range(4)])
