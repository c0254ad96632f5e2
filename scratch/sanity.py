import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
from audio_transformers_b200.urban import B200UrbanFrontEnd
# whisper: ragged batch incl. edge lengths, TMA + generic tiles
clips = [signals.whisper_clip(i, seed=3, n_samples=n) for i, n in enumerate((480000, 1, 161, 20700, 333333, 600000))]
width = (max(len(c) for c in clips) + 3) // 4 * 4
host = np.zeros((len(clips), width), np.float32)
for i, c in enumerate(clips): host[i, :len(c)] = c
out = ops.whisper_logmel(torch.from_numpy(host).cuda(), torch.tensor([len(c) for c in clips], dtype=torch.int32).cuda())
m = ops.whisper_frame_mask(torch.tensor([len(c) for c in clips], dtype=torch.int32).cuda())
# urban mel + prep
y = ops.mel_power(torch.randn(3, 88200).cuda(), 1e-9)
fe = B200UrbanFrontEnd()
names = signals.URBAN_PREP_CASES[:4]
z = fe.process_batch([signals.urban_raw_clip(*c) for c in names], [c[1] for c in names])
torch.cuda.synchronize()
print("ok", out.shape, m.shape, y.shape, z.shape, float(out.abs().max()), float(z.abs().max()))
