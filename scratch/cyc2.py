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
cyc = t[1024:1024 + 296].reshape(148, 2).astype(np.float64); cnt = t[512:512 + 296].reshape(148, 2)
print(f"kernel {ms/n*1e3:.1f} us = {ms/n*1e-3*1.965e9:.0f} cycles at 1965 MHz")
print("half 0: cycles mean %.0f, tiles mean %.1f;  half 1: cycles mean %.0f, tiles mean %.1f" % (cyc[:,0].mean(), cnt[:,0].mean(), cyc[:,1].mean(), cnt[:,1].mean()))
print("per-CTA max cycles: min %.0f median %.0f max %.0f" % (cyc.max(1).min(), np.median(cyc.max(1)), cyc.max(1).max()))
