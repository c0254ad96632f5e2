import sys, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops
for B in (32, 256, 2048):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 88200, generator=g).cuda()
    for _ in range(3): y = ops.mel_power(x, 1e-9)
    torch.cuda.synchronize()
    iters = 50
    ops.profile_begin(x.device, max_launches=iters, preset=1) if 'preset' in ops.profile_begin.__code__.co_varnames else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): y = ops.mel_power(x, 1e-9)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    cps = B / (ms * 1e-3)
    print(f"urban B={B}: {ms*1e3:.1f} us/step  {cps:,.0f} clips/s  HBM frac {cps*397088/6538.9e9:.3f}")
