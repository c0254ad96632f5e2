import sys, torch
sys.path.insert(0, '.')
from audio_transformers_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randn(B, 88200, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3): y = ops.mel_power(x, 1e-9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): y = ops.mel_power(x, 1e-9)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print(f"urban B={B}: {ms*1e3:.1f} us/step  {B/(ms*1e-3):,.0f} clips/s  per-tile {ms*1e3/(B*173/32/148):.2f} us")
