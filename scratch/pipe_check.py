"""Quick parity + timing probe of the Whisper kernel (development driver, not a test)."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from audio_transformers_b200 import ops, signals
from oracle import logmel_oracle as O

def parity():
    lens = [480000, 192000, 80000, 1, 159, 160, 161, 399, 400, 401, 479999, 5000, 4920, 4921, 0]
    clips = [signals.whisper_clip(i, seed=3, n_samples=max(L, 1))[:L] for i, L in enumerate(lens)]
    width = 480000
    host = np.zeros((len(clips), width), np.float32)
    for i, c in enumerate(clips):
        host[i, :len(c)] = c
    wave = torch.from_numpy(host).cuda()
    ln = torch.tensor(lens, dtype=torch.int32, device='cuda')
    for rep in range(3):
        out = ops.whisper_logmel(wave, ln)
    torch.cuda.synchronize()
    ref = O.whisper_logmel([c if len(c) else np.zeros(0, np.float32) for c in clips], dtype=np.float32)
    o = out.cpu().numpy()
    for i, L in enumerate(lens):
        print(f"L={L:7d} max-abs {np.abs(o[i] - ref[i]).max():.3e} finite {np.isfinite(o[i]).all()}")
    err = float(np.abs(o - ref).max())
    print("parity max-abs", err)
    return err

def timing(B):
    base = torch.from_numpy(signals.whisper_batch(8, seed=0)).cuda()
    pools = [base.repeat(B // 8, 1).contiguous() * (1.0 + 0.01 * i) for i in range(3)]
    for p in pools: ops.whisper_logmel(p, None)
    torch.cuda.synchronize()
    iters = 30
    ops.profile_begin(pools[0].device, max_launches=iters)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): out = ops.whisper_logmel(pools[i % 3], None)
    e1.record(); torch.cuda.synchronize()
    kms, n = ops.profile_end(pools[0].device)
    ms = e0.elapsed_time(e1) / iters
    ws = [w for k, w in ops._workspaces.items() if k[2] == B][0].view(torch.int64).cpu().numpy()
    print(f"   ws tail={int(ws[B]):#x} nonzero={np.nonzero(ws[:B])[0].tolist()[:8]}")
    print(f"B={B} step={ms:.4f} ms  kernel={kms/n:.4f} ms  clips/s(kernel)={B/(kms/n*1e-3):.0f} frac={B/(kms/n*1e-3)*2.88e6/6538.9e9:.3f}", flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        e = parity()
    if what in ("all", "time"):
        for B in (64, 512):
            timing(B)
