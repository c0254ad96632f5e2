"""Bring-up driver for the encoder stem: per-stage comparison against torch (run under `timeout`)."""
import sys, torch
sys.path.insert(0, '.')
import transformers as tr
from audio_transformers_b200 import B200WhisperEncoderStem, ops, signals
import numpy as np
torch.manual_seed(99)
enc = tr.WhisperModel(tr.WhisperConfig()).encoder.eval().cuda()
stem = B200WhisperEncoderStem.from_encoder(enc).cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wave = torch.from_numpy(signals.whisper_batch(max(B, 2), seed=3)[:B]).cuda()
feats = ops.whisper_logmel(wave, None)
torch.cuda.synchronize()
print("features ok", feats.shape, flush=True)
out = stem(feats)
torch.cuda.synchronize()
print("stem ran", out.shape, float(out.abs().max()), flush=True)
F = torch.nn.functional
bf = lambda t: t.to(torch.bfloat16).float()
with torch.no_grad():
    x1 = F.gelu(F.conv1d(bf(feats), bf(enc.conv1.weight), enc.conv1.bias, padding=1))
    x2 = F.gelu(F.conv1d(bf(x1), bf(enc.conv2.weight), enc.conv2.bias, stride=2, padding=1))
    ref = x2.permute(0, 2, 1) + enc.embed_positions.weight
d = (out - ref).abs()
print("max-abs", float(d.max()), "rel-fro", float((out - ref).norm() / ref.norm()), "ref max", float(ref.abs().max()))
bad = (d > 5e-3).nonzero()
print("bad count", bad.shape[0], "first", bad[:8].tolist())
if bad.shape[0]:
    # where do the bad entries sit (time tile / channel block)?
    print("bad rows mod 128:", torch.unique(bad[:, 1] % 128)[:20].tolist(), " channels // 32:", torch.unique(bad[:, 2] // 32).tolist())
if len(sys.argv) > 2:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): stem(feats)
    e0.record()
    for _ in range(10): stem(feats)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={B}: {ms*1e3:.1f} us per call, {B*1.88e9/ms/1e9:.1f} TFLOP/s")
if len(sys.argv) > 3:
    def lib_stem(x, dt):
        with torch.no_grad():
            y = F.gelu(F.conv1d(x.to(dt), enc.conv1.weight.to(dt), enc.conv1.bias.to(dt), padding=1))
            y = F.gelu(F.conv1d(y, enc.conv2.weight.to(dt), enc.conv2.bias.to(dt), stride=2, padding=1))
            return y.permute(0, 2, 1) + enc.embed_positions.weight.to(dt)
    for dt, tf32 in ((torch.float32, False), (torch.float32, True), (torch.bfloat16, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(3): lib_stem(feats, dt)
        e0.record()
        for _ in range(10): lib_stem(feats, dt)
        e1.record(); torch.cuda.synchronize()
        print(f"torch/cuDNN stem {dt} tf32={tf32}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")
