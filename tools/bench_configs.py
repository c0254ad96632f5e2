#!/usr/bin/env python
"""Secondary measurements for the workloads BASELINE.json lists besides the headline one.

    python tools/bench_configs.py [--out profiles/r01_configs.json]

config 1  urban 64-mel MelSpectrogram + log on 32 x 4 s clips: GPU (device-timed) and the torchaudio CPU path
config 3  Whisper features for 256 clips, alone and followed by the Whisper-tiny encoder (seeded random init, HF
          implementation as the consumer) in fp32 and bf16
config 4  variable-length batch (lengths ~ U{16000..480000}), 512 clips
config 5  1k .. 64k clips through one GPU in 512-clip chunks taken from a device-resident pool

One JSON object per line on stdout, all of them also written to --out.  Timing: CUDA events on the launch stream,
3 warm-up passes, inputs rotate over pools larger than L2 where the workload allows it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from audio_transformers_b200 import ops, signals  # noqa: E402

PEAK = 6538.9e9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) * 1e9
except Exception:
    pass


def dev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def config1():
    wave = torch.from_numpy(signals.urban_batch(32, seed=0))                       # (32, 1, 88200)
    dev = [wave[:, 0].cuda() * (1.0 - 0.01 * i) for i in range(4)]
    k = [0]

    # 27 us of GPU work per call is below the host cost of an eager op call: replay one captured graph per input
    for x in dev:
        ops.mel_power(x, 1e-9)
    torch.cuda.synchronize()
    graphs, keep = [], []
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for x in dev:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep.append(ops.mel_power(x, 1e-9))
            graphs.append(g)
    torch.cuda.synchronize()

    def step():
        graphs[k[0] % 4].replay()
        k[0] += 1
    t = dev_time(step, 400, warm=20)
    big = torch.randn(2048, 88200, device="cuda")
    tb = dev_time(lambda: ops.mel_power(big, 1e-9), 20)
    res = {"config": "1: urban 64-mel log-mel, 32 x 4 s clips", "gpu_s_per_batch": t, "gpu_clips_per_s": 32 / t,
           "launch": "cuda graph replay of the public op, 4 rotating inputs",
           "gpu_clips_per_s_batch2048": 2048 / tb, "hbm_frac_batch2048": 2048 / tb * 397088 / PEAK}
    try:
        import torchaudio
        tf = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64)
        with torch.no_grad():
            torch.log(tf(wave) + 1e-9)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 3.0:
                torch.log(tf(wave) + 1e-9)
                n += 1
        tc = (time.perf_counter() - t0) / n
        res.update(cpu_s_per_batch=tc, cpu_clips_per_s=32 / tc, cpu_threads=torch.get_num_threads(),
                   cpu_impl=f"torchaudio {torchaudio.__version__} MelSpectrogram + log, batched")
    except Exception as exc:  # pragma: no cover
        res["cpu_impl"] = f"unavailable: {type(exc).__name__}"
    return res


def config3():
    B = 256
    base = torch.from_numpy(signals.whisper_batch(16, seed=3)).cuda()
    wave = base.repeat(B // 16, 1).contiguous()
    t_fe = dev_time(lambda: ops.whisper_logmel(wave, None), 20)
    res = {"config": "3: Whisper features + encoder, batch 256", "features_s": t_fe, "features_clips_per_s": B / t_fe}
    try:
        from transformers import WhisperConfig, WhisperModel
        torch.manual_seed(0)
        enc = WhisperModel(WhisperConfig()).encoder.eval().cuda()
        with torch.no_grad():
            for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
                e = enc.to(dt)

                def step():
                    f = ops.whisper_logmel(wave, None)
                    for i in range(0, B, 64):
                        e(f[i:i + 64].to(dt))
                t = dev_time(step, 3, warm=1)
                res[f"features_plus_encoder_{name}_s"] = t
                res[f"features_plus_encoder_{name}_clips_per_s"] = B / t
                res[f"features_share_{name}"] = t_fe / t
    except Exception as exc:  # pragma: no cover
        res["encoder"] = f"unavailable: {type(exc).__name__}: {exc}"
    return res


def config4():
    B = 512
    rng = np.random.default_rng(1)
    lens = rng.integers(16000, 480001, size=B).astype(np.int32)
    base = torch.from_numpy(signals.whisper_batch(16, seed=4)).cuda()
    wave = base.repeat(B // 16, 1).contiguous()
    dl = torch.from_numpy(lens).cuda()
    t = dev_time(lambda: ops.whisper_logmel(wave, dl), 20)
    t_full = dev_time(lambda: ops.whisper_logmel(wave, None), 20)
    byt = float((4 * lens.astype(np.int64) + 960000).sum())
    return {"config": "4: variable-length clips (16000..480000 samples), batch 512", "s_per_batch": t, "clips_per_s": B / t,
            "full_length_clips_per_s": B / t_full, "algorithmic_GBps": byt / t / 1e9, "hbm_frac": byt / t / PEAK,
            "mean_len_s": float(lens.mean() / 16000)}


def config5():
    chunk = 512
    base = torch.from_numpy(signals.whisper_batch(16, seed=5)).cuda()
    pool = [base.repeat(chunk // 16, 1).contiguous() * (1.0 - 0.02 * i) for i in range(3)]       # 3 x 983 MB
    out = []
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        k = [0]

        def run():
            for _ in range(n // chunk):
                ops.whisper_logmel(pool[k[0] % 3], None)
                k[0] += 1
        t = dev_time(run, 1 if n >= 16384 else 3, warm=1)
        out.append({"clips": n, "s": t, "clips_per_s": n / t, "hbm_frac": n / t * 2880000 / PEAK})
    return {"config": "5: 1k..64k 30 s clips through one GPU in 512-clip chunks", "n_gpus": 1, "sweep": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_configs.json"))
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device (no CPU fallback)"
    results = []
    for fn in (config1, config3, config4, config5):
        r = fn()
        r["gpu"] = torch.cuda.get_device_name(0)
        print(json.dumps(r), flush=True)
        results.append(r)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
