#!/usr/bin/env python
"""What limits the end-to-end number when several ranks share one host?  

    python tools/host_limits.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_limits.py

Every rank runs each leg AT THE SAME TIME as the others (barrier before and after), and the slowest rank's time counts:
  h2d     123 MB pinned float32 -> device, 20 copies           (PCIe in, host memory read)
  d2h     61 MB device -> pinned, 20 copies                     (PCIe out, host memory write)
  cast    64 float64 clips -> pinned float32 (b200mel_host_pack, no CUDA), 10 passes   (host memory bandwidth, cores)
  step    the bench's e2e step: extractor on a list of 64 float64 clips + D2H of the features, two streams, 20 steps
One JSON line on stdout (rank 0): aggregate GB/s of every leg and the per-rank thread count.
"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from audio_transformers_b200 import B200WhisperFeatureExtractor, _lib, signals  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def timed(fn):
        fn()                                           # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = 64
    fe = B200WhisperFeatureExtractor(device=dev)
    clips = [signals.whisper_clip(i, seed=rank).astype(np.float64) for i in range(B)]
    host_in = torch.empty((B, 480000), dtype=torch.float32).pin_memory()
    host_out = [torch.empty((B, 80, 3000), dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_in = torch.empty((B, 480000), dtype=torch.float32, device=dev)
    dev_out = torch.zeros((B, 80, 3000), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ptrs = (ctypes.c_void_p * B)(*[c.ctypes.data for c in clips])
    lens = np.full(B, 480000, dtype=np.int64)
    res = {"n_gpus": world, "cpus_visible": len(os.sched_getaffinity(0)), "pack_threads_per_rank": fe._pack_threads}

    t = timed(lambda: [dev_in.copy_(host_in, non_blocking=True) for _ in range(20)])
    res["h2d_gbs_total"] = world * 20 * host_in.numel() * 4 / t / 1e9
    t = timed(lambda: [host_out[0].copy_(dev_out, non_blocking=True) for _ in range(20)])
    res["d2h_gbs_total"] = world * 20 * dev_out.numel() * 4 / t / 1e9

    def cast():
        for _ in range(10):
            lib.b200mel_host_pack(ptrs, lens.ctypes.data_as(ctypes.c_void_p), B, 1, 480000, ctypes.c_void_p(host_in.data_ptr()),
                                  480000, None, fe._pack_threads)
    t = timed(cast)
    res["cast_clips_per_s_total"] = world * 10 * B / t
    res["cast_gbs_total"] = world * 10 * B * 480000 * 12 / t / 1e9        # 8 bytes read + 4 written per sample

    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def steps():
        for k in range(20):
            with torch.cuda.stream(streams[k % 2]):
                f = fe(clips, sampling_rate=16000, return_tensors="pt").input_features
                host_out[k % 2].copy_(f, non_blocking=True)
    t = timed(steps)
    res["e2e_clips_per_s_total"] = world * 20 * B / t
    res["e2e_host_memory_gbs_total"] = world * 20 * B * (480000 * (8 + 4 + 4) + 240000 * 4) / t / 1e9   # cast read + write, DMA read, D2H write
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
