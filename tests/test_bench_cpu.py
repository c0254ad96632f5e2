"""bench.py on a machine without a GPU: the reference arm must run (it is the reference's CPU path) and print the
contract's JSON line; the own arm must refuse loudly instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "4", env={"OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "whisper_logmel_30s_clips_per_sec" and rec["unit"] == "clips/s"
    assert rec["higher_is_better"] is True and rec["value"] > 0 and rec["gpu_launches"] == 0
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still take every core it may run on
    assert rec["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_own_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the own arm runs (covered by the gpu tests and the driver)")
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "CUDA" in (res.stderr + res.stdout)
