"""Build libb200mel.so in-tree with nvcc for sm_100a.

    python -m audio_transformers_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "b200mel.cu")
OUT = os.path.join(HERE, "libb200mel.so")
DEPS = ([SRC, os.path.join(HERE, "csrc", "generated", "tables.inc"), os.path.join(os.path.dirname(HERE), "include", "b200mel.h")]
        + sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libb200mel.so")
    return p


STAMP = OUT + ".sources.sha256"


def sources_digest() -> str:
    """sha256 over the flags and every source the library is built from (names and bytes)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        h.update(os.path.relpath(d, HERE).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def up_to_date() -> bool:
    """True when the library on disk was built from exactly the present sources (a content hash written next to it:
    modification times say nothing about a snapshot that was copied to another box)."""
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == sources_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}")
    with open(STAMP, "w") as f:
        f.write(sources_digest() + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
