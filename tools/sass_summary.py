#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (static instruction counts).

    python tools/sass_summary.py [> profiles/rNN_sass_summary.txt]

Which Blackwell-specific instructions the kernels contain (B200_PROFILING.md, "What proves a Blackwell-native
kernel"): UTMALDG / UTMAPF / UBLKCP / UBLKPF (TMA and bulk copies / prefetches), SYNCS (mbarrier), the packed FP32
pipe (FFMA2 / FADD2 / FMUL2), MUFU.LG2, REDUX, and -- absent by design, there is no dense contraction on this path --
UTC*MMA / LDTM / STTM."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_transformers_b200", "libb200mel.so")
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU.LG2",
        "MUFU", "REDUX", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "USETMAXREG", "UTC", "LDTM", "STTM", "HMMA", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            short = re.search(r"win\d{3}\d\d([a-z_0-9]+kernel[a-z_0-9]*)(ILi(\d)E)?E", name)
            cur = (short.group(1) + (f"<{short.group(3)}>" if short.group(3) else "")) if short else name
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    print(f"# static SASS instruction counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':32s} {'total':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
    for name, ops in kernels.items():
        def count(key):
            if key in ("FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "RED"):
                return sum(1 for o in ops if o.split(".")[0] == key)
            if key == "MUFU.LG2":
                return sum(1 for o in ops if o.startswith("MUFU.LG2"))
            if key == "ATOM":
                return sum(1 for o in ops if o.split(".")[0] in ("ATOM", "ATOMG", "ATOMS"))
            return sum(1 for o in ops if o.startswith(key))
        print(f"{name:32s} {len(ops):6d} " + " ".join(f"{count(k):8d}" for k in KEYS))


if __name__ == "__main__":
    main()
