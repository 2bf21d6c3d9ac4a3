"""Compile oracle/rvc_oracle.c into oracle/_build/librvc_oracle.so (test infrastructure only).

The reference (codename0og/codename-rvc-fork-3) has no native sources for this path: its
arithmetic lives in the un-vendored faiss-cpu==1.7.3 wheel, so there is no `oracle/_ref` build
of reference sources ("Oracle").  This recipe only builds our C restatement.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rvc_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librvc_oracle.so")


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O3", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared",
           "-fPIC", "-fvisibility=hidden", SRC, "-o", OUT, "-lm"]
    # fmaf() must be a single-rounding fused multiply-add; with -mfma it is one instruction,
    # without it gcc calls libm's (correct, slow) software fmaf.
    if _cpu_has("fma") and _cpu_has("avx2"):
        cmd[1:1] = ["-mavx2", "-mfma"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
