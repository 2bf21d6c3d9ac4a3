"""Build librvcr.so in-tree with nvcc for sm_100a (B200) -- no torch, no JIT cache.

    python codename-rvc-fork-3_b200/build.py [--force] [--verbose]

Every csrc/*.cu is compiled to an object with
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
(in parallel, only when the source or a header is newer than the object) and linked into
codename-rvc-fork-3_b200/librvcr.so with a static cudart.  The .so is git-ignored but travels to
the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librvcr.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Wno-deprecated-declarations", "-diag-suppress", "20012"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; librvcr needs the CUDA 12.9 toolkit to build")
    return p


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)]
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        if stale:
            cmd = [nvcc, *ARCH, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or not os.path.exists(LIB) or force:
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined",
               "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of librvcr.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
