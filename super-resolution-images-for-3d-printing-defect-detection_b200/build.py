"""Build libsrb200.so in-tree with nvcc for sm_100a (no torch involved).

    python super-resolution-images-for-3d-printing-defect-detection_b200/build.py [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "srb200", "libsrb200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["core.cu", "metrics.cu", "metrics_mma.cu", "bicubic.cu", "tiling.cu", "conv.cu", "conv_direct.cu", "conv_head.cu", "conv_tc.cu", "conv_fold.cu", "conv_headtc.cu", "conv_up.cu", "attention_tc.cu", "conv_head8.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "srb200.h"))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        print("\n".join(logs))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if jobs or force or _stale(OUT, objs):
        run([nvcc(), "-shared", "-o", OUT, *objs, "-lcudart", "-Xlinker", "--no-undefined"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
