"""In-tree build of libssb.so (hand-written sm_100a CUDA behind the C-ABI).

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to
the sources (git-ignored, shipped to the GPU box by gpurun)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssb.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
# per-file extra flags: the float64 tracker math keeps NumPy's expression order
EXTRA = {"tracker.cu": ["-fmad=false"]}
if os.environ.get("SSB_DW_FFMA2") in ("0", "1"):      # A/B switch of the depthwise FMA form (reid_tc3.cu)
    EXTRA["reid_tc3.cu"] = ["-DSSB_DW_FFMA2=" + os.environ["SSB_DW_FFMA2"]]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ssb.h"))
    objs = []
    for src in sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
