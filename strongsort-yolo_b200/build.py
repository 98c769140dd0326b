"""In-tree build of libssb.so (hand-written sm_100a CUDA behind the C-ABI, include/ssb.h) and of
libssb_dbg.so = the same sources + the A/B baselines and diagnostics (-DSSB_BASELINES: the fp32 SIMT
OSNet, the 9-tap and round-1 OSBlock kernels, tc_probe, the debug entry points of include/ssb_debug.h)
that only the parity tests and tools load.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so files sit next to
the sources (git-ignored, shipped to the GPU box by gpurun)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssb.so")
LIB_DBG = os.path.join(HERE, "libssb_dbg.so")
# sources that exist only in the debug library (empty translation units without -DSSB_BASELINES)
DBG_ONLY = {"reid.cu", "reid_tc3.cu", "tc_probe.cu"}
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


def _build_one(lib, objdir, flags, srcs, force, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ssb.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ssb_debug.h"))
    todo, objs = [], []
    for src in srcs:
        obj = os.path.join(objdir, src[:-3] + ".o")
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + [h for h in headers if os.path.exists(h)]):
            todo.append([nvcc] + ARCH + COMMON + flags + EXTRA.get(src, []) + ["-c", path, "-o", obj])
    procs = []
    for cmd in todo:                       # the translation units are independent: compile them concurrently
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if force or _stale(lib, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", lib] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return lib


def build_variant(name, flags, verbose=False):
    """A/B build of the product library with extra -D flags -> variants/libssb_<name>.so (tools/build_variants.py)."""
    out = os.path.join(HERE, "variants", f"libssb_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    prod = [f for f in sources() if f not in DBG_ONLY]
    return _build_one(out, os.path.join(HERE, "build_var_" + name), list(flags), prod, False, verbose)


def build(force=False, verbose=False, debug=True):
    prod = [f for f in sources() if f not in DBG_ONLY]
    _build_one(LIB, os.path.join(HERE, "build"), [], prod, force, verbose)
    if debug:
        _build_one(LIB_DBG, os.path.join(HERE, "build_dbg"), ["-DSSB_BASELINES"], sources(), force, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
