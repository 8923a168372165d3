"""Builds libplayaid_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

`python -m playaid_core_b200.build` or `__graft_entry__.build()`. The .so is git-ignored but ships
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libplayaid_b200.so")

# file -> extra flags. preprocess.cu must reproduce OpenCV/Pillow rounding: no FMA contraction.
SOURCES = {
    "preprocess.cu": ["-fmad=false"],
    "conv_gemm.cu": [],
    "conv_gemm2.cu": [],
    "conv_patch.cu": [],
    "conv_patch2.cu": [],
    "conv1.cu": [],
    "small_kernels.cu": ["-fmad=false"],   # the fp64 box geometry must not be contracted (bit-equal to numpy); the other kernels here are memory-bound
    "transformer_kernels.cu": [],
    "pa_api.cu": [],
}
COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcudafe", "--diag_suppress=177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _deps_mtime() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False, verbose: bool = False, experiment: bool = False) -> str:
    """`experiment=True` compiles the tuning / debug switches in (-DPA_EXPERIMENT: PA_CONV_DEBUG, PA_NO_PAIR, PA_PP_*, PA_ST_*
    environment variables); the shipped library ignores them."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *COMMON, *extra, *(["-DPA_EXPERIMENT"] if experiment else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"])
    return LIB


if __name__ == "__main__":
    exp = "--experiment" in sys.argv
    print(build(force="--force" in sys.argv or exp, verbose="-v" in sys.argv, experiment=exp))
