"""Builds rsis_b200/lib/librsis_b200.so (sm_100a only) with nvcc.  In-tree so the .so travels with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# RSIS_B200_LIB: load / build another file instead (development: the -DRSIS_DEBUG_TIMING build of the stamp probes)
LIB_PATH = os.environ.get("RSIS_B200_LIB") or os.path.join(LIB_DIR, "librsis_b200.so")
SOURCES = ["api.cu", "pack.cu", "layout.cu", "conv_simt.cu", "conv_umma.cu", "decoder_ops.cu", "bn_train.cu", "backward.cu", "objectives.cu", "postprocess.cu", "optim.cu", "dispatch.cu"]
NVCC_FLAGS = ([f"-DRSIS_GROUP_EPI_WARPS={int(os.environ['RSIS_B200_GROUP_EPI_WARPS'])}"] if os.environ.get("RSIS_B200_GROUP_EPI_WARPS") else []) + (["-DRSIS_DEBUG_TIMING"] if os.environ.get("RSIS_B200_BUILD_DEBUG_TIMING") else []) + ([f"-DRSIS_TRACE_BLOCK={int(os.environ['RSIS_B200_TRACE_BLOCK'])}"] if os.environ.get("RSIS_B200_TRACE_BLOCK") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build librsis_b200.so")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rsis_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(os.path.dirname(LIB_PATH), src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
