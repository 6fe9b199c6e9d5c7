"""Builds libtmolb200.so in-tree with nvcc for sm_100a (no other architecture, no fallback).

    python -m tensormol_b200.csrc.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["tm_api.cu", "tm_nlist.cu", "tm_desc.cu", "tm_mlp.cu", "tm_gemm_tc.cu", "tm_pair.cu", "tm_force.cu", "tm_compat.cu"]
HEADERS = ["tm_internal.h", os.path.join("..", "..", "include", "tmolb200.h")]
LIB = os.path.join(PKG, "libtmolb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]
FLAGS += os.environ.get("TM_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DTC_PROFILE for the GEMM warp-role cycle counters


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(HERE, ".build_stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    objs = []
    logs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, obj, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
        objs.append(obj)
    log = "\n".join(logs)
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (see tensormol_b200/csrc/build.log)")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
