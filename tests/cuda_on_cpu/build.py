"""Compiles the TRAINING kernels' .cu sources for the CPU (TEST INFRASTRUCTURE; see fake/cuda_runtime.h).

    from tests.cuda_on_cpu import build; lib = build.load()     # ctypes library with the same cg3d_* entry points

Source transformation (nothing else is touched): `kernel<<<grid, block[, smem[, stream]]>>>(args)` becomes
`cpu_cuda::launch(kernel, grid, block, smem, stream, args)` and `extern __shared__ T name[];` becomes a pointer to the
launch's dynamic buffer.  Built with g++ -O1 -ffp-contract=off (the CUDA build uses -fmad=false) into tests/cuda_on_cpu/_build.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cagroup3d_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["spconv_bwd.cu", "train_bwd.cu", "train_assign.cu", "train_loss.cu", "nms.cu", "sort.cu", "coords.cu", "pool_interp.cu", "detect.cu", "proposal.cu", "train_ops.cu", "spconv_simt.cu"]

_DEFS = """
#include <cuda_runtime.h>
namespace cpu_cuda {
Barrier block_barrier;
std::vector<Warp>* warps = nullptr;
dim3 g_blockDim, g_gridDim;
char* dyn_smem = nullptr;
thread_local dim3 t_threadIdx, t_blockIdx;
thread_local int t_linear = 0;
int sync_or_acc[2] = {0, 0};
}
"""


def _match(src: str, i: int, open_ch: str, close_ch: str) -> int:
    depth = 0
    while True:
        c = src[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1


def _split_top(s: str):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def transform(src: str) -> str:
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(cpu_cuda::dyn_smem);", src)
    out, i = "", 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            return out + src[i:]
        k = j
        if src[k - 1] == ">":                                 # kernel<template args><<<...>>>: include the argument list
            depth = 0
            while True:
                k -= 1
                depth += {">": 1, "<": -1}.get(src[k], 0)
                if depth == 0:
                    break
        while src[k - 1].isalnum() or src[k - 1] == "_":
            k -= 1
        kernel = src[k:j]
        e = src.find(">>>", j)
        cfg = _split_top(src[j + 3:e])
        cfg += ["0"] * (4 - len(cfg))
        a0 = src.find("(", e)
        a1 = _match(src, a0, "(", ")")
        args = src[a0 + 1:a1].strip()
        out += src[i:k] + f"cpu_cuda::launch({kernel}, dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({cfg[2]}), (cudaStream_t)({cfg[3]})" + \
            (", " + args if args else "") + ")"
        i = a1 + 1


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libtrain_kernels_cpu.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, "fake", "cuda_runtime.h"), os.path.join(CSRC, "common.cuh"), __file__,
                   os.path.join(ROOT, "include", "cagroup3d_b200.h")]
    if not force and os.path.exists(lib) and os.path.getmtime(lib) > max(os.path.getmtime(d) for d in deps):
        return lib
    cpps = []
    for s in srcs:
        text = transform(open(s).read())
        text = text.replace('#include "common.cuh"', f'#include "{os.path.join(CSRC, "common.cuh")}"')
        text = text.replace('#include "../../include/cagroup3d_b200.h"', f'#include "{os.path.join(ROOT, "include", "cagroup3d_b200.h")}"')
        cpp = os.path.join(OUT, os.path.basename(s)[:-3] + ".cpp")
        open(cpp, "w").write(text)
        cpps.append(cpp)
    defs = os.path.join(OUT, "defs.cpp")
    open(defs, "w").write(_DEFS)
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-I", os.path.join(HERE, "fake"),
           *cpps, defs, "-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-4000:])
    return lib


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
