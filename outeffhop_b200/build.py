"""Build liboeh.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m outeffhop_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to outeffhop_b200/csrc/_obj/, the library to
outeffhop_b200/lib/liboeh.so (git-ignored; travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "lib", "liboeh.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["api.cu", "quant.cu", "softmax_rows.cu", "gate.cu", "attn_fwd_sm100.cu", "percentile.cu", "outlier_stats.cu", "ce_loss.cu", "layernorm.cu", "layernorm_f32.cu", "layernorm_f16.cu", "layernorm_bf16.cu", "gelu.cu", "gemm_sm100.cu", "gemm_ln_sm100.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(INCLUDE, "oeh.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, verbose: bool):
    spath = os.path.join(CSRC, src)
    opath = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(opath) and os.path.getmtime(opath) >= max(os.path.getmtime(spath), _deps_mtime()):
        return opath, ""
    cmd = [nvcc(), "-c", spath, "-o", opath] + NVCC_FLAGS
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = p.stdout + p.stderr
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    with open(opath + ".ptxas.log", "w") as f:
        f.write(log)
    return opath, log if verbose else ""


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in results]
    for _, log in results:
        if log:
            print(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
