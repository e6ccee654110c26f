"""Build libistvt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python 2023-tifs-istvt_b200/build.py [--force]

The shared library has a plain C ABI (include/istvt_b200.h), links cudart statically and resolves the
one driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it loads on a CPU-only box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libistvt_b200.so")
SOURCES = [
    "common.cu",
    "layernorm.cu",
    "gemm_tcgen05.cu",
    "gemm_tcgen05_2cta.cu",
    "gemm_f32.cu",
    "attn_temporal.cu",
    "attn_spatial.cu",
    "attn_joint.cu",
    "token_build.cu",
    "attn_spatial_bwd.cu",
    "backward.cu",
    "entry_flow_bwd.cu",
    "entry_flow.cu",
    "conv_stem_tc.cu",
    "conv3x3_tc.cu",
    "conv3x3_strip.cu",
    "sepconv_fused.cu",
    "xception_tail.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
] + os.environ.get("ISTVT_BUILD_DEFS", "").split()      # e.g. -DISTVT_GEMM_TRACE for tools/gemm_trace.py (debug only)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC or put it on PATH")


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/istvt_b200.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
