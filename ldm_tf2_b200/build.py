"""Builds libldm_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m ldm_tf2_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libldm_b200.so")
SOURCES = ["kernels.cu", "engine.cu", "model.cu", "api.cu", "comm.cu", "validate.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-msse4.2", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include")]


def _extra_flags():
    # e.g. LDM_B200_NVCC_FLAGS=-DLDM_GEMM_TRACE_FINE for the instrumented epilogue (profiles/trace_epilogue.py)
    return os.environ.get("LDM_B200_NVCC_FLAGS", "").split()


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/ldm_b200.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    # LDM_B200_BUILD_DIR: build objects + library somewhere else (e.g. while a gpurun snapshot of the tree is pending)
    out_dir = os.environ.get("LDM_B200_BUILD_DIR")
    lib = os.path.join(out_dir, "libldm_b200.so") if out_dir else LIB
    obj_dir = os.path.join(out_dir, "build") if out_dir else os.path.join(HERE, "build")
    stamp = os.path.join(obj_dir, "stamp")
    dig = _digest()
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    os.makedirs(obj_dir, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + _extra_flags()

    def cc(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
