"""Builds libbags_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python bundle-adjusting-gaussian-splatting_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbags_b200.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend.cu", "preprocess_bwd.cu", "knn.cu", "distort.cu", "loss.cu", "adam.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC", "-Xptxas", "-v"]


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = os.path.join(OBJ, src + ".o")
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj, *NVCC_FLAGS]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(OBJ, src + ".log"), "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
