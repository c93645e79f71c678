"""Builds, in-tree: libbags_b200.so (the C-ABI library: every CUDA kernel, nvcc for sm_100a) and
diff_gaussian_rasterization/_C*.so (the thin PyTorch C++ extension above the C ABI, g++ against libtorch).

    python bundle-adjusting-gaussian-splatting_b200/build.py [--force] [--verbose]

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbags_b200.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend.cu", "preprocess_bwd.cu", "knn.cu", "distort.cu", "loss.cu", "adam.cu", "densify.cu", "allreduce.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC", "-Xptxas", "-v"]


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def ext_path():
    import sysconfig
    return os.path.join(HERE, "diff_gaussian_rasterization", "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ext(force=False, verbose=False):
    """diff_gaussian_rasterization._C: csrc/ext.cpp compiled with g++ against libtorch and linked to libbags_b200.so
    (found at run time through an $ORIGIN-relative rpath)."""
    import sysconfig
    out = ext_path()
    src = os.path.join(CSRC, "ext.cpp")
    hdr = os.path.join(HERE, "..", "include", "bags_b200.h")
    if (not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(hdr))
            and os.path.getmtime(out) >= os.path.getmtime(LIB)):
        return out
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-variable", src, "-o", out,
           "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for inc in ce.include_paths("cuda"):
        cmd += ["-isystem", inc]
    cmd += ["-isystem", sysconfig.get_paths()["include"]]
    cmd += ["-L" + HERE, "-lbags_b200", "-Wl,-rpath,$ORIGIN/..", "-L" + tlib, "-Wl,-rpath," + tlib,
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python"]
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    if os.path.isdir(cuda_lib):
        cmd += ["-L" + cuda_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, "ext.cpp.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for ext.cpp:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return out


def build(force=False, verbose=False):
    lib = build_lib(force, verbose)
    build_ext(force, verbose)
    return lib


def build_lib(force=False, verbose=False):
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
