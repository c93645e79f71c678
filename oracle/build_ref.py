#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CUDA extensions as a checker.

Compiles the reference sources where they lie under /root/reference (3dgs-pose and simple-knn) with plain
nvcc/g++ commands (the reference's own setup.py / CMake are NOT run) into the git-ignored ``oracle/_ref/``:

    oracle/_ref/diff_gaussian_rasterization/{__init__.py,_C.so}   <- 3dgs-pose/{cuda_rasterizer/*.cu,rasterize_points.cu,ext.cpp}
    oracle/_ref/simple_knn/{__init__.py,_C.so}                    <- simple-knn/{simple_knn.cu,spatial.cu,ext.cpp}

    oracle/_ref/glue/{gaussian_renderer/__init__.py, scene/{cameras,gaussian_model}.py, utils/*.py}
                                                                  <- the reference's own Python callers of the two
                                                                     extensions, byte for byte (for the integration
                                                                     test: render(), Camera, GaussianModel, losses)

Nothing under oracle/_ref is product code: only tests/, __graft_entry__.smoke() and ``bench.py --impl reference``
import it, as the thing the product is compared against.  The only Python file copied is the reference's own
autograd wrapper (needed to call the reference through its public API on the GPU box, where /root/reference does
not exist); it lands in the ignored output directory, never in git history.

gcc-13 needs <cstdint>/<cfloat> injected (rasterizer_impl.h:24 uses std::uintptr_t, simple_knn.cu:90 FLT_MAX);
no reference source is edited.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("BAGS_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100,code=sm_100"]


def _flags():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-I", p]
    libdirs = ce.library_paths("cuda")
    common = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=_C", "-D_GLIBCXX_USE_CXX11_ABI=1"]
    return inc, libdirs, common, list(ce.COMMON_NVCC_FLAGS)


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_ext(name, srcdir, sources, extra_inc, pyfiles):
    inc, libdirs, common, nvcc_flags = _flags()
    pkg = os.path.join(OUT, name)
    obj = os.path.join(OUT, "_obj", name)
    os.makedirs(pkg, exist_ok=True)
    os.makedirs(obj, exist_ok=True)
    so = os.path.join(pkg, "_C.so")
    srcs = [os.path.join(srcdir, s) for s in sources]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
        print(f"[build_ref] {so} up to date")
        return so
    cmds, objs = [], []
    for s in srcs:
        o = os.path.join(obj, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmds.append(["nvcc", "-c", s, "-o", o, "-std=c++17", "--compiler-options", "-fPIC",
                         "-include", "cstdint", "-include", "cfloat", *ARCH, *nvcc_flags, *common, *inc, *extra_inc])
        else:
            cmds.append(["g++", "-c", s, "-o", o, "-std=c++17", "-O2", "-fPIC", "-include", "cstdint",
                         *common, *inc, *extra_inc])
    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(_run, cmds))
    link = ["g++", "-shared", "-o", so, *objs]
    for d in libdirs:
        link += ["-L", d, f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lcudart", "-lc10_cuda", "-ltorch_cuda"]
    _run(link)
    for src, dst in pyfiles:
        shutil.copyfile(src, os.path.join(pkg, dst))
    return so


GLUE_FILES = [
    "gaussian_renderer/__init__.py",         # render(): builds the settings tuple, calls the rasterizer, 5-value unpack
    "scene/cameras.py",                      # Camera: learnable delta_quaternion / delta_translation / fov -> matrices
    "scene/gaussian_model.py",               # GaussianModel: activations, create_from_pcd -> distCUDA2
    "utils/graphics_utils.py", "utils/general_utils.py", "utils/sh_utils.py", "utils/system_utils.py",
    "utils/rigid_utils.py", "utils/camera.py", "utils/loss_utils.py",
]


def stage_glue():
    """Copies the reference's Python callers of the extensions, unmodified, into the git-ignored oracle/_ref/glue so
    that tests/test_reference_glue_gpu.py can run them on the GPU box (where /root/reference does not exist) on top
    of both implementations.  scene/__init__.py is deliberately NOT staged: `scene` becomes a namespace package, so
    importing scene.cameras does not drag in the dataset readers and lens networks (out of scope, missing packages)."""
    for rel in GLUE_FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, "glue", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    print(f"[build_ref] staged {len(GLUE_FILES)} reference Python files under {os.path.join(OUT, 'glue')}")


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present -- using prebuilt oracle/_ref if any")
        return 0
    r = os.path.join(REF, "3dgs-pose")
    build_ext("diff_gaussian_rasterization", r,
              ["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu", "cuda_rasterizer/backward.cu",
               "cuda_rasterizer/utils.cu", "rasterize_points.cu", "ext.cpp"],
              ["-I", os.path.join(r, "third_party/glm")],
              [(os.path.join(r, "diff_gaussian_rasterization/__init__.py"), "__init__.py")])
    k = os.path.join(REF, "simple-knn")
    build_ext("simple_knn", k, ["simple_knn.cu", "spatial.cu", "ext.cpp"], [],
              [(os.path.join(k, "simple_knn/__init__.py"), "__init__.py")]
              if os.path.exists(os.path.join(k, "simple_knn/__init__.py")) else [])
    init = os.path.join(OUT, "simple_knn", "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()
    stage_glue()
    return 0


if __name__ == "__main__":
    sys.exit(main())
