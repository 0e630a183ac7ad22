#!/usr/bin/env python
"""Compile the reference's own CUDA rasterizers into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

This is our own recipe (plain nvcc/g++ command lines), not the reference's
setup.py.  The sources are compiled *where they lie* under /root/reference;
nothing is copied into this repository and the only outputs are the shared
objects

    oracle/_ref/ref_P_C.so   <- submodules/diff-gaussian-rasterization            (F=15, 15x15 tiles)
    oracle/_ref/ref_D_C.so   <- submodules/diff-gaussian-rasterization-disentangle-optim (F=3, 16x16)
    oracle/_ref/ref_knn_C.so <- submodules/simple-knn (distCUDA2)

Both are pybind11/torch extension modules exporting the five functions of the
reference's ext.cpp (rasterize_language_gaussians, ..._backward, ...).  They
need a GPU to *run*; they are built here (no GPU needed) and travel to the GPU
box with the gpurun snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).

The one source incompatibility with CUDA 12.9 (rasterizer_impl.h uses
std::uintptr_t without <cstdint>) is solved with a forced pre-include, so the
reference sources stay byte-for-byte untouched.

Only tests/, bench.py's reference legs and tools/ may import the results.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")

VARIANTS = {
    "ref_P_C": "submodules/diff-gaussian-rasterization",
    "ref_D_C": "submodules/diff-gaussian-rasterization-disentangle-optim",
    "ref_knn_C": "submodules/simple-knn",        # distCUDA2 (mean squared distance to the 3 nearest neighbours)
}
VARIANT_SOURCES = {"ref_knn_C": ["simple_knn.cu", "spatial.cu", "ext.cpp"]}
SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
    "ext.cpp",
]


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_variant(name, rel, arch="100"):
    import torch
    from torch.utils import cpp_extension as ce

    src_root = os.path.join(REF, rel)
    if not os.path.isdir(src_root):
        raise FileNotFoundError(src_root)
    obj_dir = os.path.join(OUT, "obj_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    target = os.path.join(OUT, name + ".so")
    incs = ce.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"],
                                                    os.path.join(src_root, "third_party/glm"), src_root]
    inc_flags = [f"-I{p}" for p in incs]
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", "-std=c++17"]
    nvcc_flags = ["-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                  "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                  "--expt-relaxed-constexpr", "--compiler-options", "-fPIC",
                  "-gencode", f"arch=compute_{arch},code=sm_{arch}",
                  "--pre-include", "cstdint", "--pre-include", "cfloat", "-w"]
    objs, jobs = [], []
    for s in VARIANT_SOURCES.get(name, SOURCES):
        o = os.path.join(obj_dir, os.path.basename(s) + ".o")
        objs.append(o)
        src = os.path.join(src_root, s)
        if os.path.exists(o) and os.path.getmtime(o) > os.path.getmtime(src):
            continue
        if s.endswith(".cu"):
            jobs.append(["nvcc", "-c", src, "-o", o] + inc_flags + common + nvcc_flags)
        else:
            jobs.append(["g++", "-c", src, "-o", o, "-fPIC", "-O2", "-w"] + inc_flags + common)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(_run, jobs))
    lib_dirs = ce.library_paths(device_type="cuda")
    link = ["g++", "-shared", "-o", target] + objs + [f"-L{p}" for p in lib_dirs] + \
           ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    _run(link)
    return target


def build_all(names=None):
    os.makedirs(OUT, exist_ok=True)
    names = names or list(VARIANTS)
    with ThreadPoolExecutor(max_workers=2) as ex:
        return list(ex.map(lambda n: build_variant(n, VARIANTS[n]), names))


if __name__ == "__main__":
    print(build_all(sys.argv[1:] or None))
