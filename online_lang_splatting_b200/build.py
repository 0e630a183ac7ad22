"""Builds the in-tree CUDA shared library (libols_b200.so) for sm_100a with plain nvcc.

The library is the C ABI of include/ols_b200.h; it is git-ignored but travels to the GPU box with
the gpurun snapshot.  `python -m online_lang_splatting_b200.build` rebuilds when sources changed.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libols_b200.so")
SOURCES = ["ols_api.cu", "ols_forward.cu", "ols_backward.cu", "ols_ae.cu", "ols_loss.cu", "ols_optim.cu", "ols_knn.cu", "ols_hr.cu", "ols_ssim.cu", "ols_densify.cu", "ols_online_ae.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "ols_b200.h"))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.join(HERE, "lib"), exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + _deps()):
            jobs.append((s, [nvcc, "-c", src, "-o", obj] + NVCC_FLAGS))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(f"--- {name}\n{r.stdout}{r.stderr}")
        else:
            with open(os.path.join(obj_dir, name + ".ptxas.log"), "w") as f:
                f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}")

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(run, jobs))
    if jobs or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                       "-lcudart", "-lcuda"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
