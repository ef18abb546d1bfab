"""In-tree build recipes (nvcc / g++), used by ``__graft_entry__.build()`` and lazily by the harness.

Everything is built next to the sources so that the artefacts travel with a repository snapshot:

* ``specfem3d_b200/libspecfem3d_b200.so`` -- the product: CUDA kernels (sm_100a) + the Fortran-callable C ABI
* ``specfem3d_b200/_meshgen.so``          -- host-side harness generator helpers (g++)
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libspecfem3d_b200.so")
MESHGEN_PATH = os.path.join(HERE, "_meshgen.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",          # the reference builds with nvcc defaults: FMA contraction on, no fast-math
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return res.stdout


def build_meshgen(force: bool = False) -> str:
    src = os.path.join(CSRC, "meshgen.cpp")
    if not force and _newer(MESHGEN_PATH, [src]):
        return MESHGEN_PATH
    _run(["g++", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", MESHGEN_PATH, src])
    return MESHGEN_PATH


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(
        os.path.join(ROOT, "include", "*.h"))
    if not force and _newer(LIB_PATH, deps):
        return LIB_PATH
    out = _run([NVCC] + NVCC_FLAGS + ["-shared", "-o", LIB_PATH] + srcs + ["-lcudart"],
               log=os.path.join(HERE, "build_cuda.log"))
    if verbose:
        print(out)
    return LIB_PATH


def build_all(force: bool = False) -> None:
    build_meshgen(force)
    build_cuda(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB_PATH, MESHGEN_PATH)
