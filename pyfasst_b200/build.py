"""Builds pyfasst_b200/libpyfasst_b200.so from csrc/*.cu with nvcc for sm_100a.

The shared library is built IN-TREE (it travels to the GPU box with the repo
snapshot; it is git-ignored).  nvcc cross-compiles without a GPU.

    python -m pyfasst_b200.build [--force] [--verbose]
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libpyfasst_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libpyfasst_b200.so")
    return nvcc


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu and link the shared library. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(INCLUDE, "*.h")))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, p.returncode, p.stdout + p.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, rc, out in ex.map(run, jobs):
            if verbose or rc != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            if rc != 0:
                raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if jobs or force or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + \
            ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        cmd, rc, out = run(cmd)
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError("link of libpyfasst_b200.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
