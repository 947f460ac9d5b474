"""In-tree build of libann_b200_f32.so / libann_b200_f64.so (nvcc for the kernels, gcc for
the C host).  The .so files stay next to this file so that they travel with the repo
snapshot to the GPU box and show up as in-tree native code."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CUDA_HOME = os.path.dirname(os.path.dirname(os.path.realpath(NVCC)))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

CU_SOURCES = ["annb_prepare.cu", "annb_leaf.cu", "annb_finish.cu", "annb_query.cu", "annb_probe.cu"]
C_SOURCES = ["ann_host.c", "ann_results.c", "ann_ingest.c", "ann_query.c", "ann_dist.c", "ann_save_io.c", "ann_multi.c"]
HEADERS = [os.path.join(INCLUDE, h) for h in os.listdir(INCLUDE)] + [
    os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]


def lib_path(suffix: str) -> str:
    return os.path.join(HERE, f"libann_b200_{suffix}.so")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build(force: bool = False, verbose: bool = False) -> None:
    """Compile (in parallel) whatever is stale, then link the two libraries."""
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    jobs, links = [], []
    for suffix, defs in (("f32", ["-DUSE_FLOAT"]), ("f64", [])):
        objs = []
        for src in CU_SOURCES:
            obj = os.path.join(HERE, "build", f"{os.path.splitext(src)[0]}_{suffix}.o")
            path = os.path.join(CSRC, src)
            if force or _stale(obj, [path] + HEADERS):
                jobs.append([NVCC, *ARCH, "-O3", "-lineinfo", "-fmad=false", "-std=c++17", *defs,
                             "-I", INCLUDE, "-I", CSRC, "-Xcompiler", "-fPIC", "-c", path, "-o", obj])
            objs.append(obj)
        for src in C_SOURCES:
            obj = os.path.join(HERE, "build", f"{os.path.splitext(src)[0]}_{suffix}.o")
            path = os.path.join(CSRC, src)
            if force or _stale(obj, [path] + HEADERS):
                jobs.append(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", *defs,
                             "-I", INCLUDE, "-I", CSRC, "-I", os.path.join(CUDA_HOME, "include"),
                             "-c", path, "-o", obj])
            objs.append(obj)
        links.append((lib_path(suffix), objs))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
    for out, objs in links:
        if force or jobs or _stale(out, objs):
            _run([NVCC, *ARCH, "-shared", "-Xlinker", "-Bsymbolic", "-o", out, *objs,
                  "-lcudart", "-lm", "-lpthread", "-ldl"], verbose)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
