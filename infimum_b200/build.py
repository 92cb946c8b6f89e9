"""Build libinfimum_b200.so (CUDA, sm_100a) and the host-emulation test library.

    python -m infimum_b200.build [--force] [--jobs N]

nvcc cross-compiles for sm_100a without a GPU.  Outputs land in-tree
(infimum_b200/_build/*.o, infimum_b200/libinfimum_b200.so,
infimum_b200/libinfimum_hostemu.so); they are git-ignored but travel to the
GPU box with the working-tree snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libinfimum_b200.so")
HOSTEMU = os.path.join(HERE, "libinfimum_hostemu.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wno-unknown-pragmas",
                     "-I", os.path.join(ROOT, "include"), "-I", CSRC]
CU_SOURCES = ["poseidon_t%d.cu" % t for t in range(2, 9)] + ["dense_generic.cu", "leaves.cu", "tree_paths.cu", "imad_peak.cu", "multi.cu", "capi.cu"]
CXX_SOURCES = ["host_params.cpp"]
HEADERS = ["fr.cuh", "poseidon.cuh", "coop.cuh", "poseidon_tu.cuh", "launch.h", "host_fr.h", "host_params.h",
           os.path.join(ROOT, "include", "infimum_b200.h")]


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p if os.path.isabs(p) else os.path.join(CSRC, p), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, force):
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    dig = _digest([src] + HEADERS, " ".join(NVCC_FLAGS))
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    else:
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "include"),
               "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True


def build(force: bool = False, jobs: int | None = None, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    jobs = jobs or min(len(CU_SOURCES) + 1, os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(jobs) as ex:
        results = list(ex.map(lambda s: _compile(s, force), CU_SOURCES + CXX_SOURCES))
    objs = [o for o, _ in results]
    if any(c for _, c in results) or not os.path.exists(LIB) or force:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            print("built", LIB)
    # host emulation of the device arithmetic, for the CPU test-suite
    he_src = [os.path.join(CSRC, "hostemu.cpp"), os.path.join(CSRC, "host_params.cpp")]
    dig = _digest(["hostemu.cpp", "host_params.cpp"] + HEADERS)
    stamp = HOSTEMU + ".sha"
    if force or not os.path.exists(HOSTEMU) or not os.path.exists(stamp) or open(stamp).read() != dig:
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", CSRC,
               "-o", HOSTEMU] + he_src + ["-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("hostemu build failed:\n%s\n%s" % (r.stdout, r.stderr))
        with open(stamp, "w") as f:
            f.write(dig)
        if verbose:
            print("built", HOSTEMU)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    build(a.force, a.jobs)
    sys.exit(0)
