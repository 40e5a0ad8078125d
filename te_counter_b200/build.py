"""Build recipes, in-tree: libtecount.so (hand-written CUDA, sm_100a only) and libtecbam.so (host BAM
decoder, C++ / zlib / threads, no CUDA)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("TEC_LIB") or os.path.join(HERE, "libtecount.so")
BAM_LIB = os.environ.get("TEC_BAM_LIB") or os.path.join(HERE, "libtecbam.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + \
        [os.path.join(os.path.dirname(HERE), "include", "tecount.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-o", LIB, os.path.join(CSRC, "tecount.cu"), "-lz"]
    cmd[1:1] = os.environ.get("TEC_NVCC_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libtecount.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_bam(force=False):
    src = [os.path.join(CSRC, "bamdecode.cpp"), os.path.join(os.path.dirname(HERE), "include", "tecbam.h")]
    if not force and os.path.exists(BAM_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(BAM_LIB) for s in src):
        return BAM_LIB
    cmd = [CXX, "-O3", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread", "-o", BAM_LIB, src[0], "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libtecbam.so")
    return BAM_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_bam(force="--force" in sys.argv))
