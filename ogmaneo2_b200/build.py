"""Builds libogn_b200.so in-tree: nvcc (sm_100a) for the kernels, g++ for the host classes, one shared library.

    python -m ogmaneo2_b200.build [--force]

nvcc cross-compiles without a GPU. The library links cudart statically, so it loads (but cannot compute) on a box
without a driver; the CUDA path is the only path.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libogn_b200.so")
OBJ = os.path.join(HERE, "_build")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA, "bin", "nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-prec-div=true",
              "-prec-sqrt=true", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", CXX]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-sign-compare", f"-I{CUDA}/include",
             f"-I{ROOT}/include"]
HOST_SOURCES = ["Device", "Helpers", "SparseCoder", "Predictor", "Actor", "Hierarchy", "ImageEncoder", "FlatApi"]


def _sources():
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))]
    srcs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h")]
    srcs += [os.path.join(ROOT, "include", "ogmaneo", f) for f in os.listdir(os.path.join(ROOT, "include", "ogmaneo"))]
    return srcs


STAMP = os.path.join(HERE, "libogn_b200.so.sha")


def _digest():
    import hashlib
    h = hashlib.sha256()
    for s in sorted(_sources()):
        h.update(os.path.relpath(s, ROOT).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).replace(ROOT, "").encode())
    return h.hexdigest()


def stale():
    """True when the library is missing or was built from different sources (content hash, not mtimes: the snapshot
    that travels to the GPU box does not preserve them)."""
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = [os.path.join(OBJ, "ogn_kernels.o")]
    _run([NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "ogn_kernels.cu"), "-o", objs[0]], verbose)
    for s in HOST_SOURCES:
        o = os.path.join(OBJ, s + ".o")
        _run([CXX] + CXX_FLAGS + ["-c", os.path.join(CSRC, "host", s + ".cpp"), "-o", o], verbose)
        objs.append(o)
    _run([NVCC, "-shared", "-cudart", "static", "-ccbin", CXX, "-o", OUT] + objs, verbose)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
