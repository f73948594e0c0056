"""Builds open_spiel_coup_b200/libcoup_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc.

The library is built IN-TREE so that it travels with the repo snapshot to the GPU box; it is
git-ignored (history stays source-only). nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcoup_b200.so")
SOURCES = ["coup_capi.cu"]
HOST_SOURCES = ["coup_host_policy.cc"]   # host compiler only (AVX2 path picked at run time)
HEADERS = ["coup_host_policy.cc"] + sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "coup_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        cc = [os.environ.get("CXX", "g++"), "-std=c++17", "-O3", "-fPIC", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cc, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("host compile failed:\n" + " ".join(cc) + "\n" + res.stdout + res.stderr)
        objs.append(obj)
    extra = os.environ.get("COUP_B200_NVCC_EXTRA", "").split()      # e.g. -DCOUP_WS_DEBUG for the cycle counters
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
