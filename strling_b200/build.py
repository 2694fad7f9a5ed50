"""In-tree build of libstrgpu.so (hand-written CUDA for sm_100a + the C ABI).  No JIT cache: the .so sits next
to this file so it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libstrgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["api.cu", "scan_kernels.cu", "cluster_kernels.cu", "comm.cu", "decode_kernels.cu"]
CXX_SOURCES = ["pack.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "--cudart", "static", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _sources():
    return [os.path.join(CSRC, s) for s in CUDA_SOURCES + CXX_SOURCES]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + [os.path.join(ROOT, "include", "strgpu.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    deps += [os.path.join(HERE, "host", "inflate_fast.hpp")]   # compiled into decode_kernels.cu as the device-side decoder
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(HERE, "host")
CLI = os.path.join(HERE, "bin", "strling")
CXX = os.environ.get("CXX", "g++")


def build_cli(force: bool = False) -> str:
    """The `strling` command line (C++ host side: BAM decode, mate pairing, .bin / bounds files) over libstrgpu.so."""
    build_lib()
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + [LIB]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = [CXX, "-O3", "-std=c++17", "-Wall", "-Wextra", "-pthread", *srcs, "-I", os.path.join(ROOT, "include"), "-I", HOST,
           "-L", HERE, "-lstrgpu", "-lz", "-Wl,-rpath,$ORIGIN/..", "-o", CLI]
    subprocess.check_call(cmd)
    return CLI


if __name__ == "__main__":
    build_cli(force="--force" in sys.argv)
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
