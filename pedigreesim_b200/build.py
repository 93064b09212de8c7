"""Build libpsim.so in-tree for sm_100a (nvcc cross-compiles without a GPU) and the host
program `pedigreesim` (host/main.cpp, plain C++17 over the C ABI)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpsim.so")
SOURCES = ["psim.cu"]
HEADERS = ["psim_meiosis.cuh", "psim_rng.cuh", "psim_k1.cuh", "psim_k0.cuh", "psim_emit_tiles.cuh", "psim_emit_rows.cuh",
           "psim_emit_generic.cuh", "psim_text.cuh", "psim_stats.cuh", "psim_misc.cuh", "psim_shard.inl", os.path.join("..", "..", "include", "psim.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


HOST_DIR = os.path.join(PKG, "host")
HOST_BIN = os.path.join(PKG, "bin", "pedigreesim")
HOST_SOURCES = ["main.cpp", "writers.hpp", "popfiles.hpp", "text.hpp"]


def build_host(force=False):
    """g++ host/main.cpp -> bin/pedigreesim, linked against libpsim.so next to it ($ORIGIN/..)."""
    deps = [os.path.join(HOST_DIR, f) for f in HOST_SOURCES] + [os.path.join(PKG, "..", "include", "psim.h"), LIB]
    if not force and os.path.exists(HOST_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_BIN) for d in deps):
        return HOST_BIN
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-o", HOST_BIN, os.path.join(HOST_DIR, "main.cpp"),
           "-L" + PKG, "-lpsim", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    return HOST_BIN


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not is_stale():
        build_host()
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    build_host(force=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
