"""Build libcircom_b200.so (CUDA kernels for sm_100a + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcircom_b200.so")
CLI = os.path.join(HERE, "circom_cuda_witness")
SOURCES = ["capi.cu", "tape_calls.cu", "flatten.cpp", "formats.cpp", "hostpack.cpp", "r1cs_compile.cpp"]
CLI_SOURCES = ["cli.cpp"]
HEADERS = ["kernels.cuh", "fr_device.cuh", "tape.h", "tape_calls.h", "u256.h", "hostpack.h", "r1cs_small.h", os.path.join("..", "..", "include", "circom_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS + CLI_SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("CW_NVCC_EXTRA", "").split()
    tmp = LIB + ".building"     # (a reader - another process, a snapshot of the tree - never sees a half-written library)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed:\n" + r.stdout[-3000:] + r.stderr[-6000:])
    os.replace(tmp, LIB)
    if verbose:
        print(r.stderr)
    # command-line calculator (client of the C ABI only)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "cli.cpp"), "-L" + HERE,
                        "-lcircom_b200", "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (cli) failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
