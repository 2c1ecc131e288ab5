"""Build recipe for libb200deflate.so (the C-ABI library with the sm_100a kernels).

One nvcc invocation, in-tree output (so the .so travels with the repo snapshot to the GPU box):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --shared -Xcompiler -fPIC
nvcc cross-compiles without a GPU.  `python deflate.hpp_b200/build.py` or `build()` from Python.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200deflate.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", "b200_deflate.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    """Compile csrc/b200_deflate.cu (which includes the .cuh kernels) into libb200deflate.so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB, os.path.join(CSRC, "b200_deflate.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
