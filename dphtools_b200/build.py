"""Build ``dphtools_b200/csrc/libdphlm.so`` in-tree with nvcc for sm_100a.

``python -m dphtools_b200.build`` (or ``__graft_entry__.build()``).  nvcc cross-compiles without a GPU.
"""

import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libdphlm.so")
SOURCES = ["lm_kernels.cu"]
DEPS = ["lm_core.cuh", os.path.join("..", "..", "include", "dphlm.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-cudart", "static",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + DEPS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared"] + SOURCES + ["-o", LIB]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
