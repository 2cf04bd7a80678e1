"""Build ``dphtools_b200/csrc/libdphlm.so`` in-tree with nvcc for sm_100a.

``python -m dphtools_b200.build [--force] [-v] [--only gauss2d] [--out PATH] [-DNAME=VALUE ...]``
(or ``__graft_entry__.build()``).  nvcc cross-compiles without a GPU; the translation units are compiled
in parallel.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libdphlm.so")
SOURCES = ["lm_kernels.cu", "roi_extract.cu", "lm_inst_gauss2d.cu", "lm_inst_gauss2df.cu", "lm_inst_ellip.cu", "lm_inst_gint.cu", "lm_inst_gintf.cu", "lm_inst_mexp1.cu", "lm_inst_mexp2.cu", "lm_inst_mexp3.cu", "lm_inst_mexp_irf.cu"]
DEPS = ["lm_core.cuh", "lm_launch.cuh", os.path.join("..", "..", "include", "dphlm.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + DEPS)


def build(force=False, verbose=False, out=LIB, defines=(), objdir=None):
    if not force and not _stale(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = objdir or os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", out] + [o for o, _ in results], cwd=CSRC)
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = os.path.abspath(argv[argv.index("--out") + 1]) if "--out" in argv else LIB
    print(build(force="--force" in argv or "--out" in argv, verbose="-v" in argv, out=out,
                defines=[a for a in argv if a.startswith("-D")],
                objdir=os.path.join(CSRC, "build", os.path.basename(out) + ".objs") if "--out" in argv else None))
