"""Recipe for ``oracle/_ref/``: the reference's own implementation of the path, made available where ``/root/reference``
does not exist (the GPU box).  TEST / BASELINE INFRASTRUCTURE.

The reference is pure Python, so "building" it is placing the four files its ``dphtools.utils.lm`` needs under
``oracle/_ref/dphtools/`` (git-ignored build output, like a compiled ``.so``; it travels with the gpurun snapshot and is
never committed).  ``bench.py --impl reference`` and the ``cpu_baseline`` leg import it from there
(``cpu_baseline.kind == "reference"``) and fall back to the float64 port ``oracle/lm_oracle.py`` when it is absent.
"""

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DPHTOOLS_REF", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["dphtools/__init__.py", "dphtools/_version.py", "dphtools/utils/__init__.py", "dphtools/utils/lm.py"]


def build():
    """Returns the directory to put on ``sys.path``, or None when the reference tree is not present and nothing was built before."""
    if os.path.isdir(os.path.join(REF_SRC, "dphtools")):
        for rel in FILES:
            dst = os.path.join(OUT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    return OUT if available() else None


def available():
    return all(os.path.exists(os.path.join(OUT, rel)) for rel in FILES)


def load():
    """``dphtools.utils.lm`` of the reference from ``oracle/_ref`` (None if it has not been built)."""
    if not available():
        return None
    import importlib
    import sys

    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    return importlib.import_module("dphtools.utils.lm")


if __name__ == "__main__":
    print(build())
