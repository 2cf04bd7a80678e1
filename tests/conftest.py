import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle.golden import load_golden  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The shared libraries are build artefacts (git-ignored): build them when a fresh checkout runs the tests.
    nvcc cross-compiles without a GPU; on the GPU box the prebuilt files travel with the snapshot."""
    import shutil

    from dphtools_b200 import build as cuda_build

    if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        cuda_build.build()
    elif not os.path.exists(cuda_build.LIB):
        pytest.exit("libdphlm.so is missing and nvcc is not available to build it", returncode=2)
    from oracle import host_emu

    host_emu.build()


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get
