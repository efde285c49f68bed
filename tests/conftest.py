import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def emu_lib():
    """tests/emu: the product's kernel sources compiled for the host (test infrastructure only)."""
    import subprocess

    from tiberate_fhe_b200 import _native

    here = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(here, "_build", "libtb200_emu.so")
    srcs = [os.path.join(ROOT, "tiberate_fhe_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "tiberate_fhe_b200", "csrc"))]
    srcs.append(os.path.join(here, "emu_runtime.cpp"))
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.run(["sh", os.path.join(here, "build_emu.sh")], check=True, capture_output=True)
    return _native.Lib(so)
