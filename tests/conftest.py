import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    _ensure_built()


def _ensure_built():
    """The built artefacts are git-ignored; a fresh checkout builds them once (nvcc cross-compiles without a GPU)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "soap_bubble_simulation_b200", "csrc", "libsbx.so")
    oracle = os.path.join(root, "oracle", "libsbx_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.dirname(so), "-j4", "-s"])
    if not os.path.exists(oracle):
        subprocess.check_call(["make", "-C", os.path.dirname(oracle), "-s"])


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
