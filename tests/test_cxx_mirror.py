"""The header-only C++ mirror (include/sbx.hpp) compiles against libsbx.so; on a GPU box it replays a few of the
reference's unit tests through the C ABI from C++."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "cxx_mirror_smoke")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cxx_mirror_smoke.cpp"),
                           "-o", exe, "-L" + CSRC, "-lsbx", "-Wl,-rpath," + CSRC])
    return exe


def test_cxx_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CXX_MIRROR_NO_GPU_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cxx_mirror_replays_reference_tests_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CXX_MIRROR_GPU_OK" in r.stdout, f"rc={r.returncode}\n{r.stdout}{r.stderr}"
