"""The C++ host layer (include/apgk_adapters.hpp): compile tests/cpp/test_adapters.cpp against
libapgk.so.  Without a GPU it must refuse loudly (no CPU fallback); on a GPU it must pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, apgk_lib):
    exe = str(tmp_path / "test_adapters")
    libdir = os.path.join(ROOT, "allpathslg_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp"), "-L" + libdir, "-lapgk",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_cpp_adapters_compile_and_refuse_without_gpu(tmp_path, apgk_lib):
    import torch

    exe = _build(tmp_path, apgk_lib)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; see the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_adapters_on_gpu(tmp_path, apgk_lib):
    exe = _build(tmp_path, apgk_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "ADAPTERS OK" in r.stdout
