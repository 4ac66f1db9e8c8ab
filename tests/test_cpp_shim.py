"""The C++ drop-in shim (include/pipsgpu_shim.hpp) driven through DoubleLinearSolver* virtual calls by
tests/cpp/shim_test.cpp. CPU: it compiles against the C ABI, links, and fails loudly without a device.
GPU (-m gpu): residual checks of the plugin-level and system-level paths pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(pkg, tmp_path):
    pkg.capi.lib()
    libdir = os.path.join(ROOT, "pips-ipmpp_b200", "csrc")
    exe = str(tmp_path / "shim_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), "-o", exe, "-L", libdir, "-lpipsgpu",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_shim_compiles_links_and_refuses_to_run_without_gpu(pkg, tmp_path):
    import torch
    exe = _build(pkg, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (covered by the gpu test)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_shim_drives_the_cuda_path(pkg, tmp_path):
    exe = _build(pkg, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim_test PASS" in r.stdout
