"""The C++ facade (ct_icp::Odometry over the C ABI) compiles with plain g++ against include/cticp.h and behaves:
without a GPU it reports CTICP_ERR_NO_DEVICE (exit 42); with one it runs the reference's box-scene integration test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def _build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
           os.path.join(ROOT, "ct_icp_b200", "include"), os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
           "-L", os.path.join(ROOT, "ct_icp_b200"), "-lcticp_b200", "-Wl,-rpath," + os.path.join(ROOT, "ct_icp_b200"),
           "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_facade_compiles_and_fails_loudly_without_gpu():
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 42, (r.returncode, r.stdout, r.stderr)
    assert "NO_DEVICE" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["CERES", "GN"])
def test_facade_box_scene(solver):
    exe = _build()
    r = subprocess.run([exe, solver], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FACADE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
