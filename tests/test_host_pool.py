"""cticp::HostPool (the host team of RegisterFrame's min/max + packing pass): built with plain g++ against the engine
library and run on CPU — coverage of ParallelFor, concurrency of ParallelRegion, poll → sleep → wake transitions."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_pool_test")


def test_host_pool():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cuda_inc = "/usr/local/cuda/include"
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-I", cuda_inc, "-I", os.path.join(ROOT, "ct_icp_b200", "csrc"),
           os.path.join(ROOT, "tests", "cpp", "host_pool_test.cpp"), "-L", os.path.join(ROOT, "ct_icp_b200"),
           "-lcticp_b200", "-Wl,-rpath," + os.path.join(ROOT, "ct_icp_b200"), "-lpthread", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST POOL OK" in r.stdout, r.stdout + r.stderr


def test_peer_exchange_protocol_model():
    """Design check of the multi-GPU exchange (csrc/peer_exchange.cuh): LL words + two parity slots per source rank are
    race-free for any interleaving and give every rank the same rank-ordered sum (host model, threads = ranks)."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(ROOT, "tests", "cpp", "peer_protocol_model")
    r = subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "peer_protocol_model.cpp"),
                        "-lpthread", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PEER PROTOCOL OK" in r.stdout, r.stdout + r.stderr


def test_engine_math_header_on_the_host():
    """csrc/se3.cuh + device_map.cuh compile with plain g++: the polynomial sin / cos, the slerp with hoisted constants, the
    reciprocal voxel coordinate, and models of the device's conversion tricks against their exact definitions."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(ROOT, "tests", "cpp", "se3_math_test")
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-I", "/usr/local/cuda/include", "-I",
                        os.path.join(ROOT, "ct_icp_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "se3_math_test.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SE3 MATH OK" in r.stdout, r.stdout + r.stderr
