"""GPU: build and run the C++ host-side mirror of the reference API (host/sdr.hpp) against the oracle."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_host_mirror():
    import oracle_lib
    oracle_lib.build()
    exe = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
    src = exe + ".cpp"
    lib = os.path.join(ROOT, "unnamed-rust-sdr_b200", "lib")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, src, "-L" + lib, "-lsdr_b200",
                           "-L" + os.path.join(ROOT, "oracle"), "-lsdr_oracle",
                           "-Wl,-rpath," + lib, "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    return exe


def test_cpp_host_mirror():
    exe = build_host_mirror()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HOST MIRROR OK" in r.stdout, r.stdout + r.stderr
