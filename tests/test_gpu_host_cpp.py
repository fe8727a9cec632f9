"""The C++ host mirror (host/phant_host.hpp) over the C ABI: compiles on CPU, runs the reference's tests on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "host_test")


def build():
    lib = os.path.join(ROOT, "phant_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", BIN, os.path.join(ROOT, "host", "host_test.cpp"), f"-L{lib}", "-lphantgpu",
                    f"-Wl,-rpath,{lib}"], check=True)


def test_host_mirror_compiles_and_links():
    build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_reference_tests_through_the_cpp_mirror():
    build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
