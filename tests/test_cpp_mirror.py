"""The C++ host mirror of the reference API (scroll-prover_b200/halo2_b200.hpp) exercised through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_halo2_b200.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_halo2_b200")


def _build():
    libdir = os.path.join(ROOT, "scroll-prover_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", BIN, SRC, "-L" + libdir, "-lb200zk",
                           "-Wl,-rpath," + libdir])


def test_cpp_mirror_compiles_against_the_abi():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_mirror_runs_on_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
