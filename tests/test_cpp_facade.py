"""Builds and runs the C++ facade smoke test (include/otters.hpp over the C ABI)."""
import os
import subprocess

import pytest

from helpers import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")


def build():
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L" + os.path.join(ROOT, "otters_b200"), "-lotters_b200", "-Wl,-rpath," + os.path.join(ROOT, "otters_b200")]
    subprocess.run(cmd, check=True)


def test_cpp_facade_host_logic():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "FACADE_TEST_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_facade_on_device():
    build()
    r = subprocess.run([EXE, "--device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "FACADE_TEST_OK (host+device)" in r.stdout, r.stdout + r.stderr
