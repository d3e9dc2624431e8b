"""Builds tests/cpp/test_facade.cpp (the reference's doc-tests and unit tests against the C++ facade
include/mesh_to_sdf.hpp) with g++, links libm2s.so and runs it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_facade_test(tmpdir):
    exe = os.path.join(tmpdir, "test_facade")
    libdir = os.path.join(ROOT, "mesh_to_sdf_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), "-o", exe, "-L", libdir, "-l:libm2s.so",
           f"-Wl,-rpath,{libdir}", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_facade_compiles_and_links(tmp_path):
    build_facade_test(str(tmp_path))


@pytest.mark.gpu
def test_cpp_facade_runs(tmp_path):
    exe = build_facade_test(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all tests passed" in r.stdout
