"""The C-ABI library loads without a GPU, exports every symbol include/m2s.h declares, and fails loudly (no CPU
fallback) when no CUDA device is usable. No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "m2s.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"M2S_API[^;(]*?\b(m2s_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["m2s_create", "m2s_destroy", "m2s_generate_grid_sdf", "m2s_generate_sdf",
                 "m2s_generate_grid_sdf_device", "m2s_generate_sdf_device", "m2s_synchronize", "m2s_last_error",
                 "m2s_last_timings", "m2s_launch_count", "m2s_expand_topology", "m2s_grid_from_bounding_box",
                 "m2s_mesh_create", "m2s_mesh_destroy", "m2s_mesh_grid_sdf", "m2s_mesh_grid_sdf_device", "m2s_mesh_sdf",
                 "m2s_host_alloc", "m2s_host_register", "m2s_device_alloc", "m2s_ipc_export", "m2s_ipc_open",
                 "m2s_set_option", "m2s_last_error_copy", "m2s_last_timings_device"]:
        assert must in syms


def test_library_exports_every_declared_symbol(m2s):
    L = m2s.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"libm2s.so does not export {name}"
    assert L.m2s_abi_version() == 2


def test_library_exports_only_the_abi(m2s):
    out = subprocess.run(["nm", "-D", "--defined-only", m2s.LIB_PATH], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("m2s_") for n in names), names


def test_library_does_not_link_the_oracle_or_python(m2s):
    out = subprocess.run(["ldd", m2s.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "python" not in out and "torch" not in out


def test_library_contains_no_library_kernels(m2s):
    """Every kernel of the product is hand-written: the radix sort and the scan are csrc/m2s_sort.cuh, no cub /
    thrust instantiation is linked in, and no source of the product includes them."""
    out = subprocess.run(["nm", "-C", m2s.LIB_PATH], capture_output=True, text=True).stdout
    assert "k_sort_onesweep" in out and "k_exclusive_scan_u32" in out
    assert "cub::" not in out and "thrust::" not in out
    csrc = os.path.join(ROOT, "mesh_to_sdf_b200", "csrc")
    for f in os.listdir(csrc):
        src = open(os.path.join(csrc, f), errors="ignore").read()
        assert not re.search(r"#include\s*<(cub|thrust)/", src), f


def test_header_compiles_as_c():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="needs a machine without a GPU")
def test_no_cpu_fallback(m2s):
    with pytest.raises(m2s.M2SError) as e:
        m2s.Context()
    assert e.value.status == m2s.M2S_ENODEV
    with pytest.raises(m2s.M2SError):
        m2s.generate_grid_sdf(np.zeros((3, 3), np.float32), m2s.Topology.TriangleList(None),
                              m2s.Grid([0, 0, 0], [1, 1, 1], [2, 2, 2]))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mesh_to_sdf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "m2s_oracle" not in src and "libm2s_oracle" not in src, f
