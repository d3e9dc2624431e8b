"""pytest configuration: registers the ``gpu`` marker, makes the repo importable, builds the oracle (test
infrastructure, gcc) and offers shared fixtures. Tests marked ``gpu`` call the CUDA path through the C ABI and
compare it with the oracle; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def m2s():
    import mesh_to_sdf_b200 as m
    return m


@pytest.fixture(scope="session")
def ctx(m2s):
    c = m2s.Context()
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def mesh_diag(verts) -> float:
    v = np.asarray(verts, np.float64).reshape(-1, 3)
    return float(np.linalg.norm(v.max(axis=0) - v.min(axis=0)))
