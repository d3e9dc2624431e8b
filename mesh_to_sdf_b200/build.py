"""Builds ``mesh_to_sdf_b200/libm2s.so`` (hand-written CUDA for sm_100a + the C ABI of include/m2s.h) in-tree with nvcc.

    python -m mesh_to_sdf_b200.build [--force] [--verbose]

The shared object has no Python / torch dependency: it links the CUDA runtime statically and exports only
the ``m2s_*`` symbols. It is git-ignored (history stays source-only) but travels to the GPU box with the
working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libm2s.so")
SOURCES = ["m2s_build.cu", "m2s_query.cu", "m2s_post.cu", "m2s_api.cu"]
HEADERS = ["m2s_geom.cuh", "m2s_internal.h", os.path.join("..", "..", "include", "m2s.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libm2s.so cannot be built (there is no CPU fallback)")
    return p


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """out: alternative output path for development A/B builds (loaded with M2S_LIB=<path>)."""
    if out is None and not force and not is_stale():
        return SO
    extra = os.environ.get("M2S_NVCC_EXTRA", "").split()  # development: e.g. -DPKT_MIN_BLOCKS=5
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", out or SO] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libm2s.so")
    if verbose:
        print(log)
    return out or SO


if __name__ == "__main__":
    _out = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=_out[0] if _out else None))
