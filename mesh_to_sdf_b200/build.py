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
SOURCES = ["m2s_build.cu", "m2s_grid.cu", "m2s_points.cu", "m2s_post.cu", "m2s_api.cu"]
HEADERS = ["m2s_geom.cuh", "m2s_search.cuh", "m2s_sort.cuh", "m2s_internal.h", os.path.join("..", "..", "include", "m2s.h")]

NVCC_COMPILE = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]
NVCC_LINK = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static", "-lpthread",
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


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra: list | None = None) -> str:
    """Compiles every translation unit in parallel (one nvcc each) and links them. out / extra: alternative output
    path and extra nvcc flags for development A/B builds (loaded with M2S_LIB=<path>)."""
    if out is None and not force and not is_stale():
        return SO
    extra = list(extra or []) + os.environ.get("M2S_NVCC_EXTRA", "").split()  # development: e.g. -DRUN_MIN_BLOCKS=6
    target = out or SO
    objdir = os.path.join(HERE, "..", "build", "obj_" + os.path.basename(target).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_COMPILE + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = "", False
    for cmd, obj, p in procs:
        o, _ = p.communicate()
        log += " ".join(cmd) + "\n" + o
        failed |= p.returncode != 0
    if not failed:
        cmd = [nvcc] + NVCC_LINK + ["-o", target] + [obj for _, obj, _ in procs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + r.stdout + r.stderr
        failed |= r.returncode != 0
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libm2s.so")
    if verbose:
        print(log)
    return target


if __name__ == "__main__":
    _out = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    _extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=_out[0] if _out else None,
                extra=_extra))
