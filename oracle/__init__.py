"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/libm2s_oracle.so`` (C++ restatement of the reference's CPU algorithm, see
``m2s_oracle_geo.hpp`` / ``m2s_oracle.cpp`` for the reference file:line citations).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package — as the checker / the timed CPU reference, never as the product path.
``mesh_to_sdf_b200`` never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libm2s_oracle.so")

_f = C.POINTER(C.c_float)
_u32 = C.POINTER(C.c_uint32)
_u64 = C.POINTER(C.c_uint64)
_d = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``). Returns the path of the shared object."""
    srcs = [os.path.join(_HERE, n) for n in ("m2s_oracle.cpp", "m2s_oracle_geo.hpp", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.m2s_oracle_point_triangle_distance.restype = C.c_float
        L.m2s_oracle_point_triangle_distance2.restype = C.c_float
        L.m2s_oracle_point_triangle_signed_distance.restype = C.c_float
        L.m2s_oracle_grid_cell_idx.restype = C.c_uint64
        L.m2s_oracle_expand_topology.restype = C.c_uint64
        L.m2s_oracle_compare_distances.argtypes = [C.c_float, C.c_float]
        L.m2s_oracle_approx_eq_f32.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float]
        L.m2s_oracle_grid_cell_coords.argtypes = [_u64, C.c_uint64, _u64]
        L.m2s_oracle_expand_topology.argtypes = [C.c_int, _u32, C.c_uint64, C.c_uint64, _u32]
        L.m2s_oracle_generate_sdf.argtypes = [_f, C.c_uint64, _u32, C.c_uint64, _f, C.c_uint64, C.c_int,
                                              C.c_int, C.c_int, _f]
        L.m2s_oracle_generate_sdf_tree.argtypes = [_f, C.c_uint64, _u32, C.c_uint64, _f, C.c_uint64, C.c_int,
                                                   C.c_int, C.c_int, _f, _d]
        L.m2s_oracle_grid_cells_exact.argtypes = [_f, C.c_uint64, _u32, C.c_uint64, _f, _f, _u64, C.c_int,
                                                  _u64, C.c_uint64, C.c_int, _f]
        L.m2s_oracle_generate_grid_sdf_faithful.argtypes = [_f, C.c_uint64, _u32, C.c_uint64, _f, _f, _u64,
                                                            C.c_int, C.c_int, _f, _d, _u64]
        _lib = L
    return _lib


# enum orders of the reference: SignMethod lib.rs:204-216, AccelerationMethod lib.rs:224-239
RAYCAST, NORMAL = 0, 1
ACCEL_NONE, ACCEL_BVH, ACCEL_RTREE, ACCEL_RTREE_BVH = 0, 1, 2, 3


class OracleError(RuntimeError):
    pass


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a, t):
    return a.ctypes.data_as(t)


def _v3(a):
    return _f32(a, (3,))


def _check(rc):
    if rc == 0:
        return
    raise OracleError({1: "invalid argument", 2: "index out of bounds (reference panics)",
                       3: "NaN distance (reference panics)", 7: "empty mesh"}.get(rc, f"rc={rc}"))


def hardware_threads() -> int:
    return int(lib().m2s_oracle_hardware_threads())


# ---- leaf arithmetic -------------------------------------------------------------------------------
def triangle_bounding_box(a, b, c):
    mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().m2s_oracle_triangle_bounding_box(_p(_v3(a), _f), _p(_v3(b), _f), _p(_v3(c), _f), _p(mn, _f), _p(mx, _f))
    return mn, mx


def closest_point_segment(p, a, b):
    out = np.zeros(3, np.float32)
    lib().m2s_oracle_closest_point_segment(_p(_v3(p), _f), _p(_v3(a), _f), _p(_v3(b), _f), _p(out, _f))
    return out


def closest_point_triangle(p, a, b, c):
    out = np.zeros(3, np.float32)
    lib().m2s_oracle_closest_point_triangle(_p(_v3(p), _f), _p(_v3(a), _f), _p(_v3(b), _f), _p(_v3(c), _f),
                                            _p(out, _f))
    return out


def point_triangle_distance(p, a, b, c) -> float:
    return float(lib().m2s_oracle_point_triangle_distance(_p(_v3(p), _f), _p(_v3(a), _f), _p(_v3(b), _f),
                                                          _p(_v3(c), _f)))


def point_triangle_distance2(p, a, b, c) -> float:
    return float(lib().m2s_oracle_point_triangle_distance2(_p(_v3(p), _f), _p(_v3(a), _f), _p(_v3(b), _f),
                                                           _p(_v3(c), _f)))


def point_triangle_signed_distance(p, a, b, c) -> float:
    return float(lib().m2s_oracle_point_triangle_signed_distance(_p(_v3(p), _f), _p(_v3(a), _f),
                                                                 _p(_v3(b), _f), _p(_v3(c), _f)))


def ray_triangle_intersection_aligned(o, v0, v1, v2, axis: int):
    """Returns t (float) for ``Some(t)`` and None for ``None`` (geo.rs:165-216)."""
    t = C.c_float(0)
    hit = lib().m2s_oracle_ray_triangle_intersection_aligned(_p(_v3(o), _f), _p(_v3(v0), _f), _p(_v3(v1), _f),
                                                             _p(_v3(v2), _f), int(axis), C.byref(t))
    return float(t.value) if hit else None


def approx_eq_f32(a, b, ulps, eps) -> bool:
    return bool(lib().m2s_oracle_approx_eq_f32(float(a), float(b), int(ulps), float(eps)))


def compare_distances(a, b) -> int:
    """-1 Less, 0 Equal, 1 Greater (lib.rs:242-259). Raises where the reference panics."""
    r = lib().m2s_oracle_compare_distances(float(a), float(b))
    if r == 2:
        raise OracleError("NaN distance")
    return r


# ---- Grid ------------------------------------------------------------------------------------------
def _cnt(count):
    return np.ascontiguousarray(count, dtype=np.uint64).reshape(3)


def grid_from_bounding_box(bmin, bmax, count):
    first, size = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().m2s_oracle_grid_from_bounding_box(_p(_v3(bmin), _f), _p(_v3(bmax), _f), _p(_cnt(count), _u64),
                                            _p(first, _f), _p(size, _f))
    return first, size


def grid_cell_idx(count, cell) -> int:
    return int(lib().m2s_oracle_grid_cell_idx(_p(_cnt(count), _u64), _p(_cnt(cell), _u64)))


def grid_cell_coords(count, idx):
    out = np.zeros(3, np.uint64)
    lib().m2s_oracle_grid_cell_coords(_p(_cnt(count), _u64), int(idx), _p(out, _u64))
    return [int(v) for v in out]


def grid_cell_center(first, size, count, cell):
    out = np.zeros(3, np.float32)
    lib().m2s_oracle_grid_cell_center(_p(_v3(first), _f), _p(_v3(size), _f), _p(_cnt(count), _u64),
                                      _p(_cnt(cell), _u64), _p(out, _f))
    return out


def grid_last_cell(first, size, count):
    out = np.zeros(3, np.float32)
    lib().m2s_oracle_grid_last_cell(_p(_v3(first), _f), _p(_v3(size), _f), _p(_cnt(count), _u64), _p(out, _f))
    return out


def grid_bounding_box(first, size, count):
    mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().m2s_oracle_grid_bounding_box(_p(_v3(first), _f), _p(_v3(size), _f), _p(_cnt(count), _u64),
                                       _p(mn, _f), _p(mx, _f))
    return mn, mx


def grid_snap(first, size, count, p):
    """Returns (inside: bool, cell: [x,y,z]) — SnapResult::Inside/Outside (grid.rs:145-170)."""
    out = np.zeros(3, np.uint64)
    r = lib().m2s_oracle_grid_snap(_p(_v3(first), _f), _p(_v3(size), _f), _p(_cnt(count), _u64),
                                   _p(_v3(p), _f), _p(out, _u64))
    return bool(r), [int(v) for v in out]


# ---- Topology --------------------------------------------------------------------------------------
def expand_topology(kind: int, indices, nv: int) -> np.ndarray:
    """Topology::get_triangles (lib.rs:175-193). kind 0 = TriangleList, 1 = TriangleStrip."""
    if indices is None:
        ip, n = None, 0
    else:
        indices = np.ascontiguousarray(indices, dtype=np.uint32).ravel()
        ip, n = _p(indices, _u32), indices.size
    cnt = lib().m2s_oracle_expand_topology(int(kind), ip, n, int(nv), None)
    out = np.zeros((cnt, 3), np.uint32)
    if cnt:
        lib().m2s_oracle_expand_topology(int(kind), ip, n, int(nv), _p(out, _u32))
    return out


# ---- drivers ---------------------------------------------------------------------------------------
def _mesh(verts, tris):
    verts = _f32(verts).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
    return verts, tris


def generate_sdf(verts, tris, queries, accel: int, sign: int = RAYCAST, threads: int = 0) -> np.ndarray:
    """Exact generate_sdf (lib.rs:291-311), every triangle is a candidate."""
    verts, tris = _mesh(verts, tris)
    queries = _f32(queries).reshape(-1, 3)
    if len(tris) == 0 and accel == ACCEL_RTREE_BVH:
        return np.zeros(0, np.float32)  # rtree_bvh.rs:104-106
    out = np.zeros(len(queries), np.float32)
    _check(lib().m2s_oracle_generate_sdf(_p(verts, _f), len(verts), _p(tris, _u32), len(tris), _p(queries, _f),
                                         len(queries), accel, sign, threads, _p(out, _f)))
    return out


def generate_sdf_tree(verts, tris, queries, accel: int, sign: int = RAYCAST, threads: int = 0):
    """Tree-accelerated CPU stand-in for Rtree / RtreeBvh / Bvh(Raycast). Returns (sdf, [build_ms, query_ms])."""
    verts, tris = _mesh(verts, tris)
    queries = _f32(queries).reshape(-1, 3)
    out = np.zeros(len(queries), np.float32)
    ms = np.zeros(2, np.float64)
    _check(lib().m2s_oracle_generate_sdf_tree(_p(verts, _f), len(verts), _p(tris, _u32), len(tris),
                                              _p(queries, _f), len(queries), accel, sign, threads, _p(out, _f),
                                              _p(ms, _d)))
    return out, ms


def grid_cells_exact(verts, tris, first, size, count, sign: int, cell_idx=None, threads: int = 0) -> np.ndarray:
    """Exact grid SDF (brute force over every triangle) at ``cell_idx`` (flat indices) or at every cell."""
    verts, tris = _mesh(verts, tris)
    count = _cnt(count)
    if cell_idx is None:
        n, ip = int(np.prod(count)), None
    else:
        cell_idx = np.ascontiguousarray(cell_idx, dtype=np.uint64).ravel()
        n, ip = cell_idx.size, _p(cell_idx, _u64)
    out = np.zeros(n, np.float32)
    _check(lib().m2s_oracle_grid_cells_exact(_p(verts, _f), len(verts), _p(tris, _u32), len(tris),
                                             _p(_v3(first), _f), _p(_v3(size), _f), _p(count, _u64), sign, ip, n,
                                             threads, _p(out, _f)))
    return out


def generate_grid_sdf_faithful(verts, tris, first, size, count, sign: int, threads: int = 0):
    """Step-by-step restatement of generate/grid.rs:265-378. Returns (sdf, phase_ms[3], steps[3])."""
    verts, tris = _mesh(verts, tris)
    count = _cnt(count)
    out = np.zeros(int(np.prod(count)), np.float32)
    ms = np.zeros(3, np.float64)
    steps = np.zeros(3, np.uint64)
    _check(lib().m2s_oracle_generate_grid_sdf_faithful(_p(verts, _f), len(verts), _p(tris, _u32), len(tris),
                                                       _p(_v3(first), _f), _p(_v3(size), _f), _p(count, _u64),
                                                       sign, threads, _p(out, _f), _p(ms, _d), _p(steps, _u64)))
    return out, ms, steps
