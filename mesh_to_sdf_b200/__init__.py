"""mesh_to_sdf_b200 — host-side mirror of the ``mesh_to_sdf`` crate's public API over ``libm2s.so``.

The names, argument meaning and error behaviour follow the reference (``mesh_to_sdf/src/lib.rs:146-311``,
``src/grid.rs``): ``generate_sdf``, ``generate_grid_sdf``, ``Grid``, ``SnapResult``, ``Topology``,
``SignMethod``, ``AccelerationMethod``. Every compute call goes through the C ABI of ``include/m2s.h`` into
hand-written CUDA (sm_100a). There is no CPU fallback: if ``libm2s.so`` is missing or no CUDA device is
usable, the call raises.

This is the binding used by tests/ and bench.py; the Rust facade a crate user would link is in
``rust/mesh_to_sdf``.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import threading
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# M2S_LIB: development override (A/B builds of the same library, e.g. with -DM2S_STATS_BUILD)
LIB_PATH = os.environ.get("M2S_LIB") or os.path.join(_HERE, "libm2s.so")

__all__ = [
    "generate_sdf", "generate_grid_sdf", "Grid", "SnapResult", "Topology", "SignMethod", "AccelerationMethod",
    "M2SError", "Context", "Mesh", "lib", "LIB_PATH", "host_alloc", "host_register", "host_unregister",
]

_f = C.POINTER(C.c_float)
_u32 = C.POINTER(C.c_uint32)
_u64 = C.POINTER(C.c_uint64)


class M2SError(RuntimeError):
    """The reference panics (``lib.rs:257``, slice-index panics, ``rtree.rs:117``); this binding raises."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


M2S_OK, M2S_EINVAL, M2S_EINDEX, M2S_ENAN, M2S_ECUDA, M2S_ENCCL, M2S_ENODEV, M2S_EEMPTY = range(8)
_STATUS_NAMES = {0: "M2S_OK", 1: "M2S_EINVAL", 2: "M2S_EINDEX", 3: "M2S_ENAN", 4: "M2S_ECUDA", 5: "M2S_ENCCL",
                 6: "M2S_ENODEV", 7: "M2S_EEMPTY"}


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("build_ms", C.c_float), ("sign_ms", C.c_float), ("dist_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("seed_ms", C.c_float), ("host_path", C.c_int)]

    def as_dict(self):
        d = {k: float(getattr(self, k)) for k, _ in self._fields_ if k != "host_path"}
        d["host_path"] = HOST_PATH_NAMES.get(int(self.host_path), str(int(self.host_path)))
        return d


# m2s_host_path_taken (m2s.h): how the result of the last call reached its host destination
HOST_PATH_NAMES = {0: "device", 1: "zerocopy", 2: "pipelined", 3: "staged", 4: "registered"}
# m2s_set_option keys / values (m2s.h)
OPT_BUILD_MODE, OPT_HOST_PATH, OPT_COPY_THREADS, OPT_RAY_BINS, OPT_BALANCE, OPT_RUN_LENGTH = 1, 2, 3, 4, 5, 6
BUILD_REPLICATED, BUILD_BROADCAST = 0, 1
HOST_AUTO, HOST_STAGED, HOST_PIPELINED, HOST_REGISTER = 0, 1, 2, 3


_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """Loads ``libm2s.so``. Raises if it has not been built — there is no fallback path."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise M2SError(M2S_ENODEV, f"{LIB_PATH} is missing: build it with `python -m mesh_to_sdf_b200.build` "
                                           "(there is no CPU fallback)")
            L = C.CDLL(LIB_PATH)
            vp = C.c_void_p
            L.m2s_abi_version.restype = C.c_int
            L.m2s_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
            L.m2s_create_on_stream.argtypes = [C.c_int, vp, C.POINTER(vp)]
            L.m2s_destroy.argtypes = [vp]
            L.m2s_destroy.restype = None
            L.m2s_last_error.argtypes = [vp]
            L.m2s_last_error.restype = C.c_char_p
            L.m2s_last_timings.argtypes = [vp, C.POINTER(Timings)]
            L.m2s_last_timings_device.argtypes = [vp, C.c_int, C.POINTER(Timings)]
            L.m2s_last_error_copy.argtypes = [vp, C.c_char_p, C.c_size_t]
            L.m2s_set_option.argtypes = [vp, C.c_int, C.c_int64]
            L.m2s_mesh_create.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(vp)]
            L.m2s_mesh_create_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(vp)]
            L.m2s_mesh_destroy.argtypes = [vp]
            L.m2s_mesh_destroy.restype = None
            L.m2s_mesh_grid_sdf.argtypes = [vp, vp, _f, _f, _u64, C.c_int, C.c_uint64, C.c_uint64, vp]
            L.m2s_mesh_grid_sdf_device.argtypes = [vp, vp, _f, _f, _u64, C.c_int, C.c_uint64, C.c_uint64, vp]
            L.m2s_mesh_sdf.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp]
            L.m2s_mesh_sdf_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp]
            L.m2s_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
            L.m2s_host_free.argtypes = [vp]
            L.m2s_host_free.restype = None
            L.m2s_host_register.argtypes = [vp, C.c_size_t]
            L.m2s_host_unregister.argtypes = [vp]
            L.m2s_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
            L.m2s_device_free.argtypes = [vp, vp]
            L.m2s_ipc_export.argtypes = [vp, vp, C.c_char_p]
            L.m2s_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
            L.m2s_ipc_close.argtypes = [vp, vp]
            L.m2s_launch_count.argtypes = [vp]
            L.m2s_launch_count.restype = C.c_uint64
            L.m2s_device_count.argtypes = [vp]
            L.m2s_synchronize.argtypes = [vp]
            L.m2s_generate_grid_sdf.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, _f, _f, _u64, C.c_int, vp]
            L.m2s_generate_grid_sdf_slab.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, _f, _f, _u64, C.c_int,
                                                     C.c_uint64, C.c_uint64, vp]
            L.m2s_debug_stats.argtypes = [vp, _u64]
            L.m2s_generate_sdf.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, C.c_int, C.c_int, vp]
            L.m2s_generate_grid_sdf_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, _f, _f, _u64, C.c_int,
                                                       C.c_uint64, C.c_uint64, vp]
            L.m2s_generate_sdf_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, C.c_int,
                                                  C.c_int, vp]
            L.m2s_grid_order.argtypes = [vp, vp, C.c_uint64, vp, vp]
            L.m2s_grid_order_device.argtypes = [vp, vp, C.c_uint64, vp, vp]
            L.m2s_sample_grid_sdf.argtypes = [vp, vp, _f, _f, _u64, vp, C.c_uint64, C.c_int, C.c_float, vp]
            L.m2s_sample_grid_sdf_device.argtypes = [vp, vp, _f, _f, _u64, vp, C.c_uint64, C.c_int, C.c_float, vp]
            L.m2s_expand_topology.argtypes = [C.c_int, vp, C.c_int, C.c_uint64, C.c_uint64, _u32]
            L.m2s_expand_topology.restype = C.c_uint64
            L.m2s_grid_from_bounding_box.argtypes = [_f, _f, _u64, _f, _f]
            L.m2s_grid_from_bounding_box.restype = None
            _lib = L
    return _lib


# ---- enums (declaration order of the reference) ----------------------------------------------------------
class SignMethod(enum.IntEnum):
    """``lib.rs:204-216``. ``Raycast`` is the default."""
    Raycast = 0
    Normal = 1


@dataclass(frozen=True)
class AccelerationMethod:
    """``lib.rs:224-239``: ``None(SignMethod)``, ``Bvh(SignMethod)``, ``Rtree``, ``RtreeBvh`` (default)."""
    kind: int
    sign: SignMethod = SignMethod.Raycast

    @staticmethod
    def none(sign: SignMethod = SignMethod.Raycast) -> "AccelerationMethod":
        return AccelerationMethod(0, SignMethod(sign))

    @staticmethod
    def bvh(sign: SignMethod = SignMethod.Raycast) -> "AccelerationMethod":
        return AccelerationMethod(1, SignMethod(sign))

    @staticmethod
    def rtree() -> "AccelerationMethod":
        return AccelerationMethod(2)

    @staticmethod
    def rtree_bvh() -> "AccelerationMethod":
        return AccelerationMethod(3)

    @staticmethod
    def default() -> "AccelerationMethod":
        return AccelerationMethod(3)


AccelerationMethod.Rtree = AccelerationMethod(2)
AccelerationMethod.RtreeBvh = AccelerationMethod(3)
AccelerationMethod.None_ = AccelerationMethod.none
AccelerationMethod.Bvh = AccelerationMethod.bvh


@dataclass(frozen=True)
class Topology:
    """``lib.rs:151-167``: ``TriangleList(Option<&[I]>)`` / ``TriangleStrip(Option<&[I]>)``; I = u16 or u32."""
    kind: int
    indices: Optional[np.ndarray] = None

    @staticmethod
    def TriangleList(indices=None) -> "Topology":
        return Topology(0, None if indices is None else _as_indices(indices))

    @staticmethod
    def TriangleStrip(indices=None) -> "Topology":
        return Topology(1, None if indices is None else _as_indices(indices))

    def get_triangles(self, n_vertices: int) -> np.ndarray:
        """``Topology::get_triangles`` (``lib.rs:175-193``) → uint32 [nt, 3]."""
        if self.kind == 0 and self.indices is not None and self.indices.dtype == np.uint32:
            # a u32 triangle list IS its own expansion (tuples() drops a trailing partial triple): no copy
            n3 = self.indices.size // 3
            return self.indices[:3 * n3].reshape(n3, 3)
        L = lib()
        if self.indices is None:
            ptr, nbytes, n = None, 4, 0
        else:
            ptr, nbytes, n = self.indices.ctypes.data, self.indices.dtype.itemsize, self.indices.size
        cnt = L.m2s_expand_topology(self.kind, ptr, nbytes, n, n_vertices, None)
        out = np.empty((cnt, 3), np.uint32)
        if cnt:
            L.m2s_expand_topology(self.kind, ptr, nbytes, n, n_vertices, out.ctypes.data_as(_u32))
        return out


def _as_indices(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.uint16:
        return np.ascontiguousarray(a).ravel()
    if a.size and (a.min() < 0 or a.max() > 0xFFFFFFFF):
        raise M2SError(M2S_EINVAL, "indices must fit u32 (I: Into<u32>)")
    return np.ascontiguousarray(a, dtype=np.uint32).ravel()


# ---- Grid (src/grid.rs) ------------------------------------------------------------------------------------
@dataclass(frozen=True)
class SnapResult:
    """``grid.rs:10-17``: ``Inside(cell)`` / ``Outside(cell)``."""
    inside: bool
    cell: tuple


class Grid:
    """``grid.rs:30-170``. Host-only value type; all arithmetic in float32 like the reference."""

    def __init__(self, first_cell, cell_size, cell_count):
        self.first_cell = np.asarray(first_cell, np.float32).reshape(3).copy()
        self.cell_size = np.asarray(cell_size, np.float32).reshape(3).copy()
        cc = [int(c) for c in cell_count]
        if len(cc) != 3 or any(c < 0 for c in cc):
            raise M2SError(M2S_EINVAL, "cell_count must be three non-negative integers")
        self.cell_count = tuple(cc)

    new = classmethod(lambda cls, first_cell, cell_size, cell_count: cls(first_cell, cell_size, cell_count))

    @classmethod
    def from_bounding_box(cls, bbox_min, bbox_max, cell_count) -> "Grid":
        """``grid.rs:59-74`` (through the C helper the Rust facade shares)."""
        mn = np.asarray(bbox_min, np.float32).reshape(3).copy()
        mx = np.asarray(bbox_max, np.float32).reshape(3).copy()
        cc = np.asarray(cell_count, np.uint64).reshape(3).copy()
        first, size = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib().m2s_grid_from_bounding_box(mn.ctypes.data_as(_f), mx.ctypes.data_as(_f), cc.ctypes.data_as(_u64),
                                         first.ctypes.data_as(_f), size.ctypes.data_as(_f))
        return cls(first, size, [int(c) for c in cc])

    def get_first_cell(self):
        return self.first_cell.copy()

    def get_last_cell(self):  # grid.rs:82-88 (first + count * size, as in the reference)
        n = np.asarray(self.cell_count, np.float32)
        return (self.first_cell + n * self.cell_size).astype(np.float32)

    def get_cell_size(self):
        return self.cell_size.copy()

    def get_cell_count(self):
        return self.cell_count

    def get_total_cell_count(self) -> int:
        return self.cell_count[0] * self.cell_count[1] * self.cell_count[2]

    def get_bounding_box(self):  # grid.rs:110-119
        mn = (self.first_cell - self.cell_size * np.float32(0.5)).astype(np.float32)
        mx = (mn + np.asarray(self.cell_count, np.float32) * self.cell_size).astype(np.float32)
        return mn, mx

    def get_cell_idx(self, cell: Sequence[int]) -> int:  # grid.rs:122-124
        return cell[2] + cell[1] * self.cell_count[2] + cell[0] * self.cell_count[1] * self.cell_count[2]

    def get_cell_integer_coordinates(self, cell_idx: int):  # grid.rs:127-132
        z = cell_idx % self.cell_count[2]
        y = (cell_idx // self.cell_count[2]) % self.cell_count[1]
        x = cell_idx // (self.cell_count[1] * self.cell_count[2])
        return [x, y, z]

    def get_cell_center(self, cell: Sequence[int]):  # grid.rs:135-141
        c = np.asarray(cell, np.float32)
        return (self.first_cell + c * self.cell_size).astype(np.float32)

    def snap_point_to_grid(self, point) -> SnapResult:  # grid.rs:145-170
        p = np.asarray(point, np.float32).reshape(3)
        with np.errstate(all="ignore"):
            q = np.floor((p - self.get_bounding_box()[0]).astype(np.float32) / self.cell_size)
        cell, res = [], []
        for i in range(3):
            v = float(q[i])
            c = 0 if v != v else int(max(min(v, 2.0 ** 63 - 1), -2.0 ** 63))  # `as isize` saturates, NaN -> 0
            cell.append(c)
            res.append(min(max(c, 0), self.cell_count[i] - 1))
        return SnapResult(cell == res, tuple(res))

    def __eq__(self, other):
        return (isinstance(other, Grid) and np.array_equal(self.first_cell, other.first_cell) and
                np.array_equal(self.cell_size, other.cell_size) and self.cell_count == other.cell_count)

    def __repr__(self):
        return f"Grid(first_cell={self.first_cell.tolist()}, cell_size={self.cell_size.tolist()}, cell_count={list(self.cell_count)})"


# ---- context ------------------------------------------------------------------------------------------------
class Context:
    """Owns an ``m2s_ctx``: device(s), stream(s), scratch arenas. One call at a time per context."""

    def __init__(self, devices: Optional[Sequence[int]] = None, stream: Optional[int] = None):
        L = lib()
        self._h = C.c_void_p()
        # one call at a time per context, and the error text is read before another thread's call can replace it
        self._lock = threading.RLock()
        if stream is not None:
            dev = 0 if not devices else int(devices[0])
            rc = L.m2s_create_on_stream(dev, C.c_void_p(stream), C.byref(self._h))
        elif devices:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = L.m2s_create(arr, len(devices), C.byref(self._h))
        else:
            rc = L.m2s_create(None, 0, C.byref(self._h))
        if rc != M2S_OK:
            self._h = C.c_void_p()
            raise M2SError(rc, "m2s_create failed: no usable CUDA device (libm2s has no CPU fallback)"
                           if rc == M2S_ENODEV else "m2s_create failed")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().m2s_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != M2S_OK:
            buf = C.create_string_buffer(512)
            lib().m2s_last_error_copy(self._h, buf, len(buf))
            raise M2SError(rc, buf.value.decode(errors="replace"))

    def _call(self, fn, *args):
        with self._lock:
            self._check(fn(self._h, *args))

    def set_option(self, option: int, value: int):
        """``m2s_set_option``: OPT_BUILD_MODE / OPT_HOST_PATH / OPT_COPY_THREADS."""
        self._call(lib().m2s_set_option, int(option), int(value))

    @property
    def launch_count(self) -> int:
        return int(lib().m2s_launch_count(self._h))

    @property
    def device_count(self) -> int:
        return int(lib().m2s_device_count(self._h))

    def timings(self, device_index: int = 0) -> dict:
        t = Timings()
        lib().m2s_last_timings_device(self._h, int(device_index), C.byref(t))
        return t.as_dict()

    def synchronize(self):
        self._call(lib().m2s_synchronize)

    # host-buffer entry points (numpy in, numpy out)
    def grid_sdf(self, verts: np.ndarray, tris: np.ndarray, grid: Grid, sign: int, out: Optional[np.ndarray] = None):
        verts, tris = _mesh_arrays(verts, tris)
        out = _out_array(out, grid.get_total_cell_count())
        cc = np.asarray(grid.cell_count, np.uint64)
        self._call(lib().m2s_generate_grid_sdf, verts.ctypes.data, len(verts), tris.ctypes.data, len(tris),
                   grid.first_cell.ctypes.data_as(_f), grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64),
                   int(sign), out.ctypes.data)
        return out

    def grid_sdf_slab(self, verts: np.ndarray, tris: np.ndarray, grid: Grid, sign: int, x_begin: int, x_end: int,
                      out: Optional[np.ndarray] = None):
        """Cells x in [x_begin, x_end) of ``grid`` (the per-rank call of a one-process-per-GPU deployment)."""
        verts, tris = _mesh_arrays(verts, tris)
        out = _out_array(out, max(0, x_end - x_begin) * grid.cell_count[1] * grid.cell_count[2])
        cc = np.asarray(grid.cell_count, np.uint64)
        self._call(lib().m2s_generate_grid_sdf_slab, verts.ctypes.data, len(verts), tris.ctypes.data, len(tris),
                   grid.first_cell.ctypes.data_as(_f), grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64),
                   int(sign), x_begin, x_end, out.ctypes.data)
        return out

    def debug_stats(self):
        a = (C.c_uint64 * 4)()
        lib().m2s_debug_stats(self._h, a)
        return [int(x) for x in a]

    def sdf(self, verts: np.ndarray, tris: np.ndarray, queries: np.ndarray, accel: int, sign: int,
            out: Optional[np.ndarray] = None):
        verts, tris = _mesh_arrays(verts, tris)
        queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        out = _out_array(out, len(queries))
        self._call(lib().m2s_generate_sdf, verts.ctypes.data, len(verts), tris.ctypes.data, len(tris),
                   queries.ctypes.data, len(queries), int(accel), int(sign), out.ctypes.data)
        return out

    # device-buffer entry points (raw device pointers as ints; enqueue only)
    def grid_sdf_device(self, d_verts: int, nv: int, d_tris: int, nt: int, grid: Grid, sign: int, x_begin: int,
                        x_end: int, d_out: int):
        cc = np.asarray(grid.cell_count, np.uint64)
        self._call(lib().m2s_generate_grid_sdf_device, d_verts, nv, d_tris, nt, grid.first_cell.ctypes.data_as(_f),
                   grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64), int(sign), x_begin, x_end, d_out)

    def sdf_device(self, d_verts: int, nv: int, d_tris: int, nt: int, d_queries: int, nq: int, accel: int, sign: int,
                   d_out: int):
        self._call(lib().m2s_generate_sdf_device, d_verts, nv, d_tris, nt, d_queries, nq, int(accel), int(sign), d_out)

    # mesh handles: upload + build once, query many times
    def mesh(self, verts: np.ndarray, tris: np.ndarray) -> "Mesh":
        return Mesh(self, verts, tris)

    def mesh_device(self, d_verts: int, nv: int, d_tris: int, nt: int) -> "Mesh":
        return Mesh(self, None, None, device_ptrs=(d_verts, nv, d_tris, nt))

    # device memory that other processes can map (one process per GPU): see m2s.h
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._call(lib().m2s_device_alloc, nbytes, C.byref(p))
        return int(p.value)

    def device_free(self, d_ptr: int):
        self._call(lib().m2s_device_free, d_ptr)

    def ipc_export(self, d_ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._call(lib().m2s_ipc_export, d_ptr, buf)
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._call(lib().m2s_ipc_open, C.create_string_buffer(bytes(handle), 64), C.byref(p))
        return int(p.value)

    def ipc_close(self, d_ptr: int):
        self._call(lib().m2s_ipc_close, d_ptr)

    # post-passes on a finished grid (what the reference's in-repo caller runs next, mesh_to_sdf_client/src/sdf.rs)
    def grid_order(self, sdf, want_order: bool = True, want_minmax: bool = True):
        """(ordered cell indices by ascending distance (stable, f32::total_cmp), (min, max)) - sdf.rs:65-68, :123."""
        sdf = np.ascontiguousarray(sdf, np.float32).reshape(-1)
        order = np.empty(len(sdf), np.uint32) if want_order else None
        mm = np.zeros(2, np.float32) if want_minmax else None
        self._call(lib().m2s_grid_order, sdf.ctypes.data, len(sdf), order.ctypes.data if want_order else None,
                   mm.ctypes.data if want_minmax else None)
        return order, (None if mm is None else (mm[0], mm[1]))

    def grid_order_device(self, d_sdf: int, n: int, d_order: int, d_minmax: int):
        self._call(lib().m2s_grid_order_device, d_sdf, n, d_order or None, d_minmax or None)

    def sample_grid_sdf(self, sdf, grid: Grid, points, mode: int = 1, iso: float = 0.0) -> np.ndarray:
        """sdf_grid() of draw_raymarching.wgsl:118-200 at arbitrary points; mode 0 snap, 1 trilinear, 2 tetrahedral."""
        sdf = np.ascontiguousarray(sdf, np.float32).reshape(-1)
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        cc = np.asarray(grid.cell_count, np.uint64)
        if len(sdf) != int(np.prod(cc)):
            raise ValueError("sdf length does not match the grid")
        out = np.empty(len(pts), np.float32)
        self._call(lib().m2s_sample_grid_sdf, sdf.ctypes.data, grid.first_cell.ctypes.data_as(_f),
                   grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64), pts.ctypes.data, len(pts), int(mode),
                   float(iso), out.ctypes.data)
        return out

    def sample_grid_sdf_device(self, d_sdf: int, grid: Grid, d_points: int, n_points: int, mode: int, iso: float,
                               d_out: int):
        cc = np.asarray(grid.cell_count, np.uint64)
        self._call(lib().m2s_sample_grid_sdf_device, d_sdf, grid.first_cell.ctypes.data_as(_f),
                   grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64), d_points, n_points, int(mode),
                   float(iso), d_out)


def _mesh_arrays(verts, tris):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    return verts, tris


def _out_array(out, n: int) -> np.ndarray:
    """The destination libm2s writes ``n`` float32 into: allocated here, or the caller's array after checking that
    it is exactly that (a wrong dtype / size / stride would make the library write past the buffer)."""
    if out is None:
        return np.empty(n, np.float32)
    if not isinstance(out, np.ndarray) or out.dtype != np.float32 or out.size != n or not out.flags.c_contiguous \
            or not out.flags.writeable:
        raise ValueError(f"out must be a writeable C-contiguous float32 array of {n} elements")
    return out


class Mesh:
    """``m2s_mesh``: a mesh uploaded once with its LBVH on every device of the context. Calls on it pay neither the
    upload nor the build (the reference's viewer regenerates the grid of one mesh on every parameter change,
    mesh_to_sdf_client/src/sdf_program.rs:679-721). Destroy it before its context."""

    def __init__(self, ctx: "Context", verts, tris, device_ptrs=None):
        self._ctx = ctx
        self._h = C.c_void_p()
        if device_ptrs is not None:
            d_verts, nv, d_tris, nt = device_ptrs
            ctx._call(lib().m2s_mesh_create_device, d_verts, nv, d_tris, nt, C.byref(self._h))
        else:
            verts, tris = _mesh_arrays(verts, tris)
            ctx._call(lib().m2s_mesh_create, verts.ctypes.data, len(verts), tris.ctypes.data, len(tris),
                      C.byref(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and self._ctx._h.value:
            lib().m2s_mesh_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def grid_sdf(self, grid: Grid, sign: int, x_begin: int = 0, x_end: Optional[int] = None,
                 out: Optional[np.ndarray] = None) -> np.ndarray:
        x_end = grid.cell_count[0] if x_end is None else x_end
        out = _out_array(out, max(0, x_end - x_begin) * grid.cell_count[1] * grid.cell_count[2])
        cc = np.asarray(grid.cell_count, np.uint64)
        self._ctx._call(lib().m2s_mesh_grid_sdf, self._h, grid.first_cell.ctypes.data_as(_f),
                        grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64), int(sign), x_begin, x_end,
                        out.ctypes.data)
        return out

    def grid_sdf_device(self, grid: Grid, sign: int, x_begin: int, x_end: int, d_out: int):
        cc = np.asarray(grid.cell_count, np.uint64)
        self._ctx._call(lib().m2s_mesh_grid_sdf_device, self._h, grid.first_cell.ctypes.data_as(_f),
                        grid.cell_size.ctypes.data_as(_f), cc.ctypes.data_as(_u64), int(sign), x_begin, x_end, d_out)

    def sdf(self, queries, accel: int, sign: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        out = _out_array(out, len(queries))
        self._ctx._call(lib().m2s_mesh_sdf, self._h, queries.ctypes.data, len(queries), int(accel), int(sign),
                        out.ctypes.data)
        return out

    def sdf_device(self, d_queries: int, nq: int, accel: int, sign: int, d_out: int):
        self._ctx._call(lib().m2s_mesh_sdf_device, self._h, d_queries, nq, int(accel), int(sign), d_out)


class PinnedArray:
    """float32 array in page-locked, mapped host memory from ``m2s_host_alloc``: a destination the distance kernel
    writes in place (no staging, no copy). ``.array`` is the numpy view; free with ``close()``."""

    def __init__(self, n: int):
        self._p = C.c_void_p()
        rc = lib().m2s_host_alloc(max(1, n) * 4, C.byref(self._p))
        if rc != M2S_OK:
            raise M2SError(rc, "m2s_host_alloc failed")
        self.array = np.ctypeslib.as_array(C.cast(self._p, _f), shape=(n,))

    def close(self):
        if self._p.value:
            self.array = None
            lib().m2s_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_alloc(n: int) -> PinnedArray:
    return PinnedArray(n)


def host_register(a: np.ndarray):
    """Page-locks memory the caller owns (``m2s_host_register``); unregister before freeing it."""
    rc = lib().m2s_host_register(a.ctypes.data, a.nbytes)
    if rc != M2S_OK:
        raise M2SError(rc, "m2s_host_register failed")


def host_unregister(a: np.ndarray):
    rc = lib().m2s_host_unregister(a.ctypes.data)
    if rc != M2S_OK:
        raise M2SError(rc, "m2s_host_unregister failed")


_default_ctx: Optional[Context] = None
_default_lock = threading.Lock()


def default_context() -> Context:
    """Process-global lazily created context (the Rust facade keeps a ``OnceLock<Mutex<..>>`` the same way).
    ``M2S_DEVICES=0,1,..`` selects the devices (default: device 0)."""
    global _default_ctx
    with _default_lock:
        if _default_ctx is None:
            env = os.environ.get("M2S_DEVICES", "").strip()
            devices = [int(t) for t in env.split(",") if t.strip() != ""] if env else None
            _default_ctx = Context(devices)
    return _default_ctx


class SampleMode(enum.IntEnum):
    """``raymarch_mode`` of mesh_to_sdf_client/shaders/draw_raymarching.wgsl."""
    Snap = 0
    Trilinear = 1
    Tetrahedral = 2


# ---- the two public free functions ----------------------------------------------------------------------------
def generate_grid_sdf(vertices, indices: Topology, grid: Grid, sign_method: SignMethod = SignMethod.Raycast,
                      ctx: Optional[Context] = None) -> np.ndarray:
    """``generate_grid_sdf(vertices, indices, grid, sign_method) -> Vec<f32>`` (``generate/grid.rs:265-378``).

    Returns ``nx*ny*nz`` float32 in ``Grid::get_cell_idx`` order (z fastest, x slowest)."""
    verts = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    tris = indices.get_triangles(len(verts))
    return (ctx or default_context()).grid_sdf(verts, tris, grid, int(sign_method))


def generate_sdf(vertices, indices: Topology, query_points,
                 acceleration_method: AccelerationMethod = AccelerationMethod.RtreeBvh,
                 ctx: Optional[Context] = None) -> np.ndarray:
    """``generate_sdf(vertices, indices, query_points, acceleration_method) -> Vec<f32>`` (``lib.rs:291-311``)."""
    verts = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    tris = indices.get_triangles(len(verts))
    if len(tris) == 0 and acceleration_method.kind == 3:
        return np.zeros(0, np.float32)  # rtree_bvh.rs:104-106: empty mesh -> empty Vec
    return (ctx or default_context()).sdf(verts, tris, query_points, acceleration_method.kind,
                                          int(acceleration_method.sign))
