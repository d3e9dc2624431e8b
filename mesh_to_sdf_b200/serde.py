"""Serialization of signed distance fields — the ``mesh_to_sdf::serde`` module (mesh_to_sdf/src/serde.rs), format V1.

The reference writes ``rmp_serde::to_vec(&SerializeVersion::V1(sdf))`` (serde.rs:156-160): MessagePack with enums as
one-entry maps keyed by the variant name, structs as arrays in field order, ``f32`` as 0xca + big-endian bits,
``usize`` in the shortest unsigned form and sequences with the shortest array header:

    {"V1": {"Generic": [[[x, y, z], ...], [d, ...]]}}                      serde.rs:84-95
    {"V1": {"Grid": [[[fx, fy, fz], [sx, sy, sz], [nx, ny, nz]], [d, ...]]}}  serde.rs:97-106, grid.rs:30-37

Host-only data format on the output side of the hot path: no kernels. The encoder / decoder
below are self-contained (no msgpack dependency) and are pinned byte for byte by the reference's own fixtures
``tests/sdf_generic_v1.bin`` / ``tests/sdf_grid_v1.bin`` (serde.rs:313-372).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Union

import numpy as np

from . import Grid


class SerdeError(Exception):
    """serde.rs:43-52 (SerializationFailed / DeserializationFailed / IoError)."""


@dataclass
class Generic:
    """DeserializeGeneric, serde.rs:122-131."""
    query_points: np.ndarray  # (n, 3) float32
    distances: np.ndarray     # (n,) float32


@dataclass
class GridSdf:
    """DeserializeGrid, serde.rs:133-142."""
    grid: Grid
    distances: np.ndarray


def _array_header(n: int) -> bytes:
    if n < 16:
        return bytes([0x90 | n])
    if n < 1 << 16:
        return b"\xdc" + struct.pack(">H", n)
    if n < 1 << 32:
        return b"\xdd" + struct.pack(">I", n)
    raise SerdeError("sequence longer than 2^32-1")


def _uint(v: int) -> bytes:
    if v < 0:
        raise SerdeError("negative cell count")
    if v < 128:
        return bytes([v])
    if v < 1 << 8:
        return b"\xcc" + struct.pack(">B", v)
    if v < 1 << 16:
        return b"\xcd" + struct.pack(">H", v)
    if v < 1 << 32:
        return b"\xce" + struct.pack(">I", v)
    return b"\xcf" + struct.pack(">Q", v)


def _str(s: str) -> bytes:
    b = s.encode()
    assert len(b) < 32
    return bytes([0xa0 | len(b)]) + b


_F32 = np.dtype([("tag", "u1"), ("v", ">f4")])
_PT = np.dtype([("hdr", "u1"), ("x", _F32), ("y", _F32), ("z", _F32)])


def _f32_seq(a) -> bytes:
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    rec = np.empty(len(a), _F32)
    rec["tag"] = 0xca
    rec["v"] = a
    return _array_header(len(a)) + rec.tobytes()


def _point_seq(p) -> bytes:
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
    rec = np.empty(len(p), _PT)
    rec["hdr"] = 0x93
    for k, name in enumerate("xyz"):
        rec[name]["tag"] = 0xca
        rec[name]["v"] = p[:, k]
    return _array_header(len(p)) + rec.tobytes()


def serialize(sdf: Union[Generic, GridSdf]) -> bytes:
    """serde.rs:156-160: always the latest version (V1)."""
    if isinstance(sdf, Generic):
        body = b"\x92" + _point_seq(sdf.query_points) + _f32_seq(sdf.distances)
        name = "Generic"
    elif isinstance(sdf, GridSdf):
        g = sdf.grid
        grid = (b"\x93" + _f32_seq(g.first_cell) + _f32_seq(g.cell_size)
                + b"\x93" + b"".join(_uint(int(c)) for c in g.cell_count))
        body = b"\x92" + grid + _f32_seq(sdf.distances)
        name = "Grid"
    else:
        raise SerdeError("expected Generic or GridSdf")
    return b"\x81" + _str("V1") + b"\x81" + _str(name) + body


class _Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        self.i = 0

    def take(self, n: int) -> memoryview:
        if self.i + n > len(self.b):
            raise SerdeError("truncated input")
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def byte(self) -> int:
        return self.take(1)[0]

    def map1_key(self) -> str:
        if self.byte() != 0x81:
            raise SerdeError("expected a one-entry map (enum variant)")
        t = self.byte()
        if t & 0xe0 == 0xa0:
            n = t & 0x1f
        elif t == 0xd9:
            n = self.byte()
        else:
            raise SerdeError("expected a string key")
        return bytes(self.take(n)).decode()

    def array(self) -> int:
        t = self.byte()
        if t & 0xf0 == 0x90:
            return t & 0x0f
        if t == 0xdc:
            return struct.unpack(">H", self.take(2))[0]
        if t == 0xdd:
            return struct.unpack(">I", self.take(4))[0]
        raise SerdeError("expected an array")

    def expect_array(self, n: int):
        if self.array() != n:
            raise SerdeError(f"expected an array of {n}")

    def uint(self) -> int:
        t = self.byte()
        if t < 0x80:
            return t
        fmt = {0xcc: ">B", 0xcd: ">H", 0xce: ">I", 0xcf: ">Q"}.get(t)
        if fmt is None:
            raise SerdeError("expected an unsigned integer")
        return struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]

    def f32_seq(self) -> np.ndarray:
        n = self.array()
        rec = np.frombuffer(self.take(n * _F32.itemsize), _F32)
        if n and not np.all(rec["tag"] == 0xca):
            raise SerdeError("expected f32 elements")
        return rec["v"].astype(np.float32)

    def point_seq(self) -> np.ndarray:
        n = self.array()
        rec = np.frombuffer(self.take(n * _PT.itemsize), _PT)
        if n and not (np.all(rec["hdr"] == 0x93) and all(np.all(rec[k]["tag"] == 0xca) for k in "xyz")):
            raise SerdeError("expected [f32; 3] points")
        return np.stack([rec[k]["v"] for k in "xyz"], axis=1).astype(np.float32)


def deserialize(data: bytes) -> Union[Generic, GridSdf]:
    """serde.rs:162-170. Raises SerdeError (DeserializationFailed) on anything that is not a V1 document."""
    r = _Reader(data)
    if r.map1_key() != "V1":
        raise SerdeError("unknown format version")
    kind = r.map1_key()
    r.expect_array(2)
    if kind == "Generic":
        out = Generic(r.point_seq(), r.f32_seq())
    elif kind == "Grid":
        r.expect_array(3)
        first, size = r.f32_seq(), r.f32_seq()
        r.expect_array(3)
        count = [r.uint() for _ in range(3)]
        if len(first) != 3 or len(size) != 3:
            raise SerdeError("grid vectors must have 3 components")
        out = GridSdf(Grid(first, size, count), r.f32_seq())
    else:
        raise SerdeError(f"unknown variant {kind!r}")
    if r.i != len(r.b):
        raise SerdeError("trailing bytes")
    return out


def save_to_file(sdf: Union[Generic, GridSdf], path) -> None:
    """serde.rs:187-193."""
    data = serialize(sdf)
    try:
        with open(path, "wb") as f:
            f.write(data)
    except OSError as e:
        raise SerdeError(f"IoError: {e}") from e


def read_from_file(path) -> Union[Generic, GridSdf]:
    """serde.rs:217-221."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise SerdeError(f"IoError: {e}") from e
    return deserialize(data)
