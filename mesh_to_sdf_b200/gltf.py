"""Binary glTF (.glb) -> vertices / indices for the harness.

The reference's benches, example and viewer read their meshes with the ``easy_gltf`` crate
(mesh_to_sdf/benches/generate_grid_sdf.rs:8-31, benches/generate_sdf.rs, examples/demo.rs): ``easy_gltf::load(path)``
returns scenes whose ``models`` are the mesh primitives met on a depth-first walk of the scene's nodes, every
position already multiplied by its node's global transform (f32), with optional indices. This module restates that
for ``.glb`` files with one embedded binary chunk — all the reference ships — without any dependency: host-only,
no kernels.

    models = gltf.load_glb("knight.glb")      # scene 0, traversal order
    verts, idx = models[0].vertices, models[0].triangle_indices()
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}
TRIANGLES, TRIANGLE_STRIP, TRIANGLE_FAN = 4, 5, 6


class GltfError(ValueError):
    pass


@dataclass
class Model:
    """One mesh primitive (easy_gltf ``Model``): positions in world space, optional indices, primitive mode."""
    vertices: np.ndarray            # (n, 3) float32, node transform applied
    indices: Optional[np.ndarray]   # (m,) uint32 or None
    mode: int = TRIANGLES

    def triangle_indices(self) -> np.ndarray:
        """u32 index triples as a flat list, whatever the primitive mode (strips / fans expanded, no winding flip —
        the same convention as ``Topology::TriangleStrip``, mesh_to_sdf/src/lib.rs:183-192)."""
        idx = self.indices if self.indices is not None else np.arange(len(self.vertices), dtype=np.uint32)
        if self.mode == TRIANGLES:
            return idx[: len(idx) // 3 * 3].astype(np.uint32)
        if self.mode == TRIANGLE_STRIP:
            if len(idx) < 3:
                return np.zeros(0, np.uint32)
            return np.stack([idx[:-2], idx[1:-1], idx[2:]], axis=1).reshape(-1).astype(np.uint32)
        if self.mode == TRIANGLE_FAN:
            if len(idx) < 3:
                return np.zeros(0, np.uint32)
            first = np.full(len(idx) - 2, idx[0], np.uint32)
            return np.stack([first, idx[1:-1], idx[2:]], axis=1).reshape(-1).astype(np.uint32)
        raise GltfError(f"primitive mode {self.mode} is not a triangle mode")


def _chunks(data: bytes):
    if len(data) < 12:
        raise GltfError("not a GLB file (too short)")
    magic, version, length = struct.unpack_from("<III", data, 0)
    if magic != 0x46546C67 or version != 2:
        raise GltfError("not a glTF 2.0 binary file")
    js, bin_chunk, off = None, None, 12
    while off + 8 <= min(length, len(data)):
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942 and bin_chunk is None:
            bin_chunk = chunk
        off += 8 + clen
    if js is None:
        raise GltfError("GLB without a JSON chunk")
    return js, bin_chunk if bin_chunk is not None else b""


def _accessor(js, bin_chunk, idx) -> np.ndarray:
    acc = js["accessors"][idx]
    if "bufferView" not in acc or "sparse" in acc:
        raise GltfError("sparse / view-less accessors are not supported")
    bv = js["bufferViews"][acc["bufferView"]]
    if bv.get("buffer", 0) != 0:
        raise GltfError("only the embedded binary chunk is supported")
    comp, ncomp, n = _COMPONENT[acc["componentType"]], _NCOMP[acc["type"]], acc["count"]
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    item = np.dtype(comp).itemsize * ncomp
    stride = bv.get("byteStride", 0) or item
    if start + (n - 1) * stride + item > len(bin_chunk) and n > 0:
        raise GltfError("accessor reaches past the binary chunk")
    if stride == item:
        return np.frombuffer(bin_chunk, comp, n * ncomp, start).reshape(n, ncomp).copy()
    raw = np.frombuffer(bin_chunk, np.uint8, (n - 1) * stride + item, start) if n else np.zeros(0, np.uint8)
    rows = np.lib.stride_tricks.as_strided(raw, (n, item), (stride, 1)) if n else np.zeros((0, item), np.uint8)
    return np.ascontiguousarray(rows).view(comp).reshape(n, ncomp)


def _node_matrix(node) -> np.ndarray:
    """Local transform, column-vector convention: glTF stores ``matrix`` column-major; otherwise T * R * S."""
    if "matrix" in node:
        return np.array(node["matrix"], np.float32).reshape(4, 4).T
    m = np.eye(4, dtype=np.float32)
    if "scale" in node:
        m = np.diag(np.array(list(node["scale"]) + [1.0], np.float32)) @ m
    if "rotation" in node:
        x, y, z, w = (np.float32(v) for v in node["rotation"])
        r = np.eye(4, dtype=np.float32)
        r[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        m = r @ m
    if "translation" in node:
        t = np.eye(4, dtype=np.float32)
        t[:3, 3] = node["translation"]
        m = t @ m
    return m.astype(np.float32)


def load_glb_bytes(data: bytes, scene: Optional[int] = None) -> List[Model]:
    js, bin_chunk = _chunks(data)
    scenes = js.get("scenes", [])
    if not scenes:
        return []
    sc = scenes[js.get("scene", 0) if scene is None else scene]
    models: List[Model] = []

    def walk(ni, parent):
        node = js["nodes"][ni]
        m = (parent @ _node_matrix(node)).astype(np.float32)
        if "mesh" in node:
            for prim in js["meshes"][node["mesh"]]["primitives"]:
                if "POSITION" not in prim.get("attributes", {}):
                    continue
                pos = _accessor(js, bin_chunk, prim["attributes"]["POSITION"]).astype(np.float32)
                if pos.shape[1] != 3:
                    raise GltfError("POSITION must be VEC3")
                if not np.array_equal(m, np.eye(4, dtype=np.float32)):
                    p4 = np.concatenate([pos, np.ones((len(pos), 1), np.float32)], axis=1)
                    pos = (p4 @ m.T)[:, :3].astype(np.float32)  # f32, like cgmath Matrix4 * Vector4
                idx = None
                if "indices" in prim:
                    idx = _accessor(js, bin_chunk, prim["indices"]).reshape(-1).astype(np.uint32)
                models.append(Model(np.ascontiguousarray(pos), idx, int(prim.get("mode", TRIANGLES))))
        for c in node.get("children", []):
            walk(c, m)

    for ni in sc.get("nodes", []):
        walk(ni, np.eye(4, dtype=np.float32))
    return models


def load_glb(path, scene: Optional[int] = None) -> List[Model]:
    """The models of one scene of a .glb file (default: the file's default scene), in easy_gltf's traversal order."""
    with open(path, "rb") as f:
        return load_glb_bytes(f.read(), scene)
