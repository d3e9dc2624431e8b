"""Generates the committed fixtures of tests/golden/ from the reference's own test assets and fixtures.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (small, committed):
  suzanne.npz   vertices/indices of mesh_to_sdf/assets/suzanne.glb  (968 triangles; used by default.rs:83-109,
                bvh.rs:154-310, rtree.rs:135-242, rtree_bvh.rs:183-274)
  ferris3d.npz  model 0 of mesh_to_sdf/assets/ferris3d.glb          (generate/grid.rs:728-843)
  annoted_cube.npz  mesh_to_sdf/assets/annoted_cube.glb (12 triangles)
  kat.json      the known-answer vectors of the reference's doc-tests / unit tests, transcribed with their
                file:line, plus the two proptest regression seeds (proptest-regressions/geo.txt:7-8)
The GLB reader handles exactly what these assets use: one binary chunk, float32 POSITION, u16/u32 indices,
node transforms ignored like easy_gltf's `model.vertices()` … except that easy_gltf APPLIES node transforms;
the assets used here have identity/translation-free mesh nodes for model 0 (checked below).
"""
import json
import os
import struct
import sys

import numpy as np

REF = "/root/reference/mesh_to_sdf"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_glb(path):
    with open(path, "rb") as f:
        data = f.read()
    magic, version, length = struct.unpack_from("<III", data, 0)
    assert magic == 0x46546C67 and version == 2
    off = 12
    js = None
    bin_chunk = None
    while off < length:
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942:
            bin_chunk = chunk
        off += 8 + clen
    return js, bin_chunk


def accessor(js, bin_chunk, idx):
    acc = js["accessors"][idx]
    bv = js["bufferViews"][acc["bufferView"]]
    comp = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}[
        acc["componentType"]]
    ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}[acc["type"]]
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = bv.get("byteStride", 0)
    item = np.dtype(comp).itemsize * ncomp
    n = acc["count"]
    if stride and stride != item:
        out = np.empty((n, ncomp), comp)
        for i in range(n):
            out[i] = np.frombuffer(bin_chunk, comp, ncomp, start + i * stride)
        return out
    return np.frombuffer(bin_chunk, comp, n * ncomp, start).reshape(n, ncomp).copy()


def node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], np.float64).reshape(4, 4).T
    m = np.eye(4)
    if "scale" in node:
        m = np.diag(list(node["scale"]) + [1.0]) @ m
    if "rotation" in node:
        x, y, z, w = node["rotation"]
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        rm = np.eye(4)
        rm[:3, :3] = r
        m = rm @ m
    if "translation" in node:
        t = np.eye(4)
        t[:3, 3] = node["translation"]
        m = t @ m
    return m


def first_model(path):
    """scene 0 -> first mesh primitive in node traversal order, with the node transform applied (easy_gltf)."""
    js, bin_chunk = read_glb(path)
    scene = js["scenes"][js.get("scene", 0)]
    found = []

    def walk(ni, parent):
        node = js["nodes"][ni]
        m = parent @ node_matrix(node)
        if "mesh" in node:
            for prim in js["meshes"][node["mesh"]]["primitives"]:
                found.append((prim, m))
        for c in node.get("children", []):
            walk(c, m)

    for ni in scene["nodes"]:
        walk(ni, np.eye(4))
    prim, m = found[0]
    pos = accessor(js, bin_chunk, prim["attributes"]["POSITION"]).astype(np.float32)
    idx = accessor(js, bin_chunk, prim["indices"]).ravel().astype(np.uint32)
    if not np.allclose(m, np.eye(4)):
        # easy_gltf transforms positions in f32 (cgmath Matrix4 * Vector4)
        m32 = m.astype(np.float32)
        p4 = np.concatenate([pos, np.ones((len(pos), 1), np.float32)], axis=1)
        pos = (p4 @ m32.T)[:, :3].astype(np.float32)
    return pos, idx, len(found), not np.allclose(m, np.eye(4))


def main():
    for name in ("suzanne", "ferris3d", "annoted_cube"):
        pos, idx, nprims, transformed = first_model(os.path.join(REF, "assets", name + ".glb"))
        assert idx.size % 3 == 0 and idx.max() < len(pos)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), vertices=pos, indices=idx)
        print(name, "verts", len(pos), "tris", idx.size // 3, "primitives", nprims, "node transform", transformed)

    kat = {
        "doc_generate_sdf_rtree_bvh": {  # lib.rs:13-31
            "src": "mesh_to_sdf/src/lib.rs:13-31",
            "vertices": [[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]], "indices": [0, 1, 2],
            "query_points": [[0.5, 0.5, 0.5]], "accel": "RtreeBvh", "expect": [1.0]},
        "doc_generate_grid_sdf": {  # lib.rs:34-58
            "src": "mesh_to_sdf/src/lib.rs:34-58",
            "vertices": [[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]], "indices": [0, 1, 2],
            "bbox_min": [0., 0., 0.], "bbox_max": [10., 10., 10.], "cell_count": [10, 10, 10], "sign": "Raycast",
            "expect_index": 0, "expect": 1.0},
        "doc_generate_sdf_fn": {  # lib.rs:269-289
            "src": "mesh_to_sdf/src/lib.rs:269-289",
            "vertices": [[0., 1., 0.], [1., 2., 3.], [1., 3., 4.]], "indices": [0, 1, 2],
            "query_points": [[0., 0., 0.]], "accel": "RtreeBvh", "expect": [1.0]},
        "doc_generate_grid_sdf_fn": {  # generate/grid.rs:205-231
            "src": "mesh_to_sdf/src/generate/grid.rs:205-231",
            "vertices": [[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 4.]], "indices": [0, 1, 2],
            "bbox_min": [0., 0., 0.], "bbox_max": [10., 10., 10.], "cell_count": [10, 10, 10], "sign": "Raycast",
            "expect_index": 0, "expect": 1.0},
        "segment": [  # geo.rs:311-323
            {"src": "mesh_to_sdf/src/geo.rs:316-318", "a": [0., 0., 0.], "b": [1., 0., 0.], "p": [0.3, 1.0, 0.0],
             "expect": [0.3, 0.0, 0.0]},
            {"src": "mesh_to_sdf/src/geo.rs:320-322", "a": [0., 0., 0.], "b": [1., 0., 0.], "p": [10.3, 1.0, 10.0],
             "expect": [1.0, 0.0, 0.0]}],
        "proptest_regressions": [  # proptest-regressions/geo.txt:7-8
            {"p": [0.0, -8.055119, 1.1846914], "a": [0.0, 0.0, 0.0], "b": [0.0, 0.0, 8.367966],
             "c": [-7.806354, 9.330519, 0.0]},
            {"p": [0.0, -5.8359632, 4.405388], "a": [0.0, 0.9572999, 9.758267], "b": [6.9999175, -4.739112, 7.5462694],
             "c": [0.0, -9.673183, 0.52112055]}],
        "grid_from_bounding_box": {  # grid.rs:203-213 (test_from_bounding_box)
            "src": "mesh_to_sdf/src/grid.rs:180-297"},
        "suzanne_python_baseline": {  # default.rs:83-109, tests/generate_python_baseline.py:11-23
            "src": "mesh_to_sdf/src/generate/generic/default.rs:83-109",
            "query_points": [[0., 0., 0.], [1., 1., 1.], [0.1, 0.2, 0.2]], "baseline": [-0.42, 0.69, -0.46],
            "pysdf": [0.45216727, -0.6997909, 0.45411023], "python_mesh_to_sdf": [-0.40961263, 0.6929414, -0.46345082],
            "tolerance": 0.1},
        "point_array": {"src": "mesh_to_sdf/src/point/impl_array.rs:41-63"},
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("kat.json written")


if __name__ == "__main__":
    sys.exit(main())
