"""Deterministic synthetic inputs for the configs of BASELINE.json.

No RNG for meshes: the "bumpy torus" T(Nu, Nv) is watertight, has 2*Nu*Nv triangles / Nu*Nv vertices,
and is wound so that ``cross(b - a, c - a)`` points outward (the orientation the reference's Normal
sign method assumes, geo.rs:40-42).
"""
from __future__ import annotations

import numpy as np

# name -> (Nu, Nv, grid n, sign/method)   
CONFIGS = {
    "C2": dict(nu=64, nv=40, n=128, sign="Normal"),
    "C3": dict(nu=256, nv=196, n=256, sign="Raycast"),
    "C4": dict(nu=640, nv=392, nq=1_000_000, accel="RtreeBvh"),
    "C5": dict(nu=1024, nv=490, n=512, sign="Raycast"),
}


def bumpy_torus(nu: int, nv: int):
    """Returns (vertices float32 [nu*nv,3], triangles uint32 [2*nu*nv,3])."""
    i = np.arange(nu, dtype=np.float64)[:, None]
    j = np.arange(nv, dtype=np.float64)[None, :]
    u = 2.0 * np.pi * i / nu
    v = 2.0 * np.pi * j / nv
    R = 1.0
    r = 0.35 * (1.0 + 0.15 * np.sin(5.0 * u) * np.cos(3.0 * v))
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v) + 0.0 * u
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    ii = np.arange(nu)[:, None]
    jj = np.arange(nv)[None, :]
    i1 = (ii + 1) % nu
    j1 = (jj + 1) % nv
    p00 = (ii * nv + jj).ravel()
    p10 = (i1 * nv + jj).ravel()
    p01 = (ii * nv + j1).ravel()
    p11 = (i1 * nv + j1).ravel()
    # d/du x d/dv points outward on a torus -> (p00, p10, p11) and (p00, p11, p01) are outward-wound.
    t0 = np.stack([p00, p10, p11], axis=-1)
    t1 = np.stack([p00, p11, p01], axis=-1)
    tris = np.empty((2 * nu * nv, 3), dtype=np.uint32)
    tris[0::2] = t0
    tris[1::2] = t1
    return verts, tris


def icosphere(subdiv: int, radius: float = 1.0):
    """Outward-wound icosphere: 20*4^subdiv triangles. Analytic SDF = |p| - radius (O(h^2) error)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = (np.array(verts) * radius).astype(np.float32)
    F = np.array(faces, dtype=np.uint32)
    return V, F


def padded_grid_box(verts: np.ndarray, lo: float = 0.20, hi: float = 0.23):
    """Mesh AABB expanded asymmetrically (0.20 / 0.23 x extent) so no cell centre sits on a symmetry plane."""
    mn = verts.min(axis=0).astype(np.float32)
    mx = verts.max(axis=0).astype(np.float32)
    ext = mx - mn
    return (mn - np.float32(lo) * ext).astype(np.float32), (mx + np.float32(hi) * ext).astype(np.float32)


def splitmix64_points(n: int, bmin, bmax, seed: int = 0x6D32734446) -> np.ndarray:
    """n query points uniform in [bmin, bmax): splitmix64 -> top 24 bits -> [0,1) float32."""
    m = 3 * n
    mask = (1 << 64) - 1
    idx = np.arange(1, m + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & mask) + idx * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))
    u = u.reshape(n, 3)
    bmin = np.asarray(bmin, np.float32)
    bmax = np.asarray(bmax, np.float32)
    return (bmin + u * (bmax - bmin)).astype(np.float32)


def mesh_diag(verts: np.ndarray) -> float:
    return float(np.linalg.norm(verts.max(axis=0).astype(np.float64) - verts.min(axis=0).astype(np.float64)))
