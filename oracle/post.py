"""ORACLE — TEST INFRASTRUCTURE ONLY: numpy restatement of what the reference's in-repo caller does with a
finished grid SDF. float32 numpy arithmetic never fuses a*b+c, like rustc.

PARITY UNPINNED for ``sample_grid``: the reference implements it only as a WGSL fragment shader
(mesh_to_sdf_client/shaders/draw_raymarching.wgsl), which cannot run here and has no golden vectors; the
restatement follows the shader's evaluation order and is anchored by properties (values at cell centres,
exactness on affine fields). ``grid_order`` / ``minmax`` restate std / itertools semantics that are pinned by
hand-made known answers in tests/test_post.py.
"""
from __future__ import annotations

import numpy as np


def total_order_key(a: np.ndarray) -> np.ndarray:
    """f32::total_cmp as an unsigned key (core::f32::total_cmp: flip all bits of negatives, the sign bit of the rest)."""
    b = np.ascontiguousarray(a, np.float32).view(np.uint32)
    return np.where(b >> 31 != 0, ~b, b ^ np.uint32(0x80000000)).astype(np.uint32)


def grid_order(sdf: np.ndarray) -> np.ndarray:
    """mesh_to_sdf_client/src/sdf.rs:65-68: ``(0..n).sorted_by(|i, j| data[i].total_cmp(&data[j]))`` — itertools
    ``sorted_by`` is the stable ``Vec::sort_by``."""
    return np.argsort(total_order_key(sdf), kind="stable").astype(np.uint32)


def minmax(sdf: np.ndarray):
    """mesh_to_sdf_client/src/sdf.rs:123: ``data.iter().copied().minmax()`` (itertools::minmax over PartialOrd:
    the FIRST of several equal minima and the LAST of several equal maxima; -0.0 == +0.0)."""
    a = np.ascontiguousarray(sdf, np.float32)
    lo, hi = a.min(), a.max()
    first_min = int(np.flatnonzero(a == lo)[0])
    last_max = int(np.flatnonzero(a == hi)[-1])
    return a[first_min], a[last_max]


def _fetch(sdf, count, c, iso):
    """draw_raymarching.wgsl:92-99: clamp the cell to the grid, idx = z + y*nz + x*nz*ny, minus iso."""
    c = np.minimum(np.maximum(c, 0), np.asarray(count, np.int64) - 1)
    idx = c[:, 2] + c[:, 1] * count[2] + c[:, 0] * count[2] * count[1]
    return (sdf[idx] - np.float32(iso)).astype(np.float32)


SNAP, TRILINEAR, TETRAHEDRAL = 0, 1, 2


def sample_grid(sdf, first_cell, cell_size, cell_count, points, mode, iso=0.0) -> np.ndarray:
    """sdf_grid(position, iso), draw_raymarching.wgsl:118-200 (+ compute_tetrahedral_barycenter :585-650)."""
    sdf = np.ascontiguousarray(sdf, np.float32)
    first = np.asarray(first_cell, np.float32)
    size = np.asarray(cell_size, np.float32)
    count = [int(c) for c in cell_count]
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    one = np.float32(1.0)
    # Grid::get_last_cell, mesh_to_sdf/src/grid.rs:82-88: first + count * size
    end = (first + np.asarray(count, np.float32) * size).astype(np.float32)
    outside = np.any(p < first, axis=1) | np.any(p > end, axis=1)
    with np.errstate(all="ignore"):
        if mode == SNAP:
            start_grid = (first - size * np.float32(0.5)).astype(np.float32)
            c = np.floor(((p - start_grid).astype(np.float32) / size).astype(np.float32)).astype(np.int64)
            d = _fetch(sdf, count, c, iso)
        else:
            ci = ((p - first).astype(np.float32) / size).astype(np.float32)
            fl = np.floor(ci).astype(np.float32)
            f = (ci - fl).astype(np.float32)
            c = fl.astype(np.int64)

            def at(off):
                return _fetch(sdf, count, c + np.asarray(off, np.int64), iso)

            def lerp(a, b, t):
                return (a * (one - t) + b * t).astype(np.float32)

            if mode == TRILINEAR:
                x00 = lerp(at([0, 0, 0]), at([1, 0, 0]), f[:, 0])
                x01 = lerp(at([0, 0, 1]), at([1, 0, 1]), f[:, 0])
                x10 = lerp(at([0, 1, 0]), at([1, 1, 0]), f[:, 0])
                x11 = lerp(at([0, 1, 1]), at([1, 1, 1]), f[:, 0])
                d = lerp(lerp(x00, x10, f[:, 1]), lerp(x01, x11, f[:, 1]), f[:, 2])
            else:
                r, g, b = f[:, 0], f[:, 1], f[:, 2]
                n = len(p)
                w = np.zeros((n, 4), np.float32)
                v2 = np.zeros((n, 3), np.int64)
                v3 = np.zeros((n, 3), np.int64)
                # six sequential `if`s: a later match overrides an earlier one
                cases = [
                    ((g >= b) & (b >= r), (one - g, g - b, b - r, r), (0, 1, 0), (0, 1, 1)),
                    ((b > r) & (r > g), (one - b, b - r, r - g, g), (0, 0, 1), (1, 0, 1)),
                    ((b > g) & (g >= r), (one - b, b - g, g - r, r), (0, 0, 1), (0, 1, 1)),
                    ((r >= g) & (g > b), (one - r, r - g, g - b, b), (1, 0, 0), (1, 1, 0)),
                    ((g > r) & (r >= b), (one - g, g - r, r - b, b), (0, 1, 0), (1, 1, 0)),
                    ((r >= b) & (b >= g), (one - r, r - b, b - g, g), (1, 0, 0), (1, 0, 1)),
                ]
                for m, bary, a2, a3 in cases:
                    for k in range(4):
                        w[m, k] = np.asarray(bary[k], np.float32)[m]
                    v2[m] = a2
                    v3[m] = a3
                s0, s3 = at([0, 0, 0]), at([1, 1, 1])
                s1 = _fetch(sdf, count, c + v2, iso)
                s2 = _fetch(sdf, count, c + v3, iso)
                d = (((w[:, 0] * s0 + w[:, 1] * s1).astype(np.float32) + w[:, 2] * s2).astype(np.float32)
                     + w[:, 3] * s3).astype(np.float32)
    return np.where(outside, np.float32(100.0), d).astype(np.float32)
