"""Measures, on the CPU, how many triangles a TILE-level conservative test lets through on workload C3, against what
the per-voxel radii of the packet walk let through (the kernel's own counter: 27 triangles per 64-voxel tile).

For a tile with centre c and half-diagonal h every voxel v has d(v) <= d(c) + h, and a triangle T can be nearest to some
voxel only if dist(c, T) - h <= d(c) + h. So the tile-level candidate set is {T : dist(c, T) <= d(c) + 2h} - what a
block-cooperative walk against the tile's box would have to put into shared memory.
usage: python scripts/tile_candidates.py [n_tiles]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mesh_to_sdf_b200 import synth


def point_tri_dist(p, a, b, c):
    """distance from ONE point p to many triangles (float64, Ericson's closest point, vectorised over triangles)"""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = p - b
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = p - c
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    q = np.empty_like(a)
    done = np.zeros(len(a), bool)

    def put(mask, val):
        m = mask & ~done
        q[m] = val[m]
        done[m] = True

    put((d1 <= 0) & (d2 <= 0), a)
    put((d3 >= 0) & (d4 <= d3), b)
    with np.errstate(divide="ignore", invalid="ignore"):
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[:, None] * ab)
        put((d6 >= 0) & (d5 <= d6), c)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[:, None] * ac)
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + w[:, None] * (c - b))
        den = 1.0 / (va + vb + vc)
        put(np.ones(len(a), bool), a + (vb * den)[:, None] * ab + (vc * den)[:, None] * ac)
    return np.linalg.norm(p - q, axis=1)


def main(n_tiles):
    verts, tris = synth.bumpy_torus(256, 196)
    mn, mx = synth.padded_grid_box(verts)
    n = 256
    size = (mx - mn).astype(np.float64) / n
    v = verts.astype(np.float64)
    a, b, c = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    rng = np.random.default_rng(5)
    out = {}
    for name, shape in (("warp tile 2x4x8 (64 voxels)", (2, 4, 8)), ("brick 4x8x16 (512 voxels)", (4, 8, 16))):
        h = 0.5 * np.linalg.norm(np.array(shape) * size)
        counts, far = [], []
        for _ in range(n_tiles):
            cell = rng.integers(0, n, 3) // np.array(shape) * np.array(shape)
            centre = mn + (cell + 0.5 * np.array(shape)) * size
            d = point_tri_dist(centre, a, b, c)
            dc = d.min()
            counts.append(int((d <= dc + 2 * h).sum()))
            far.append(dc)
        counts, far = np.array(counts), np.array(far)
        out[name] = (h, counts, far)
        print(f"{name}: half-diagonal {h:.4f}; tile-level candidates per tile over {n_tiles} random tiles: "
              f"mean {counts.mean():.0f}, median {np.median(counts):.0f}, p10 {np.percentile(counts, 10):.0f}, "
              f"p90 {np.percentile(counts, 90):.0f}, max {counts.max()}; "
              f"tiles within 2h of the surface: {np.mean(far < 2 * h) * 100:.0f} % "
              f"(their mean {counts[far < 2 * h].mean() if np.any(far < 2 * h) else 0:.0f}); "
              f"far tiles' mean {counts[far >= 2 * h].mean():.0f}")
    print("per-voxel radii (k_grid_nearest_run's counter, M2S_STATS build): 27.3 triangles queued per 64-voxel tile, "
          "98.9 node visits")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
