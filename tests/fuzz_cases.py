"""Random meshes x random grids / queries against the exact oracle (shared by tests/test_gpu_api.py and
scripts/fuzz_gpu.py). Raycast-signed results must be bit-identical; Normal within the near-tie window with (almost)
equal signs."""
import numpy as np

import mesh_to_sdf_b200 as m2s
import oracle
from mesh_to_sdf_b200 import synth


def run_fuzz(ctx, n_cases: int, seed: int = 0, log=print) -> int:
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        kind = rng.integers(0, 4)
        if kind == 0:  # soup
            nv, nt = int(rng.integers(3, 200)), int(rng.integers(1, 500))
            verts = rng.uniform(-1, 1, (nv, 3)).astype(np.float32)
            tris = rng.integers(0, nv, (nt, 3)).astype(np.uint32)
        elif kind == 1:  # torus, maybe with degenerate / duplicated triangles
            verts, tris = synth.bumpy_torus(int(rng.integers(4, 40)), int(rng.integers(3, 30)))
            tris = tris.copy()
            k = int(rng.integers(0, max(1, len(tris) // 10)))
            if k:
                tris[rng.integers(0, len(tris), k), 2] = tris[rng.integers(0, len(tris), k), 1]
        elif kind == 2:  # thin / flat cluster: many near-coplanar triangles
            nv = int(rng.integers(10, 150))
            verts = rng.uniform(-1, 1, (nv, 3)).astype(np.float32)
            verts[:, int(rng.integers(0, 3))] *= np.float32(rng.choice([0.0, 1e-6, 1e-3, 0.05]))
            tris = rng.integers(0, nv, (int(rng.integers(1, 300)), 3)).astype(np.uint32)
        else:  # tiny mesh
            verts = rng.uniform(-1, 1, (int(rng.integers(3, 6)), 3)).astype(np.float32)
            tris = rng.integers(0, len(verts), (int(rng.integers(1, 4)), 3)).astype(np.uint32)
        # random similarity: scale over 12 decades, offset far from the origin sometimes
        scale = np.float32(10.0 ** rng.uniform(-6, 6))
        offset = (rng.uniform(-1, 1, 3) * scale * rng.choice([0.0, 1.0, 100.0])).astype(np.float32)
        verts = (verts * scale + offset).astype(np.float32)
        lo, hi = verts.min(axis=0), verts.max(axis=0)
        ext = np.maximum(hi - lo, scale * np.float32(1e-3))
        pad = rng.uniform(-0.2, 1.5, 3).astype(np.float32)
        bmin = (lo - pad * ext).astype(np.float32)
        bmax = (hi + rng.uniform(-0.2, 1.5, 3).astype(np.float32) * ext).astype(np.float32)
        bmax = np.maximum(bmax, bmin + ext * np.float32(1e-3))
        dims = [int(rng.integers(1, 40)) for _ in range(3)]
        if case % 10 == 9:  # now and then a grid with many brick planes (neighbour seeds, stragglers)
            dims = [int(rng.integers(40, 110)) for _ in range(3)]
        grid = m2s.Grid.from_bounding_box(bmin, bmax, dims)
        tag = f"case {case} kind {kind} nt {len(tris)} dims {dims} scale {scale:.3g}"
        for sign in (0, 1):
            got = ctx.grid_sdf(verts, tris, grid, sign)
            want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, sign)
            if sign == 0:
                if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                    bad += 1
                    log("MISMATCH raycast grid", tag, int(np.sum(got.view(np.uint32) != want.view(np.uint32))))
            else:
                tol = max(4e-6, 4e-6 * float(np.max(np.abs(want))))
                mag_ok = np.max(np.abs(np.abs(got) - np.abs(want))) <= tol
                sign_bad = float(np.mean(np.signbit(got) != np.signbit(want)))
                if not mag_ok or (kind == 1 and sign_bad > 0) or sign_bad > 0.01:
                    bad += 1
                    log("MISMATCH normal grid", tag, float(np.max(np.abs(np.abs(got) - np.abs(want)))), sign_bad)
        q = (bmin + rng.uniform(-0.2, 1.2, (int(rng.integers(1, 3000)), 3)) * (bmax - bmin)).astype(np.float32)
        for accel, sign in ((0, 0), (1, 0), (3, 0), (2, 0)):
            got = ctx.sdf(verts, tris, q, accel, sign)
            want = oracle.generate_sdf(verts, tris, q, accel, sign)
            same = np.array_equal(np.abs(got).view(np.uint32), np.abs(want).view(np.uint32)) if accel == 2 else \
                np.array_equal(got.view(np.uint32), want.view(np.uint32))
            if not same:
                bad += 1
                log("MISMATCH points", tag, accel, int(np.sum(got.view(np.uint32) != want.view(np.uint32))))
    return bad
