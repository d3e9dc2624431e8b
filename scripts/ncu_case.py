"""One short run per kernel for `ncu --set full` captures (development aid).
usage: python scripts/ncu_case.py grid|C2|C5|points"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

what = sys.argv[1] if len(sys.argv) > 1 else "grid"
with m2s.Context([0]) as c:
    if what in ("grid", "C2", "C5"):
        nu, nv, n, sign = {"grid": (256, 196, 256, 0), "C2": (64, 40, 128, 1), "C5": (1024, 490, 512, 0)}[what]
        verts, tris = synth.bumpy_torus(nu, nv)
        mn, mx = synth.padded_grid_box(verts)
        grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
        c.set_option(m2s.OPT_HOST_PATH, m2s.HOST_STAGED)  # device destination: what bench.py's `value` times
        for _ in range(3):
            c.grid_sdf(verts, tris, grid, sign)
    else:
        verts, tris = synth.bumpy_torus(640, 392)
        mn, mx = synth.padded_grid_box(verts)
        q = synth.splitmix64_points(1_000_000, mn, mx)
        for _ in range(3):
            c.sdf(verts, tris, q, 3, 0)
print("done")
