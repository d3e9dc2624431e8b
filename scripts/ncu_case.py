"""One short run per kernel for `ncu --set full` captures (development aid).
usage: python scripts/ncu_case.py grid|points"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

what = sys.argv[1] if len(sys.argv) > 1 else "grid"
with m2s.Context([0]) as c:
    if what == "grid":
        verts, tris = synth.bumpy_torus(256, 196)
        mn, mx = synth.padded_grid_box(verts)
        grid = m2s.Grid.from_bounding_box(mn, mx, [256, 256, 256])
        out = m2s.host_alloc(256 ** 3)
        c.set_option(m2s.OPT_HOST_PATH, m2s.HOST_STAGED)  # device destination: what bench.py's `value` times
        for _ in range(3):
            c.grid_sdf(verts, tris, grid, 0)
    else:
        verts, tris = synth.bumpy_torus(640, 392)
        mn, mx = synth.padded_grid_box(verts)
        q = synth.splitmix64_points(1_000_000, mn, mx)
        for _ in range(3):
            c.sdf(verts, tris, q, 3, 0)
print("done")
