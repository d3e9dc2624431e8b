"""Timing of one 64-plane rank slab of config C5 (1 003 520 triangles, 512^3) through the host entry point (development aid)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(1024, 490)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [512, 512, 512])
ctx = m2s.default_context()
out = np.empty(64 * 512 * 512, np.float32)
for r in range(3):
    ctx.grid_sdf_slab(verts, tris, grid, 0, 192, 256, out)
    print({k: round(v, 3) for k, v in ctx.timings().items()}, flush=True)
