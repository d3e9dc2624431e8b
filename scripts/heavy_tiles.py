"""Development measurement (needs a library built with -DM2S_STATS_BUILD -DM2S_STATS_HEAVY=<node visits>; M2S_LIB=...,
M2S_STATS=1): how much of the distance kernel's walk sits in its heaviest warp tiles, and how long the longest one is.
usage: M2S_LIB=build/libm2s_h1.so M2S_STATS=1 python scripts/heavy_tiles.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

for name, nu, nv, n in (("C3", 256, 196, 256), ("C5", 1024, 490, 512)):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
    d_out = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    with m2s.Context([0]) as ctx, ctx.mesh(verts, tris) as mesh:
        ctx.debug_stats()
        mesh.grid_sdf_device(grid, 0, 0, n, d_out.data_ptr())
        ctx.synchronize()
        s = ctx.debug_stats()
        print(f"{name}: tiles {s[2]}, heavy tiles {s[1]} ({100.0 * s[1] / max(s[2], 1):.3f} %), node visits in heavy tiles "
              f"{s[0]} (mean {s[0] / max(s[1], 1):.0f} each), longest tile {s[3]} node visits, kernel "
              f"{ctx.timings()['dist_ms']:.2f} ms", flush=True)
