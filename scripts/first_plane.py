"""Development measurement: fixed cost of an x-slab launch of the distance kernel (C5 on one GPU; the multi-GPU path
starts one slab per device, so every device pays it once). Times slabs [x0, x0 + planes) for both run lengths and
prints them beside planes x (whole-grid time / 512); with a -DM2S_STATS_BUILD library (M2S_LIB=..., M2S_STATS=1)
also node visits / exact evaluations per tile.
usage: python scripts/first_plane.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

verts, tris = synth.bumpy_torus(1024, 490)
mn, mx = synth.padded_grid_box(verts)
n = 512
grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
d_out = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
stats = bool(os.environ.get("M2S_STATS"))
with m2s.Context([0]) as ctx, ctx.mesh(verts, tris) as mesh:
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":  # two launches of one brick plane for a profiler capture
        for _ in range(2):
            mesh.grid_sdf_device(grid, 0, 192, 196, d_out.data_ptr() + 4 * 192 * n * n)
            ctx.synchronize()
        sys.exit(0)
    for v in (4, 2):
        ctx.set_option(m2s.OPT_RUN_LENGTH, v)
        whole = None
        for x0, planes in ((0, 512), (192, 4), (192, 16), (192, 64), (192, 128), (64, 4), (64, 64), (448, 64)):
            best = 1e9
            for _ in range(3):
                if stats:
                    ctx.debug_stats()  # reads and clears
                mesh.grid_sdf_device(grid, 0, x0, x0 + planes, d_out.data_ptr() + 4 * x0 * n * n)
                ctx.synchronize()
                best = min(best, ctx.timings()["dist_ms"])
            if whole is None:
                whole = best
            line = f"V={v} x0={x0:3d} planes={planes:3d}: kernel {best:7.3f} ms | planes x whole / 512 = {whole * planes / 512:7.3f}"
            if stats:
                s = ctx.debug_stats()
                line += f" | nodes/tile {s[0] / max(s[2], 1):.1f} leaves-or-evals/tile {s[1] / max(s[2], 1):.1f} seedless {s[3]} of {s[2]}"
            print(line, flush=True)
