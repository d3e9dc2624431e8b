"""Kernel time of the grid distance kernel with runs of 2 and 4 voxels per lane over mesh sizes, grid sizes and cell
shapes (development aid behind the heuristic in grid_run_length, m2s_grid.cu). usage: python scripts/run_length_matrix.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

ctx = m2s.default_context()
print("| mesh | triangles | grid | cells | sign | V=2 ms | V=4 ms | V=2 B ms | V=4 B ms | best |\n|---|---:|---|---|---|---:|---:|---:|---:|---|")
for nu, nv in ((64, 40), (100, 64), (128, 98), (180, 140), (256, 196), (512, 392), (1024, 490)):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    with ctx.mesh(verts, tris) as mesh:
        for n in (128, 256):
            if n == 128 and len(tris) > 300000:
                continue
            for shape in ("flat", "cubic"):
                if shape == "flat":
                    grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
                else:  # cubic cells: same cell size on every axis, fewer cells along z
                    cs = float((mx - mn).max()) / n
                    cnt = [max(8, int(np.ceil((mx[i] - mn[i]) / cs))) for i in range(3)]
                    grid = m2s.Grid(mn + 0.5 * cs, [cs, cs, cs], cnt)
                out = m2s.host_alloc(grid.get_total_cell_count())
                for sign in (0,):
                    t = {}
                    for v in (2, 4, 18, 20):
                        ctx.set_option(m2s.OPT_RUN_LENGTH, v)
                        best = 1e9
                        for _ in range(3):
                            mesh.grid_sdf(grid, sign, out=out.array)
                            best = min(best, ctx.timings()["dist_ms"])
                        t[v] = best
                    print(f"| T({nu},{nv}) | {len(tris)} | {'x'.join(map(str, grid.cell_count))} | {shape} | "
                          f"{'Raycast' if sign == 0 else 'Normal'} | {t[2]:.3f} | {t[4]:.3f} | {t[18]:.3f} | {t[20]:.3f} | "
                          f"{min(t, key=t.get)} ({min(t.values()) / t[2]:.3f}) |", flush=True)
                out.close()
ctx.set_option(m2s.OPT_RUN_LENGTH, 0)
