"""C5 through the host ABI, twice (for `ncu --metrics gpu__time_duration.sum` launch lists of the 1M-triangle build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(1024, 490)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [512, 512, 512])
with m2s.Context([0]) as c:
    out = m2s.host_alloc(512 ** 3)
    for _ in range(2):
        c.grid_sdf_slab(verts, tris, grid, 0, 0, 64, out.array[:64 * 512 * 512])
print("done")
