import os, sys, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["M2S_STATS"] = "1"
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(256, 196)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [256, 256, 256])
import torch
dv = torch.from_numpy(verts).cuda(); dt = torch.from_numpy(tris.view(np.int32)).cuda()
out = torch.empty(256**3, dtype=torch.float32, device="cuda")
with m2s.Context() as ctx:
    for r in range(3):
        ctx.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, 0, 0, 256, out.data_ptr())
        ctx.synchronize()
        st = ctx.debug_stats()
        print("tiles", st[2], "nodes/tile", st[0] / st[2], "leaves/tile", st[1] / st[2], "fallback tiles", st[3], f"({st[3]/st[2]*100:.2f}%)", ctx.timings()["dist_ms"])
