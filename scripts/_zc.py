import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(256, 196)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [256, 256, 256])
n = 256 ** 3
dv = torch.from_numpy(verts).cuda(); dt = torch.from_numpy(tris.view(np.int32)).cuda()
dev_out = torch.empty(n, dtype=torch.float32, device="cuda")
pin_out = torch.empty(n, dtype=torch.float32).pin_memory()
ctx = m2s.Context()
for name, ptr in (("device", dev_out.data_ptr()), ("pinned-host (zero-copy stores)", pin_out.data_ptr())):
    for r in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, 0, 0, 256, ptr)
        ctx.synchronize(); dt_ms = (time.perf_counter() - t0) * 1e3
    print(name, f"wall {dt_ms:.3f} ms", {k: round(v, 3) for k, v in ctx.timings().items()}, flush=True)
print("equal:", np.array_equal(dev_out.cpu().numpy().view(np.uint32), pin_out.numpy().view(np.uint32)))
