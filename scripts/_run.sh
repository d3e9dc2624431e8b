python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
v, t = synth.bumpy_torus(640, 392)
mn, mx = synth.padded_grid_box(v)
q = synth.splitmix64_points(1_000_000, mn, mx)
ctx = m2s.default_context()
for name, accel, sign in (("RtreeBvh", 3, 0), ("Rtree", 2, 0), ("Bvh/Raycast", 1, 0), ("Bvh/Normal", 1, 1), ("None/Raycast(100k q)", 0, 0)):
    qq = q[:100_000] if accel == 0 else q
    for r in range(3):
        out = ctx.sdf(v, t, qq, accel, sign)
    tm = ctx.timings()
    print(name, " ".join(f"{k}={x:.3f}" for k, x in tm.items()), flush=True)
PY
