import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["M2S_STATS"] = "1"
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
L = m2s.lib()
L.m2s_debug_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
def stats(ctx):
    a = (C.c_uint64 * 4)()
    L.m2s_debug_stats(ctx._h, a)
    return [int(x) for x in a]
import itertools
for dyn, lev, leaf in itertools.product((1,), (0, 1), (2, 4)):
    os.environ["M2S_SEED_LEVELS"] = str(lev); os.environ["M2S_LEAF_SIZE"] = str(leaf); os.environ["M2S_PACKET"] = str(dyn)
    with m2s.Context() as ctx:
        for name, nu, nv, n, sign in (("C2", 64, 40, 128, 1), ("C3", 256, 196, 256, 0)):
            verts, tris = synth.bumpy_torus(nu, nv)
            mn, mx = synth.padded_grid_box(verts)
            grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
            out = ctx.grid_sdf(verts, tris, grid, sign)
            out = ctx.grid_sdf(verts, tris, grid, sign)
            s = stats(ctx)
            print(f"packet={dyn} seeds={lev} leaf={leaf} {name}: searches {s[2]} nodes/search {s[0]/max(s[2],1):.1f} leaves/search {s[1]/max(s[2],1):.1f} dist_ms {ctx.timings()['dist_ms']:.2f}", flush=True)
            if name == "C3" and lev == 0 and leaf == 4:
                # far vs near split: histogram of |d|
                a = np.abs(out)
                print("   |d| quantiles", np.quantile(a, [0.1, 0.5, 0.9]))
