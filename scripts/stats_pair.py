"""Traversal counters of the grid kernels on C3 (needs a -DM2S_STATS_BUILD library: M2S_LIB=build/libm2s_stats.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["M2S_STATS"] = "1"
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
L = m2s.lib()
L.m2s_debug_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
verts, tris = synth.bumpy_torus(256, 196)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [256, 256, 256])
for pair in sys.argv[1:] or ["0", "2", "3"]:
    os.environ["M2S_PAIR"] = pair
    with m2s.Context() as ctx:
        ctx.grid_sdf(verts, tris, grid, 0)
        ctx.grid_sdf(verts, tris, grid, 0)
        a = (C.c_uint64 * 4)()
        L.m2s_debug_stats(ctx._h, a)
        s = [int(x) for x in a]
        if hasattr(L, "m2s_debug_hist"):
            h = (C.c_uint64 * 32)()
            L.m2s_debug_hist(ctx._h, h)
            h = [int(x) for x in h]
            if sum(h):
                print("   visits per tile by subtree size 2^k leaves:", " ".join(f"{k}:{v/max(s[2],1):.1f}" for k, v in enumerate(h) if v))
        print(f"pair={pair}: no-seed tiles {s[3]} tiles {s[2]} nodes/tile {s[0]/max(s[2],1):.1f} leaves/tile {s[1]/max(s[2],1):.1f} "
              f"dist_ms {ctx.timings()['dist_ms']:.2f}", flush=True)
