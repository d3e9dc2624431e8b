"""Development sweep: leaf size x seed levels on C3 (and C2 / C4), kernel timings from m2s_last_timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

def grid_case(nu, nv, n, sign):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    return verts, tris, m2s.Grid.from_bounding_box(mn, mx, [n, n, n])

cases = {"C2": grid_case(64, 40, 128, 1) + (1,), "C3": grid_case(256, 196, 256, 0) + (0,)}
v4, t4 = synth.bumpy_torus(640, 392)
mn, mx = synth.padded_grid_box(v4)
q4 = synth.splitmix64_points(1_000_000, mn, mx)
ref = {}
for leaf in (1, 2, 4, 8):
    for lev in (0, 1, 2):
        os.environ["M2S_LEAF_SIZE"] = str(leaf)
        os.environ["M2S_SEED_LEVELS"] = str(lev)
        with m2s.Context() as ctx:
            line = f"leaf={leaf} seeds={lev}:"
            for name, (verts, tris, grid, sign) in cases.items():
                best = None
                for r in range(3):
                    out = ctx.grid_sdf(verts, tris, grid, sign)
                    t = ctx.timings()
                    k = t["build_ms"] + t["sign_ms"] + t["dist_ms"]
                    best = k if best is None else min(best, k)
                if name not in ref: ref[name] = out.copy()
                same = np.array_equal(ref[name].view(np.uint32), out.view(np.uint32))
                line += f"  {name} kernels {best:.3f} ms (dist {t['dist_ms']:.3f}, same={same})"
            best = None
            for r in range(3):
                out = ctx.sdf(v4, t4, q4, 3, 0)
                t = ctx.timings()
                k = t["build_ms"] + t["sign_ms"] + t["dist_ms"]
                best = k if best is None else min(best, k)
            if "C4" not in ref: ref["C4"] = out.copy()
            same = np.array_equal(ref["C4"].view(np.uint32), out.view(np.uint32))
            line += f"  C4 kernels {best:.3f} ms (build+sort {t['build_ms']:.3f}, same={same})"
            # distance-only variant of C4 to split nearest vs ray parity: Rtree mode has no rays
            out = ctx.sdf(v4, t4, q4, 2, 0); out = ctx.sdf(v4, t4, q4, 2, 0)
            line += f"  C4-rtree dist {ctx.timings()['dist_ms']:.3f}"
            print(line, flush=True)
