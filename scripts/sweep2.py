"""Sweep of build/traversal knobs on C2/C3 (kernel timings)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
def case(nu, nv, n):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    return verts, tris, m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
cases = {"C2": case(64, 40, 128) + (1,), "C3": case(256, 196, 256) + (0,)}
ref = {}
knobs = [dict(zip(("M2S_SEED_PACKET", "M2S_SEED_STRIDE", "M2S_LEAF_SIZE"), k)) for k in itertools.product(("0",), ("4",), ("1", "2"))]
os.environ["M2S_HOST_CHUNKS"] = "1"
for kn in knobs:
    os.environ.update(kn)
    with m2s.Context() as ctx:
        line = " ".join(f"{k[4:]}={v}" for k, v in kn.items()) + ":"
        for name, (verts, tris, grid, sign) in cases.items():
            best = None
            for r in range(3):
                out = ctx.grid_sdf(verts, tris, grid, sign)
                t = ctx.timings()
                k = t["total_ms"]
                if best is None or k < best[0]: best = (k, t)
            if name not in ref: ref[name] = out.copy()
            same = np.array_equal(ref[name].view(np.uint32), out.view(np.uint32))
            t = best[1]
            st = ctx.debug_stats()
            if st[2]: line += f" [n/w {st[0]/st[2]:.0f} l/w {st[1]/st[2]:.0f}]"
            line += f"  {name} total {best[0]:.3f} ms (d2h {t['d2h_ms']:.2f} build {t['build_ms']:.2f} sign {t['sign_ms']:.2f} seed {t['seed_ms']:.2f} dist {t['dist_ms']:.2f} same={same})"
        print(line, flush=True)
