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
knobs = [dict(zip(("M2S_SEED_LEVELS", "M2S_LEAF_SIZE", "M2S_FLAT"), k)) for k in itertools.product(("1", "2"), ("2", "4"), ("0", "0.1", "0.3", "1.0"))]
for kn in knobs:
    os.environ.update(kn)
    with m2s.Context() as ctx:
        line = " ".join(f"{k[4:]}={v}" for k, v in kn.items()) + ":"
        for name, (verts, tris, grid, sign) in cases.items():
            best = None
            for r in range(3):
                out = ctx.grid_sdf(verts, tris, grid, sign)
                t = ctx.timings()
                k = t["build_ms"] + t["sign_ms"] + t["seed_ms"] + t["dist_ms"]
                if best is None or k < best[0]: best = (k, t)
            if name not in ref: ref[name] = out.copy()
            same = np.array_equal(ref[name].view(np.uint32), out.view(np.uint32))
            t = best[1]
            line += f"  {name} {best[0]:.3f} ms (build {t['build_ms']:.2f} sign {t['sign_ms']:.2f} seed {t['seed_ms']:.2f} dist {t['dist_ms']:.2f} same={same})"
        print(line, flush=True)
