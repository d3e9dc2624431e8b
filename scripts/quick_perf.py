"""Quick timing of the host entry points on the BASELINE configs (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth

def run_grid(name, nu, nv, n, sign, reps=int(os.environ.get("REPS", "3")), cubic=False):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    if cubic:  # isotropic cells: the box grown to a cube around its centre
        c, h = 0.5 * (mn + mx), 0.5 * float(np.max(mx - mn))
        mn, mx = (c - h).astype(np.float32), (c + h).astype(np.float32)
    grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
    ctx = m2s.default_context()
    out = np.empty(n ** 3, np.float32)
    for r in range(reps):
        t0 = time.perf_counter()
        ctx.grid_sdf(verts, tris, grid, sign, out)
        dt = time.perf_counter() - t0
        t = ctx.timings()
        print(f"{name} rep{r}: wall {dt*1e3:.2f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in t.items()) +
              f"  -> {n**3/ (t['build_ms']+t['sign_ms']+t['dist_ms'])/1e3:.1f} Mvox/s (kernels)", flush=True)
    print(f"  neg frac {np.mean(out<0):.4f} min {out.min():.4f} max {out.max():.4f}")

def run_points(name, nu, nv, nq, accel, sign, reps=3):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    q = synth.splitmix64_points(nq, mn, mx)
    ctx = m2s.default_context()
    for r in range(reps):
        t0 = time.perf_counter()
        out = ctx.sdf(verts, tris, q, accel, sign)
        dt = time.perf_counter() - t0
        t = ctx.timings()
        print(f"{name} rep{r}: wall {dt*1e3:.2f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in t.items()), flush=True)
    print(f"  neg frac {np.mean(out<0):.4f}")

if __name__ == "__main__":
    which = sys.argv[1:] or ["C2", "C3", "C4"]
    if "C2" in which: run_grid("C2", 64, 40, 128, 1)
    if "C3" in which: run_grid("C3", 256, 196, 256, 0)
    if "C3I" in which: run_grid("C3I", 256, 196, 256, 0, cubic=True)
    if "C3N" in which: run_grid("C3N", 256, 196, 256, 1)
    if "C4" in which: run_points("C4", 640, 392, 1_000_000, 3, 0)
    if "C5" in which: run_grid("C5", 1024, 490, 512, 0, reps=2)
