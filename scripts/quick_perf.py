"""Quick timing of the entry points on the BASELINE configs (development aid; bench.py is the contract).
usage: python scripts/quick_perf.py [C2 C3 C3N C4 C5 ...] ; LIBS="a.so,b.so" compares builds (one process each)."""
import sys, time, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def run_grid(m2s, synth, name, nu, nv, n, sign, reps=3):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
    ctx = m2s.default_context()
    pinned = m2s.host_alloc(n ** 3)
    pageable = np.empty(n ** 3, np.float32)
    pageable[:] = 0  # touched: page faults are not what is compared here
    for label, out, opt in [("zerocopy", pinned.array, m2s.HOST_AUTO), ("pipelined", pageable, m2s.HOST_AUTO),
                            ("staged", pageable, m2s.HOST_STAGED), ("registered", pageable, m2s.HOST_REGISTER)]:
        ctx.set_option(m2s.OPT_HOST_PATH, opt)
        best = None
        for r in range(reps + 1):
            t0 = time.perf_counter()
            ctx.grid_sdf(verts, tris, grid, sign, out)
            dt = (time.perf_counter() - t0) * 1e3
            t = ctx.timings()
            if r and (best is None or dt < best[0]):
                best = (dt, t)
        dt, t = best
        print(f"{name} {label:10s}: wall {dt:7.2f} ms | " + " ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in t.items()) +
              f" | {n**3 / dt / 1e3:.0f} Mvox/s e2e, {n**3 / t['dist_ms'] / 1e3:.0f} kernel", flush=True)
    ctx.set_option(m2s.OPT_HOST_PATH, m2s.HOST_AUTO)
    with ctx.mesh(verts, tris) as mesh:
        for r in range(2):
            t0 = time.perf_counter()
            mesh.grid_sdf(grid, sign, out=pinned.array)
            dt = (time.perf_counter() - t0) * 1e3
        print(f"{name} handle    : wall {dt:7.2f} ms | " + " ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in ctx.timings().items()), flush=True)
    print(f"  neg frac {np.mean(pinned.array < 0):.4f} checksum {float(np.abs(pinned.array[::4097]).sum()):.6f}")
    if os.environ.get("M2S_STATS"):
        ctx.debug_stats()
        ctx.grid_sdf(verts, tris, grid, sign, pinned.array)
        st = ctx.debug_stats()
        if st[2]:
            print(f"  stats per tile: nodes {st[0] / st[2]:.1f} leaves {st[1] / st[2]:.1f} tiles {st[2]} seedless {st[3] / st[2]:.3f}")
    pinned.close()


def run_points(m2s, synth, name, nu, nv, nq, accel, sign, reps=3):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    q = synth.splitmix64_points(nq, mn, mx)
    ctx = m2s.default_context()
    for r in range(reps):
        t0 = time.perf_counter()
        out = ctx.sdf(verts, tris, q, accel, sign)
        dt = time.perf_counter() - t0
        t = ctx.timings()
        print(f"{name} rep{r}: wall {dt*1e3:.2f} ms | " + " ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in t.items()), flush=True)
    print(f"  neg frac {np.mean(out<0):.4f} checksum {float(np.abs(out[::97]).sum()):.6f}")
    if os.environ.get("M2S_STATS"):
        ctx.debug_stats()
        ctx.sdf(verts, tris, q, accel, sign)
        st = ctx.debug_stats()
        if st[2]:
            print(f"  stats per packet: nodes {st[0] / st[2]:.1f} leaves {st[1] / st[2]:.1f} packets {st[2]}")


def main(which):
    import mesh_to_sdf_b200 as m2s
    from mesh_to_sdf_b200 import synth
    print("lib:", m2s.LIB_PATH, flush=True)
    if "C2" in which: run_grid(m2s, synth, "C2", 64, 40, 128, 1)
    if "C3" in which: run_grid(m2s, synth, "C3", 256, 196, 256, 0)
    if "C3N" in which: run_grid(m2s, synth, "C3N", 256, 196, 256, 1)
    if "C4" in which: run_points(m2s, synth, "C4", 640, 392, 1_000_000, 3, 0)
    if "C4N" in which: run_points(m2s, synth, "C4N", 640, 392, 1_000_000, 1, 1)
    if "C5" in which: run_grid(m2s, synth, "C5", 1024, 490, 512, 0, reps=2)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C2", "C3", "C4"]
    libs = [l for l in os.environ.get("LIBS", "").split(",") if l]
    if libs:
        for l in libs:
            env = dict(os.environ, M2S_LIB=os.path.abspath(l))
            env.pop("LIBS")
            subprocess.run([sys.executable, os.path.abspath(__file__)] + which, env=env)
    else:
        main(which)
