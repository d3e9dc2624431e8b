"""Multi-GPU development measurements (bench.py is the contract; this is the phase table behind it).

  python scripts/multi_gpu.py balance          one GPU: kernel time of every x-slab of C5 for N = 2, 4, 8
                                               (how uneven equal-width slabs are: predicts the strong-scaling loss)
  python scripts/multi_gpu.py context [N]      ONE process driving N GPUs (m2s_create(devices, n)): C5 into one
                                               device grid on GPU 0 (peer-mapped stores) and into one host buffer
                                               (pinned / pageable), replicated build vs build-once + broadcast
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth


def c5(n=512, nu=1024, nv=490):
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    return verts, tris, m2s.Grid.from_bounding_box(mn, mx, [n, n, n])


def balance():
    import torch

    verts, tris, grid = c5()
    nx, ny, nz = grid.cell_count
    d_out = torch.empty(nx * ny * nz, dtype=torch.float32, device="cuda")
    with m2s.Context([0]) as ctx, ctx.mesh(verts, tris) as mesh:
        for world in (1, 2, 4, 8, 16):
            ms = []
            for r in range(world):
                x0, x1 = nx * r // world, nx * (r + 1) // world
                for _ in range(2):
                    mesh.grid_sdf_device(grid, 0, x0, x1, d_out.data_ptr() + 4 * x0 * ny * nz)
                    ctx.synchronize()
                t = ctx.timings()
                ms.append(t["dist_ms"] + t["sign_ms"])
            print(f"N={world:2d}: slab kernel+rows ms = {' '.join(f'{m:.2f}' for m in ms)} | sum {sum(ms):.2f} "
                  f"max {max(ms):.2f} -> balance efficiency {sum(ms) / (world * max(ms)):.3f}", flush=True)


def context(n):
    import torch

    verts, tris, grid = c5()
    nx, ny, nz = grid.cell_count
    cells = nx * ny * nz
    devs = list(range(n))
    d_verts = torch.from_numpy(verts).to("cuda:0")
    d_tris = torch.from_numpy(tris.view(np.int32)).to("cuda:0")
    d_out = torch.empty(cells, dtype=torch.float32, device="cuda:0")
    pinned = m2s.host_alloc(cells)
    pageable = np.zeros(cells, np.float32)
    ref = None
    with m2s.Context(devs) as ctx:
        for mode, label in ((m2s.BUILD_REPLICATED, "replicated build"), (m2s.BUILD_BROADCAST, "build once + broadcast")):
            ctx.set_option(m2s.OPT_BUILD_MODE, mode)
            best = 1e9
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.grid_sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), grid, 0, 0, nx,
                                    d_out.data_ptr())
                ctx.synchronize()
                best = min(best, (time.perf_counter() - t0) * 1e3)
            rows = [ctx.timings(i) for i in range(n)]
            print(f"[{n} GPUs, {label}] device grid on GPU 0: wall {best:.2f} ms = {cells / best / 1e3:.0f} Mvox/s | "
                  + " | ".join(f"d{i}: build {t['build_ms']:.2f} rows {t['sign_ms']:.2f} dist {t['dist_ms']:.2f} "
                               f"total {t['total_ms']:.2f}" for i, t in enumerate(rows)), flush=True)
            got = d_out.cpu().numpy()
            if ref is None:
                ref = got
            else:
                print("   equals the replicated-build grid bitwise:", bool(np.array_equal(ref.view(np.uint32), got.view(np.uint32))))
            for out, lab in ((pinned.array, "pinned host"), (pageable, "pageable host")):
                best = 1e9
                for _ in range(3):
                    t0 = time.perf_counter()
                    ctx.grid_sdf(verts, tris, grid, 0, out)
                    best = min(best, (time.perf_counter() - t0) * 1e3)
                rows = [ctx.timings(i) for i in range(n)]
                print(f"[{n} GPUs, {label}] {lab}: wall {best:.2f} ms = {cells / best / 1e3:.0f} Mvox/s | "
                      + " | ".join(f"d{i}: {t['host_path']} h2d {t['h2d_ms']:.2f} build {t['build_ms']:.2f} dist "
                                   f"{t['dist_ms']:.2f} d2h {t['d2h_ms']:.2f}" for i, t in enumerate(rows)), flush=True)
                print("   equals the device grid bitwise:", bool(np.array_equal(ref.view(np.uint32), out.view(np.uint32))))
    pinned.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "balance"
    if what == "balance":
        balance()
    else:
        context(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
