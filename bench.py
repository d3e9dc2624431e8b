#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: Mvoxels/s of generate_grid_sdf on BASELINE config C3
(~100k-triangle watertight synthetic mesh, 256^3 grid, SignMethod::Raycast), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3|C2|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one full pass of the hot path over one slab: triangle records + Morton sort + LBVH + oriented boxes, the
Raycast row parities, the seeding passes and the nearest-triangle kernel with the sign epilogue.
  value : inputs already resident in HBM (m2s_generate_grid_sdf_device), timed with CUDA events on the stream
          the kernels are launched on; L2 is flushed between steps (256 MiB write, outside the event pairs).
  e2e   : the same step through the host-buffer C ABI call a facade user makes (m2s_generate_grid_sdf_slab):
          pinned host buffers, H2D of vertices+indices and D2H of the slab inside the timed region.
  N > 1 : weak scaling — every rank computes a 256-plane x-slab of a (256*N) x 256 x 256 grid over the same
          box (slabs along x, the slowest axis of Grid::get_cell_idx); no data-path collective.
  --impl reference : the reference's CPU algorithm (restated C++, oracle/, all host threads) on a bounded
          sample of the same workload; rank 0 only.
torch is plumbing here: device buffers, the stream, torch.distributed. The compute is libm2s.so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Nu, Nv, n, sign)  — BASELINE.md §4
    "C2": (64, 40, 128, 1),
    "C3": (256, 196, 256, 0),
    "C5": (1024, 490, 512, 0),
}
METRIC = "Mvoxels/s at 256^3 grid, 1/2/4/8 GPU vs ref CPU; HBM GB/s % peak"


def make_workload(name: str, world: int, scaling: str):
    from mesh_to_sdf_b200 import synth
    import mesh_to_sdf_b200 as m2s

    nu, nv, n, sign = WORKLOADS[name]
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    nx = n * world if scaling == "weak" else n
    grid = m2s.Grid.from_bounding_box(mn, mx, [nx, n, n])
    return verts, tris, grid, sign, n


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("k_grid_nearest_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step(oracle, verts, tris, grid, sign, planes: int, threads: int = 0):
    """The reference CPU path (generate/grid.rs restated, oracle/) on `planes` x-planes cut from the middle of
    the grid — same mesh, same cell size, same sign method. Returns (seconds, voxels)."""
    nx, ny, nz = grid.cell_count
    planes = min(planes, nx)
    x0 = (nx - planes) // 2
    first = grid.first_cell.copy()
    first[0] = np.float32(first[0] + np.float32(x0) * grid.cell_size[0])
    t0 = time.perf_counter()
    oracle.generate_grid_sdf_faithful(verts, tris, first, grid.cell_size, [planes, ny, nz], sign, threads)
    return time.perf_counter() - t0, planes * ny * nz


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    oracle.build()
    verts, tris, grid, sign, n = make_workload(args.workload, max(1, args.gpus), "weak")
    cores = oracle.hardware_threads()
    planes = args.ref_planes
    for _ in range(args.warmup):
        cpu_reference_step(oracle, verts, tris, grid, sign, planes)
    tot_s, tot_v = 0.0, 0
    for _ in range(args.steps):
        s, v = cpu_reference_step(oracle, verts, tris, grid, sign, planes)
        tot_s += s
        tot_v += v
    value = tot_v / tot_s / 1e6
    sample = (f"{planes} of {grid.cell_count[0]} x-planes (x {planes}x{n}x{n} voxels, middle of the grid) per step, "
              f"full mesh; faithful restatement of generate/grid.rs (restated C++, not rustc-built)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, verts, tris, grid, sign, max(1, args.gpus), "weak"),
        "cpu_baseline": {"value": value, "unit": "Mvoxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(name, verts, tris, grid, sign, world, scaling):
    return {
        "workload": f"{name}: bumpy torus T({WORKLOADS[name][0]},{WORKLOADS[name][1]}) {len(tris)} triangles / "
                    f"{len(verts)} vertices, grid {grid.cell_count[0]}x{grid.cell_count[1]}x{grid.cell_count[2]}, "
                    f"SignMethod::{'Raycast' if sign == 0 else 'Normal'}",
        "triangles": int(len(tris)), "vertices": int(len(verts)), "grid": list(grid.cell_count),
        "sign_method": "Raycast" if sign == 0 else "Normal",
        "partition": f"x-slabs, {world} rank(s), {scaling} scaling" if world > 1 else "single GPU, whole grid",
        "l2": "flushed between steps (256 MiB device write outside the per-step event pairs)",
        "step": "records + Morton sort + LBVH + oriented boxes + row parities + nearest kernel (neighbour seeds), per step",
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    import mesh_to_sdf_b200 as m2s
    from mesh_to_sdf_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: libm2s has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    verts, tris, grid, sign, n = make_workload(args.workload, world, args.scaling)
    nx, ny, nz = grid.cell_count
    x0, x1 = sharding.slab_bounds(nx, world)[rank]
    slab_cells = (x1 - x0) * ny * nz
    total_cells = nx * ny * nz

    stream = torch.cuda.current_stream(dev)
    ctx = m2s.Context([local_rank], stream=stream.cuda_stream)
    d_verts = torch.from_numpy(verts).to(dev)
    d_tris = torch.from_numpy(tris.view(np.int32)).to(dev)
    d_out = torch.empty(slab_cells, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        ctx.grid_sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), grid, sign, x0, x1,
                            d_out.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm ("value") ----
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step_device()
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, phase = [], {"build_ms": 0.0, "sign_ms": 0.0, "seed_ms": 0.0, "dist_ms": 0.0}
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        ctx.synchronize()  # also surfaces deferred data errors; outside the event pair's GPU time
        t = ctx.timings()
        kernel_ms.append(t["dist_ms"])
        for key in phase:
            phase[key] += t[key] / args.steps
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kern = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = total_cells * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end-to-end arm through the host-buffer ABI (pinned host memory) ----
    h_verts = torch.from_numpy(verts).pin_memory()
    h_tris = torch.from_numpy(tris.view(np.int32)).pin_memory()
    h_out = torch.empty(slab_cells, dtype=torch.float32).pin_memory()
    np_verts, np_tris, np_out = h_verts.numpy(), h_tris.numpy().view(np.uint32), h_out.numpy()
    for _ in range(2):
        ctx.grid_sdf_slab(np_verts, np_tris, grid, sign, x0, x1, np_out)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ctx.grid_sdf_slab(np_verts, np_tris, grid, sign, x0, x1, np_out)  # synchronous: returns with the slab on the host
        e2e_s += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    barrier()
    e2e_value = total_cells * args.steps / float(e2e_t.item()) / 1e6
    e2e_phase = ctx.timings()
    checksum = float(np.abs(np_out[:: max(1, slab_cells // 4096)]).sum())

    # ---- roofline of the dominant kernel (k_grid_nearest_run) ----
    b_alg = 4 * slab_cells + 12 * len(verts) + 12 * len(tris)  # output once + raw mesh once
    kern_ms = float(kern.item())
    peak, peak_src = measured_peak()
    achieved = b_alg / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "k_grid_nearest_run<Raycast, V=2>", "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": b_alg, "peak_source": peak_src,
                "note": "exact nearest-triangle search is issue/L1-bound, not HBM-bound (see DESIGN.md, profiles/)"}

    line = {
        "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, verts, tris, grid, sign, world, args.scaling),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mvoxels/s", "h2d_bytes_per_step": int(12 * len(verts) + 12 * len(tris)),
                "d2h_bytes_per_step": int(4 * slab_cells), "ms_per_step": float(e2e_t.item()) / args.steps * 1e3,
                "api": "m2s_generate_grid_sdf_slab (host buffers, pinned; H2D copy of the mesh, the result written into the pinned destination by the kernel's own stores over PCIe)", "phases_ms_last": e2e_phase,
                "checksum": checksum},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "phases_ms": phase,
    }

    # ---- CPU baseline beside it (rank 0, N == 1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        oracle.build()
        planes = args.cpu_planes
        secs, vox = cpu_reference_step(oracle, verts, tris, grid, sign, planes)
        line["cpu_baseline"] = {
            "value": vox / secs / 1e6, "unit": "Mvoxels/s", "cores": oracle.hardware_threads(), "kind": "port",
            "seconds": secs,
            "sample": f"{min(planes, nx)} of {nx} x-planes ({vox} voxels, middle of the grid), full mesh, one run; "
                      "faithful restatement of generate/grid.rs (restated C++, not rustc-built)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-planes", type=int, default=256, help="x-planes of the CPU baseline sample (ours arm)")
    ap.add_argument("--ref-planes", type=int, default=48, help="x-planes per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
