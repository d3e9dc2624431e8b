#!/usr/bin/env python
"""bench.py — benchmark of the hot path of mesh_to_sdf on B200: generate_grid_sdf / generate_sdf on the BASELINE
configs, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

N = 1 (default): workload C3 (the configuration BASELINE.json's metric is quoted on: ~100k triangles, 256^3 grid,
  Raycast); the other single-GPU configurations (C2, C4, C5 on one GPU) ride along in `extra_configs`.
N > 1: workload C5 (1M triangles, 512^3 grid, Raycast; C3 rides along in `extra_configs`), STRONG scaling: the
  x-slabs of ONE grid are computed by the N ranks and assembled into ONE flat result — device-resident on rank 0 (every rank's distance kernel stores its
  slab straight into rank 0's buffer over NVLink: cudaIpc mapping, no gather step) for `value`, one host buffer
  shared by the ranks for `e2e`. The timed region ends when the whole grid is there.

A step = one full pass of the hot path: triangle records + Morton sort + LBVH + oriented boxes (rebuilt every step,
like the reference rebuilds its structures every call), the Raycast row parities, the nearest-triangle kernel with
the sign epilogue.
  value : inputs already resident in HBM (m2s_generate_grid_sdf_device / m2s_generate_sdf_device), timed with CUDA
          events on the stream the kernels are launched on; L2 flushed between steps (256 MiB write).
  e2e   : the same step through the call a facade user makes, with the buffers the facade has: PAGEABLE host
          vertices / indices / queries and a pageable destination (include/mesh_to_sdf.hpp allocates a std::vector,
          the Rust facade a Vec) — H2D and D2H inside the timed region. `e2e.pinned` is the same call with a
          page-locked destination (m2s_host_alloc: the kernel's own stores write it in place).
  --impl reference : the reference's CPU algorithm (restated C++, oracle/, all host threads) on a bounded sample
          of the same workload; rank 0 only; never loads libm2s.so.
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

# name: (Nu, Nv, grid n, sign) / (Nu, Nv, queries) — BASELINE.md §4, SURVEY §8d
GRID_WORKLOADS = {"C2": (64, 40, 128, 1), "C3": (256, 196, 256, 0), "C5": (1024, 490, 512, 0)}
POINT_WORKLOADS = {"C4": (640, 392, 1_000_000)}
METRIC = "Mvoxels/s at 256^3 grid, 1/2/4/8 GPU vs ref CPU; HBM GB/s % peak"
RAYCAST, NORMAL, ACCEL_RTREE_BVH = 0, 1, 3


# ---------------------------------------------------------------------------------------------------------------
# workloads (numpy only: the reference arm must not load libm2s)
# ---------------------------------------------------------------------------------------------------------------
class PlainGrid:
    """Grid::from_bounding_box (src/grid.rs:59-74) in numpy float32 — same arithmetic as the library helper."""

    def __init__(self, bmin, bmax, count):
        bmin, bmax = np.asarray(bmin, np.float32), np.asarray(bmax, np.float32)
        self.cell_count = [int(c) for c in count]
        self.cell_size = ((bmax - bmin) / np.asarray(self.cell_count, np.float32)).astype(np.float32)
        self.first_cell = (bmin + self.cell_size * np.float32(0.5)).astype(np.float32)


def synth_module():
    """mesh_to_sdf_b200/synth.py loaded by path, so that the reference arm does not import the package."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("m2s_synth", os.path.join(ROOT, "mesh_to_sdf_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_grid_workload(name: str):
    synth = synth_module()
    nu, nv, n, sign = GRID_WORKLOADS[name]
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    return verts, tris, PlainGrid(mn, mx, [n, n, n]), sign, synth.mesh_diag(verts)


def make_point_workload(name: str):
    synth = synth_module()
    nu, nv, nq = POINT_WORKLOADS[name]
    verts, tris = synth.bumpy_torus(nu, nv)
    mn, mx = synth.padded_grid_box(verts)
    return verts, tris, synth.splitmix64_points(nq, mn, mx), synth.mesh_diag(verts)


def slab_bounds(nx: int, world: int):
    return [(nx * r // world, nx * (r + 1) // world) for r in range(world)]


def rebalance(bounds, times, align: int = 4):
    """New slab boundaries from the measured time of every slab: the cost per x-plane is taken as uniform inside a
    slab, the new cuts sit at equal shares of the cumulative cost, rounded to whole brick planes (`align` x-planes).
    bounds: [(x0, x1)] per rank, times: seconds or ms per rank."""
    world, nx = len(bounds), bounds[-1][1]
    total = float(sum(times))
    if world == 1 or total <= 0:
        return list(bounds)
    cuts, acc, r = [0], 0.0, 0
    for k in range(1, world):
        want = total * k / world
        while r < world - 1 and acc + times[r] < want:
            acc += times[r]
            r += 1
        x0, x1 = bounds[r]
        x = x0 + (x1 - x0) * (want - acc) / max(times[r], 1e-12)
        x = int(round(x / align)) * align
        cuts.append(min(max(x, cuts[-1] + align), nx - align * (world - k)))
    cuts.append(nx)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def grid_config(name, verts, tris, grid, sign, world):
    nu, nv = GRID_WORKLOADS[name][:2]
    return {
        "workload": f"{name}: bumpy torus T({nu},{nv}) {len(tris)} triangles / {len(verts)} vertices, grid "
                    f"{grid.cell_count[0]}x{grid.cell_count[1]}x{grid.cell_count[2]}, "
                    f"SignMethod::{'Raycast' if sign == 0 else 'Normal'}",
        "triangles": int(len(tris)), "vertices": int(len(verts)), "grid": list(grid.cell_count),
        "sign_method": "Raycast" if sign == 0 else "Normal",
        "partition": (f"x-slabs of ONE grid over {world} ranks, strong scaling, assembled into one flat result "
                      f"(device: peer-mapped stores into rank 0's buffer; host: one shared buffer)")
        if world > 1 else "single GPU, whole grid",
        "l2": "flushed between steps (256 MiB device write outside the per-step event pairs)",
        "step": "records + Morton sort + LBVH + oriented boxes + row parities + nearest kernel, rebuilt every step",
    }


def points_config(name, verts, tris, nq):
    nu, nv = POINT_WORKLOADS[name][:2]
    return {
        "workload": f"{name}: generate_sdf, {nq} scattered queries (splitmix64, uniform in the padded box) on a bumpy "
                    f"torus T({nu},{nv}) {len(tris)} triangles / {len(verts)} vertices, AccelerationMethod::RtreeBvh",
        "triangles": int(len(tris)), "vertices": int(len(verts)), "queries": int(nq),
        "l2": "flushed between steps (256 MiB device write outside the per-step event pairs)",
        "step": "records + Morton sort + LBVH + oriented boxes + query Morton sort + nearest kernel with the three "
                "axis-ray parities, rebuilt every step",
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(key: str):
    """dram bytes per launch of a kernel from the committed ncu capture (profiles/roofline_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
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
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout: float = 8.0):
        """nvidia-smi takes a second or two to deliver its first sample on an 8-GPU box: wait for it, so that a
        short timed region is not left without samples."""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Start of the timed region: only samples from here on are reported."""
        self.t_mark = time.perf_counter()

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
        t_mark = getattr(self, "t_mark", 0.0)
        timed = [l for t, l in self.lines if t >= t_mark] or [l for _, l in self.lines[-3:]]
        for l in timed:
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


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (oracle/ = test infrastructure: only here, as the checker / the reported baseline)
# ---------------------------------------------------------------------------------------------------------------
def cpu_grid_sample(oracle, verts, tris, grid, sign, planes: int, threads: int = 0):
    """The reference CPU path (generate/grid.rs restated, oracle/) on `planes` x-planes cut from the middle of the
    grid — same mesh, same cell size, same sign method. Returns (seconds, voxels, sdf of the sample, first plane)."""
    nx, ny, nz = grid.cell_count
    planes = min(planes, nx)
    x0 = (nx - planes) // 2
    first = grid.first_cell.copy()
    first[0] = np.float32(first[0] + np.float32(x0) * grid.cell_size[0])
    t0 = time.perf_counter()
    sdf = oracle.generate_grid_sdf_faithful(verts, tris, first, grid.cell_size, [planes, ny, nz], sign, threads)[0]
    return time.perf_counter() - t0, planes * ny * nz, sdf, x0


def grid_parity(oracle, verts, tris, grid, sign, diag, gpu_sdf, faithful, fx0, n_exact=3000):
    """the three deviation numbers of a grid config. gpu_sdf: our full grid (host); faithful: the
    reference restatement on the x-planes [fx0, fx0 + planes) (the reference's grid driver propagates candidates
    between neighbouring cells and can miss the true nearest triangle: generic/bvh.rs:237-239 'sometimes fails')."""
    nx, ny, nz = grid.cell_count
    tol = 1e-4 * diag
    rng = np.random.default_rng(20261017)
    idx = rng.choice(nx * ny * nz, size=min(n_exact, nx * ny * nz), replace=False).astype(np.uint64)
    exact = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, sign, idx)
    got = gpu_sdf[idx.astype(np.int64)]
    out = {"tolerance": tol, "tolerance_rule": "1e-4 * mesh bbox diagonal",
           "exact_sample_cells": int(len(idx)),
           "max_abs_vs_exact_sample": float(np.max(np.abs(got - exact))),
           "bit_exact_vs_exact_sample": bool(np.array_equal(got.view(np.uint32), exact.view(np.uint32))),
           "sign_mismatch_vs_exact_sample": int(np.sum(np.signbit(got) != np.signbit(exact)))}
    planes = len(faithful) // (ny * nz)
    ours = gpu_sdf[fx0 * ny * nz:(fx0 + planes) * ny * nz]
    diff = np.abs(faithful) - np.abs(ours)
    off = np.abs(diff) > tol
    out.update({
        "faithful_cells": int(len(faithful)),
        "frac_gt_tol_vs_faithful": float(np.mean(off)),
        "one_sided": bool(np.all(diff[off] > 0)),  # wherever they differ the reference's |d| is the LARGER one
        "max_excess_of_faithful": float(diff.max()) if len(diff) else 0.0,
        "sign_mismatch_frac_vs_faithful": float(np.mean(np.signbit(faithful[~off]) != np.signbit(ours[~off]))),
        "note": "faithful = step-by-step restatement of generate/grid.rs (restated C++, not rustc-built); its "
                "propagation over-estimates some cells, ours is the exact minimum",
    })
    return out


def libm2s_mapped() -> bool:
    try:
        with open("/proc/self/maps") as f:
            return any("libm2s.so" in line for line in f)
    except OSError:
        return False


def run_reference(args):
    """The reference arm: rank 0 only, CPU only; never imports mesh_to_sdf_b200 (asserted on /proc/self/maps)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    oracle.build()
    cores = oracle.hardware_threads()
    name = args.workload or ("C3" if args.gpus <= 1 else "C5")
    if name in POINT_WORKLOADS:
        verts, tris, queries, _ = make_point_workload(name)
        nq = min(len(queries), args.ref_queries)
        unit, cfg = "Mqueries/s", points_config(name, verts, tris, len(queries))
        sample = (f"SAMPLED: the first {nq} of {len(queries)} queries per step, full mesh, both trees rebuilt per step; "
                  "tree-accelerated restatement of generic/rtree_bvh.rs (restated C++, not rustc-built)")

        def step():
            t0 = time.perf_counter()
            oracle.generate_sdf_tree(verts, tris, queries[:nq], ACCEL_RTREE_BVH, RAYCAST, 0)
            return time.perf_counter() - t0, nq
    else:
        verts, tris, grid, sign, _ = make_grid_workload(name)
        planes = min(args.ref_planes if name != "C5" else min(args.ref_planes, 16), grid.cell_count[0])
        n = grid.cell_count[1]
        unit, cfg = "Mvoxels/s", grid_config(name, verts, tris, grid, sign, max(1, args.gpus))
        sample = (f"SAMPLED: {planes} of {grid.cell_count[0]} x-planes per step ({planes}x{n}x{n} voxels from the "
                  f"middle of the grid), full mesh; faithful restatement of generate/grid.rs (restated C++, not "
                  f"rustc-built)")

        def step():
            s, v, _, _ = cpu_grid_sample(oracle, verts, tris, grid, sign, planes)
            return s, v
    cfg["workload"] += " — reference arm " + sample
    for _ in range(args.warmup):
        step()
    tot_s, tot_v = 0.0, 0
    for _ in range(args.steps):
        s, v = step()
        tot_s += s
        tot_v += v
    value = tot_v / tot_s / 1e6
    assert not libm2s_mapped(), "the reference arm must not load the product library"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "libm2s_mapped": False,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one process per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("no CUDA device: libm2s has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_list(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def timed_device_steps(env, ctx, step, steps, warmup, phase_keys=("build_ms", "sign_ms", "seed_ms", "dist_ms"),
                       sampler=None):
    """W untimed + K timed steps of `step()` (enqueue only) with an L2 flush before each, CUDA events on the launching
    stream. Returns (sum of step ms [max over ranks], mean kernel ms [max over ranks], phases of this rank, launches)."""
    torch = env.torch
    for _ in range(warmup):
        env.flush.fill_(1)
        step()
    ctx.synchronize()
    if sampler is not None:
        sampler.wait_ready()
    env.barrier()
    if sampler is not None:
        sampler.mark()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    phases = {k: 0.0 for k in phase_keys}
    kernel_ms = []
    for k in range(steps):
        env.flush.fill_(k & 0xff)
        ev[k][0].record(env.stream)
        step()
        ev[k][1].record(env.stream)
        ctx.synchronize()  # also surfaces deferred data errors; outside the event pair's GPU time
        t = ctx.timings()
        kernel_ms.append(t["dist_ms"])
        for key in phases:
            phases[key] += t[key] / steps
    env.barrier()
    launches = ctx.launch_count - launches0
    total_ms = env.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    kern_ms = env.max_over_ranks(float(np.mean(kernel_ms)))
    return total_ms, kern_ms, phases, launches


def timed_host_steps(env, call, steps, warmup=2):
    """K synchronous host-buffer calls bracketed by barriers; returns seconds (max over ranks)."""
    for _ in range(warmup):
        call()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    s = time.perf_counter() - t0
    s = env.max_over_ranks(s)
    env.barrier()
    return s


def run_kernel_name(sign, grid, n_tris):
    """Name of the distance kernel's instantiation: the library picks the run length V (voxels per lane) from the
    shape of the WHOLE grid and the mesh, also for a slab of it — csrc/m2s_grid.cu grid_run_length, restated here
    for the label only."""
    cells = float(grid.cell_count[0]) * grid.cell_count[1] * grid.cell_count[2]
    sx, sy, sz = (abs(float(v)) for v in grid.cell_size)
    thin_z = 16.0 * sz <= 1.5 * max(2.0 * sx, 4.0 * sy)
    fine = cells >= 64.0 * n_tris
    if cells >= 8.0e6 and fine and thin_z:
        v = "V=4"
    elif not thin_z and not fine:
        v = "V=2, lane layout 1"
    else:
        v = "V=2"
    return f"k_grid_nearest_run<{'Raycast' if sign == 0 else 'Normal'}, {v}>"


def roofline_block(kernel, kern_ms, b_alg, traffic_key, note):
    peak, peak_src = measured_peak()
    achieved = b_alg / (kern_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(traffic_key), "kernel": kernel, "kernel_ms": kern_ms,
            "algorithmic_bytes_per_launch": int(b_alg), "peak_source": peak_src, "note": note}


ISSUE_NOTE = ("exact nearest-triangle search is issue-bound, not HBM-bound: the instruction count per voxel, not "
              "bytes, sets its time (DESIGN.md §3, profiles/)")


def bench_grid_single(env, m2s, name, steps, warmup, want_cpu, cpu_planes, want_parity, host_steps=None,
                      want_post=False):
    """One grid config on ONE GPU: device-resident value, e2e through the facade's pageable buffers (+ pinned),
    roofline, optional CPU baseline + parity numbers."""
    torch = env.torch
    verts, tris, pgrid, sign, diag = make_grid_workload(name)
    grid = m2s.Grid(pgrid.first_cell, pgrid.cell_size, pgrid.cell_count)
    nx, ny, nz = grid.cell_count
    cells = nx * ny * nz
    host_steps = host_steps or steps
    ctx = m2s.Context([env.local_rank], stream=env.stream.cuda_stream)
    d_verts = torch.from_numpy(verts).to(env.dev)
    d_tris = torch.from_numpy(tris.view(np.int32)).to(env.dev)
    d_out = torch.empty(cells, dtype=torch.float32, device=env.dev)

    def step_device():
        ctx.grid_sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), grid, sign, 0, nx,
                            d_out.data_ptr())

    sampler = ClockSampler(env.local_rank)
    sampler.start()
    total_ms, kern_ms, phases, launches = timed_device_steps(env, ctx, step_device, steps, warmup, sampler=sampler)
    clocks = sampler.stop()
    value = cells * steps / (total_ms * 1e-3) / 1e6
    post = bench_post_passes(env, ctx, m2s, grid, d_out, cells) if want_post else None
    del d_out

    # ---- e2e: the facade call, pageable everything (what include/mesh_to_sdf.hpp / the Rust facade hand over) ----
    topo = m2s.Topology.TriangleList(tris.reshape(-1))  # the caller's flat u32 index list
    sign_m = m2s.SignMethod(sign)
    result = {}

    def call_facade():  # allocates its own result like the reference (generate/grid.rs:376 returns a fresh Vec)
        result["sdf"] = m2s.generate_grid_sdf(verts, topo, grid, sign_m, ctx=ctx)

    s_facade = timed_host_steps(env, call_facade, host_steps)
    facade_phase = ctx.timings()
    gpu_sdf = result["sdf"]
    reuse = np.empty(cells, np.float32)
    reuse[:] = 0

    def call_reuse():  # the same pageable path into a destination the caller keeps across calls
        ctx.grid_sdf(verts, tris, grid, sign, reuse)

    s_reuse = timed_host_steps(env, call_reuse, host_steps)
    pinned = m2s.host_alloc(cells)

    def call_pinned():
        ctx.grid_sdf(verts, tris, grid, sign, pinned.array)

    s_pinned = timed_host_steps(env, call_pinned, host_steps)
    pinned_phase = ctx.timings()
    same = bool(np.array_equal(pinned.array.view(np.uint32), gpu_sdf.view(np.uint32)))
    pinned.close()
    h2d = int(12 * len(verts) + 12 * len(tris))
    e2e = {
        "value": cells * host_steps / s_facade / 1e6, "unit": "Mvoxels/s", "h2d_bytes_per_step": h2d,
        "d2h_bytes_per_step": int(4 * cells), "ms_per_step": s_facade / host_steps * 1e3,
        "api": "generate_grid_sdf(vertices, Topology::TriangleList(indices), grid, sign) -> fresh pageable array: "
               "m2s_generate_grid_sdf with pageable inputs and destination (index expansion, allocation and the "
               "first-touch page faults of the result inside the timed region)",
        "host_path": facade_phase["host_path"], "phases_ms_last": facade_phase,
        "reused_pageable_destination": {"value": cells * host_steps / s_reuse / 1e6,
                                        "ms_per_step": s_reuse / host_steps * 1e3},
        "pinned": {"value": cells * host_steps / s_pinned / 1e6, "ms_per_step": s_pinned / host_steps * 1e3,
                   "host_path": pinned_phase["host_path"],
                   "api": "m2s_generate_grid_sdf into an m2s_host_alloc destination (written in place by the "
                          "kernel's stores over PCIe)"},
        "pinned_equals_pageable_bitwise": same,
        "checksum": float(np.abs(gpu_sdf[:: max(1, cells // 4096)]).sum()),
    }
    b_alg = 4 * cells + 12 * len(verts) + 12 * len(tris)  # output once + raw mesh once
    line = {
        "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": grid_config(name, verts, tris, grid, sign, 1),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline_block(run_kernel_name(sign, grid, len(tris)), kern_ms, b_alg,
                                   f"k_grid_nearest_dram_bytes_per_launch_{name}", ISSUE_NOTE),
        "phases_ms": phases,
    }
    if post is not None:
        line["post_passes"] = post
    if want_cpu:
        import oracle

        oracle.build()
        secs, vox, faithful, fx0 = cpu_grid_sample(oracle, verts, tris, grid, sign, cpu_planes)
        line["cpu_baseline"] = {
            "value": vox / secs / 1e6, "unit": "Mvoxels/s", "cores": oracle.hardware_threads(), "kind": "port",
            "seconds": secs,
            "sample": f"{min(cpu_planes, nx)} of {nx} x-planes ({vox} voxels, middle of the grid), full mesh, one run; "
                      "faithful restatement of generate/grid.rs (restated C++, not rustc-built)"}
        if want_parity:
            line["parity"] = grid_parity(oracle, verts, tris, grid, sign, diag, gpu_sdf, faithful, fx0)
    ctx.close()
    return line


def bench_post_passes(env, ctx, m2s, grid, d_sdf, cells, reps=10):
    """What the reference's in-repo caller does with the grid next, on the device-resident grid:
    render order + iso limits (mesh_to_sdf_client/src/sdf.rs:65-68, :123) and sampling with interpolation
    (shaders/draw_raymarching.wgsl:118-200). HBM-bound passes: algorithmic bytes over the measured copy peak."""
    torch = env.torch
    peak, _ = measured_peak()
    out = {}
    d_order = torch.empty(cells, dtype=torch.int32, device=env.dev)
    d_mm = torch.empty(2, dtype=torch.float32, device=env.dev)

    def timed(fn):
        for _ in range(3):
            fn()
        ctx.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = []
        for _ in range(reps):
            env.flush.fill_(7)
            ev0.record(env.stream)
            fn()
            ev1.record(env.stream)
            ctx.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        return float(np.median(ms))

    ms = timed(lambda: ctx.grid_order_device(d_sdf.data_ptr(), cells, d_order.data_ptr(), d_mm.data_ptr()))
    b = 8 * cells + 8  # the grid read once, the order written once, (min, max)
    sort_b = 60 * cells  # what the four radix passes move (csrc/m2s_sort.cuh): 4 + 12 | 16 | 16 | 12 bytes per cell
    out["grid_order"] = {"ms": ms, "Mcells_per_s": cells / ms / 1e3, "algorithmic_bytes": int(b),
                         "roofline_frac_hbm": b / (ms * 1e-3) / 1e9 / peak,
                         "sort_pass_bytes": int(sort_b), "sort_pass_frac_hbm": sort_b / (ms * 1e-3) / 1e9 / peak,
                         "what": "m2s_grid_order_device: the library's own onesweep radix sort (4 x 8 bits, keys from the "
                                 "distances on the fly, cell indices as payloads) + min / max; no library kernel"}
    n_pts = 4_000_000
    lo = torch.tensor(np.asarray(grid.first_cell), device=env.dev)
    hi = torch.tensor(np.asarray(grid.get_last_cell(), np.float32), device=env.dev)
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(1234)
    pts = (lo + (hi - lo) * torch.rand((n_pts, 3), device=env.dev, generator=gen)).contiguous()
    d_val = torch.empty(n_pts, dtype=torch.float32, device=env.dev)
    for mode, mname in ((1, "trilinear"), (2, "tetrahedral")):
        ms = timed(lambda: ctx.sample_grid_sdf_device(d_sdf.data_ptr(), grid, pts.data_ptr(), n_pts, mode, 0.0,
                                                      d_val.data_ptr()))
        b = 16 * n_pts  # 12 B point + 4 B result; the gathered corners hit L2 (the grid is 64 MiB)
        out["sample_" + mname] = {"ms": ms, "Msamples_per_s": n_pts / ms / 1e3, "algorithmic_bytes": int(b),
                                  "roofline_frac_hbm": b / (ms * 1e-3) / 1e9 / peak,
                                  "what": f"m2s_sample_grid_sdf_device, {n_pts} uniform points, {mname} on the dual grid"}
    return out


def bench_points_single(env, m2s, name, steps, warmup, want_cpu, cpu_queries):
    torch = env.torch
    verts, tris, queries, diag = make_point_workload(name)
    nq = len(queries)
    ctx = m2s.Context([env.local_rank], stream=env.stream.cuda_stream)
    d_verts = torch.from_numpy(verts).to(env.dev)
    d_tris = torch.from_numpy(tris.view(np.int32)).to(env.dev)
    d_q = torch.from_numpy(queries).to(env.dev)
    d_out = torch.empty(nq, dtype=torch.float32, device=env.dev)

    def step_device():
        ctx.sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), d_q.data_ptr(), nq,
                       ACCEL_RTREE_BVH, RAYCAST, d_out.data_ptr())

    total_ms, kern_ms, phases, launches = timed_device_steps(env, ctx, step_device, steps, warmup)
    value = nq * steps / (total_ms * 1e-3) / 1e6
    topo = m2s.Topology.TriangleList(tris.reshape(-1))
    result = {}

    def call_facade():
        result["sdf"] = m2s.generate_sdf(verts, topo, queries, m2s.AccelerationMethod.RtreeBvh, ctx=ctx)

    s_facade = timed_host_steps(env, call_facade, steps)
    facade_phase = ctx.timings()
    gpu = result["sdf"]
    b_alg = 16 * nq + 12 * len(verts) + 12 * len(tris)  # B_points
    line = {
        "metric": "Mqueries/s, generate_sdf 1M scattered queries x 500k triangles (RtreeBvh), 1 GPU vs ref CPU",
        "value": value, "unit": "Mqueries/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": points_config(name, verts, tris, nq),
        "e2e": {"value": nq * steps / s_facade / 1e6, "unit": "Mqueries/s",
                "h2d_bytes_per_step": int(12 * len(verts) + 12 * len(tris) + 12 * nq), "d2h_bytes_per_step": int(4 * nq),
                "ms_per_step": s_facade / steps * 1e3,
                "api": "generate_sdf(vertices, Topology::TriangleList(indices), queries, RtreeBvh) -> fresh pageable "
                       "array: m2s_generate_sdf with pageable inputs and destination",
                "phases_ms_last": facade_phase, "checksum": float(np.abs(gpu[::97]).sum())},
        "gpu_launches": int(launches),
        "roofline": roofline_block("k_points_run<UNSIGNED, 3 axis rays>", kern_ms, b_alg,
                                   "k_points_run_dram_bytes_per_launch_C4", ISSUE_NOTE),
        "phases_ms": phases,
    }
    if want_cpu:
        import oracle

        oracle.build()
        n = min(nq, cpu_queries)
        t0 = time.perf_counter()
        ref, ms = oracle.generate_sdf_tree(verts, tris, queries[:n], ACCEL_RTREE_BVH, RAYCAST, 0)
        secs = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": n / secs / 1e6, "unit": "Mqueries/s", "cores": oracle.hardware_threads(), "kind": "port",
            "seconds": secs, "build_ms": float(ms[0]), "query_ms": float(ms[1]),
            "sample": f"the first {n} of {nq} queries, full mesh, trees built once; tree-accelerated restatement of "
                      "generic/rtree_bvh.rs (restated C++, not rustc-built)"}
        line["parity"] = {"cells": int(n), "bit_exact_vs_cpu_tree_restatement":
                          bool(np.array_equal(ref.view(np.uint32), gpu[:n].view(np.uint32))),
                          "max_abs": float(np.max(np.abs(ref - gpu[:n]))), "tolerance": 1e-4 * diag}
    ctx.close()
    return line


class SharedHostBuffer:
    """One host buffer that every rank maps (POSIX shared memory; a file under /tmp if /dev/shm is too small)."""

    def __init__(self, env, nbytes: int, tag: str):
        from multiprocessing import shared_memory

        self.env, self.shm, self.mm, self.path = env, None, None, None
        name = None
        if env.rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
                name = ("shm", self.shm.name)
            except Exception:
                self.path = f"/tmp/m2s_bench_{tag}_{os.getpid()}.bin"
                with open(self.path, "wb") as f:
                    f.truncate(nbytes)
                name = ("file", self.path)
        box = [name]
        if env.world > 1:
            env.dist.broadcast_object_list(box, src=0)
        kind, ident = box[0]
        if kind == "shm":
            if env.rank != 0:
                self.shm = shared_memory.SharedMemory(name=ident)
                try:  # only the creating rank owns the segment (Python < 3.13 would unlink it from every process)
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
            self.array = np.ndarray((nbytes // 4,), dtype=np.float32, buffer=self.shm.buf)
        else:
            self.mm = np.memmap(ident, dtype=np.float32, mode="r+", shape=(nbytes // 4,))
            self.array = self.mm
        self.kind = kind

    def close(self):
        self.array = None
        self.env.barrier()
        if self.shm is not None:
            self.shm.close()
            if self.env.rank == 0:
                self.shm.unlink()
        if self.mm is not None:
            del self.mm
        if self.path and self.env.rank == 0:
            os.unlink(self.path)


def bench_grid_multi(env, m2s, name, steps, warmup, balance=True):
    """Strong scaling of ONE grid over the ranks, assembled into one flat result."""
    torch, dist = env.torch, env.dist
    world, rank = env.world, env.rank
    verts, tris, pgrid, sign, diag = make_grid_workload(name)
    grid = m2s.Grid(pgrid.first_cell, pgrid.cell_size, pgrid.cell_count)
    nx, ny, nz = grid.cell_count
    plane, cells = ny * nz, nx * ny * nz
    bounds = slab_bounds(nx, world)
    ctx = m2s.Context([env.local_rank], stream=env.stream.cuda_stream)
    d_verts = torch.from_numpy(verts).to(env.dev)
    d_tris = torch.from_numpy(tris.view(np.int32)).to(env.dev)

    # rank 0 owns the flat device grid; the others map it (cudaIpc) and their kernels store into it over NVLink
    base = ctx.device_alloc(4 * cells) if rank == 0 else 0
    box = [ctx.ipc_export(base) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    mapped = base if rank == 0 else ctx.ipc_open(box[0])
    cut = {"x0": bounds[rank][0], "x1": bounds[rank][1]}
    sampler = ClockSampler(env.local_rank)
    sampler.start()

    def step_device():
        ctx.grid_sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), grid, sign, cut["x0"],
                            cut["x1"], mapped + 4 * cut["x0"] * plane)

    # Equal-width slabs are uneven in cost (profiles/r2b_c5_slab_balance.log: 0.86 at 8 slabs). A service that
    # regenerates grids keeps the split of its last call: the cuts are moved to equal shares of the measured kernel
    # time during up to 8 UNTIMED steps, then frozen for the warm-up and the timed steps.
    balance_log = [[list(b) for b in bounds]]
    if balance:
        for _ in range(8):
            env.flush.fill_(3)
            step_device()
            ctx.synchronize()
            times = env.gather_list(float(ctx.timings()["dist_ms"]))
            if max(times) <= 1.02 * min(times):
                break
            bounds = rebalance(bounds, times)
            cut["x0"], cut["x1"] = bounds[rank]
            balance_log.append([list(b) for b in bounds])
        env.barrier()
    x0, x1 = cut["x0"], cut["x1"]

    total_ms, kern_ms, phases, launches = timed_device_steps(env, ctx, step_device, steps, warmup, sampler=sampler)
    clocks = sampler.stop()
    value = cells * steps / (total_ms * 1e-3) / 1e6
    rank_table = env.gather_list({"rank": rank, "x": [x0, x1], **{k: round(v, 4) for k, v in phases.items()}})
    all_launches = sum(env.gather_list(int(launches)))

    # the assembled device grid, checked on rank 0 against the same grid computed by rank 0 alone (also the strong-
    # scaling base: the same workload on one GPU, in the same run)
    single = None
    if rank == 0:
        d_ref = torch.empty(cells, dtype=torch.float32, device=env.dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = []
        for k in range(3):
            env.flush.fill_(k)
            ev0.record(env.stream)
            ctx.grid_sdf_device(d_verts.data_ptr(), len(verts), d_tris.data_ptr(), len(tris), grid, sign, 0, nx,
                                d_ref.data_ptr())
            ev1.record(env.stream)
            ctx.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        try:
            from cuda.bindings import runtime as cudart
        except ImportError:
            from cuda import cudart
        assembled = torch.empty(cells, dtype=torch.float32, device=env.dev)
        rc = cudart.cudaMemcpy(assembled.data_ptr(), base, 4 * cells, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
        assert int(rc[0]) == 0
        single = {"ms_per_step": float(np.mean(ms[1:])), "value": cells / (float(np.mean(ms[1:])) * 1e-3) / 1e6,
                  "unit": "Mvoxels/s", "what": "the same workload, whole grid, rank 0's GPU alone (device-resident)",
                  "assembled_equals_single_gpu_bitwise": bool(torch.equal(assembled.view(torch.int32),
                                                                          d_ref.view(torch.int32)))}
        del d_ref, assembled
    env.barrier()

    # ---- e2e: pageable host inputs on every rank, ONE shared host buffer as the destination ----
    shared = SharedHostBuffer(env, 4 * cells, "grid")
    mine = shared.array[x0 * plane:x1 * plane]

    def call_host():
        ctx.grid_sdf_slab(verts, tris, grid, sign, x0, x1, mine)

    s_pageable = timed_host_steps(env, call_host, steps)
    path_pageable = ctx.timings()["host_path"]
    checksum = float(np.abs(shared.array[:: max(1, cells // 4096)]).sum()) if rank == 0 else 0.0
    m2s.host_register(mine)  # page-locked once: every rank's kernel then stores its slab in place
    s_registered = timed_host_steps(env, call_host, steps)
    path_registered = ctx.timings()["host_path"]
    e2e_rank = env.gather_list({"rank": rank, **{k: (round(v, 4) if isinstance(v, float) else v)
                                                 for k, v in ctx.timings().items()}})
    m2s.host_unregister(mine)
    del mine
    shared.close()
    h2d = int(12 * len(verts) + 12 * len(tris))
    e2e = {
        "value": cells * steps / s_pageable / 1e6, "unit": "Mvoxels/s", "h2d_bytes_per_step": h2d * world,
        "d2h_bytes_per_step": int(4 * cells), "ms_per_step": s_pageable / steps * 1e3,
        "api": f"m2s_generate_grid_sdf_slab on every rank: pageable mesh inputs, destination = the rank's slab of ONE "
               f"pageable host buffer shared by the ranks ({shared.kind})",
        "host_path": path_pageable,
        "registered": {"value": cells * steps / s_registered / 1e6, "ms_per_step": s_registered / steps * 1e3,
                       "host_path": path_registered,
                       "api": "the same shared buffer page-locked once with m2s_host_register: written in place by "
                              "the kernels' stores"},
        "per_rank_phases_ms_last": e2e_rank, "checksum": checksum,
    }
    slab_cells = (x1 - x0) * plane
    b_alg = 4 * slab_cells + 12 * len(verts) + 12 * len(tris)
    line = {
        "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": grid_config(name, verts, tris, grid, sign, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(all_launches),
        # traffic: no ncu capture of a slab launch exists (the committed one is the whole grid in one launch) -> null
        "roofline": roofline_block(run_kernel_name(sign, grid, len(tris)) + " (slowest rank's slab)", kern_ms, b_alg,
                                   None, ISSUE_NOTE),
        "phases_ms": phases, "per_rank_phases_ms": rank_table, "single_gpu_same_workload": single,
        "slab_cuts": {"method": "equal shares of the measured per-slab kernel time over up to 8 untimed steps, frozen before "
                                "the warm-up" if balance else "equal widths",
                      "history": balance_log},
        "assembly": "device: every rank's distance kernel stores its x-slab into rank 0's flat buffer through a "
                    "cudaIpc peer mapping (NVLink), no gather step; the LBVH is built on every rank (replicated: "
                    "a broadcast cannot start before rank 0's build ends, DESIGN.md §6)",
    }
    if rank != 0:
        ctx.ipc_close(mapped)
    env.barrier()
    if rank == 0:
        ctx.device_free(base)
    ctx.close()
    return line


def run_ours(args):
    env = Env(args)
    import mesh_to_sdf_b200 as m2s

    warmup = max(args.warmup, 3)
    if env.world > 1:
        name = args.workload or "C5"
        if name not in GRID_WORKLOADS:
            raise SystemExit("multi-GPU runs take a grid workload (C2, C3, C5)")
        line = bench_grid_multi(env, m2s, name, args.steps, warmup, balance=not args.equal_slabs)
        if not args.workload and not args.no_extra:
            # the metric's own grid (256^3, C3) over the same ranks, strong scaling into one assembled grid as well
            line["extra_configs"] = [bench_grid_multi(env, m2s, "C3", args.steps, warmup, balance=not args.equal_slabs)]
    else:
        name = args.workload or "C3"
        cpu = not args.no_cpu_baseline
        if name in POINT_WORKLOADS:
            line = bench_points_single(env, m2s, name, args.steps, warmup, cpu, args.cpu_queries)
        else:
            planes = args.cpu_planes if name != "C5" else min(args.cpu_planes, 16)
            line = bench_grid_single(env, m2s, name, args.steps, warmup, cpu, planes, cpu, want_post=name == "C3")
        if not args.workload and not args.no_extra:
            extras = []
            extras.append(bench_grid_single(env, m2s, "C2", args.steps, warmup, cpu, 128, cpu))
            extras.append(bench_points_single(env, m2s, "C4", args.steps, warmup, cpu, args.cpu_queries))
            extras.append(bench_grid_single(env, m2s, "C5", max(3, args.steps // 3), warmup, cpu, 8, False,
                                            host_steps=3))
            line["extra_configs"] = extras
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(GRID_WORKLOADS) + sorted(POINT_WORKLOADS),
                    help="default: C3 on one GPU (+ C2, C4, C5 in extra_configs), C5 on several")
    ap.add_argument("--cpu-planes", type=int, default=256, help="x-planes of the CPU baseline sample (ours arm)")
    ap.add_argument("--cpu-queries", type=int, default=200_000, help="queries of the C4 CPU baseline sample")
    ap.add_argument("--ref-planes", type=int, default=48, help="x-planes per step of the --impl reference arm")
    ap.add_argument("--ref-queries", type=int, default=200_000, help="queries per step of the reference arm (C4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs")
    ap.add_argument("--equal-slabs", action="store_true", help="N > 1: equal-width x-slabs instead of balanced cuts")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
