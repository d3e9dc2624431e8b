"""CPU tests of bench.py's host logic: the numpy Grid of the reference arm, the shared host buffer the ranks of the
multi-GPU e2e arm fill slab by slab (world_size-2 gloo), and that the reference arm never maps the product library."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_plain_grid_matches_the_oracle_helper(oracle):
    rng = np.random.default_rng(7)
    for _ in range(50):
        mn = rng.uniform(-3, 3, 3).astype(np.float32)
        mx = (mn + rng.uniform(0.01, 5, 3)).astype(np.float32)
        cnt = [int(c) for c in rng.integers(1, 600, 3)]
        g = bench.PlainGrid(mn, mx, cnt)
        first, size = oracle.grid_from_bounding_box(mn, mx, cnt)
        assert np.array_equal(g.first_cell.view(np.uint32), np.asarray(first, np.float32).view(np.uint32))
        assert np.array_equal(g.cell_size.view(np.uint32), np.asarray(size, np.float32).view(np.uint32))


def test_slab_bounds_match_the_library_split():
    from mesh_to_sdf_b200 import sharding
    for nx in (1, 7, 256, 512, 513):
        for world in (1, 2, 4, 8):
            assert bench.slab_bounds(nx, world) == sharding.slab_bounds(nx, world)


def test_reference_arm_line_and_no_product_library():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C2",
                        "--steps", "1", "--warmup", "0", "--ref-planes", "4"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["libm2s_mapped"] is False and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and "SAMPLED" in line["config"]["workload"]
    assert line["e2e"]["value"] == line["value"] and line["unit"] == "Mvoxels/s"


_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
class E:  # the part of bench.Env the shared buffer uses
    pass
env = E(); env.rank, env.world, env.dist = rank, world, dist
env.barrier = dist.barrier
nx, plane = 7, 30
buf = bench.SharedHostBuffer(env, 4 * nx * plane, "test")
x0, x1 = bench.slab_bounds(nx, world)[rank]
buf.array[x0 * plane:x1 * plane] = np.arange(x0 * plane, x1 * plane, dtype=np.float32) + 0.5
dist.barrier()
ok = bool(np.array_equal(buf.array, np.arange(nx * plane, dtype=np.float32) + 0.5))
buf.close()
dist.destroy_process_group()
print("rank", rank, "ok" if ok else "BAD")
"""


def test_shared_host_buffer_world2_gloo():
    pytest.importorskip("torch")
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", _WORKER, ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for rank, p in enumerate(procs):
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0 and f"rank {rank} ok" in out, err[-2000:]


def test_rebalance_moves_cuts_towards_equal_cost():
    # slab costs measured on C5 at 8 equal slabs (profiles/r2b_c5_slab_balance.log)
    bounds = bench.slab_bounds(512, 8)
    times = [6.67, 7.15, 8.42, 8.41, 8.82, 8.02, 6.74, 6.58]
    new = bench.rebalance(bounds, times)
    assert new[0][0] == 0 and new[-1][1] == 512
    assert all(a[1] == b[0] for a, b in zip(new, new[1:])) and all(x1 > x0 for x0, x1 in new)
    assert all(x0 % 4 == 0 for x0, _ in new)  # whole brick planes
    dens = np.concatenate([[t / (x1 - x0)] * (x1 - x0) for t, (x0, x1) in zip(times, bounds)])
    est = [float(dens[x0:x1].sum()) for x0, x1 in new]
    assert max(est) / (sum(est) / 8) < max(times) / (sum(times) / 8)  # better than the equal split
    assert max(est) / (sum(est) / 8) < 1.06
    # degenerate inputs keep a valid partition
    for t in ([0.0] * 8, [1.0] + [0.0] * 7, [1e-9] * 7 + [5.0]):
        nb = bench.rebalance(bounds, t)
        assert nb[0][0] == 0 and nb[-1][1] == 512 and all(x1 - x0 >= 4 for x0, x1 in nb)
    assert bench.rebalance([(0, 100)], [3.0]) == [(0, 100)]
