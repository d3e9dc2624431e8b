"""world_size-2 gloo test of the multi-rank host logic (slab split + reassembly); the per-rank slab values come
from the oracle here because there is no GPU — the GPU tests check the same entry point on real slabs."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from mesh_to_sdf_b200 import sharding, synth


def test_slab_bounds_cover_and_match_the_c_split():
    for nx in (1, 2, 5, 8, 31, 256, 257):
        for world in (1, 2, 3, 4, 8):
            b = sharding.slab_bounds(nx, world)
            assert b[0][0] == 0 and b[-1][1] == nx
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, plane, full, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x0, x1 = sharding.slab_bounds(nx, world)[rank]
        local = torch.from_numpy(full[x0 * plane: x1 * plane].copy())
        out = sharding.all_gather_slabs(local, nx, plane, rank, world)
        ret[rank] = bool(np.array_equal(out.numpy().view(np.uint32), full.view(np.uint32)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx", [8, 7])
def test_all_gather_slabs_world2(nx):
    import oracle
    verts, tris = synth.bumpy_torus(8, 6)
    mn, mx = synth.padded_grid_box(verts)
    first, size = oracle.grid_from_bounding_box(mn, mx, [nx, 5, 6])
    full = oracle.grid_cells_exact(verts, tris, first, size, [nx, 5, 6], 0)
    # slab-wise oracle == whole-grid oracle (independent units)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), nx, 30, full, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
