"""Multi-GPU plumbing: one process per GPU, the grid sharded by contiguous slabs along x — the slowest axis
of ``Grid::get_cell_idx`` (``src/grid.rs:122-124``), so every rank owns one contiguous range of the flat output —
and query points sharded by contiguous index ranges. Every voxel / query depends only on the (replicated) mesh,
so there is no data-path collective; the all-gather below only reassembles the flat ``Vec<f32>`` when a caller
wants the whole grid on every rank.

torch.distributed is plumbing here (NCCL on the GPUs, gloo in the CPU tests); the compute stays in libm2s.
"""
from __future__ import annotations

from typing import List, Tuple


def slab_bounds(nx: int, world: int) -> List[Tuple[int, int]]:
    """x ranges ``[x0, x1)`` per rank; identical to the split inside ``m2s_generate_grid_sdf`` (m2s_api.cu)."""
    return [(nx * r // world, nx * (r + 1) // world) for r in range(world)]


def range_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous query ranges per rank (``m2s_generate_sdf`` uses the same split)."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def all_gather_slabs(local, nx: int, plane: int, rank: int, world: int, group=None):
    """Reassembles the flat grid from per-rank slabs. ``local`` holds ``(x1 - x0) * plane`` float32 values of
    this rank's slab. Slabs are padded to the largest plane count so one ``all_gather_into_tensor`` suffices
    (NCCL needs equal contributions); returns a tensor of ``nx * plane`` values on every rank."""
    import torch
    import torch.distributed as dist

    bounds = slab_bounds(nx, world)
    max_planes = max(b - a for a, b in bounds)
    x0, x1 = bounds[rank]
    assert local.numel() == (x1 - x0) * plane, "local slab has the wrong size"
    if world == 1:
        return local
    send = local
    if x1 - x0 != max_planes:
        send = torch.empty(max_planes * plane, dtype=local.dtype, device=local.device)
        send[: local.numel()] = local
        send[local.numel():] = 0
    recv = torch.empty(world * max_planes * plane, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(b - a == max_planes for a, b in bounds):
        return recv
    parts = [recv[r * max_planes * plane: r * max_planes * plane + (b - a) * plane] for r, (a, b) in enumerate(bounds)]
    return torch.cat(parts)


def all_gather_ranges(local, n: int, rank: int, world: int, group=None):
    """Same for per-query results sharded by ``range_bounds``."""
    return all_gather_slabs(local, n, 1, rank, world, group)
