"""GPU parity of generate_grid_sdf (CUDA path through the C ABI) against the oracle.

Bars: Raycast |d| and sign are bit-exact against the exact oracle (same un-fused fp32 leaf arithmetic, exact
min); Normal is within 1e-4 * mesh diagonal (north-star tolerance; in practice <= 2e-6: the near-tie fold of
lib.rs:242-259 is order dependent) with identical signs."""
import numpy as np
import pytest

from mesh_to_sdf_b200 import synth
from conftest import mesh_diag

pytestmark = pytest.mark.gpu

RAYCAST, NORMAL = 0, 1


def _grid_for(m2s, verts, n):
    mn, mx = synth.padded_grid_box(verts)
    return m2s.Grid.from_bounding_box(mn, mx, n)


def test_doc_kat_grid_raycast(m2s):
    # lib.rs:34-58: sdf[0] == 1.0
    verts = np.array([[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]], np.float32)
    grid = m2s.Grid.from_bounding_box([0., 0., 0.], [10., 10., 10.], [10, 10, 10])
    sdf = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(np.array([0, 1, 2], np.uint32)), grid,
                                m2s.SignMethod.Raycast)
    assert sdf.shape == (1000,)
    assert sdf[0] == 1.0


def test_doc_kat_grid_rs(m2s):
    # generate/grid.rs:205-231
    verts = np.array([[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 4.]], np.float32)
    grid = m2s.Grid.from_bounding_box([0., 0., 0.], [10., 10., 10.], [10, 10, 10])
    sdf = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(np.array([0, 1, 2], np.uint32)), grid,
                                m2s.SignMethod.Raycast)
    assert sdf[0] == 1.0


def test_c1_single_triangle_normal_bit_exact(m2s, oracle):
    # BASELINE config C1: single triangle, 8^3, Normal
    verts = np.array([[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]], np.float32)
    tris = np.array([[0, 1, 2]], np.uint32)
    grid = m2s.Grid.from_bounding_box([0., 0., 0.], [10., 10., 10.], [8, 8, 8])
    got = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(tris), grid, m2s.SignMethod.Normal)
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, NORMAL)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_generate_grid_equals_generic_none_raycast(m2s, oracle):
    # generate/grid.rs:692-724 (assert_eq!, exact)
    verts = np.array([[0., 1., 0.], [1., 2., 3.], [1., 3., 4.], [2., 0., 0.]], np.float32)
    idx = np.array([0, 1, 2, 1, 2, 3], np.uint32)
    grid = m2s.Grid.from_bounding_box([0., 0., 0.], [5., 5., 5.], [5, 5, 5])
    q = np.array([grid.get_cell_center([x, y, z]) for x in range(5) for y in range(5) for z in range(5)], np.float32)
    sdf = m2s.generate_sdf(verts, m2s.Topology.TriangleList(idx), q, m2s.AccelerationMethod.none(m2s.SignMethod.Raycast))
    grid_sdf = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(idx), grid, m2s.SignMethod.Raycast)
    assert np.array_equal(sdf, grid_sdf)
    want = oracle.generate_sdf(verts, idx.reshape(-1, 3), q, oracle.ACCEL_NONE, oracle.RAYCAST)
    assert np.array_equal(sdf.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("nu,nv,n", [(16, 10, 24), (48, 30, 40)])
def test_torus_raycast_bit_exact(m2s, oracle, nu, nv, n):
    verts, tris = synth.bumpy_torus(nu, nv)
    grid = _grid_for(m2s, verts, [n, n + 3, n - 5])
    got = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(tris), grid, m2s.SignMethod.Raycast)
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, RAYCAST)
    assert np.array_equal(np.abs(got).view(np.uint32), np.abs(want).view(np.uint32))
    assert np.array_equal(np.signbit(got), np.signbit(want))
    assert (got < 0).any() and (got > 0).any()


@pytest.mark.parametrize("nu,nv,n", [(16, 10, 24), (48, 30, 40)])
def test_torus_normal_within_tolerance(m2s, oracle, nu, nv, n):
    verts, tris = synth.bumpy_torus(nu, nv)
    grid = _grid_for(m2s, verts, [n, n - 2, n + 1])
    got = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(tris), grid, m2s.SignMethod.Normal)
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, NORMAL)
    tol = 1e-4 * mesh_diag(verts)  # north-star tolerance
    assert np.max(np.abs(got - want)) <= tol
    assert np.max(np.abs(np.abs(got) - np.abs(want))) <= 4e-6
    assert np.array_equal(np.signbit(got), np.signbit(want))


def test_empty_mesh_fills_max(m2s):
    grid = m2s.Grid([0., 0., 0.], [1., 1., 1.], [3, 4, 5])
    out = m2s.generate_grid_sdf(np.zeros((0, 3), np.float32), m2s.Topology.TriangleList(np.zeros(0, np.uint32)), grid,
                                m2s.SignMethod.Raycast)
    assert out.shape == (60,) and np.all(out == np.finfo(np.float32).max)


def test_slab_invariance_device_entry(m2s, oracle):
    torch = pytest.importorskip("torch")
    verts, tris = synth.bumpy_torus(32, 20)
    grid = _grid_for(m2s, verts, [20, 17, 19])
    whole = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(tris), grid, m2s.SignMethod.Raycast)
    dv = torch.from_numpy(verts).cuda()
    dt = torch.from_numpy(tris.astype(np.int64)).to(torch.int32).cuda()  # same bits as u32
    plane = 17 * 19
    with m2s.Context() as c:
        parts = []
        for x0, x1 in [(0, 7), (7, 8), (8, 20)]:
            out = torch.empty((x1 - x0) * plane, dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            c.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, RAYCAST, x0, x1, out.data_ptr())
            c.synchronize()
            parts.append(out.cpu().numpy())
    assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))


def test_device_entry_defers_data_errors(m2s):
    # device-buffer calls only enqueue; a bad index surfaces at the next m2s_synchronize (M2S_EINDEX)
    torch = pytest.importorskip("torch")
    verts, tris = synth.bumpy_torus(8, 6)
    bad = tris.copy()
    bad[5, 1] = 10_000
    grid = _grid_for(m2s, verts, [8, 8, 8])
    dv = torch.from_numpy(verts).cuda()
    out = torch.empty(512, dtype=torch.float32, device="cuda")
    with m2s.Context() as c:
        dt = torch.from_numpy(bad.view(np.int32)).cuda()
        torch.cuda.synchronize()
        c.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(bad), grid, RAYCAST, 0, 8, out.data_ptr())
        with pytest.raises(m2s.M2SError) as e:
            c.synchronize()
        assert e.value.status == m2s.M2S_EINDEX
        # the flags were cleared: a good call afterwards synchronises cleanly and is correct
        dt = torch.from_numpy(tris.view(np.int32)).cuda()
        torch.cuda.synchronize()
        c.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, RAYCAST, 0, 8, out.data_ptr())
        c.synchronize()
        want = c.grid_sdf(verts, tris, grid, RAYCAST)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
        t = c.timings()
        assert t["total_ms"] > 0 and c.launch_count > 0


@pytest.mark.parametrize("dims", [(37, 21, 30), (9, 50, 5), (64, 64, 64), (5, 9, 131)])
def test_ragged_grids_and_slabs_bit_exact(m2s, oracle, dims):
    # the distance kernel works on 2 x 4 x 8 voxel tiles (runs of 2 voxels per lane): grids whose sides are not
    # multiples of the tile (runs cut by the grid's end, partial bricks) and slabs that start inside a brick must
    # give the exact oracle's bits
    verts, tris = synth.bumpy_torus(40, 24)
    grid = _grid_for(m2s, verts, list(dims))
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, RAYCAST)
    want_n = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, NORMAL)
    with m2s.Context() as c:
        got = c.grid_sdf(verts, tris, grid, RAYCAST)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # Normal: magnitudes inside the near-tie window of compare_distances, identical signs
        got_n = c.grid_sdf(verts, tris, grid, NORMAL)
        assert np.max(np.abs(np.abs(got_n) - np.abs(want_n))) <= 4e-6
        assert np.array_equal(np.signbit(got_n), np.signbit(want_n))
        torch = pytest.importorskip("torch")
        dv = torch.from_numpy(verts).cuda()
        dt = torch.from_numpy(tris.view(np.int32)).cuda()
        plane = dims[1] * dims[2]
        x0, x1 = dims[0] // 3, dims[0] - 2
        out = torch.empty((x1 - x0) * plane, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        c.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, RAYCAST, x0, x1, out.data_ptr())
        c.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want[x0 * plane:x1 * plane].view(np.uint32))


@pytest.mark.parametrize("run_length", [2, 4, 18, 20])
@pytest.mark.parametrize("dims", [(37, 21, 30), (5, 9, 131), (48, 40, 33)])
def test_every_run_length_and_lane_layout_bit_exact(m2s, oracle, run_length, dims):
    # the grid kernel exists with runs of 2 and 4 voxels per lane and two lane layouts (M2S_OPT_RUN_LENGTH: 2, 4,
    # +16 = the 4 x 4 x 2V layout); the library picks one by grid / mesh shape. Every variant must give the exact
    # oracle's bits on ragged grids and slabs, with both sign methods, through the pipelined host path too
    verts, tris = synth.bumpy_torus(40, 24)
    grid = _grid_for(m2s, verts, list(dims))
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, RAYCAST)
    want_n = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, NORMAL)
    with m2s.Context() as c:
        c.set_option(m2s.OPT_RUN_LENGTH, run_length)
        got = c.grid_sdf(verts, tris, grid, RAYCAST)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        got_n = c.grid_sdf(verts, tris, grid, NORMAL)
        assert np.max(np.abs(np.abs(got_n) - np.abs(want_n))) <= 4e-6
        assert np.array_equal(np.signbit(got_n), np.signbit(want_n))
        c.set_option(m2s.OPT_HOST_PATH, m2s.HOST_PIPELINED)
        assert np.array_equal(c.grid_sdf(verts, tris, grid, RAYCAST).view(np.uint32), want.view(np.uint32))
        plane = dims[1] * dims[2]
        x0, x1 = dims[0] // 3, dims[0] - 2
        part = c.grid_sdf_slab(verts, tris, grid, RAYCAST, x0, x1)
        assert np.array_equal(part.view(np.uint32), want[x0 * plane:x1 * plane].view(np.uint32))


def test_normal_sign_near_ties_positive_wins(m2s, oracle):
    # lib.rs:242-254: approximately equal |d| (2 ulps / 1e-6) -> the positive one wins. Queries in the mid-plane of
    # a thin slab see two faces at (almost) the same distance from opposite sides; queries straight above a shared
    # edge of a convex / concave fold see two triangles at exactly the same distance.
    zb, zd = 0.0625, np.float32(0.0625) + np.float32(5e-7)
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0],          # face A, z = 0, normal +z
                      [0, 0, zb], [1, 0, zb], [0, 1, zb], [1, 1, zb],      # face B, z = 1/16, normal +z: the mid-plane
                                                                           # z = 1/32 is a cell centre -> exact tie, + and -
                      [2, 0, 0], [3, 0, 0], [2, 1, 0], [3, 1, 0],          # face C, z = 0
                      [2, 0, zd], [3, 0, zd], [2, 1, zd], [3, 1, zd],      # face D, 5e-7 higher: near-tie inside 1e-6
                      [4, 0, 0], [5, 0, 1], [5, 1, 1], [4, 1, 0], [6, 0, 0], [6, 1, 0]], np.float32)  # a roof fold
    tris = np.array([[0, 1, 2], [1, 3, 2], [4, 5, 6], [5, 7, 6], [8, 9, 10], [9, 11, 10], [12, 13, 14], [13, 15, 14],
                     [16, 17, 18], [16, 18, 19], [17, 20, 21], [17, 21, 18]], np.uint32)
    grid = m2s.Grid([-0.25, -0.25, -0.5], [0.125, 0.125, 0.03125], [56, 14, 64])
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, NORMAL)
    with m2s.Context() as c:
        got = c.grid_sdf(verts, tris, grid, NORMAL)
    assert np.max(np.abs(np.abs(got) - np.abs(want))) <= 4e-6
    assert np.array_equal(np.signbit(got), np.signbit(want))
    mid = got.reshape(56, 14, 64)[:, :, 17]  # z = 1/32: between A and B (exact tie) and between C and D (near-tie)
    assert np.all(mid[3:9, 3:9] == np.float32(0.03125)) and np.all(mid[19:25, 3:9] > 0)
