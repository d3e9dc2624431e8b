"""Pins the oracle's Grid / Topology restatement and the host-side mirror (mesh_to_sdf_b200.Grid, Topology)
against the reference's unit tests (mesh_to_sdf/src/grid.rs:176-297) and lib.rs:175-193."""
import numpy as np
import pytest


def both_grids(oracle, m2s, mn, mx, count):
    first, size = oracle.grid_from_bounding_box(mn, mx, count)
    g = m2s.Grid.from_bounding_box(mn, mx, count)
    assert np.array_equal(first, g.first_cell) and np.array_equal(size, g.cell_size)
    return first, size, g


def test_new(m2s):
    # grid.rs:180-189
    g = m2s.Grid.new([0.1, 0.2, 0.3], [1.1, 1.2, 1.3], [11, 12, 13])
    assert g.get_first_cell().tolist() == np.float32([0.1, 0.2, 0.3]).tolist()
    assert g.get_cell_size().tolist() == np.float32([1.1, 1.2, 1.3]).tolist()
    assert g.get_cell_count() == (11, 12, 13)
    assert g.get_total_cell_count() == 11 * 12 * 13


def test_first_last_cells(oracle, m2s):
    # grid.rs:191-199
    g = m2s.Grid([0., 1., 2.], [1., 2., 3.], [10, 20, 30])
    assert g.get_last_cell().tolist() == [10., 41., 92.]
    assert oracle.grid_last_cell([0., 1., 2.], [1., 2., 3.], [10, 20, 30]).tolist() == [10., 41., 92.]


def test_from_bounding_box(oracle, m2s):
    # grid.rs:201-213
    first, size, g = both_grids(oracle, m2s, [-1., 0., 1.], [0., 2., 5.], [2, 2, 2])
    assert first.tolist() == [-0.75, 0.5, 2.]
    assert size.tolist() == [0.5, 1., 2.]
    mn, mx = g.get_bounding_box()
    assert mn.tolist() == [-1., 0., 1.] and mx.tolist() == [0., 2., 5.]
    omn, omx = oracle.grid_bounding_box(first, size, [2, 2, 2])
    assert omn.tolist() == [-1., 0., 1.] and omx.tolist() == [0., 2., 5.]


def test_snap_point_to_grid(oracle, m2s):
    # grid.rs:215-241
    first, size, g = both_grids(oracle, m2s, [0., 0., 0.], [1., 1., 1.], [2, 2, 2])
    cases = [([0.4, 0.8, 0.1], True, [0, 1, 0]), ([-0.5, 0.8, 0.8], False, [0, 1, 1]),
             ([0.8, 0.8, 0.8], True, [1, 1, 1]), ([0.8, 1.5, 0.8], False, [1, 1, 1])]
    for p, inside, cell in cases:
        assert oracle.grid_snap(first, size, [2, 2, 2], p) == (inside, cell)
        r = g.snap_point_to_grid(p)
        assert (r.inside, list(r.cell)) == (inside, cell)


def test_get_cell_idx(oracle, m2s):
    # grid.rs:243-256
    _, _, g = both_grids(oracle, m2s, [0., 0., 0.], [1., 1., 1.], [2, 3, 4])
    table = {(0, 0, 0): 0, (0, 0, 1): 1, (0, 1, 0): 4, (0, 1, 1): 5, (1, 0, 0): 12, (1, 0, 1): 13, (1, 1, 0): 16,
             (1, 1, 1): 17}
    for cell, idx in table.items():
        assert oracle.grid_cell_idx([2, 3, 4], cell) == idx
        assert g.get_cell_idx(cell) == idx


def test_get_cell_integer_coordinates(oracle, m2s):
    # grid.rs:258-281
    g = m2s.Grid.from_bounding_box([0., 0., 0.], [1., 1., 1.], [5, 10, 15])
    for i in range(750):
        c = g.get_cell_integer_coordinates(i)
        assert g.get_cell_idx(c) == i
        assert oracle.grid_cell_coords([5, 10, 15], i) == c
    for x in range(5):
        for y in range(10):
            for z in range(15):
                assert g.get_cell_integer_coordinates(g.get_cell_idx([x, y, z])) == [x, y, z]


def test_get_cell_center(oracle, m2s):
    # grid.rs:283-297
    first, size, g = both_grids(oracle, m2s, [0., 0., 0.], [1., 1., 1.], [2, 2, 2])
    for x in range(2):
        for y in range(2):
            for z in range(2):
                want = [0.25 + 0.5 * x, 0.25 + 0.5 * y, 0.25 + 0.5 * z]
                assert g.get_cell_center([x, y, z]).tolist() == want
                assert oracle.grid_cell_center(first, size, [2, 2, 2], [x, y, z]).tolist() == want


def test_topology_expansion(oracle, m2s):
    # lib.rs:175-193: list -> tuples() (partial tail dropped); strip -> tuple_windows() (no winding flip);
    # None -> 0..vertices.len()
    idx = np.array([0, 1, 2, 1, 2, 3, 7], np.uint32)
    want_list = [[0, 1, 2], [1, 2, 3]]
    want_strip = [[0, 1, 2], [1, 2, 1], [2, 1, 2], [1, 2, 3], [2, 3, 7]]
    assert oracle.expand_topology(0, idx, 8).tolist() == want_list
    assert oracle.expand_topology(1, idx, 8).tolist() == want_strip
    assert m2s.Topology.TriangleList(idx).get_triangles(8).tolist() == want_list
    assert m2s.Topology.TriangleStrip(idx).get_triangles(8).tolist() == want_strip
    assert m2s.Topology.TriangleList(idx.astype(np.uint16)).get_triangles(8).tolist() == want_list
    assert m2s.Topology.TriangleStrip(idx.astype(np.uint16)).get_triangles(8).tolist() == want_strip
    assert oracle.expand_topology(0, None, 7).tolist() == [[0, 1, 2], [3, 4, 5]]
    assert m2s.Topology.TriangleList(None).get_triangles(7).tolist() == [[0, 1, 2], [3, 4, 5]]
    assert m2s.Topology.TriangleStrip(None).get_triangles(4).tolist() == [[0, 1, 2], [1, 2, 3]]
    assert oracle.expand_topology(1, None, 2).shape == (0, 3)
    assert m2s.Topology.TriangleStrip(None).get_triangles(2).shape == (0, 3)
