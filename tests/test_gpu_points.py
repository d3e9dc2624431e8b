"""GPU parity of generate_sdf (all four AccelerationMethods) against the exact oracle, through the C ABI.

Raycast-signed methods (None(Raycast), Bvh(Raycast), RtreeBvh) are bit-exact: |d| is the exact fp32 minimum
and the ray predicates are evaluated un-fused like the reference. Normal-signed methods are within the
north-star tolerance 1e-4 * mesh diagonal (observed <= 2e-6) with identical signs. Rtree returns the signed
distance of the single nearest triangle; ties between equidistant triangles are unspecified in the reference
(rstar returns "a" nearest neighbour, rtree.rs:116), so |d| is compared bit-exactly and signs off ties."""
import json
import os

import numpy as np
import pytest

from mesh_to_sdf_b200 import synth
from conftest import mesh_diag

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def load_mesh(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return z["vertices"].astype(np.float32), z["indices"].astype(np.uint32).reshape(-1, 3)


def methods(m2s):
    A, S = m2s.AccelerationMethod, m2s.SignMethod
    return [("None(Raycast)", A.none(S.Raycast), 0, 0), ("None(Normal)", A.none(S.Normal), 0, 1),
            ("Bvh(Raycast)", A.bvh(S.Raycast), 1, 0), ("Bvh(Normal)", A.bvh(S.Normal), 1, 1),
            ("Rtree", A.Rtree, 2, 0), ("RtreeBvh", A.RtreeBvh, 3, 0)]


@pytest.mark.parametrize("key", ["doc_generate_sdf_rtree_bvh", "doc_generate_sdf_fn"])
def test_doc_kats(m2s, key):
    k = KAT[key]  # lib.rs:13-31, lib.rs:269-289: exactly [1.0]
    for name, method, _, _ in methods(m2s):
        sdf = m2s.generate_sdf(np.array(k["vertices"], np.float32), m2s.Topology.TriangleList(np.array(k["indices"], np.uint32)),
                               np.array(k["query_points"], np.float32), method)
        assert sdf.tolist() == k["expect"], name
    # default method is RtreeBvh
    sdf = m2s.generate_sdf(np.array(k["vertices"], np.float32), m2s.Topology.TriangleList(np.array(k["indices"], np.uint32)),
                           np.array(k["query_points"], np.float32))
    assert sdf.tolist() == k["expect"]


def check_against_oracle(m2s, oracle, verts, tris, q, tol):
    for name, method, accel, sign in methods(m2s):
        got = m2s.generate_sdf(verts, m2s.Topology.TriangleList(tris), q, method)
        want = oracle.generate_sdf(verts, tris, q, accel, sign)
        assert got.shape == want.shape
        if name in ("None(Raycast)", "Bvh(Raycast)", "RtreeBvh"):
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        elif name == "Rtree":
            assert np.array_equal(np.abs(got).view(np.uint32), np.abs(want).view(np.uint32)), name
            # sign may differ only where two triangles are exactly equidistant (tie -> unspecified element)
            assert np.mean(np.signbit(got) != np.signbit(want)) < 0.02, name
        else:
            assert np.max(np.abs(got - want)) <= tol, name
            assert np.max(np.abs(np.abs(got) - np.abs(want))) <= 4e-6, name
            assert np.array_equal(np.signbit(got), np.signbit(want)), name


def test_torus_all_methods(m2s, oracle):
    verts, tris = synth.bumpy_torus(40, 24)
    mn, mx = synth.padded_grid_box(verts)
    q = synth.splitmix64_points(6000, mn, mx)
    check_against_oracle(m2s, oracle, verts, tris, q, 1e-4 * mesh_diag(verts))


def test_suzanne_all_methods(m2s, oracle):
    # the reference's own fixture (968 triangles, not watertight, has degenerate-ish parts)
    verts, tris = load_mesh("suzanne")
    mn, mx = verts.min(axis=0) - 0.3, verts.max(axis=0) + 0.3
    q = synth.splitmix64_points(5000, mn, mx, seed=7)
    check_against_oracle(m2s, oracle, verts, tris, q, 1e-4 * mesh_diag(verts))
    # default.rs:83-109: python baseline within 0.1
    k = KAT["suzanne_python_baseline"]
    sdf = m2s.generate_sdf(verts, m2s.Topology.TriangleList(tris), np.array(k["query_points"], np.float32),
                           m2s.AccelerationMethod.none(m2s.SignMethod.Normal))
    for got, base in zip(sdf, k["baseline"]):
        assert abs(got - base) < k["tolerance"]


def test_bvh_five_points(m2s, oracle):
    # generic/bvh.rs:154-189
    verts, tris = load_mesh("suzanne")
    q = np.array([[0.01, 0.01, 0.5], [1., 1., 1.], [0.1, 0.2, 0.2], [1.1, 2.2, 5.2], [-0.1, 0.2, -0.2]], np.float32)
    A, S = m2s.AccelerationMethod, m2s.SignMethod
    t = m2s.Topology.TriangleList(tris)
    bvh = m2s.generate_sdf(verts, t, q, A.bvh(S.Raycast))
    none = m2s.generate_sdf(verts, t, q, A.none(S.Raycast))
    rtree = m2s.generate_sdf(verts, t, q, A.Rtree)
    rtree_bvh = m2s.generate_sdf(verts, t, q, A.RtreeBvh)
    assert np.all(np.abs(bvh - none) < 0.01)
    assert np.all(np.abs(np.abs(rtree) - np.abs(bvh)) < 0.01)
    assert np.all(np.abs(rtree_bvh - bvh) < 0.01)
    assert np.array_equal(rtree_bvh, oracle.generate_sdf(verts, tris, q, 3))


def test_degenerate_triangles(m2s, oracle):
    # geo.rs:73-88 guards: zero-area triangles (two or three equal vertices) mixed into a mesh
    verts, tris = synth.bumpy_torus(12, 8)
    extra = np.array([[0, 0, 5], [3, 7, 7], [9, 9, 9], [2, 4, 2]], np.uint32)
    tris = np.concatenate([tris[:40], extra, tris[40:]])
    mn, mx = synth.padded_grid_box(verts)
    q = synth.splitmix64_points(3000, mn, mx, seed=11)
    for accel, sign in [(0, 0), (3, 0), (2, 0)]:
        got = m2s.default_context().sdf(verts, tris, q, accel, sign)
        want = oracle.generate_sdf(verts, tris, q, accel, sign)
        assert np.array_equal(np.abs(got).view(np.uint32), np.abs(want).view(np.uint32)), (accel, sign)
    grid = m2s.Grid.from_bounding_box(mn, mx, [12, 11, 10])
    got = m2s.generate_grid_sdf(verts, m2s.Topology.TriangleList(tris), grid, m2s.SignMethod.Raycast)
    want = oracle.grid_cells_exact(verts, tris, grid.first_cell, grid.cell_size, grid.cell_count, 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_topologies(m2s, oracle):
    # generate/grid.rs:846-904: list / strip, indexed / not, u16 / u32
    verts, tris = load_mesh("annoted_cube")
    flat = verts[tris.ravel()]
    mn, mx = verts.min(axis=0) - 0.1, verts.max(axis=0) + 0.1
    grid = m2s.Grid.from_bounding_box(mn, mx, [25, 25, 25])
    T, S = m2s.Topology, m2s.SignMethod
    a = m2s.generate_grid_sdf(verts, T.TriangleList(tris), grid, S.Normal)
    b = m2s.generate_grid_sdf(verts, T.TriangleList(tris.astype(np.uint16)), grid, S.Normal)
    c = m2s.generate_grid_sdf(flat, T.TriangleList(None), grid, S.Normal)
    assert np.array_equal(a, b)
    assert np.allclose(a, c, atol=1e-6)
    # a strip: 0 1 2 3 -> (0,1,2), (1,2,3)
    quad = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.2]], np.float32)
    s1 = m2s.generate_grid_sdf(quad, T.TriangleStrip(None), grid, S.Normal)
    s2 = m2s.generate_grid_sdf(quad, T.TriangleStrip(np.array([0, 1, 2, 3], np.uint16)), grid, S.Normal)
    s3 = m2s.generate_grid_sdf(quad, T.TriangleList(np.array([0, 1, 2, 1, 2, 3], np.uint32)), grid, S.Normal)
    assert np.array_equal(s1, s2) and np.array_equal(s1, s3)
    want = oracle.grid_cells_exact(quad, np.array([[0, 1, 2], [1, 2, 3]], np.uint32), grid.first_cell, grid.cell_size,
                                   grid.cell_count, 1)
    assert np.max(np.abs(s1 - want)) <= 4e-6 and np.array_equal(np.signbit(s1), np.signbit(want))


def test_error_behaviour(m2s):
    # the reference panics; the ABI returns a status and the binding raises
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    q = np.zeros((2, 3), np.float32)
    T, A = m2s.Topology, m2s.AccelerationMethod
    with pytest.raises(m2s.M2SError) as e:  # slice index out of bounds
        m2s.generate_sdf(verts, T.TriangleList(np.array([0, 1, 7], np.uint32)), q, A.RtreeBvh)
    assert e.value.status == m2s.M2S_EINDEX
    with pytest.raises(m2s.M2SError) as e:  # rtree.rs:117 unwrap on an empty tree
        m2s.generate_sdf(np.zeros((0, 3), np.float32), T.TriangleList(np.zeros(0, np.uint32)), q, A.Rtree)
    assert e.value.status == m2s.M2S_EEMPTY
    # rtree_bvh.rs:104-106: empty mesh -> empty Vec
    assert m2s.generate_sdf(np.zeros((0, 3), np.float32), T.TriangleList(np.zeros(0, np.uint32)), q, A.RtreeBvh).shape == (0,)
    # None / Bvh on an empty mesh: f32::MAX per query
    out = m2s.generate_sdf(np.zeros((0, 3), np.float32), T.TriangleList(np.zeros(0, np.uint32)), q, A.none())
    assert np.all(out == np.finfo(np.float32).max)
    bad = verts.copy()
    bad[1, 1] = np.nan
    with pytest.raises(m2s.M2SError) as e:  # lib.rs:257 "NaN distance"
        m2s.generate_sdf(bad, T.TriangleList(None), q, A.none(m2s.SignMethod.Normal))
    assert e.value.status == m2s.M2S_ENAN
    # the context stays usable after an error
    ok = m2s.generate_sdf(verts, T.TriangleList(None), q, A.RtreeBvh)
    assert np.all(np.isfinite(ok))
    # no queries -> empty result
    assert m2s.generate_sdf(verts, T.TriangleList(None), np.zeros((0, 3), np.float32), A.RtreeBvh).shape == (0,)


def test_query_order_preserved(m2s, oracle):
    # output order = query order (par_iter().map().collect()); the GPU sorts queries internally
    verts, tris = synth.bumpy_torus(16, 10)
    mn, mx = synth.padded_grid_box(verts)
    q = synth.splitmix64_points(5000, mn, mx, seed=3)
    perm = np.random.default_rng(0).permutation(len(q))
    a = m2s.default_context().sdf(verts, tris, q, 3, 0)
    b = m2s.default_context().sdf(verts, tris, q[perm], 3, 0)
    assert np.array_equal(a[perm], b)


def _raycast_methods():
    return [(0, 0), (1, 0), (3, 0)]  # None(Raycast): +X parity; Bvh(Raycast), RtreeBvh: best of three axes


def test_ray_bins_equal_the_box_tree_walk_and_the_oracle(m2s, oracle):
    # the axis-ray parities come from per-axis 2-D triangle bins by default and from the packet walk of the box tree
    # with M2S_OPT_RAY_BINS = 0: same candidate filter (padded boxes, geo.rs:4-22), same predicates -> same bits
    rng = np.random.default_rng(11)
    cases = []
    verts, tris = synth.bumpy_torus(96, 64)
    mn, mx = synth.padded_grid_box(verts)
    cases.append(("torus", verts, tris, synth.splitmix64_points(20000, mn, mx)))
    # a few big triangles (a box around everything: every one covers far more than 64 cells) + the small ones
    bmn, bmx = mn - 0.1, mx + 0.1
    c = np.array([[bmn[0], bmn[1], bmn[2]], [bmx[0], bmn[1], bmn[2]], [bmx[0], bmx[1], bmn[2]], [bmn[0], bmx[1], bmn[2]],
                  [bmn[0], bmn[1], bmx[2]], [bmx[0], bmn[1], bmx[2]], [bmx[0], bmx[1], bmx[2]], [bmn[0], bmx[1], bmx[2]]],
                 np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5],
                  [0, 4, 7], [0, 7, 3]], np.uint32)
    v2 = np.concatenate([verts, c]).astype(np.float32)
    t2 = np.concatenate([tris, f + len(verts)]).astype(np.uint32)
    q2 = synth.splitmix64_points(8000, bmn - 0.2, bmx + 0.2)
    cases.append(("torus in a box (12 big triangles)", v2, t2, q2))
    # more big overlapping triangles than the per-axis budget: the library falls back to the tree walk by itself
    fan_v = (rng.uniform(-1.0, 1.0, (900, 3)) * np.array([3.0, 3.0, 1.0])).astype(np.float32)
    fan_t = np.arange(900, dtype=np.uint32).reshape(-1, 3)
    cases.append(("300 big random triangles (over budget)", fan_v, fan_t, synth.splitmix64_points(3000, [-3, -3, -1], [3, 3, 1])))
    with m2s.Context() as c:
        for name, v, t, q in cases:
            for accel, sign in _raycast_methods():
                c.set_option(m2s.OPT_RAY_BINS, 1)
                a = c.sdf(v, t, q, accel, sign)
                c.set_option(m2s.OPT_RAY_BINS, 0)
                b = c.sdf(v, t, q, accel, sign)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, accel)
                want = oracle.generate_sdf(v, t, q, accel, sign)
                assert np.array_equal(a.view(np.uint32), want.view(np.uint32)), (name, accel)
        # a mesh handle builds its bins once and keeps them
        c.set_option(m2s.OPT_RAY_BINS, 1)
        name, v, t, q = cases[1]
        with c.mesh(v, t) as mesh:
            first = mesh.sdf(q, 3)
            again = mesh.sdf(q[::-1].copy(), 3)
            assert np.array_equal(first.view(np.uint32), again[::-1].view(np.uint32))
            assert np.array_equal(first.view(np.uint32), oracle.generate_sdf(v, t, q, 3, 0).view(np.uint32))
