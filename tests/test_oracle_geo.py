"""Pins the oracle's leaf arithmetic (oracle/m2s_oracle_geo.hpp) against the reference's own unit tests,
property tests and saved proptest regressions (mesh_to_sdf/src/geo.rs:218-323, proptest-regressions/geo.txt).

The two property tests are re-run with hypothesis against independent float64 baselines written here (the
reference uses an SDFGen-style distance, geo.rs:329-379, and a generic Moller-Trumbore ray, geo.rs:396-454)."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def approx_eq(a, b, ulps=5, eps=1e-3):
    """float-cmp approx_eq!(f32, a, b, ulps, epsilon)."""
    a, b = np.float32(a), np.float32(b)
    if a == b or abs(float(a) - float(b)) <= eps:
        return True
    return abs(int(a.view(np.int32)) - int(b.view(np.int32))) <= ulps


def baseline_distance_f64(p, a, b, c):
    """Independent point-triangle distance in float64: project on the plane, else min of the three edges."""
    p, a, b, c = (np.asarray(v, np.float64) for v in (p, a, b, c))

    def seg(p, u, v):
        d = v - u
        t = np.clip(np.dot(p - u, d) / np.dot(d, d), 0.0, 1.0)
        return np.linalg.norm(p - (u + t * d))

    n = np.cross(b - a, c - a)
    nn = np.dot(n, n)
    if nn > 0:
        q = p - n * (np.dot(p - a, n) / nn)
        # barycentric inside test
        w0 = np.dot(np.cross(b - q, c - q), n)
        w1 = np.dot(np.cross(c - q, a - q), n)
        w2 = np.dot(np.cross(a - q, b - q), n)
        if w0 >= 0 and w1 >= 0 and w2 >= 0:
            return abs(np.dot(p - a, n)) / np.sqrt(nn)
    return min(seg(p, a, b), seg(p, b, c), seg(p, c, a))


def moller_trumbore_f64(o, d, a, b, c):
    o, d, a, b, c = (np.asarray(v, np.float64) for v in (o, d, a, b, c))
    e1, e2 = b - a, c - a
    h = np.cross(d, e2)
    det = np.dot(e1, h)
    if det == 0:
        return None, 0.0
    s = o - a
    u = np.dot(s, h) / det
    q = np.cross(s, e1)
    v = np.dot(d, q) / det
    t = np.dot(e2, q) / det
    margin = min(u, v, 1 - u - v)  # > 0 strictly inside
    if u < 0 or v < 0 or u + v > 1 or t <= 0:
        return None, margin if t > 0 else -abs(t)
    return t, margin


def test_segment_kats(oracle):
    # geo.rs:311-323 (assert_eq!, exact)
    for k in KAT["segment"]:
        got = oracle.closest_point_segment(k["p"], k["a"], k["b"])
        assert got.tolist() == np.asarray(k["expect"], np.float32).tolist(), k["src"]


def test_proptest_regressions(oracle):
    # proptest-regressions/geo.txt:7-8, property of geo.rs:225-256
    for k in KAT["proptest_regressions"]:
        d = oracle.point_triangle_distance(k["p"], k["a"], k["b"], k["c"])
        base = baseline_distance_f64(k["p"], k["a"], k["b"], k["c"])
        assert not np.isnan(d)
        assert approx_eq(d, base), (d, base)
        # the same inputs are regression seeds of the ray property (geo.rs:258-287) too
        for axis, direction in enumerate(np.eye(3)):
            t = oracle.ray_triangle_intersection_aligned(k["p"], k["a"], k["b"], k["c"], axis)
            tb, margin = moller_trumbore_f64(k["p"], direction, k["a"], k["b"], k["c"])
            if abs(margin) > 1e-4:
                assert (t is None) == (tb is None)
                if t is not None:
                    assert approx_eq(t, tb)


coords = st.floats(min_value=-10.0, max_value=10.0, allow_nan=False, width=32)
vec3 = st.tuples(coords, coords, coords)


@settings(max_examples=1000, deadline=None)
@given(p=vec3, a=vec3, b=vec3, c=vec3)
def test_closest_point_triangle_property(oracle, p, a, b, c):
    # geo.rs:225-256
    def near(u, v):
        return any(approx_eq(u[i], v[i]) for i in range(3))

    if near(a, b) or near(a, c) or near(b, c):
        return
    q = oracle.closest_point_triangle(p, a, b, c)
    d = float(np.linalg.norm(np.asarray(p, np.float32) - q))
    d_direct = oracle.point_triangle_distance(p, a, b, c)
    base = baseline_distance_f64(p, a, b, c)
    assert not np.isnan(d_direct)
    assert approx_eq(d_direct, base), (d_direct, base)
    assert approx_eq(d, d_direct, ulps=2, eps=1e-5)
    assert approx_eq(oracle.point_triangle_distance2(p, a, b, c), base * base, ulps=8, eps=2e-2)
    assert abs(oracle.point_triangle_signed_distance(p, a, b, c)) == np.float32(d_direct)


@settings(max_examples=1000, deadline=None)
@given(p=vec3, a=vec3, b=vec3, c=vec3)
def test_ray_triangle_intersection_property(oracle, p, a, b, c):
    # geo.rs:258-287: aligned vs generic ray on all three axes, hit / no-hit agreement and t
    for axis, direction in enumerate(np.eye(3)):
        t = oracle.ray_triangle_intersection_aligned(p, a, b, c, axis)
        tb, margin = moller_trumbore_f64(p, direction, a, b, c)
        if abs(margin) < 1e-4:
            continue  # knife edge: fp32 and fp64 may legitimately disagree
        assert (t is None) == (tb is None), (axis, t, tb, margin)
        if t is not None:
            assert approx_eq(t, tb), (t, tb)


def test_ray_kat_generic_cases(oracle):
    # geo.rs:289-309 uses the generic routine with four directions; the two axis-aligned ones carry over
    a, b, c = [0., 1., 0.], [1., 0., 0.], [0., 0., 1.]
    o = [0.2, 0.2, 0.2]
    assert oracle.ray_triangle_intersection_aligned(o, a, b, c, 2) is not None  # +Z hits
    # -Z is not expressible (aligned rays are +axis only); from above the plane +Z must miss
    assert oracle.ray_triangle_intersection_aligned([0.2, 0.2, 0.9], a, b, c, 2) is None


def test_triangle_bounding_box_padding(oracle):
    # geo.rs:4-22: min/max -/+ 1e-4
    mn, mx = oracle.triangle_bounding_box([0., 1., 2.], [3., -1., 5.], [1., 0., -2.])
    e = np.float32(0.0001)
    assert mn.tolist() == [np.float32(0.) - e, np.float32(-1.) - e, np.float32(-2.) - e]
    assert mx.tolist() == [np.float32(3.) + e, np.float32(1.) + e, np.float32(5.) + e]


def test_degenerate_guards(oracle):
    # geo.rs:73-88
    p = [0.3, 1.0, 0.0]
    a, b = [0., 0., 0.], [1., 0., 0.]
    assert oracle.closest_point_triangle(p, a, a, a).tolist() == [0., 0., 0.]
    assert oracle.closest_point_triangle(p, a, a, b).tolist() == [np.float32(0.3), 0., 0.]  # a==b -> seg(a,c)
    assert oracle.closest_point_triangle(p, a, b, b).tolist() == [np.float32(0.3), 0., 0.]  # b==c -> seg(a,b)
    assert oracle.closest_point_triangle(p, a, b, a).tolist() == [np.float32(0.3), 0., 0.]  # a==c -> seg(a,b)
    assert not np.isnan(oracle.point_triangle_distance(p, a, a, a))


def test_signed_distance_convention(oracle):
    # geo.rs:43-56: positive on the side of cross(b-a, c-a); dot == 0 -> negative
    a, b, c = [0., 0., 0.], [1., 0., 0.], [0., 1., 0.]  # normal +Z
    assert oracle.point_triangle_signed_distance([0.2, 0.2, 1.0], a, b, c) == 1.0
    assert oracle.point_triangle_signed_distance([0.2, 0.2, -1.0], a, b, c) == -1.0
    assert oracle.point_triangle_signed_distance([2.0, 0.0, 0.0], a, b, c) == -1.0  # in-plane: dot == 0


def test_compare_distances(oracle):
    # lib.rs:242-259 with float-cmp approx_eq!(ulps = 2, epsilon = 1e-6)
    assert oracle.compare_distances(1.0, 2.0) == -1
    assert oracle.compare_distances(-1.0, 2.0) == -1
    assert oracle.compare_distances(-3.0, 2.0) == 1
    assert oracle.compare_distances(1.0, 1.0) == 0
    assert oracle.compare_distances(-1.0, 1.0) == 1      # tie: positive wins
    assert oracle.compare_distances(1.0, -1.0) == -1
    assert oracle.compare_distances(-1.0, 1.0 + 5e-7) == 1   # within epsilon: still a tie
    assert oracle.compare_distances(1.0 + 5e-7, -1.0) == -1
    assert oracle.compare_distances(-1.0, 1.00001) == -1     # outside the window: plain |.| order
    big = np.float32(1000.0)
    up2 = np.nextafter(np.nextafter(big, np.float32(2000)), np.float32(2000))
    up3 = np.nextafter(up2, np.float32(2000))
    assert oracle.compare_distances(-big, up2) == 1          # 2 ulps apart: tie
    assert oracle.compare_distances(-big, up3) == -1         # 3 ulps (> 1e-6 apart): no tie
    with pytest.raises(oracle.OracleError):
        oracle.compare_distances(float("nan"), 1.0)
    assert oracle.approx_eq_f32(1.0, 1.0 + 1e-7, 2, 1e-6)
    assert not oracle.approx_eq_f32(1.0, 1.1, 2, 1e-6)


def test_point_array_kats():
    # point/impl_array.rs:41-63: exact f32 results of the default Point methods the oracle restates
    k = KAT["point_array"]["src"]
    v = np.array([1., 2., 3.], np.float32)
    w = np.array([4., 5., 6.], np.float32)
    length = np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=np.float32)
    assert length == np.float32(3.7416575), k
    d = w - v
    dist = np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2], dtype=np.float32)
    assert dist == np.float32(5.196152), k
