// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the leaf arithmetic of Azkellas/mesh_to_sdf (lib v0.4.0, commit edc25fb).
// It is the *checker* for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it. The product library (libm2s.so) never links,
// loads or calls anything in oracle/.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/mesh_to_sdf/). The reference is Rust; rustc never contracts a*b+c into an FMA,
// so this file MUST be compiled with -ffp-contract=off (the Makefile does) and every expression
// keeps the reference's evaluation order (Rust evaluates a*b + c*d + e*f left to right).
//
// Parity pinning: the reference cannot be compiled here (no rustc/cargo, un-vendored crates), so
// this restatement is pinned by the reference's own known-answer tests, doc-tests, proptest
// regressions and fixtures — see tests/test_oracle_*.py and DESIGN.md "Oracle".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace m2s_oracle {

struct V3 {
    float x, y, z;
};

// src/point.rs:82-141 — default Point methods (used by the [f32;3] impl, point/impl_array.rs).
static inline V3 v_add(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 v_sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float v_dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 v_cross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3 v_fmul(const V3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float v_length(const V3& a) { return std::sqrt(v_dot(a, a)); }
static inline float v_dist(const V3& a, const V3& b) { return v_length(v_sub(a, b)); }
static inline float v_dist2(const V3& a, const V3& b) {
    V3 d = v_sub(a, b);
    return v_dot(d, d);
}
static inline bool v_eq(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline float v_get(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// src/geo.rs:4-22 — AABB padded by -/+ 1e-4.
static inline void triangle_bounding_box(const V3& a, const V3& b, const V3& c, V3* mn, V3* mx) {
    const float EPSILONF = 0.0001f;
    V3 lo = {std::fmin(a.x, std::fmin(b.x, c.x)), std::fmin(a.y, std::fmin(b.y, c.y)),
             std::fmin(a.z, std::fmin(b.z, c.z))};
    V3 hi = {std::fmax(a.x, std::fmax(b.x, c.x)), std::fmax(a.y, std::fmax(b.y, c.y)),
             std::fmax(a.z, std::fmax(b.z, c.z))};
    V3 e = {EPSILONF, EPSILONF, EPSILONF};
    *mn = v_sub(lo, e);
    *mx = v_add(hi, e);
}

// src/geo.rs:141-151 — project p on [ab]. Rust's f32::clamp(0,1) propagates NaN; 0/0 here is NaN
// only when a == b, which the callers (geo.rs:73-88) route so that it cannot happen unless all
// three points coincide (handled before).
static inline V3 closest_point_segment(const V3& p, const V3& a, const V3& b) {
    V3 ab = v_sub(b, a);
    float m = v_dot(ab, ab);
    V3 ap = v_sub(p, a);
    float s12 = v_dot(ab, ap) / m;
    if (s12 < 0.0f) s12 = 0.0f;
    if (s12 > 1.0f) s12 = 1.0f;
    return v_add(a, v_fmul(ab, s12));
}

// src/geo.rs:70-138 — Embree-style closest point on triangle with degenerate guards.
static inline V3 closest_point_triangle(const V3& p, const V3& a, const V3& b, const V3& c) {
    bool ab_eq = v_eq(a, b), bc_eq = v_eq(b, c), ac_eq = v_eq(a, c);
    if (ab_eq && bc_eq && ac_eq) return a;        // geo.rs:74-76
    if (ab_eq) return closest_point_segment(p, a, c);  // geo.rs:77-79
    if (bc_eq) return closest_point_segment(p, a, b);  // geo.rs:80-82
    if (ac_eq) return closest_point_segment(p, a, b);  // geo.rs:83-85

    V3 ab = v_sub(b, a);
    V3 ac = v_sub(c, a);
    V3 ap = v_sub(p, a);

    float d1 = v_dot(ab, ap);
    float d2 = v_dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) return a;  // geo.rs:97-99

    V3 bp = v_sub(p, b);
    float d3 = v_dot(ab, bp);
    float d4 = v_dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) return b;  // geo.rs:104-106

    V3 cp = v_sub(p, c);
    float d5 = v_dot(ab, cp);
    float d6 = v_dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) return c;  // geo.rs:111-113

    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {  // geo.rs:116-119
        float v = d1 / (d1 - d3);
        return v_add(a, v_fmul(ab, v));
    }

    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {  // geo.rs:122-125
        float v = d2 / (d2 - d6);
        return v_add(a, v_fmul(ac, v));
    }

    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && d4 - d3 >= 0.0f && d5 - d6 >= 0.0f) {  // geo.rs:128-132
        float v = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        V3 bc = v_sub(c, b);
        return v_add(b, v_fmul(bc, v));
    }

    float denom = 1.0f / (va + vb + vc);  // geo.rs:134-137
    float v = vb * denom;
    float w = vc * denom;
    return v_add(v_add(a, v_fmul(ab, v)), v_fmul(ac, w));
}

// src/geo.rs:26-30
static inline float point_triangle_distance(const V3& p, const V3& a, const V3& b, const V3& c) {
    V3 n = closest_point_triangle(p, a, b, c);
    return v_dist(p, n);
}
// src/geo.rs:33-37
static inline float point_triangle_distance2(const V3& p, const V3& a, const V3& b, const V3& c) {
    V3 n = closest_point_triangle(p, a, b, c);
    return v_dist2(p, n);
}
// src/geo.rs:43-56 (+ triangle_normal :60-64). dot == 0 -> negative.
static inline float point_triangle_signed_distance(const V3& p, const V3& a, const V3& b,
                                                   const V3& c) {
    V3 nearest = closest_point_triangle(p, a, b, c);
    V3 direction = v_sub(p, nearest);
    V3 normal = v_cross(v_sub(b, a), v_sub(c, a));
    float distance = v_dist(p, nearest);
    return (v_dot(direction, normal) > 0.0f) ? distance : -distance;
}

// src/geo.rs:165-216 — axis-aligned ray/triangle test. axis: 0=X 1=Y 2=Z (GridAlign order :155-160).
// Returns true and writes *t iff the reference returns Some(t).
static inline bool ray_triangle_intersection_aligned(const V3& o, const V3& v0, const V3& v1,
                                                     const V3& v2, int axis, float* t_out) {
    V3 e01 = v_sub(v1, v0);
    V3 e12 = v_sub(v2, v1);
    V3 e20 = v_sub(v0, v2);
    V3 p0 = v_sub(o, v0);
    V3 p1 = v_sub(o, v1);
    V3 p2 = v_sub(o, v2);
    // geo.rs:181-195: X:(x; y,z)  Y:(y; z,x)  Z:(z; x,y)
    const int ix = axis, iy = (axis + 1) % 3, iz = (axis + 2) % 3;
    float w0 = v_get(p1, iz) * v_get(e12, iy) - v_get(p1, iy) * v_get(e12, iz);  // geo.rs:199
    float w1 = v_get(p2, iz) * v_get(e20, iy) - v_get(p2, iy) * v_get(e20, iz);  // geo.rs:200
    float w2 = v_get(p0, iz) * v_get(e01, iy) - v_get(p0, iy) * v_get(e01, iz);  // geo.rs:201
    if ((w0 < 0.0f && w1 < 0.0f && w2 < 0.0f) || (w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) {
        float t = -(w0 * v_get(p0, ix) + w2 * v_get(p2, ix) + w1 * v_get(p1, ix)) /
                  (w0 + w1 + w2);  // geo.rs:208
        if (t > 0.0f) {
            *t_out = t;
            return true;
        }
    }
    return false;
}

// float-cmp 0.9.0 (Cargo.lock:1386-1387; source not vendored — published algorithm restated):
// approx_eq!(f32, a, b, ulps = U, epsilon = E)  <=>  a == b || |a-b| <= E || |bits(a)-bits(b)| <= U
static inline bool approx_eq_f32(float a, float b, int32_t ulps, float epsilon) {
    if (a == b) return true;
    float eps = std::fabs(a - b);
    if (eps <= epsilon) return true;
    int32_t ai, bi;
    std::memcpy(&ai, &a, 4);
    std::memcpy(&bi, &b, 4);
    int32_t diff = (int32_t)((uint32_t)ai - (uint32_t)bi);  // wrapping_sub
    int32_t ad = diff == std::numeric_limits<int32_t>::min() ? std::numeric_limits<int32_t>::max()
                                                             : (diff < 0 ? -diff : diff);
    return ad <= ulps;
}

// src/lib.rs:242-259. Returns -1 (Less), 0 (Equal), +1 (Greater); sets *nan when the reference
// would panic ("NaN distance" :257 / unwrap :253).
static inline int compare_distances(float a, float b, bool* nan = nullptr) {
    float aa = std::fabs(a), bb = std::fabs(b);
    if (approx_eq_f32(aa, bb, 2, 1e-6f)) {
        bool an = std::signbit(a), bn = std::signbit(b);
        if (an && !bn) return 1;
        if (!an && bn) return -1;
    }
    if (aa < bb) return -1;
    if (aa > bb) return 1;
    if (aa == bb) return 0;
    if (nan) *nan = true;
    return 0;
}

}  // namespace m2s_oracle
