// Device leaf arithmetic of the hot path. Every expression reproduces the reference's evaluation
// order with *un-contracted* IEEE fp32 operations (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn /
// __fsqrt_rn never fuse into FMA), because rustc does not contract a*b+c. That makes per-triangle
// results bit-identical to mesh_to_sdf/src/geo.rs and keeps the strict predicates
// (geo.rs:203 `w < 0 / w > 0`, :210 `t > 0`) reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m2s {

struct f3 {
    float x, y, z;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ f3 v_sub(f3 a, f3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ f3 v_add(f3 a, f3 b) { return {fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ f3 v_fmul(f3 a, float s) { return {fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }
// src/point.rs:99-101: x*x' + y*y' + z*z' evaluated left to right.
__device__ __forceinline__ float v_dot(f3 a, f3 b) {
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
// src/point.rs:104-110
__device__ __forceinline__ f3 v_cross(f3 a, f3 b) {
    return {fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
            fsub(fmul(a.x, b.y), fmul(a.y, b.x))};
}
__device__ __forceinline__ bool v_eq(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// src/grid.rs:135-141 — first + (idx as f32) * size, mul then add, not fused.
__device__ __forceinline__ float cell_center(float first, float size, uint32_t idx) {
    return fadd(first, fmul((float)idx, size));
}

// src/geo.rs:141-151
__device__ __forceinline__ f3 closest_point_segment(f3 p, f3 a, f3 b) {
    f3 ab = v_sub(b, a);
    float m = v_dot(ab, ab);
    f3 ap = v_sub(p, a);
    float s = fdiv(v_dot(ab, ap), m);
    if (s < 0.0f) s = 0.0f;  // f32::clamp keeps NaN
    if (s > 1.0f) s = 1.0f;
    return v_add(a, v_fmul(ab, s));
}

// src/geo.rs:90-137 — the Embree closest point for a NON-degenerate triangle (a, b, c pairwise
// different; the :73-88 guards are applied by the caller through closest_point_triangle_any).
__device__ __forceinline__ f3 closest_point_triangle(f3 p, f3 a, f3 b, f3 c) {
    const f3 ab = v_sub(b, a);
    const f3 ac = v_sub(c, a);
    const f3 ap = v_sub(p, a);
    const float d1 = v_dot(ab, ap);
    const float d2 = v_dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) return a;
    const f3 bp = v_sub(p, b);
    const float d3 = v_dot(ab, bp);
    const float d4 = v_dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) return b;
    const f3 cp = v_sub(p, c);
    const float d5 = v_dot(ab, cp);
    const float d6 = v_dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) return c;
    const float vc = fsub(fmul(d1, d4), fmul(d3, d2));
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        const float v = fdiv(d1, fsub(d1, d3));
        return v_add(a, v_fmul(ab, v));
    }
    const float vb = fsub(fmul(d5, d2), fmul(d1, d6));
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        const float v = fdiv(d2, fsub(d2, d6));
        return v_add(a, v_fmul(ac, v));
    }
    const float va = fsub(fmul(d3, d6), fmul(d5, d4));
    const float d43 = fsub(d4, d3);
    const float d56 = fsub(d5, d6);
    if (va <= 0.0f && d43 >= 0.0f && d56 >= 0.0f) {
        const float v = fdiv(d43, fadd(d43, d56));
        const f3 bc = v_sub(c, b);
        return v_add(b, v_fmul(bc, v));
    }
    const float denom = fdiv(1.0f, fadd(fadd(va, vb), vc));
    const float v = fmul(vb, denom);
    const float w = fmul(vc, denom);
    return v_add(v_add(a, v_fmul(ab, v)), v_fmul(ac, w));
}

// src/geo.rs:70-138 including the degenerate guards :73-88.
__device__ __forceinline__ f3 closest_point_triangle_any(f3 p, f3 a, f3 b, f3 c) {
    const bool ab_eq = v_eq(a, b), bc_eq = v_eq(b, c), ac_eq = v_eq(a, c);
    if (ab_eq && bc_eq && ac_eq) return a;
    if (ab_eq) return closest_point_segment(p, a, c);
    if (bc_eq) return closest_point_segment(p, a, b);
    if (ac_eq) return closest_point_segment(p, a, b);
    return closest_point_triangle(p, a, b, c);
}

// squared distance p -> q: Point::dist2, src/point.rs:123-126
__device__ __forceinline__ float v_dist2(f3 p, f3 q) {
    f3 d = v_sub(p, q);
    return v_dot(d, d);
}

// src/geo.rs:165-216. axis: 0=X 1=Y 2=Z. Returns true iff the reference returns Some(t).
template <int AXIS>
__device__ __forceinline__ bool ray_aligned(f3 o, f3 v0, f3 v1, f3 v2, float* t_out) {
    auto gx = [](f3 v) { return AXIS == 0 ? v.x : (AXIS == 1 ? v.y : v.z); };
    auto gy = [](f3 v) { return AXIS == 0 ? v.y : (AXIS == 1 ? v.z : v.x); };
    auto gz = [](f3 v) { return AXIS == 0 ? v.z : (AXIS == 1 ? v.x : v.y); };
    const f3 e01 = v_sub(v1, v0);
    const f3 e12 = v_sub(v2, v1);
    const f3 e20 = v_sub(v0, v2);
    const f3 p0 = v_sub(o, v0);
    const f3 p1 = v_sub(o, v1);
    const f3 p2 = v_sub(o, v2);
    const float w0 = fsub(fmul(gz(p1), gy(e12)), fmul(gy(p1), gz(e12)));
    const float w1 = fsub(fmul(gz(p2), gy(e20)), fmul(gy(p2), gz(e20)));
    const float w2 = fsub(fmul(gz(p0), gy(e01)), fmul(gy(p0), gz(e01)));
    if ((w0 < 0.0f && w1 < 0.0f && w2 < 0.0f) || (w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) {
        const float num = fadd(fadd(fmul(w0, gx(p0)), fmul(w2, gx(p2))), fmul(w1, gx(p1)));
        const float t = fdiv(-num, fadd(fadd(w0, w1), w2));
        if (t > 0.0f) {
            *t_out = t;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool ray_aligned_dyn(int axis, f3 o, f3 v0, f3 v1, f3 v2, float* t_out) {
    if (axis == 0) return ray_aligned<0>(o, v0, v1, v2, t_out);
    if (axis == 1) return ray_aligned<1>(o, v0, v1, v2, t_out);
    return ray_aligned<2>(o, v0, v1, v2, t_out);
}

// float-cmp 0.9 approx_eq!(f32, a, b, ulps = 2, epsilon = 1e-6) for NON-NEGATIVE a, b.
__device__ __forceinline__ bool approx_eq_abs(float a, float b) {
    if (a == b) return true;
    if (fabsf(fsub(a, b)) <= 1e-6f) return true;
    int32_t d = (int32_t)((uint32_t)__float_as_int(a) - (uint32_t)__float_as_int(b));
    if (d == INT32_MIN) return false;
    return (d < 0 ? -d : d) <= 2;
}

// src/lib.rs:242-259. -1 Less, 0 Equal, +1 Greater; *nan set where the reference panics.
__device__ __forceinline__ int compare_distances(float a, float b, bool* nan) {
    const float aa = fabsf(a), bb = fabsf(b);
    if (approx_eq_abs(aa, bb)) {
        const bool an = signbit(a), bn = signbit(b);
        if (an && !bn) return 1;
        if (!an && bn) return -1;
    }
    if (aa < bb) return -1;
    if (aa > bb) return 1;
    if (aa == bb) return 0;
    *nan = true;
    return 0;
}

// Rust `floor(t / cs) as usize` then `.min(n - 1)` (generate/grid.rs:604-607): saturating, NaN -> 0.
__device__ __forceinline__ uint32_t row_last_cell(float t, float cs, uint32_t n_axis) {
    const float q = floorf(fdiv(t, cs));
    if (!(q > 0.0f)) return 0u;  // NaN, negative, zero
    if (q >= (float)n_axis) return n_axis - 1u;
    uint32_t k = (uint32_t)q;
    return k < n_axis - 1u ? k : n_axis - 1u;
}

// squared distance from p to the box [lo, hi] (0 inside). Plain (contractible) arithmetic: this is a
// pruning bound, never a result; callers compare it against a slackened bound.
__device__ __forceinline__ float box_dist2(float px, float py, float pz, float lx, float ly, float lz,
                                           float hx, float hy, float hz) {
    const float dx = fmaxf(fmaxf(lx - px, px - hx), 0.0f);
    const float dy = fmaxf(fmaxf(ly - py, py - hy), 0.0f);
    const float dz = fmaxf(fmaxf(lz - pz, pz - hz), 0.0f);
    return dx * dx + dy * dy + dz * dz;
}

}  // namespace m2s
