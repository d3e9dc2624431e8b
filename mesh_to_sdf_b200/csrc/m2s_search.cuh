// Device code shared by the grid and the point query kernels: pruning bounds over the search nodes, the greedy
// start descent, the exact leaf evaluation and the packed-fp32 helpers of the run kernels.
//
// The tree only prunes; results come from the reference-order un-fused arithmetic of m2s_geom.cuh
// (mesh_to_sdf/src/geo.rs:26-138), so |d| equals the brute-force minimum of generic/default.rs bit for bit.
#pragma once
#include "m2s_geom.cuh"
#include "m2s_internal.h"

namespace m2s {

constexpr int PKT_STACK = 128;                   // >= depth of a Karras tree over 48-bit keys + index tie-break bits
constexpr uint32_t TRAVERSAL_DONE = 0xffffffffu;  // has LEAF_BIT set; never a real leaf ref (nt < 2^30)

__device__ __forceinline__ float ord2f_q(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// Largest |coordinate| of the scene (mesh, plus queries once k_point_bounds ran): scales the
// pruning slack (absolute rounding error of the leaf arithmetic is a few ulp(M)).
__device__ __forceinline__ float scene_magnitude(const BuildStatus* __restrict__ st) {
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float lo = ord2f_q(st->lo[i]), hi = ord2f_q(st->hi[i]);
        if (lo <= hi) m = fmaxf(m, fmaxf(fabsf(lo), fabsf(hi)));
    }
    return m;
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// squared lower bound of the distance from p to anything inside the oriented box
// (centre.xyz, *) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew); see m2s_build.cu. Plain fp32: a pruning bound,
// never a result (the extents carry the slack for its rounding).
__device__ __forceinline__ float obb_dist2(const f3 p, const float4 c, const float4 u, const float4 v, const float4 w) {
    const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
    const float a = fmaxf(fabsf(dx * u.x + dy * u.y + dz * u.z) - u.w, 0.0f);
    const float b = fmaxf(fabsf(dx * v.x + dy * v.y + dz * v.z) - v.w, 0.0f);
    const float g = fmaxf(fabsf(dx * w.x + dy * w.y + dz * w.z) - w.w, 0.0f);
    return a * a + b * b + g * g;
}

// Greedy descent (no backtracking): follows the child with the smaller lower bound down to one leaf and returns
// its triangle slot. ~depth node visits; gives a search that has no neighbour seed a finite radius to start with
// (a from-infinity packet walk over 64 spread-out voxels visits thousands of nodes).
__device__ __forceinline__ uint32_t greedy_leaf(const Bvh& bvh, const f3 p) {
    uint32_t cur = 0u;  // the root is always an internal node
    for (int guard = 0; guard < 256 && !(cur & LEAF_BIT); ++guard) {
        const float4* nd = bvh.nodes + NODE_F4 * (size_t)cur;
        const float4 l0 = ldg4(nd), l1 = ldg4(nd + 1), l2 = ldg4(nd + 2), l3 = ldg4(nd + 3);
        const float4 r0 = ldg4(nd + 4), r1 = ldg4(nd + 5), r2 = ldg4(nd + 6), r3 = ldg4(nd + 7);
        const float dl = obb_dist2(p, l0, l1, l2, l3), dr = obb_dist2(p, r0, r1, r2, r3);
        cur = (dl <= dr ? __float_as_uint(l0.w) : __float_as_uint(r0.w));
    }
    return (cur & LEAF_BIT) ? (cur & LEAF_INDEX_MASK) : 0u;
}

// Exact squared distance from p to triangle slot j (geo.rs:70-138 + Point::dist2, un-fused) and, on request, the
// sign test of geo.rs:43-56: dot(p - nearest, ab x ac) > 0 is positive.
template <bool WANT_SIGN>
__device__ __forceinline__ float exact_d2_sign(const Bvh& bvh, uint32_t j, bool degen, const f3 p, bool* negative) {
    const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
    const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
    const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
    const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
    const f3 q = degen ? closest_point_triangle_any(p, a, bb, c) : closest_point_triangle(p, a, bb, c);
    const f3 dir = v_sub(p, q);
    if (WANT_SIGN) {
        const f3 n = {r2.y, r2.z, r2.w};
        *negative = !(v_dot(dir, n) > 0.0f);
    }
    return v_dot(dir, dir);
}

// ---- packed-fp32 helpers (FFMA2 / FMUL2, sm_100a): low half = left child, high half = right child ----
__device__ __forceinline__ float2 f2lo(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 f2hi(const float4 v) { return make_float2(v.z, v.w); }
// max(|t| - e, 0) per half as one saturating add each: projections and extents are pre-scaled so that 1 is
// out of reach inside the scene (k_nodes_interleave)
__device__ __forceinline__ float2 excess2(const float2 t, const float2 e) {
    return make_float2(__saturatef(fabsf(t.x) - e.x), __saturatef(fabsf(t.y) - e.y));
}
__device__ __forceinline__ float2 sumsq2(const float2 a, const float2 b, const float2 c) {
    return __ffma2_rn(c, c, __ffma2_rn(b, b, __fmul2_rn(a, a)));
}
__device__ __forceinline__ unsigned long long pack_best(float d2, uint32_t payload) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | payload;
}

constexpr uint32_t RUN_NEG_BIT = 0x80000000u;     // packed best word of the Normal rule: the nearest triangle is negative

// traversal counters cost ~2 % of the kernel: only with -DM2S_STATS_BUILD
#ifdef M2S_STATS_BUILD
#define PKT_COUNT(x) ++(x)
#else
#define PKT_COUNT(x) ((void)0)
#endif

}  // namespace m2s
