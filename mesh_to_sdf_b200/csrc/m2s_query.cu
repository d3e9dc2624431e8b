// Query kernels of libm2s.so: exact nearest-triangle search over the LBVH, the grid Raycast row
// toggles, the per-query ray parity walks and the fused sign epilogues.
//
// Replaces, for every voxel / query point at once:
//   - the splat + heap propagation of generate_grid_sdf (mesh_to_sdf/src/generate/grid.rs:383-558),
//   - compute_raycasts / generate_raycasts (generate/grid.rs:568-684),
//   - the per-query closures of generate_sdf_{default,bvh,rtree,rtree_bvh}
//     (generate/generic/default.rs:28-73, bvh.rs:77-143, rtree.rs:114-124, rtree_bvh.rs:124-172)
//     including bvh_ext.rs:59-169 (nearest candidates) and the bvh / rstar crate searches.
// The leaf arithmetic (m2s_geom.cuh) is bit-identical to src/geo.rs; the tree only prunes, with a
// conservative slack, so |d| equals the brute-force minimum of default.rs bit for bit.
#include "m2s_geom.cuh"
#include "m2s_internal.h"

namespace m2s {
namespace {

constexpr int STACK_DEPTH = 128;  // >= depth of a Karras tree over 63-bit keys + index tie-break bits

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

// lower bound for one child slot of a search node (4 x float4). Every slot is stored as an oriented box
// (a padded axis-aligned box is the same record with the identity frame), so there is one code path.
__device__ __forceinline__ float child_dist2(const f3 p, const float4 c0, const float4 c1, const float4 c2,
                                             const float4 c3) {
    return obb_dist2(p, c0, c1, c2, c3);
}

// ---------------------------------------------------------------------------------------------------
// Nearest search state. bound2 is the squared pruning radius: a subtree / triangle whose squared
// distance lower bound exceeds bound2 cannot change the result.
// ---------------------------------------------------------------------------------------------------
template <int MODE>
struct Near {
    float best2;    // UNSIGNED / ARGMIN: min squared distance so far
    float bound2;   // pruning radius^2
    float m;        // NORMAL: current signed minimum under compare_distances (lib.rs:242-259)
    uint32_t id;    // ARGMIN: original index of the arg-min triangle (ties -> smallest index)
    uint32_t slot;  // ARGMIN: its position in leaf order
    float eps;      // absolute length slack
    bool nan;

    __device__ __forceinline__ void init(float eps_len) {
        best2 = INFINITY;
        bound2 = INFINITY;
        m = 3.402823466e+38f;  // f32::MAX, default.rs:54 / bvh.rs:83
        id = 0xffffffffu;
        slot = 0;
        eps = eps_len;
        nan = false;
    }
    __device__ __forceinline__ void set_bound(float dist) {
        // (dist + slack)^2, rounded up a little. NORMAL keeps everything inside the near-tie window of
        // compare_distances (2 ulps or 1e-6) alive: a positive near-tie must still be able to win.
        float r = dist + eps;
        if (MODE == MODE_NORMAL) r += fmaxf(1.0e-6f, dist * 2.4e-7f) * 1.5f;
        bound2 = r * r * 1.000001f;
    }
};

// Exact squared distance from p to triangle slot j: geo.rs:70-138 + Point::dist2, un-fused.
__device__ __forceinline__ float exact_d2(const Bvh& bvh, uint32_t j, bool degen, const f3 p) {
    const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
    const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
    const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
    const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
    const f3 q = degen ? closest_point_triangle_any(p, a, bb, c) : closest_point_triangle(p, a, bb, c);
    const f3 dir = v_sub(p, q);
    return v_dot(dir, dir);
}

// One triangle (leaf-order slot j) against the query: the leaf arithmetic of geo.rs:26-56.
template <int MODE>
__device__ __forceinline__ void visit_tri(const Bvh& bvh, uint32_t j, bool degen, const f3 p, Near<MODE>& s) {
    const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
    const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
    const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
    const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
    const f3 q = degen ? closest_point_triangle_any(p, a, bb, c) : closest_point_triangle(p, a, bb, c);
    const f3 dir = v_sub(p, q);
    const float d2 = v_dot(dir, dir);
    if (MODE == MODE_UNSIGNED) {
        if (d2 < s.best2) {
            s.best2 = d2;
            s.slot = j;
            s.set_bound(sqrt_approx(d2));
        }
    } else if (MODE == MODE_ARGMIN) {
        if (d2 <= s.best2) {
            const uint32_t id = bvh.tri_id[j] & ~TRI_DEGEN_BIT;
            if (d2 < s.best2 || id < s.id) {
                if (d2 < s.best2) s.set_bound(sqrt_approx(d2));
                s.best2 = d2;
                s.id = id;
                s.slot = j;
            }
        }
    } else {  // MODE_NORMAL
        if (d2 <= s.bound2 || !(d2 == d2)) {
            // geo.rs:43-56: distance = |p - nearest|, sign = dot(p - nearest, ab x ac) > 0
            const float dist = __fsqrt_rn(d2);
            const f3 n = {r2.y, r2.z, r2.w};
            const float sd = v_dot(dir, n) > 0.0f ? dist : -dist;
            bool nan = false;
            if (compare_distances(sd, s.m, &nan) < 0) {
                s.m = sd;
                s.slot = j;
                s.set_bound(dist);
            }
            s.nan |= nan;
        }
    }
}

template <int MODE>
__device__ __forceinline__ void visit_leaf(const Bvh& bvh, uint32_t ref, const f3 p, Near<MODE>& s) {
    const uint32_t leaf = ref & LEAF_INDEX_MASK;
    const bool degen = (ref & LEAF_DEGEN_BIT) != 0u;
    const uint32_t b = leaf * bvh.leaf_size;
    const uint32_t e = min(bvh.nt, b + bvh.leaf_size);
    for (uint32_t j = b; j < e; ++j) {
        // plane-disc pretest: skips the exact (un-fused, ~10x costlier) leaf arithmetic for triangles
        // that provably cannot change the result
        const float4* tb = bvh.tobb + 4 * (size_t)j;
        if (obb_dist2(p, ldg4(tb), ldg4(tb + 1), ldg4(tb + 2), ldg4(tb + 3)) > s.bound2) continue;
        visit_tri<MODE>(bvh, j, degen, p, s);
    }
}

// Seed: evaluate one known-near triangle first so the traversal starts with a tight radius (the
// triangle is the nearest one of a neighbouring voxel / query, found by a coarser pass).
template <int MODE>
__device__ __forceinline__ void seed_tri(const Bvh& bvh, uint32_t j, const f3 p, Near<MODE>& s) {
    if (j >= bvh.nt) return;
    visit_tri<MODE>(bvh, j, (bvh.tri_id[j] & TRI_DEGEN_BIT) != 0u, p, s);
}

// Greedy descent (no backtracking): follows the child with the smaller lower bound down to one leaf and
// evaluates its triangles. ~depth node visits; gives a search that has no seed a finite radius to
// start with (a from-infinity packet walk over 32 spread-out voxels visits thousands of nodes).
template <int MODE>
__device__ __forceinline__ void greedy_seed(const Bvh& bvh, const f3 p, Near<MODE>& s) {
    if (bvh.nt == 0) return;
    uint32_t cur = bvh.root;
    for (int guard = 0; guard < 256 && !(cur & LEAF_BIT); ++guard) {
        const float4* nd = bvh.nodes + NODE_F4 * (size_t)cur;
        const float4 l0 = ldg4(nd), l1 = ldg4(nd + 1), l2 = ldg4(nd + 2), l3 = ldg4(nd + 3);
        const float4 r0 = ldg4(nd + 4), r1 = ldg4(nd + 5), r2 = ldg4(nd + 6), r3 = ldg4(nd + 7);
        const float dl = child_dist2(p, l0, l1, l2, l3), dr = child_dist2(p, r0, r1, r2, r3);
        cur = (dl <= dr ? __float_as_uint(l0.w) : __float_as_uint(r0.w));
    }
    if (!(cur & LEAF_BIT)) return;
    const uint32_t leaf = cur & LEAF_INDEX_MASK;
    const bool degen = (cur & LEAF_DEGEN_BIT) != 0u;
    const uint32_t b = leaf * bvh.leaf_size, e = min(bvh.nt, b + bvh.leaf_size);
    for (uint32_t j = b; j < e; ++j) visit_tri<MODE>(bvh, j, degen, p, s);
}

constexpr uint32_t TRAVERSAL_DONE = 0xffffffffu;  // has LEAF_BIT set; never a real leaf ref (nt < 2^30)

// Ordered depth-first traversal with a (ref, lower bound) stack; entries are re-checked against the
// current radius when popped, so a subtree pushed early is skipped without touching memory.
template <int MODE>
__device__ __forceinline__ uint32_t pop_next(const uint2* stack, int& sp, const Near<MODE>& s) {
    while (sp > 0) {
        const uint2 e = stack[--sp];
        if (__uint_as_float(e.y) <= s.bound2) return e.x;
    }
    return TRAVERSAL_DONE;
}

// One internal node: returns the next ref to visit (a child, a popped entry or TRAVERSAL_DONE).
template <int MODE>
__device__ __forceinline__ uint32_t node_step(const Bvh& bvh, uint32_t cur, const f3 p, const Near<MODE>& s,
                                              uint2* stack, int& sp, int* overflow) {
    const float4* nd = bvh.nodes + NODE_F4 * (size_t)cur;
    const float4 l0 = ldg4(nd), l1 = ldg4(nd + 1), l2 = ldg4(nd + 2), l3 = ldg4(nd + 3);
    const float4 r0 = ldg4(nd + 4), r1 = ldg4(nd + 5), r2 = ldg4(nd + 6), r3 = ldg4(nd + 7);
    const float dl = child_dist2(p, l0, l1, l2, l3);
    const float dr = child_dist2(p, r0, r1, r2, r3);
    const uint32_t lref = __float_as_uint(l0.w), rref = __float_as_uint(r0.w);
    const bool hl = dl <= s.bound2, hr = dr <= s.bound2;
    if (hl && hr) {
        const bool left_first = dl <= dr;
        if (sp < STACK_DEPTH) {
            stack[sp++] = left_first ? make_uint2(rref, __float_as_uint(dr)) : make_uint2(lref, __float_as_uint(dl));
        } else {
            *overflow = 1;
        }
        return left_first ? lref : rref;
    }
    if (hl) return lref;
    if (hr) return rref;
    return pop_next<MODE>(stack, sp, s);
}

// Phase-batched traversal. Measured on config C3: the 32 voxels of a warp do almost the same total
// work (sum / (32 * max) = 0.92) but interleave node and leaf steps differently, so a plain
// while-while loop ran with 11 of 32 threads active. Here every lane first walks internal nodes
// until it has collected LEAF_BATCH candidate leaves (or is done), then all lanes run the cheap
// plane-disc / thin-box pretests of their leaves and collect the surviving triangles, then all lanes
// run the exact (un-fused, reference-order) arithmetic on their survivors. The seed makes the radius
// tight from the start, so postponing the leaves costs almost no extra nodes.
constexpr int LEAF_BATCH = 6;
constexpr int TRI_BATCH = 24;  // >= LEAF_BATCH * typical survivors; flushed when full

template <int MODE>
__device__ __forceinline__ void nearest(const Bvh& bvh, const f3 p, Near<MODE>& s, int* overflow,
                                        uint32_t* work = nullptr) {
    if (bvh.nt == 0) return;
    uint2 stack[STACK_DEPTH];
    uint32_t leafbuf[LEAF_BATCH];
    uint32_t tribuf[TRI_BATCH];
    int sp = 0;
    uint32_t cur = bvh.root;
    uint32_t n_nodes = 0, n_leaves = 0;
    for (;;) {
        // ---- phase 1: internal nodes ----
        int nl = 0;
        while (cur != TRAVERSAL_DONE && nl < LEAF_BATCH) {
            if (cur & LEAF_BIT) {
                leafbuf[nl++] = cur;
                cur = pop_next<MODE>(stack, sp, s);
            } else {
                ++n_nodes;
                cur = node_step<MODE>(bvh, cur, p, s, stack, sp, overflow);
            }
        }
        if (nl == 0) break;
        n_leaves += nl;
        // ---- phase 2: pretests ----
        int ntri = 0;
        for (int k = 0; k < nl; ++k) {
            const uint32_t ref = leafbuf[k];
            const uint32_t leaf = ref & LEAF_INDEX_MASK;
            const uint32_t dg = (ref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u;
            const uint32_t b = leaf * bvh.leaf_size;
            const uint32_t e = min(bvh.nt, b + bvh.leaf_size);
            for (uint32_t j = b; j < e; ++j) {
                const float4* tb = bvh.tobb + 4 * (size_t)j;
                if (obb_dist2(p, ldg4(tb), ldg4(tb + 1), ldg4(tb + 2), ldg4(tb + 3)) > s.bound2) continue;
                if (ntri == TRI_BATCH) {  // rare: flush
                    for (int t = 0; t < ntri; ++t)
                        visit_tri<MODE>(bvh, tribuf[t] & ~TRI_DEGEN_BIT, (tribuf[t] & TRI_DEGEN_BIT) != 0u, p, s);
                    ntri = 0;
                }
                tribuf[ntri++] = j | dg;
            }
        }
        // ---- phase 3: exact leaf arithmetic ----
        for (int t = 0; t < ntri; ++t)
            visit_tri<MODE>(bvh, tribuf[t] & ~TRI_DEGEN_BIT, (tribuf[t] & TRI_DEGEN_BIT) != 0u, p, s);
    }
    if (bvh.stats) {
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
    if (work) *work = n_nodes | (n_leaves << 16);
}

// Parity of the reference hits of the ray o + t * e_AXIS (t > 0) over all triangles: the sign vote of
// bvh.rs:106-134 / rtree_bvh.rs:136-164 / default.rs:34-38. The tree walk plays the role of
// bvh.traverse (a conservative box filter); hits are decided by geo.rs:165-216 alone.
template <int AXIS>
__device__ __forceinline__ uint32_t ray_parity(const Bvh& bvh, const f3 o, int* overflow) {
    if (bvh.nt == 0) return 0u;
    constexpr int IY = (AXIS + 1) % 3, IZ = (AXIS + 2) % 3;
    const float oc[3] = {o.x, o.y, o.z};
    const float ox = oc[AXIS], oy = oc[IY], oz = oc[IZ];
    uint32_t stack[STACK_DEPTH];
    int sp = 0;
    uint32_t cur = bvh.root;
    uint32_t count = 0;
    for (;;) {
        if (cur & LEAF_BIT) {
            const uint32_t leaf = cur & LEAF_INDEX_MASK;
            const uint32_t b = leaf * bvh.leaf_size;
            const uint32_t e = min(bvh.nt, b + bvh.leaf_size);
            for (uint32_t j = b; j < e; ++j) {
                const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
                const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
                const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
                const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
                float t;
                if (ray_aligned<AXIS>(o, a, bb, c, &t)) ++count;
            }
        } else {
            const float4* nd = bvh.boxes + BOX_F4 * (size_t)cur;
            const float4 n0 = ldg4(nd), n1 = ldg4(nd + 1), n2 = ldg4(nd + 2), n3 = ldg4(nd + 3);
            const float llo[3] = {n0.x, n0.y, n0.z}, lhi[3] = {n1.x, n1.y, n1.z};
            const float rlo[3] = {n2.x, n2.y, n2.z}, rhi[3] = {n3.x, n3.y, n3.z};
            const bool hl = oy >= llo[IY] && oy <= lhi[IY] && oz >= llo[IZ] && oz <= lhi[IZ] && ox <= lhi[AXIS];
            const bool hr = oy >= rlo[IY] && oy <= rhi[IY] && oz >= rlo[IZ] && oz <= rhi[IZ] && ox <= rhi[AXIS];
            const uint32_t lref = __float_as_uint(n0.w), rref = __float_as_uint(n2.w);
            if (hl && hr) {
                if (sp < STACK_DEPTH) stack[sp++] = rref;
                else *overflow = 1;
                cur = lref;
                continue;
            }
            if (hl) { cur = lref; continue; }
            if (hr) { cur = rref; continue; }
        }
        if (sp == 0) return count & 1u;
        cur = stack[--sp];
    }
}

template <int MODE>
__device__ __forceinline__ float finish(const Bvh& bvh, const f3 p, const Near<MODE>& s) {
    if (bvh.nt == 0) return 3.402823466e+38f;
    if (MODE == MODE_UNSIGNED) return __fsqrt_rn(s.best2);  // sqrt is monotone: min sqrt = sqrt min
    if (MODE == MODE_NORMAL) return s.m;
    // ARGMIN: point_triangle_signed_distance of THE nearest triangle (rtree.rs:118-123)
    const uint32_t j = s.slot;
    const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
    const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
    const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
    const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
    const f3 q = closest_point_triangle_any(p, a, bb, c);
    const f3 dir = v_sub(p, q);
    const float dist = __fsqrt_rn(v_dot(dir, dir));
    const f3 n = {r2.y, r2.z, r2.w};
    return v_dot(dir, n) > 0.0f ? dist : -dist;
}

// ---------------------------------------------------------------------------------------------------
// Grid kernels. Block = 256 threads = a 4 x 8 x 8 (x, y, z) voxel brick; each warp owns a compact
// 2 x 4 x 4 sub-brick so its 32 traversals stay coherent. Bricks are numbered z-fastest so that
// consecutive blocks share tree nodes in L1/L2.
//
// Seeding hierarchy: the grid is searched coarse to fine. Level L visits one representative voxel
// per S_L^3 block (S = 16, 4, 1) and starts each search with the nearest triangle its parent block's
// representative found, so the pruning radius is tight from the first node on and the 32 lanes of a
// warp walk almost the same root-to-leaf path. The seed only initialises the radius: the search
// itself stays exact.
// ---------------------------------------------------------------------------------------------------
constexpr int BX = 4, BY = 8, BZ = 8;

__device__ __forceinline__ void brick_coords(uint32_t nby, uint32_t nbz, uint32_t* x, uint32_t* y, uint32_t* z) {
    uint32_t bid = blockIdx.x;
    const uint32_t bz = bid % nbz;
    bid /= nbz;
    const uint32_t by = bid % nby;
    const uint32_t bx = bid / nby;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp (wx, wy, wz) in 2 x 2 x 2; lane (lx, ly, lz) in 2 x 4 x 4, z fastest
    *x = bx * BX + ((warp >> 2) & 1u) * 2u + (lane >> 4);
    *y = by * BY + ((warp >> 1) & 1u) * 4u + ((lane >> 2) & 3u);
    *z = bz * BZ + (warp & 1u) * 4u + (lane & 3u);
}

__device__ __forceinline__ uint32_t parent_seed(const SeedLevel& L, uint32_t xr, uint32_t y, uint32_t z) {
    // (xr, y, z): voxel coordinates, x relative to the slab start
    const uint32_t cx = min(xr / L.pstride, L.px - 1), cy = min(y / L.pstride, L.py - 1),
                   cz = min(z / L.pstride, L.pz - 1);
    return L.parent[((size_t)cx * L.py + cy) * L.pz + cz];
}

// Coarse level: one thread per S^3 block of the slab; searches the block's representative voxel and
// stores the nearest triangle's slot.
#ifndef SEED_MIN_BLOCKS
#define SEED_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(256, SEED_MIN_BLOCKS)
k_grid_seed(const Bvh bvh, const GridParams g, const float grid_mag, const uint32_t stride, const uint32_t cx,
            const uint32_t cy, const uint32_t cz, const SeedLevel L, uint32_t* __restrict__ seeds,
            BuildStatus* __restrict__ st) {
    uint32_t bx, by, bz;
    brick_coords((cy + BY - 1) / BY, (cz + BZ - 1) / BZ, &bx, &by, &bz);
    if (bx >= cx || by >= cy || bz >= cz) return;
    // representative voxel: the block centre, clamped into the slab / grid
    const uint32_t xr = min(bx * stride + stride / 2, g.x1 - g.x0 - 1);
    const uint32_t y = min(by * stride + stride / 2, g.ny - 1), z = min(bz * stride + stride / 2, g.nz - 1);
    const f3 p = {cell_center(g.fx, g.sx, g.x0 + xr), cell_center(g.fy, g.sy, y), cell_center(g.fz, g.sz, z)};
    Near<MODE_UNSIGNED> s;
    s.init(4.0e-6f * fmaxf(scene_magnitude(st), grid_mag));
    if (L.parent) seed_tri<MODE_UNSIGNED>(bvh, parent_seed(L, xr, y, z), p, s);
    else greedy_seed<MODE_UNSIGNED>(bvh, p, s);
    int overflow = 0;
    nearest<MODE_UNSIGNED>(bvh, p, s, &overflow);
    seeds[((size_t)bx * cy + by) * cz + bz] = s.slot;
    if (overflow) atomicExch(&st->stack_overflow, 1);
}

template <int MODE, bool RAYSIGN>
__global__ void __launch_bounds__(256)
k_grid_nearest(const Bvh bvh, const GridParams g, const float grid_mag, const SeedLevel L,
               const uint32_t* __restrict__ px, const uint32_t* __restrict__ py, const uint32_t* __restrict__ pz,
               float* __restrict__ out, BuildStatus* __restrict__ st) {
    uint32_t xr, y, z;
    brick_coords((g.ny + BY - 1) / BY, (g.nz + BZ - 1) / BZ, &xr, &y, &z);
    xr += g.xa - g.x0;
    const uint32_t x = g.x0 + xr;
    if (x >= g.xb || y >= g.ny || z >= g.nz) return;

    const f3 p = {cell_center(g.fx, g.sx, x), cell_center(g.fy, g.sy, y), cell_center(g.fz, g.sz, z)};
    Near<MODE> s;
    s.init(4.0e-6f * fmaxf(scene_magnitude(st), grid_mag));
    if (L.parent) seed_tri<MODE>(bvh, parent_seed(L, xr, y, z), p, s);
    int overflow = 0;
    uint32_t work = 0;
    nearest<MODE>(bvh, p, s, &overflow, &work);
    float d = finish<MODE>(bvh, p, s);
    if (bvh.stats && bvh.stats[3] == 2ull) {  // M2S_STATS=2: emit the per-voxel work instead of the distance
        out[((size_t)xr * g.ny + y) * g.nz + z] = __uint_as_float(work);
        return;
    }

    if (RAYSIGN) {
        // generate/grid.rs:622-639: negative iff >= 2 of the 3 per-axis hit counts are odd. The parity
        // bitmaps hold, per row, bit i = parity of the hits whose increment range 0..=k covers cell i.
        const uint32_t rowx = y * g.nz + z, rowy = x * g.nz + z, rowz = x * g.ny + y;
        const uint32_t rows_x = g.ny * g.nz, rows_y = g.nx * g.nz, rows_z = g.nx * g.ny;
        const uint32_t ox = (px[(size_t)(x >> 5) * rows_x + rowx] >> (x & 31)) & 1u;
        const uint32_t oy = (py[(size_t)(y >> 5) * rows_y + rowy] >> (y & 31)) & 1u;
        const uint32_t oz = (pz[(size_t)(z >> 5) * rows_z + rowz] >> (z & 31)) & 1u;
        if (ox + oy + oz >= 2u) d = -d;
    }
    out[((size_t)xr * g.ny + y) * g.nz + z] = d;
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (MODE == MODE_NORMAL && s.nan) atomicExch(&st->nan_distance, 1);
}

// ---------------------------------------------------------------------------------------------------
// Packet variant of the grid kernel (the default). The 32 voxels of a warp tile (2 x 4 x 4 cells)
// need almost the same nodes, so the warp walks the tree ONCE with a shared stack in shared memory:
// node loads are warp-uniform (one L1 wavefront instead of up to 32 — the per-lane kernel saturated
// the L1 data pipe at 94 %), control flow is uniform (no SIMT divergence in the node loop) and every
// lane still prunes with its own radius, so each voxel's result is the same exact minimum. A child
// is entered if any lane needs it; children are ordered by the warp-min lower bound; a stack entry
// is dropped when its warp-min bound exceeds the warp-max radius. Triangles that survive a lane's
// plane-disc pretest are queued per lane and evaluated with the exact arithmetic in batches.
// ---------------------------------------------------------------------------------------------------
// traversal counters of the packet walk (m2s_debug_stats) cost ~2 % of the kernel: only with -DM2S_STATS_BUILD
#ifdef M2S_STATS_BUILD
#define PKT_COUNT(x) ++(x)
#else
#define PKT_COUNT(x) ((void)0)
#endif
#ifndef PKT_MIN_BLOCKS
#define PKT_MIN_BLOCKS 4
#endif
constexpr int PKT_STACK = 128;
constexpr int PKT_TRI_BATCH = 16;   // per-lane queue of surviving triangles
constexpr int PKT_FLUSH_AT = 6;     // flush all lanes' queues once any lane holds this many
constexpr int PKT_QCAP = 96;        // warp-shared work queue (UNSIGNED): < 32 left over + 2 x 32 appended per node

template <int MODE>
__device__ __forceinline__ float warp_max_bound(const Near<MODE>& s, bool valid) {
    const unsigned b = valid ? __float_as_uint(s.bound2) : 0u;  // bound2 >= 0: uint order == float order
    return __uint_as_float(__reduce_max_sync(0xffffffffu, b));
}

// The packet walk itself: all 32 lanes call it together (valid == false lanes only help). On return every
// valid lane's state `s` holds its own exact result. stack / queue / best are this warp's shared arrays.
struct PacketCounters {
    uint32_t nodes, leaves;
    int overflow;
};

template <int MODE>
__device__ __forceinline__ void packet_search(const Bvh& bvh, const f3 p, const bool valid, Near<MODE>& s,
                                              uint2* stack, uint2* queue, unsigned long long* best,
                                              PacketCounters* ctr) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int overflow = 0;
    uint32_t n_nodes = 0, n_leaves = 0;
    if (!valid) s.bound2 = -1.0f;  // never wants a child or a triangle; its state is never updated
    float max_b = warp_max_bound<MODE>(s, valid);

    // Exact-arithmetic work. NORMAL: per-lane queue (the compare_distances fold is order dependent and
    // stays with its voxel). UNSIGNED: one warp-shared queue of (triangle, owner) items processed 32 at a
    // time, any lane working for any voxel of the tile (the owner's position comes by shuffle, the
    // minimum goes back through a shared-memory atomicMin) - the expensive un-fused code then runs with
    // all lanes busy instead of ~10 of 32.
    uint32_t tribuf[MODE == MODE_UNSIGNED ? 1 : PKT_TRI_BATCH];
    int ntri = 0;
    int qn = 0;  // warp-uniform
    int sp = 0;
    uint32_t cur = bvh.nt ? bvh.root : TRAVERSAL_DONE;

    auto pop = [&]() -> uint32_t {
        uint32_t r = TRAVERSAL_DONE;
        while (sp > 0) {
            const uint2 e = stack[--sp];
            if (__uint_as_float(e.y) <= max_b) {
                r = e.x;
                break;
            }
        }
        __syncwarp();  // every lane has read its entry before lane 0 may overwrite the slot
        return r;
    };
    // every lane calls enqueue (uniform); `want` says whether this lane's voxel needs triangle `item`
    auto enqueue = [&](bool want, uint32_t item) {
        if (MODE == MODE_UNSIGNED) {
            const unsigned m = __ballot_sync(full, want);
            if (want) queue[qn + __popc(m & lt_mask)] = make_uint2(item, lane);
            qn += __popc(m);
        } else if (want) {
            tribuf[ntri++] = item;
        }
    };
    // everything == false: only full batches of 32 (UNSIGNED) / only when a lane's queue is filling (NORMAL)
    auto flush = [&](bool everything) {
        if (MODE == MODE_UNSIGNED) {
            const int nb = everything ? (qn + 31) >> 5 : qn >> 5;
            if (nb == 0) return;
            best[lane] = ((unsigned long long)__float_as_uint(s.best2) << 32) | s.slot;
            __syncwarp();
            for (int b = 0; b < nb; ++b) {
                const int idx = b * 32 + (int)lane;
                const bool act = idx < qn;
                const uint2 it = act ? queue[idx] : make_uint2(0u, lane);
                const f3 po = {__shfl_sync(full, p.x, it.y), __shfl_sync(full, p.y, it.y), __shfl_sync(full, p.z, it.y)};
                if (act) {
                    const uint32_t j = it.x & ~TRI_DEGEN_BIT;
                    const float d2 = exact_d2(bvh, j, (it.x & TRI_DEGEN_BIT) != 0u, po);
                    atomicMin(best + it.y, ((unsigned long long)__float_as_uint(d2) << 32) | j);
                }
            }
            __syncwarp();
            const int done = min(nb * 32, qn), rem = qn - done;
            const uint2 keep = (int)lane < rem ? queue[done + lane] : make_uint2(0u, 0u);
            const unsigned long long v = best[lane];
            __syncwarp();
            if ((int)lane < rem) queue[lane] = keep;
            qn = rem;
            const float nb2 = __uint_as_float((unsigned)(v >> 32));
            if (nb2 < s.best2) {
                s.best2 = nb2;
                s.slot = (uint32_t)v;
                s.set_bound(sqrt_approx(nb2));
            }
            __syncwarp();
        } else {
            const int most = __reduce_max_sync(full, ntri);
            if (!everything && most < PKT_FLUSH_AT) return;
            for (int t = 0; t < most; ++t)
                if (t < ntri) visit_tri<MODE>(bvh, tribuf[t] & ~TRI_DEGEN_BIT, (tribuf[t] & TRI_DEGEN_BIT) != 0u, p, s);
            ntri = 0;
        }
        max_b = warp_max_bound<MODE>(s, valid);
    };

    while (cur != TRAVERSAL_DONE) {
        if (!(cur & LEAF_BIT)) {
            PKT_COUNT(n_nodes);
            const float4* nd = bvh.nodes + NODE_F4 * (size_t)cur;  // warp-uniform address
            const float4 l0 = ldg4(nd), l1 = ldg4(nd + 1), l2 = ldg4(nd + 2), l3 = ldg4(nd + 3);
            const float4 r0 = ldg4(nd + 4), r1 = ldg4(nd + 5), r2 = ldg4(nd + 6), r3 = ldg4(nd + 7);
            const float dl = child_dist2(p, l0, l1, l2, l3);
            const float dr = child_dist2(p, r0, r1, r2, r3);
            const bool hl = dl <= s.bound2, hr = dr <= s.bound2;  // lanes without a voxel carry bound2 = -1
            unsigned bl = __ballot_sync(full, hl), br = __ballot_sync(full, hr);
            const uint32_t lref = __float_as_uint(l0.w), rref = __float_as_uint(r0.w);
            if (bvh.leaf_size == 1u) {
                // single-triangle leaves: the child's box IS the triangle's box, so the lanes that want it
                // queue the triangle right here and the leaf is never pushed / popped / re-tested
                bool queued = false;
                if ((lref & LEAF_BIT) && bl) {
                    enqueue(hl, (lref & LEAF_INDEX_MASK) | ((lref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    bl = 0u;
                    queued = true;
                    PKT_COUNT(n_leaves);
                }
                if ((rref & LEAF_BIT) && br) {
                    enqueue(hr, (rref & LEAF_INDEX_MASK) | ((rref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    br = 0u;
                    queued = true;
                    PKT_COUNT(n_leaves);
                }
                if (queued) flush(false);
            }
            if (bl && br) {
                // warp-min lower bounds over the lanes that want the child
                const unsigned ml = __reduce_min_sync(full, hl ? __float_as_uint(dl) : 0x7f800000u);
                const unsigned mr = __reduce_min_sync(full, hr ? __float_as_uint(dr) : 0x7f800000u);
                const bool left_first = ml <= mr;
                if (sp < PKT_STACK) {
                    if ((threadIdx.x & 31) == 0) stack[sp] = left_first ? make_uint2(rref, mr) : make_uint2(lref, ml);
                    ++sp;
                    __syncwarp();
                } else {
                    overflow = 1;
                }
                cur = left_first ? lref : rref;
            } else if (bl) {
                cur = lref;
            } else if (br) {
                cur = rref;
            } else {
                cur = pop();
            }
        } else {
            PKT_COUNT(n_leaves);
            const uint32_t leaf = cur & LEAF_INDEX_MASK;
            const uint32_t dg = (cur & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u;
            const uint32_t b = leaf * bvh.leaf_size;
            const uint32_t e = min(bvh.nt, b + bvh.leaf_size);
            for (uint32_t j = b; j < e; ++j) {  // warp-uniform loop, uniform loads
                const float4* tb = bvh.tobb + 4 * (size_t)j;
                const bool want = obb_dist2(p, ldg4(tb), ldg4(tb + 1), ldg4(tb + 2), ldg4(tb + 3)) <= s.bound2;
                enqueue(want, j | dg);
                // UNSIGNED: at most 31 + 32 items are queued here, below PKT_QCAP. NORMAL: a lane's queue holds
                // PKT_TRI_BATCH; drain it when a big leaf (K > batch - flush level) could overflow it
                if (MODE == MODE_UNSIGNED) flush(false);
                else if (bvh.leaf_size > PKT_TRI_BATCH - PKT_FLUSH_AT && __any_sync(full, ntri == PKT_TRI_BATCH)) flush(true);
            }
            flush(false);
            cur = pop();
        }
    }
    flush(true);
    ctr->nodes = n_nodes;
    ctr->leaves = n_leaves;
    ctr->overflow = overflow;
}

// SEEDPASS: the same walk over the representative voxels of the stride^3 blocks (cdim = block counts),
// storing the nearest triangle's slot instead of a distance.
template <int MODE, bool RAYSIGN, bool SEEDPASS>
__global__ void __launch_bounds__(256, PKT_MIN_BLOCKS)
k_grid_nearest_pkt(const Bvh bvh, const GridParams g, const float grid_mag, const SeedLevel L,
                   const uint32_t* __restrict__ px, const uint32_t* __restrict__ py,
                   const uint32_t* __restrict__ pz, float* __restrict__ out, BuildStatus* __restrict__ st,
                   const uint32_t stride, const uint3 cdim, uint32_t* tile_slot) {
    __shared__ uint2 s_stack[8][PKT_STACK];
    __shared__ uint2 s_queue[8][MODE == MODE_UNSIGNED ? PKT_QCAP : 1];       // (triangle slot | degen, owner lane)
    __shared__ unsigned long long s_best[8][MODE == MODE_UNSIGNED ? 32 : 1];  // per owner: (d2 bits << 32) | slot
    const unsigned full = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t xr, y, z;
    uint32_t bxs = 0, bys = 0, bzs = 0;
    bool valid;
    if (SEEDPASS) {
        brick_coords((cdim.y + BY - 1) / BY, (cdim.z + BZ - 1) / BZ, &bxs, &bys, &bzs);
        valid = bxs < cdim.x && bys < cdim.y && bzs < cdim.z;
        xr = min(bxs * stride + stride / 2, g.x1 - g.x0 - 1);
        y = min(bys * stride + stride / 2, g.ny - 1);
        z = min(bzs * stride + stride / 2, g.nz - 1);
    } else {
        brick_coords((g.ny + BY - 1) / BY, (g.nz + BZ - 1) / BZ, &xr, &y, &z);
        xr += g.xa - g.x0;  // this launch covers planes [xa, xb) of the slab
        valid = g.x0 + xr < g.xb && y < g.ny && z < g.nz;
    }
    const uint32_t x = g.x0 + xr;
    if (!__any_sync(full, valid)) return;  // warp-uniform

    const f3 p = {cell_center(g.fx, g.sx, x), cell_center(g.fy, g.sy, y), cell_center(g.fz, g.sz, z)};
    Near<MODE> s;
    s.init(4.0e-6f * fmaxf(scene_magnitude(st), grid_mag));
    // Seed = one known-near triangle that gives the search a tight radius from its first node on.
    //   tile_slot (Raycast / unsigned grids): the nearest triangle of the tile four (two) cells back in x,
    //     published by the warp that computed it (four representatives per tile, one per (y, z) quadrant).
    //     Bricks are dispatched in x-major order and ~600 are resident, so the brick 1024 dispatches back
    //     has practically always finished; if its entry is still empty (first brick plane of a launch, or
    //     a straggler) the lane falls back to a greedy descent. The seed only initialises the radius - the
    //     result is the same exact minimum either way, so this benign race cannot change a bit of output.
    //   L.parent: slots from a separate coarse pass (Normal sign: its compare_distances fold depends on
    //     the visiting order, so it keeps the deterministic seed source).
    uint32_t nseed = 0xffffffffu;
    if (!SEEDPASS && tile_slot) {
        const uint32_t plane_bricks = ((g.ny + BY - 1) / BY) * ((g.nz + BZ - 1) / BZ);
        const uint32_t lane = threadIdx.x & 31u, quad = ((lane >> 3) & 1u) * 2u + ((lane >> 1) & 1u);
        if (blockIdx.x >= plane_bricks)
            nseed = __ldcg(tile_slot + ((size_t)(blockIdx.x - plane_bricks) * 8u + (warp | 4u)) * 4u + quad);
    }
#ifdef M2S_STATS_BUILD
    if (!SEEDPASS && tile_slot && bvh.stats && (threadIdx.x & 31u) == 0 && nseed == 0xffffffffu) atomicAdd(bvh.stats + 4, 1ull);
#endif
    if (valid) {
        if (nseed != 0xffffffffu) seed_tri<MODE>(bvh, nseed, p, s);
        else if (L.parent) seed_tri<MODE>(bvh, parent_seed(L, xr, y, z), p, s);
        else greedy_seed<MODE>(bvh, p, s);
    }
    PacketCounters ctr;
    packet_search<MODE>(bvh, p, valid, s, s_stack[warp], s_queue[warp], s_best[warp], &ctr);
    const int overflow = ctr.overflow;
    const uint32_t n_nodes = ctr.nodes, n_leaves = ctr.leaves;
    if (!SEEDPASS && tile_slot) {
        // publish this tile's representatives: the x-far lanes next to the centre of each (y, z) quadrant
        // (lanes 21, 22, 25, 26 = lx 1, ly 1|2, lz 1|2); an entry stays empty if that lane has no voxel
        const uint32_t lane = threadIdx.x & 31u;
        if (valid && (lane == 21u || lane == 22u || lane == 25u || lane == 26u)) {
            const uint32_t quad = ((lane >> 3) & 1u) * 2u + ((lane >> 1) & 1u);
            __stcg(tile_slot + ((size_t)blockIdx.x * 8u + warp) * 4u + quad, s.slot);
        }
    }

    if (SEEDPASS) {
        if (valid) reinterpret_cast<uint32_t*>(out)[((size_t)bxs * cdim.y + bys) * cdim.z + bzs] = s.slot;
    } else if (valid) {
        float d = finish<MODE>(bvh, p, s);
        if (RAYSIGN) {
            // generate/grid.rs:622-639: negative iff >= 2 of the 3 per-axis hit counts are odd.
            const uint32_t rowx = y * g.nz + z, rowy = x * g.nz + z, rowz = x * g.ny + y;
            const uint32_t rows_x = g.ny * g.nz, rows_y = g.nx * g.nz, rows_z = g.nx * g.ny;
            const uint32_t hx = (px[(size_t)(x >> 5) * rows_x + rowx] >> (x & 31)) & 1u;
            const uint32_t hy = (py[(size_t)(y >> 5) * rows_y + rowy] >> (y & 31)) & 1u;
            const uint32_t hz = (pz[(size_t)(z >> 5) * rows_z + rowz] >> (z & 31)) & 1u;
            if (hx + hy + hz >= 2u) d = -d;
        }
        out[((size_t)xr * g.ny + y) * g.nz + z] = d;
        if (MODE == MODE_NORMAL && s.nan) atomicExch(&st->nan_distance, 1);
    }
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (bvh.stats && (threadIdx.x & 31) == 0) {
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------------
// Run kernel: the packet walk with a RUN of V consecutive voxels along z per lane (V = 2 by default) and both
// children of a node evaluated by one packed-fp32 instruction stream (FFMA2 / FMUL2, sm_100a). Default for
// Raycast / unsigned grids.
//
// ncu on k_grid_nearest_pkt (profiles/r1e) showed two co-limiters at ~76 %: issue slots and the L1 -> register
// write-back of the warp-uniform node loads (8 x LDG.128 broadcast = 4 KB of register writes per node visit).
// Here a warp owns 32 V voxels, so one node load and one round of votes / stack traffic serve V times the
// voxels, and the bound arithmetic shrinks from 4 x 22 scalar instructions to 33 per (2 voxels x 2 children):
//   * the node is stored with its children interleaved (Bvh::nodes_il): the low half of every packed
//     operation is the left child, the high half the right child;
//   * voxel i of a lane differs from voxel 0 by the constant step z_i - z_0 along one grid axis, so its
//     projections are one FFMA2 each: (p_i - c).u = (p_0 - c).u + (z_i - z_0) u.z;
//   * frames and extents are pre-scaled (k_nodes_interleave) so that max(|t| - e, 0) is one FADD.SAT.
// Only the pruning bounds are computed this way (plain fp32, conservative: the extents carry the slack);
// results still come from the reference-order un-fused arithmetic of exact_d2, so the output is the same
// exact minimum, bit for bit. Requires single-triangle leaves (K = 1) and at least one internal node.
// LAYOUT 0: lanes 2 x 4 x 4 runs -> tile 2 x 4 x 4V;  LAYOUT 1: lanes 4 x 4 x 2 runs -> tile 4 x 4 x 2V.
// A block of 4 warps covers a 4 x 8 x 4V brick either way. Measured on C3 (flat cells) and on a cubic-cell
// grid: (V, LAYOUT) = (2, 0) is the best or within 1 % of it on both; V = 4 executes fewer instructions
// but loses as much to its 96 registers (fewer resident warps); kept as M2S_PAIR = 5..7 for A/B runs.
// ---------------------------------------------------------------------------------------------------
#ifndef RUN2_MIN_BLOCKS
#define RUN2_MIN_BLOCKS 5
#endif

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
__device__ __forceinline__ unsigned long long pack_best(float d2, uint32_t slot) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | slot;
}

#ifndef RUN4_MIN_BLOCKS
#define RUN4_MIN_BLOCKS 4
#endif
#ifndef RUN_SEED_BLOCKS
#define RUN_SEED_BLOCKS 5  // resident blocks per SM assumed when choosing how far back the seeds come from
#endif
#ifndef RUN_WARPS
#define RUN_WARPS 4  // warps per block: 4 (brick 4 x 8 x 4V) or 8 (two such bricks stacked in z)
#endif

// SIGN: how the sign is found. RUN_SIGN_NONE / RUN_SIGN_RAYCAST search min |d| (Raycast reads the row parities in
// the epilogue). RUN_SIGN_NORMAL restates the compare_distances fold (lib.rs:242-259) without its dependence on the
// visiting order: per voxel it keeps the nearest triangle (a positive one wins an exact tie) AND the nearest
// positive triangle; the result is the positive one if it is approximately equal (2 ulps / 1e-6) to the nearest,
// else the nearest with its own sign. Everything inside the near-tie window of the radius stays alive, as in
// Near<MODE_NORMAL>::set_bound. Values can differ from a triangle-order fold by the width of that window
// (compare_distances is not transitive); signs agree (tests).
enum : int { RUN_SIGN_NONE = 0, RUN_SIGN_RAYCAST = 1, RUN_SIGN_NORMAL = 2 };
constexpr uint32_t RUN_NEG_BIT = 0x80000000u;  // in the packed best word of RUN_SIGN_NORMAL: nearest is negative

// Exact squared distance and (NORMAL) the sign test of geo.rs:43-56: dot(p - nearest, ab x ac) > 0 is positive.
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

template <int SIGN, int V, int LAYOUT>
__global__ void __launch_bounds__(32 * RUN_WARPS, (V == 4 ? RUN4_MIN_BLOCKS : RUN2_MIN_BLOCKS) * 4 / RUN_WARPS)
k_grid_nearest_run(const Bvh bvh, const GridParams g, const float grid_mag, const uint32_t* __restrict__ px,
                   const uint32_t* __restrict__ py, const uint32_t* __restrict__ pz, float* __restrict__ out,
                   BuildStatus* __restrict__ st, uint32_t* tile_slot, const uint32_t seed_planes) {
    constexpr int NV = 32 * V;         // voxels per tile
    constexpr int QCAP = 32 + 2 * NV;  // < 32 items left over + at most 2 leaves x NV voxels appended by one node
    constexpr uint32_t BZT = 4u * V;   // extent in z of the 4 warps that share an (x, y) footprint
    constexpr uint32_t BZR = BZT * (RUN_WARPS / 4);  // brick extent in z
    __shared__ uint2 s_stack[RUN_WARPS][PKT_STACK];
    __shared__ uint2 s_queue[RUN_WARPS][QCAP];            // (triangle slot | degen, owner voxel = i * 32 + lane)
    __shared__ unsigned long long s_best[RUN_WARPS][NV];  // per owner voxel: (d2 bits << 32) | [negative bit] | slot
    __shared__ uint32_t s_pos[RUN_WARPS][SIGN == RUN_SIGN_NORMAL ? NV : 1];  // NORMAL: d2 bits of the nearest positive triangle
    constexpr bool NORMAL = SIGN == RUN_SIGN_NORMAL;
    const unsigned full = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2* const stack = s_stack[warp];
    uint2* const queue = s_queue[warp];
    unsigned long long* const best = s_best[warp];
    uint32_t* const pos = s_pos[warp];

    // bricks numbered z fastest, x slowest: consecutive blocks share tree nodes in L1 / L2
    const uint32_t nby = (g.ny + BY - 1) / BY, nbz = (g.nz + BZR - 1) / BZR;
    uint32_t bid = blockIdx.x;
    const uint32_t bz = bid % nbz;
    bid /= nbz;
    const uint32_t by = bid % nby, bx = bid / nby;
    uint32_t xr, y, z0;  // first voxel of this lane's run (x relative to the slab start)
    if (LAYOUT == 0) {   // warp (wx, wy), lane (lx:2, ly:4, run:4)
        xr = bx * BX + ((warp >> 1) & 1u) * 2u + (lane >> 4);
        y = by * BY + (warp & 1u) * 4u + ((lane >> 2) & 3u);
        z0 = bz * BZR + (warp >> 2) * BZT + (lane & 3u) * V;
    } else {             // warp (wy, wz), lane (lx:4, ly:4, run:2)
        xr = bx * BX + (lane >> 3);
        y = by * BY + ((warp >> 1) & 1u) * 4u + ((lane >> 1) & 3u);
        z0 = bz * BZR + (warp >> 2) * BZT + (warp & 1u) * 2u * V + (lane & 1u) * V;
    }
    xr += g.xa - g.x0;  // this launch covers planes [xa, xb) of the slab
    const uint32_t x = g.x0 + xr;
    const bool valid_xy = x < g.xb && y < g.ny;
    bool valid[V];
#pragma unroll
    for (int i = 0; i < V; ++i) valid[i] = valid_xy && z0 + i < g.nz;
    if (!__any_sync(full, valid[0])) return;  // warp-uniform (voxel 0 is the first of the run to be valid)

    const f3 p0 = {cell_center(g.fx, g.sx, x), cell_center(g.fy, g.sy, y), cell_center(g.fz, g.sz, z0)};
    float step[V];  // z_i - z_0 (step[0] unused)
#pragma unroll
    for (int i = 1; i < V; ++i) step[i] = cell_center(g.fz, g.sz, z0 + i) - p0.z;

    const float mag = fmaxf(scene_magnitude(st), grid_mag);
    const float eps = 4.0e-6f * mag;
    const float inv_s = pair_inv_scale(mag), inv_s2 = inv_s * inv_s;
    // (dist + slack)^2, rounded up a little (Near::set_bound), in the squared units of the scaled nodes
    auto bound_of = [&](float d2) {
        const float dist = sqrt_approx(d2);
        float r = dist + eps;
        if (NORMAL) r += fmaxf(1.0e-6f, dist * 2.4e-7f) * 1.5f;  // the near-tie window of compare_distances
        return r * r * 1.000001f * inv_s2;
    };
    float best2[V], bnd[V];
    uint32_t slot[V];            // NORMAL: | RUN_NEG_BIT if that triangle sees the voxel from behind
    float pos2[NORMAL ? V : 1];  // NORMAL: squared distance of the nearest positive triangle
    bool nan = false;
#pragma unroll
    for (int i = 0; i < V; ++i) { best2[i] = INFINITY; slot[i] = 0u; }
#pragma unroll
    for (int i = 0; i < (NORMAL ? V : 1); ++i) pos2[i] = INFINITY;

    // Seed: the nearest triangle of the voxel with the same (y, z run) on the x-far face of the brick
    // `seed_planes` steps back in x, published by the warp that computed it; see k_grid_nearest_pkt for why this
    // benign race cannot change the output.
    uint32_t nseed = 0xffffffffu;
    const uint32_t back = seed_planes * nby * nbz;  // dispatch distance of that brick
    const uint32_t src_warp = LAYOUT == 0 ? (warp | 2u) : warp, src_idx = LAYOUT == 0 ? (lane & 15u) : (lane & 7u);
    if (tile_slot && blockIdx.x >= back) {
        nseed = __ldcg(tile_slot + ((size_t)(blockIdx.x - back) * RUN_WARPS + src_warp) * 16u + src_idx);
        // a straggler: the brick twice as far back has certainly finished (still a good radius)
        if (nseed == 0xffffffffu && blockIdx.x >= 2u * back)
            nseed = __ldcg(tile_slot + ((size_t)(blockIdx.x - 2u * back) * RUN_WARPS + src_warp) * 16u + src_idx);
    }
#ifdef M2S_STATS_BUILD
    if (tile_slot && bvh.stats && lane == 0 && nseed == 0xffffffffu) atomicAdd(bvh.stats + 4, 1ull);
#endif
    if (__any_sync(full, nseed >= bvh.nt)) {
        // no neighbour result (first brick planes of a launch): one greedy descent for a voxel in the middle of
        // the tile, the same on every lane (uniform loads, no divergence); its triangle seeds the lanes without one
        const f3 pc = {__shfl_sync(full, p0.x, 13), __shfl_sync(full, p0.y, 13), __shfl_sync(full, p0.z, 13)};
        Near<MODE_UNSIGNED> s;
        s.init(eps);
        greedy_seed<MODE_UNSIGNED>(bvh, pc, s);
        if (nseed >= bvh.nt) nseed = s.best2 < INFINITY ? s.slot : 0u;
    }
    if (nseed < bvh.nt) {
        const bool degen = (bvh.tri_id[nseed] & TRI_DEGEN_BIT) != 0u;
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (valid[i]) {
                const f3 pi = {p0.x, p0.y, i == 0 ? p0.z : cell_center(g.fz, g.sz, z0 + i)};
                bool neg = false;
                best2[i] = exact_d2_sign<NORMAL>(bvh, nseed, degen, pi, &neg);
                slot[i] = nseed | (NORMAL && neg ? RUN_NEG_BIT : 0u);
                if (NORMAL && !neg) pos2[i] = best2[i];
                if (NORMAL) nan |= !(best2[i] == best2[i]);
            }
    }
    // voxels outside the grid never want a child or a triangle
#pragma unroll
    for (int i = 0; i < V; ++i) bnd[i] = valid[i] ? bound_of(best2[i]) : -1.0f;
    auto warp_max_b = [&]() {
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) m = fmaxf(m, bnd[i]);
        return __uint_as_float(__reduce_max_sync(full, __float_as_uint(m)));
    };
    float max_b = warp_max_b();

    int qn = 0, sp = 0;  // warp-uniform
    int overflow = 0;
    uint32_t n_nodes = 0, n_leaves = 0;

    // every lane calls it; w[i]: this lane's voxel i needs triangle `item`
    auto enqueue = [&](const bool (&w)[V], uint32_t item) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const unsigned m = __ballot_sync(full, w[i]);
            if (w[i]) queue[qn + __popc(m & lt_mask)] = make_uint2(item, lane + 32u * i);
            qn += __popc(m);
        }
    };
    // exact arithmetic on the queued (triangle, voxel) items, 32 at a time, any lane for any voxel of the tile
    auto flush = [&](bool everything) {
        const int nb = everything ? (qn + 31) >> 5 : qn >> 5;
        if (nb == 0) return;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            best[lane + 32u * i] = pack_best(best2[i], slot[i]);
            if (NORMAL) pos[lane + 32u * i] = __float_as_uint(pos2[i]);
        }
        __syncwarp();
        for (int b = 0; b < nb; ++b) {
            const int idx = b * 32 + (int)lane;
            const bool act = idx < qn;
            const uint2 it = act ? queue[idx] : make_uint2(0u, lane);
            const int ow = (int)(it.y & 31u);
            // the owner's position: its z is recomputed exactly as the owner computed it (Grid::get_cell_center)
            const f3 po = {__shfl_sync(full, p0.x, ow), __shfl_sync(full, p0.y, ow),
                           cell_center(g.fz, g.sz, __shfl_sync(full, z0, ow) + (it.y >> 5))};
            if (act) {
                const uint32_t j = it.x & ~TRI_DEGEN_BIT;
                bool neg = false;
                const float d2 = exact_d2_sign<NORMAL>(bvh, j, (it.x & TRI_DEGEN_BIT) != 0u, po, &neg);
                atomicMin(best + it.y, pack_best(d2, j | (NORMAL && neg ? RUN_NEG_BIT : 0u)));  // tie: positive first
                if (NORMAL && !neg) atomicMin(pos + it.y, __float_as_uint(d2));
                if (NORMAL) nan |= !(d2 == d2);
            }
        }
        __syncwarp();
        const int done = min(nb * 32, qn), rem = qn - done;
        const uint2 keep = (int)lane < rem ? queue[done + lane] : make_uint2(0u, 0u);
        unsigned long long v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = best[lane + 32u * i];
            if (NORMAL) pos2[i] = __uint_as_float(pos[lane + 32u * i]);
        }
        __syncwarp();
        if ((int)lane < rem) queue[lane] = keep;
        qn = rem;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float n2 = __uint_as_float((unsigned)(v[i] >> 32));
            if (n2 < best2[i]) bnd[i] = bound_of(n2);
            best2[i] = n2;  // the packed minimum: never larger than before; NORMAL: an equal d2 may have turned positive
            slot[i] = (uint32_t)v[i];
        }
        __syncwarp();
        max_b = warp_max_b();
    };

    uint32_t cur = bvh.root;  // always an internal node: leaves are consumed at their parent
    for (;;) {
        if (qn >= 32) flush(false);  // here, where the loop-carried state merges anyway
        PKT_COUNT(n_nodes);
#ifdef M2S_STATS_BUILD
        if (bvh.stats && lane == 0) {
            const uint2 nr = bvh.node_range[cur];
            atomicAdd(bvh.stats + 8 + (31 - __clz(nr.y - nr.x + 1u)), 1ull);
        }
#endif
        const float4* nd = bvh.nodes_il + NODE_F4 * (size_t)cur;  // warp-uniform address
        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
        const float4 q4 = ldg4(nd + 4), q5 = ldg4(nd + 5), q6 = ldg4(nd + 6), q7 = ldg4(nd + 7);
        const float2 m1 = make_float2(-1.0f, -1.0f);
        // low half: left child, high half: right child
        const float2 dx = __ffma2_rn(f2lo(q0), m1, make_float2(p0.x, p0.x));
        const float2 dy = __ffma2_rn(f2hi(q0), m1, make_float2(p0.y, p0.y));
        const float2 dz = __ffma2_rn(f2lo(q1), m1, make_float2(p0.z, p0.z));
        const float2 tu = __ffma2_rn(dz, f2lo(q3), __ffma2_rn(dy, f2hi(q2), __fmul2_rn(dx, f2lo(q2))));
        const float2 tv = __ffma2_rn(dz, f2lo(q5), __ffma2_rn(dy, f2hi(q4), __fmul2_rn(dx, f2lo(q4))));
        const float2 tw = __ffma2_rn(dz, f2lo(q7), __ffma2_rn(dy, f2hi(q6), __fmul2_rn(dx, f2lo(q6))));
        const float2 eu = f2hi(q3), ev = f2hi(q5), ew = f2hi(q7);
        float2 dd[V];  // squared lower bounds of voxel i: (left child, right child)
        dd[0] = sumsq2(excess2(tu, eu), excess2(tv, ev), excess2(tw, ew));
#pragma unroll
        for (int i = 1; i < V; ++i) {
            const float2 s2 = make_float2(step[i], step[i]);
            dd[i] = sumsq2(excess2(__ffma2_rn(s2, f2lo(q3), tu), eu), excess2(__ffma2_rn(s2, f2lo(q5), tv), ev),
                           excess2(__ffma2_rn(s2, f2lo(q7), tw), ew));
        }
        bool wl[V], wr[V];  // voxel i wants the left / right child
        bool any_l = false, any_r = false;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            wl[i] = dd[i].x <= bnd[i];
            wr[i] = dd[i].y <= bnd[i];
            any_l |= wl[i];
            any_r |= wr[i];
        }
        unsigned bl = __ballot_sync(full, any_l), br = __ballot_sync(full, any_r);
        const uint32_t lref = __float_as_uint(q1.z), rref = __float_as_uint(q1.w);
        if ((lref | rref) & LEAF_BIT) {
            if (lref & LEAF_BIT) {
                if (bl) {
                    enqueue(wl, (lref & LEAF_INDEX_MASK) | ((lref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT(n_leaves);
                }
                bl = 0u;
            }
            if (rref & LEAF_BIT) {
                if (br) {
                    enqueue(wr, (rref & LEAF_INDEX_MASK) | ((rref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT(n_leaves);
                }
                br = 0u;
            }
        }
        if (bl && br) {
            // the child most lanes are nearer to goes first (two votes: short latency on the path to the next node
            // load); the other one is pushed with its warp-min lower bound over the voxels that want it
            float kl = INFINITY, kr = INFINITY;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                kl = fminf(kl, wl[i] ? dd[i].x : INFINITY);
                kr = fminf(kr, wr[i] ? dd[i].y : INFINITY);
            }
            const unsigned pref_l = __ballot_sync(full, kl < kr), pref_r = __ballot_sync(full, kr < kl);
            const bool left_first = __popc(pref_l) >= __popc(pref_r);
            const unsigned mfar = __reduce_min_sync(full, __float_as_uint(left_first ? kr : kl));
            if (sp < PKT_STACK) {
                if (lane == 0) stack[sp] = make_uint2(left_first ? rref : lref, mfar);
                ++sp;
                __syncwarp();
            } else {
                overflow = 1;
            }
            cur = left_first ? lref : rref;
        } else if (bl) {
            cur = lref;
        } else if (br) {
            cur = rref;
        } else {
            // pop: entries are re-checked against the current warp-max radius, so a subtree pushed early is
            // dropped without touching memory
            uint32_t r = TRAVERSAL_DONE;
            while (sp > 0) {
                const uint2 e = stack[--sp];
                if (__uint_as_float(e.y) <= max_b) {
                    r = e.x;
                    break;
                }
            }
            __syncwarp();  // every lane has read its entry before lane 0 may overwrite the slot
            if (r == TRAVERSAL_DONE) break;
            cur = r;
        }
    }
    flush(true);

    // publish the x-far voxels' nearest triangles for the bricks further in x
    if (tile_slot && (LAYOUT == 0 ? (lane >= 16u && (warp & 2u)) : lane >= 24u)) {
        // the middle voxel of the run (the first one where the run is cut by the grid's end)
        if (valid[0])
            __stcg(tile_slot + ((size_t)blockIdx.x * RUN_WARPS + warp) * 16u + src_idx,
                   (valid[V / 2] ? slot[V / 2] : slot[0]) & ~(NORMAL ? RUN_NEG_BIT : 0u));
    }

    // sqrt is monotone: min sqrt = sqrt min (finish<MODE_UNSIGNED>)
    float res[V];
#pragma unroll
    for (int i = 0; i < V; ++i) res[i] = __fsqrt_rn(best2[i]);
    if (NORMAL) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (slot[i] & RUN_NEG_BIT) {
                // lib.rs:242-254: an approximately equal positive distance beats the negative one
                const float dp = __fsqrt_rn(pos2[NORMAL ? i : 0]);
                res[i] = approx_eq_abs(dp, res[i]) ? dp : -res[i];
            }
        }
        if (__any_sync(full, nan) && lane == 0) atomicExch(&st->nan_distance, 1);  // lib.rs:257 "NaN distance"
    }
    if (SIGN == RUN_SIGN_RAYCAST && valid[0]) {
        // generate/grid.rs:622-639: negative iff >= 2 of the 3 per-axis hit counts are odd. The Z row of the run
        // is one row: its bits z0 .. z0+V-1 sit in one word (V divides 32)
        const uint32_t rows_x = g.ny * g.nz, rows_y = g.nx * g.nz, rows_z = g.nx * g.ny;
        const uint32_t wz = pz[(size_t)(z0 >> 5) * rows_z + (x * g.ny + y)] >> (z0 & 31u);
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (valid[i]) {
                const uint32_t z = z0 + i;
                const uint32_t hx = (px[(size_t)(x >> 5) * rows_x + (y * g.nz + z)] >> (x & 31u)) & 1u;
                const uint32_t hy = (py[(size_t)(y >> 5) * rows_y + (x * g.nz + z)] >> (y & 31u)) & 1u;
                if (hx + hy + ((wz >> i) & 1u) >= 2u) res[i] = -res[i];
            }
    }
    float* const o = out + ((size_t)xr * g.ny + y) * g.nz + z0;
    if (V == 4 && valid[3] && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
    } else if (V == 2 && valid[1] && (reinterpret_cast<uintptr_t>(o) & 7u) == 0) {
        *reinterpret_cast<float2*>(o) = make_float2(res[0], res[1]);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (valid[i]) o[i] = res[i];
    }
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (bvh.stats && lane == 0) {
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------------
// Grid Raycast rows (generate/grid.rs:568-684). Instead of walking a tree per ray, every triangle
// finds the few rows whose start-cell centre falls inside its padded, projected box (the box is
// the bvh crate's filter, geo.rs:4-22), evaluates geo.rs:165-216 there and toggles bit k
// (k = last incremented cell, grid.rs:604-607) of that row. k_rows_scan turns toggles into parities.
// ---------------------------------------------------------------------------------------------------
struct RowRange {
    uint32_t j0, j1, k0, k1;  // inclusive ranges along the two in-plane axes (IY, IZ); empty if j0 > j1
};

// candidate index range [i0, i1] of cells whose centre first + i*size lies in [lo, hi]; conservative
// (callers re-test each centre exactly). Restricted to [c0, c1).
__device__ __forceinline__ void axis_range(float first, float size, uint32_t c0, uint32_t c1, float lo, float hi,
                                           uint32_t* i0, uint32_t* i1) {
    if (c0 >= c1) { *i0 = 1; *i1 = 0; return; }
    if (!(size > 0.0f) || !isfinite((hi - first) / size)) {  // zero / negative cell size: test every cell
        *i0 = c0;
        *i1 = c1 - 1;
        return;
    }
    // exact index set is [ceil(xlo), floor(xhi)]; floor / ceil the other way absorbs the rounding of the
    // quotient (far below one cell unless the grid has > 2^20 cells per axis) — callers re-test exactly
    const float a = floorf((lo - first) / size) - ((hi - lo) > 1048576.0f * size ? 1.0f : 0.0f);
    const float b = ceilf((hi - first) / size) + ((hi - lo) > 1048576.0f * size ? 1.0f : 0.0f);
    if (b < (float)c0 || a > (float)(c1 - 1)) { *i0 = 1; *i1 = 0; return; }
    *i0 = a <= (float)c0 ? c0 : (uint32_t)a;
    *i1 = b >= (float)(c1 - 1) ? c1 - 1 : (uint32_t)b;
}

struct TriAxis {
    f3 a, b, c;
    float lo[3], hi[3];
};

__device__ __forceinline__ TriAxis load_tri(const float4* __restrict__ rec, uint32_t t) {
    const float4 r0 = ldg4(rec + 3 * (size_t)t), r1 = ldg4(rec + 3 * (size_t)t + 1), r2 = ldg4(rec + 3 * (size_t)t + 2);
    TriAxis T;
    T.a = {r0.x, r0.y, r0.z};
    T.b = {r0.w, r1.x, r1.y};
    T.c = {r1.z, r1.w, r2.x};
    const float EPS = 0.0001f;  // geo.rs:5,20-21
    T.lo[0] = fsub(fminf(T.a.x, fminf(T.b.x, T.c.x)), EPS);
    T.lo[1] = fsub(fminf(T.a.y, fminf(T.b.y, T.c.y)), EPS);
    T.lo[2] = fsub(fminf(T.a.z, fminf(T.b.z, T.c.z)), EPS);
    T.hi[0] = fadd(fmaxf(T.a.x, fmaxf(T.b.x, T.c.x)), EPS);
    T.hi[1] = fadd(fmaxf(T.a.y, fmaxf(T.b.y, T.c.y)), EPS);
    T.hi[2] = fadd(fmaxf(T.a.z, fmaxf(T.b.z, T.c.z)), EPS);
    return T;
}

struct RowCtx {
    float first[3], size[3];
    uint32_t n[3];
    uint32_t x0, x1;
};

__device__ __forceinline__ RowRange row_range(const RowCtx& g, const TriAxis& T, int axis) {
    const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
    RowRange r;
    // rows of the Y and Z axes are only needed for the slab's own x range
    const uint32_t y0 = iy == 0 ? g.x0 : 0u, y1 = iy == 0 ? g.x1 : g.n[iy];
    const uint32_t z0 = iz == 0 ? g.x0 : 0u, z1 = iz == 0 ? g.x1 : g.n[iz];
    axis_range(g.first[iy], g.size[iy], y0, y1, T.lo[iy], T.hi[iy], &r.j0, &r.j1);
    axis_range(g.first[iz], g.size[iz], z0, z1, T.lo[iz], T.hi[iz], &r.k0, &r.k1);
    if (r.k0 > r.k1) { r.j0 = 1; r.j1 = 0; }
    // the ray starts at the centre of cell 0 and only sees what lies ahead: box must reach past it
    const float o_ax = cell_center(g.first[axis], g.size[axis], 0u);
    if (!(o_ax <= T.hi[axis])) { r.j0 = 1; r.j1 = 0; }
    return r;
}

// one (row, triangle) test + toggle. (j, k) are the cell indices along (IY, IZ).
__device__ __forceinline__ void row_test(const RowCtx& g, const TriAxis& T, int axis, uint32_t j, uint32_t k,
                                         uint32_t* __restrict__ bits, uint32_t rows) {
    const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
    const float cy = cell_center(g.first[iy], g.size[iy], j), cz = cell_center(g.first[iz], g.size[iz], k);
    if (!(cy >= T.lo[iy] && cy <= T.hi[iy] && cz >= T.lo[iz] && cz <= T.hi[iz])) return;
    float oc[3];
    oc[axis] = cell_center(g.first[axis], g.size[axis], 0u);
    oc[iy] = cy;
    oc[iz] = cz;
    const f3 o = {oc[0], oc[1], oc[2]};
    float t;
    if (!ray_aligned_dyn(axis, o, T.a, T.b, T.c, &t)) return;
    const uint32_t last = row_last_cell(t, g.size[axis], g.n[axis]);
    // row index: X: y*nz + z   Y: x*nz + z   Z: x*ny + y   (the in-plane pair in (x,y,z) order)
    uint32_t row;
    if (axis == 0) row = j * g.n[2] + k;        // (iy, iz) = (y, z)
    else if (axis == 1) row = k * g.n[2] + j;   // (iy, iz) = (z, x)
    else row = j * g.n[1] + k;                  // (iy, iz) = (x, y)
    atomicXor(bits + (size_t)(last >> 5) * rows + row, 1u << (last & 31));
}

constexpr uint32_t ROWS_INLINE_MAX = 96;

__global__ void __launch_bounds__(256)
k_rows_small(const float4* __restrict__ rec, uint32_t nt, const RowCtx g, uint32_t* __restrict__ b0,
             uint32_t* __restrict__ b1, uint32_t* __restrict__ b2, uint32_t* __restrict__ big_list,
             uint32_t* __restrict__ big_count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const TriAxis T = load_tri(rec, t);
#pragma unroll 1
    for (int axis = 0; axis < 3; ++axis) {
        const RowRange r = row_range(g, T, axis);
        if (r.j0 > r.j1) continue;
        const uint64_t cnt = (uint64_t)(r.j1 - r.j0 + 1) * (uint64_t)(r.k1 - r.k0 + 1);
        if (cnt > ROWS_INLINE_MAX) {
            big_list[atomicAdd(big_count, 1u)] = t * 4u + (uint32_t)axis;
            continue;
        }
        uint32_t* bits = axis == 0 ? b0 : (axis == 1 ? b1 : b2);
        const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
        const uint32_t rows = g.n[iy] * g.n[iz];
        for (uint32_t j = r.j0; j <= r.j1; ++j)
            for (uint32_t k = r.k0; k <= r.k1; ++k) row_test(g, T, axis, j, k, bits, rows);
    }
}

// triangles that cover many rows: one block per (triangle, axis), threads stride over the rows
__global__ void __launch_bounds__(256)
k_rows_big(const float4* __restrict__ rec, const RowCtx g, uint32_t* __restrict__ b0, uint32_t* __restrict__ b1,
           uint32_t* __restrict__ b2, const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_count) {
    const uint32_t n = *big_count;
    for (uint32_t e = blockIdx.x; e < n; e += gridDim.x) {
        const uint32_t code = big_list[e];
        const uint32_t t = code >> 2;
        const int axis = (int)(code & 3u);
        const TriAxis T = load_tri(rec, t);
        const RowRange r = row_range(g, T, axis);
        if (r.j0 > r.j1) continue;
        uint32_t* bits = axis == 0 ? b0 : (axis == 1 ? b1 : b2);
        const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
        const uint32_t rows = g.n[iy] * g.n[iz];
        const uint64_t wk = (uint64_t)(r.k1 - r.k0 + 1);
        const uint64_t cnt = (uint64_t)(r.j1 - r.j0 + 1) * wk;
        for (uint64_t i = threadIdx.x; i < cnt; i += blockDim.x)
            row_test(g, T, axis, r.j0 + (uint32_t)(i / wk), r.k0 + (uint32_t)(i % wk), bits, rows);
    }
}

// toggles -> parities, in place: bit i <- XOR of the toggle bits k >= i of the row
__global__ void __launch_bounds__(256)
k_rows_scan(uint32_t* __restrict__ bits, uint32_t rows, uint32_t words) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint32_t carry = 0u;
    for (int w = (int)words - 1; w >= 0; --w) {
        uint32_t v = bits[(size_t)w * rows + r];
        v ^= v >> 1;
        v ^= v >> 2;
        v ^= v >> 4;
        v ^= v >> 8;
        v ^= v >> 16;
        if (carry) v = ~v;
        carry = v & 1u;
        bits[(size_t)w * rows + r] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// Scattered query points (generate_sdf). Queries arrive Morton-sorted (xyz + original index) so a
// warp's 32 traversals are coherent; results are scattered back to query order.
// SIGN: 0 = value already signed (Normal / Rtree) 1 = +X parity (default.rs:65-72)
//       3 = best of the 3 axes (bvh.rs:137-141, rtree_bvh.rs:167-171)
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t POINT_SEED_STRIDE = 32;  // one representative per warp of sorted queries

// Coarse level over the Morton-sorted queries: thread t searches query min(t*stride + stride/2, nq-1).
__global__ void __launch_bounds__(256)
k_points_seed(const Bvh bvh, const float4* __restrict__ q_sorted, uint32_t nq, uint32_t stride, uint32_t count,
              const uint32_t* __restrict__ parent, uint32_t parent_count, uint32_t* __restrict__ seeds,
              BuildStatus* __restrict__ st) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t i = min(t * stride + stride / 2, nq - 1);
    const float4 q = q_sorted[i];
    const f3 p = {q.x, q.y, q.z};
    Near<MODE_UNSIGNED> s;
    s.init(4.0e-6f * scene_magnitude(st));
    if (parent) seed_tri<MODE_UNSIGNED>(bvh, parent[min(t / POINT_SEED_STRIDE, parent_count - 1)], p, s);
    int overflow = 0;
    nearest<MODE_UNSIGNED>(bvh, p, s, &overflow);
    seeds[t] = s.slot;
    if (overflow) atomicExch(&st->stack_overflow, 1);
}

template <int MODE, int SIGN>
__global__ void __launch_bounds__(256)
k_points(const Bvh bvh, const float4* __restrict__ q_sorted, uint32_t nq, const uint32_t* __restrict__ parent,
         uint32_t parent_count, float* __restrict__ out, BuildStatus* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float4 q = q_sorted[i];
    const f3 p = {q.x, q.y, q.z};
    Near<MODE> s;
    s.init(4.0e-6f * scene_magnitude(st));
    if (parent) seed_tri<MODE>(bvh, parent[min(i / POINT_SEED_STRIDE, parent_count - 1)], p, s);
    int overflow = 0;
    nearest<MODE>(bvh, p, s, &overflow);
    float d = finish<MODE>(bvh, p, s);
    if (SIGN == 1) {
        if (ray_parity<0>(bvh, p, &overflow)) d = -d;
    } else if (SIGN == 3) {
        const uint32_t insides = ray_parity<0>(bvh, p, &overflow) + ray_parity<1>(bvh, p, &overflow) +
                                 ray_parity<2>(bvh, p, &overflow);
        if (insides > 1u) d = -d;
    }
    out[__float_as_uint(q.w)] = d;
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (MODE == MODE_NORMAL && s.nan) atomicExch(&st->nan_distance, 1);
}

// Packet variant for scattered queries: 32 consecutive Morton-sorted queries are close together, so they
// share one tree walk exactly like a voxel tile. No seed pass: every lane starts from a greedy descent.
template <int MODE, int SIGN>
__global__ void __launch_bounds__(256, 4)
k_points_pkt(const Bvh bvh, const float4* __restrict__ q_sorted, uint32_t nq, float* __restrict__ out,
             BuildStatus* __restrict__ st) {
    __shared__ uint2 s_stack[8][PKT_STACK];
    __shared__ uint2 s_queue[8][MODE == MODE_UNSIGNED ? PKT_QCAP : 1];
    __shared__ unsigned long long s_best[8][MODE == MODE_UNSIGNED ? 32 : 1];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < nq;
    if (!__any_sync(0xffffffffu, valid)) return;  // warp-uniform
    const float4 q = valid ? q_sorted[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const f3 p = {q.x, q.y, q.z};
    Near<MODE> s;
    s.init(4.0e-6f * scene_magnitude(st));
    if (valid) greedy_seed<MODE>(bvh, p, s);
    PacketCounters ctr;
    packet_search<MODE>(bvh, p, valid, s, s_stack[warp], s_queue[warp], s_best[warp], &ctr);
    int overflow = ctr.overflow;
    if (valid) {
        float d = finish<MODE>(bvh, p, s);
        if (SIGN == 1) {
            if (ray_parity<0>(bvh, p, &overflow)) d = -d;
        } else if (SIGN == 3) {
            const uint32_t insides = ray_parity<0>(bvh, p, &overflow) + ray_parity<1>(bvh, p, &overflow) +
                                     ray_parity<2>(bvh, p, &overflow);
            if (insides > 1u) d = -d;
        }
        out[__float_as_uint(q.w)] = d;
        if (MODE == MODE_NORMAL && s.nan) atomicExch(&st->nan_distance, 1);
    }
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (bvh.stats && (threadIdx.x & 31) == 0) {
        atomicAdd(bvh.stats + 0, (unsigned long long)ctr.nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)ctr.leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------------
// Run variant for scattered queries (the default, V = 1): the packet walk of k_grid_nearest_run for 32 V consecutive
// Morton-sorted queries per warp - both children of a node per packed-fp32 instruction (interleaved, pre-scaled
// nodes, FADD.SAT excess), one warp-shared queue of (triangle, query) items for the exact arithmetic. The queries
// of a lane are independent points (no lattice step), so every query pays its own projections.
//   MODE_UNSIGNED: min |d| (+ the ray-parity sign rules below)
//   MODE_ARGMIN:   signed distance of THE nearest triangle, ties -> lowest original index (rtree.rs:116-123):
//                  the packed word carries (d2, original index, sign) so the winner's sign comes with it
//   MODE_NORMAL:   the order-independent restatement of the compare_distances fold (see k_grid_nearest_run)
// ---------------------------------------------------------------------------------------------------
#ifndef PRUN_MIN_BLOCKS
#define PRUN_MIN_BLOCKS 5
#endif

template <int MODE, int SIGN, int V>
__global__ void __launch_bounds__(128, PRUN_MIN_BLOCKS)
k_points_run(const Bvh bvh, const float4* __restrict__ q_sorted, uint32_t nq, float* __restrict__ out,
             BuildStatus* __restrict__ st) {
    constexpr int NV = 32 * V;
    constexpr int QCAP = 32 + 2 * NV;
    constexpr bool NORMAL = MODE == MODE_NORMAL, ARGMIN = MODE == MODE_ARGMIN;
    __shared__ uint2 s_stack[4][PKT_STACK];
    __shared__ uint2 s_queue[4][QCAP];            // (triangle slot | degen, owner query = i * 32 + lane)
    __shared__ unsigned long long s_best[4][NV];  // per owner: (d2 bits << 32) | payload (see pack below)
    __shared__ uint32_t s_pos[4][NORMAL ? NV : 1];
    const unsigned full = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2* const stack = s_stack[warp];
    uint2* const queue = s_queue[warp];
    unsigned long long* const best = s_best[warp];
    uint32_t* const pos = s_pos[warp];

    const uint32_t base = (blockIdx.x * 4u + warp) * (uint32_t)NV;
    if (base >= nq) return;  // warp-uniform
    bool valid[V];
    f3 p[V];
    uint32_t orig[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const uint32_t idx = base + 32u * i + lane;
        valid[i] = idx < nq;
        const float4 q = valid[i] ? q_sorted[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        p[i] = {q.x, q.y, q.z};
        orig[i] = __float_as_uint(q.w);
    }
    const float mag = scene_magnitude(st);
    const float eps = 4.0e-6f * mag;
    const float inv_s = pair_inv_scale(mag), inv_s2 = inv_s * inv_s;
    auto bound_of = [&](float d2) {
        const float dist = sqrt_approx(d2);
        float r = dist + eps;
        if (NORMAL) r += fmaxf(1.0e-6f, dist * 2.4e-7f) * 1.5f;
        return r * r * 1.000001f * inv_s2;
    };
    // payload of the packed best word: UNSIGNED slot; NORMAL [negative bit 31] | slot (a positive triangle wins an
    // exact tie); ARGMIN (original index << 1) | negative (the lowest original index wins an exact tie)
    auto payload = [&](uint32_t j, bool neg) -> uint32_t {
        if (ARGMIN) return ((bvh.tri_id[j] & ~TRI_DEGEN_BIT) << 1) | (neg ? 1u : 0u);
        if (NORMAL) return j | (neg ? RUN_NEG_BIT : 0u);
        return j;
    };
    float best2[V], bnd[V];
    uint32_t pay[V];
    float pos2[NORMAL ? V : 1];
    bool nan = false;
#pragma unroll
    for (int i = 0; i < V; ++i) { best2[i] = INFINITY; pay[i] = 0u; }
#pragma unroll
    for (int i = 0; i < (NORMAL ? V : 1); ++i) pos2[i] = INFINITY;

    // start: a greedy descent for the lane's first query; its triangle seeds all queries of the lane
    {
        Near<MODE_UNSIGNED> s0;
        s0.init(eps);
        if (valid[0]) greedy_seed<MODE_UNSIGNED>(bvh, p[0], s0);
        if (s0.best2 < INFINITY) {
            const uint32_t j = s0.slot;
            const bool degen = (bvh.tri_id[j] & TRI_DEGEN_BIT) != 0u;
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (valid[i]) {
                    bool neg = false;
                    best2[i] = exact_d2_sign<NORMAL || ARGMIN>(bvh, j, degen, p[i], &neg);
                    pay[i] = payload(j, neg);
                    if (NORMAL && !neg) pos2[i] = best2[i];
                    if (NORMAL) nan |= !(best2[i] == best2[i]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) bnd[i] = valid[i] ? bound_of(best2[i]) : -1.0f;
    auto warp_max_b = [&]() {
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) m = fmaxf(m, bnd[i]);
        return __uint_as_float(__reduce_max_sync(full, __float_as_uint(m)));
    };
    float max_b = warp_max_b();

    int qn = 0, sp = 0;
    int overflow = 0;
    uint32_t n_nodes = 0, n_leaves = 0;
    auto enqueue = [&](const bool (&w)[V], uint32_t item) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const unsigned m = __ballot_sync(full, w[i]);
            if (w[i]) queue[qn + __popc(m & lt_mask)] = make_uint2(item, lane + 32u * i);
            qn += __popc(m);
        }
    };
    auto flush = [&](bool everything) {
        const int nb = everything ? (qn + 31) >> 5 : qn >> 5;
        if (nb == 0) return;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            best[lane + 32u * i] = pack_best(best2[i], pay[i]);
            if (NORMAL) pos[lane + 32u * i] = __float_as_uint(pos2[i]);
        }
        __syncwarp();
        for (int b = 0; b < nb; ++b) {
            const int idx = b * 32 + (int)lane;
            const bool act = idx < qn;
            const uint2 it = act ? queue[idx] : make_uint2(0u, lane);
            const int ow = (int)(it.y & 31u);
            f3 po = {__shfl_sync(full, p[0].x, ow), __shfl_sync(full, p[0].y, ow), __shfl_sync(full, p[0].z, ow)};
#pragma unroll
            for (int i = 1; i < V; ++i) {
                const f3 pi = {__shfl_sync(full, p[i].x, ow), __shfl_sync(full, p[i].y, ow), __shfl_sync(full, p[i].z, ow)};
                if ((int)(it.y >> 5) == i) po = pi;
            }
            if (act) {
                const uint32_t j = it.x & ~TRI_DEGEN_BIT;
                bool neg = false;
                const float d2 = exact_d2_sign<NORMAL || ARGMIN>(bvh, j, (it.x & TRI_DEGEN_BIT) != 0u, po, &neg);
                atomicMin(best + it.y, pack_best(d2, payload(j, neg)));
                if (NORMAL && !neg) atomicMin(pos + it.y, __float_as_uint(d2));
                if (NORMAL) nan |= !(d2 == d2);
            }
        }
        __syncwarp();
        const int done = min(nb * 32, qn), rem = qn - done;
        const uint2 keep = (int)lane < rem ? queue[done + lane] : make_uint2(0u, 0u);
        unsigned long long v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = best[lane + 32u * i];
            if (NORMAL) pos2[i] = __uint_as_float(pos[lane + 32u * i]);
        }
        __syncwarp();
        if ((int)lane < rem) queue[lane] = keep;
        qn = rem;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float n2 = __uint_as_float((unsigned)(v[i] >> 32));
            if (n2 < best2[i]) bnd[i] = bound_of(n2);
            best2[i] = n2;
            pay[i] = (uint32_t)v[i];
        }
        __syncwarp();
        max_b = warp_max_b();
    };

    uint32_t cur = bvh.root;  // an internal node (the launcher sends single-leaf trees to k_points_pkt)
    for (;;) {
        if (qn >= 32) flush(false);  // here, where the loop-carried state merges anyway
        PKT_COUNT(n_nodes);
        const float4* nd = bvh.nodes_il + NODE_F4 * (size_t)cur;  // warp-uniform address
        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
        const float4 q4 = ldg4(nd + 4), q5 = ldg4(nd + 5), q6 = ldg4(nd + 6), q7 = ldg4(nd + 7);
        const float2 m1 = make_float2(-1.0f, -1.0f);
        const float2 eu = f2hi(q3), ev = f2hi(q5), ew = f2hi(q7);
        float2 dd[V];  // squared lower bounds of query i: (left child, right child)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float2 dx = __ffma2_rn(f2lo(q0), m1, make_float2(p[i].x, p[i].x));
            const float2 dy = __ffma2_rn(f2hi(q0), m1, make_float2(p[i].y, p[i].y));
            const float2 dz = __ffma2_rn(f2lo(q1), m1, make_float2(p[i].z, p[i].z));
            const float2 tu = __ffma2_rn(dz, f2lo(q3), __ffma2_rn(dy, f2hi(q2), __fmul2_rn(dx, f2lo(q2))));
            const float2 tv = __ffma2_rn(dz, f2lo(q5), __ffma2_rn(dy, f2hi(q4), __fmul2_rn(dx, f2lo(q4))));
            const float2 tw = __ffma2_rn(dz, f2lo(q7), __ffma2_rn(dy, f2hi(q6), __fmul2_rn(dx, f2lo(q6))));
            dd[i] = sumsq2(excess2(tu, eu), excess2(tv, ev), excess2(tw, ew));
        }
        bool wl[V], wr[V];  // query i wants the left / right child
        bool any_l = false, any_r = false;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            wl[i] = dd[i].x <= bnd[i];
            wr[i] = dd[i].y <= bnd[i];
            any_l |= wl[i];
            any_r |= wr[i];
        }
        unsigned bl = __ballot_sync(full, any_l), br = __ballot_sync(full, any_r);
        const uint32_t lref = __float_as_uint(q1.z), rref = __float_as_uint(q1.w);
        if ((lref | rref) & LEAF_BIT) {
            if (lref & LEAF_BIT) {
                if (bl) {
                    enqueue(wl, (lref & LEAF_INDEX_MASK) | ((lref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT(n_leaves);
                }
                bl = 0u;
            }
            if (rref & LEAF_BIT) {
                if (br) {
                    enqueue(wr, (rref & LEAF_INDEX_MASK) | ((rref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT(n_leaves);
                }
                br = 0u;
            }
        }
        if (bl && br) {
            float kl = INFINITY, kr = INFINITY;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                kl = fminf(kl, wl[i] ? dd[i].x : INFINITY);
                kr = fminf(kr, wr[i] ? dd[i].y : INFINITY);
            }
            // the child most lanes are nearer to goes first; the other is pushed with its warp-min lower bound
            const unsigned pref_l = __ballot_sync(full, kl < kr), pref_r = __ballot_sync(full, kr < kl);
            const bool left_first = __popc(pref_l) >= __popc(pref_r);
            const unsigned mfar = __reduce_min_sync(full, __float_as_uint(left_first ? kr : kl));
            if (sp < PKT_STACK) {
                if (lane == 0) stack[sp] = make_uint2(left_first ? rref : lref, mfar);
                ++sp;
                __syncwarp();
            } else {
                overflow = 1;
            }
            cur = left_first ? lref : rref;
        } else if (bl) {
            cur = lref;
        } else if (br) {
            cur = rref;
        } else {
            uint32_t r = TRAVERSAL_DONE;
            while (sp > 0) {
                const uint2 e = stack[--sp];
                if (__uint_as_float(e.y) <= max_b) {
                    r = e.x;
                    break;
                }
            }
            __syncwarp();
            if (r == TRAVERSAL_DONE) break;
            cur = r;
        }
    }
    flush(true);

#pragma unroll
    for (int i = 0; i < V; ++i) {
        if (!valid[i]) continue;
        float d = __fsqrt_rn(best2[i]);
        if (ARGMIN) {
            if (pay[i] & 1u) d = -d;  // sign of THE nearest triangle (rtree.rs:118-123)
        } else if (NORMAL) {
            if (pay[i] & RUN_NEG_BIT) {
                const float dp = __fsqrt_rn(pos2[NORMAL ? i : 0]);
                d = approx_eq_abs(dp, d) ? dp : -d;
            }
        }
        if (SIGN == 1) {
            if (ray_parity<0>(bvh, p[i], &overflow)) d = -d;
        } else if (SIGN == 3) {
            const uint32_t insides = ray_parity<0>(bvh, p[i], &overflow) + ray_parity<1>(bvh, p[i], &overflow) +
                                     ray_parity<2>(bvh, p[i], &overflow);
            if (insides > 1u) d = -d;
        }
        out[orig[i]] = d;
    }
    if (NORMAL && __any_sync(full, nan) && lane == 0) atomicExch(&st->nan_distance, 1);
    if (overflow) atomicExch(&st->stack_overflow, 1);
    if (bvh.stats && lane == 0) {
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
}

__global__ void __launch_bounds__(256) k_fill(float* __restrict__ out, uint64_t n, float v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = v;
}

}  // namespace

#define CK(x)                               \
    do {                                    \
        cudaError_t e__ = (x);              \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

static RowCtx make_row_ctx(const GridParams& g) {
    RowCtx c;
    c.first[0] = g.fx; c.first[1] = g.fy; c.first[2] = g.fz;
    c.size[0] = g.sx; c.size[1] = g.sy; c.size[2] = g.sz;
    c.n[0] = g.nx; c.n[1] = g.ny; c.n[2] = g.nz;
    c.x0 = g.x0; c.x1 = g.x1;
    return c;
}

// Row parity bitmaps for the slab [g.x0, g.x1) (all X rows; the Y and Z rows of the slab's planes).
cudaError_t launch_grid_rows(Device& d, const GridParams& g, RowBits* rb, cudaStream_t stream, const float4* rec) {
    cudaStream_t s = stream ? stream : d.stream;
    if (!rec) rec = d.bvh.rec;
    const uint32_t n[3] = {g.nx, g.ny, g.nz};
    for (int a = 0; a < 3; ++a) {
        const int iy = (a + 1) % 3, iz = (a + 2) % 3;
        rb->rows[a] = n[iy] * n[iz];
        rb->words[a] = (n[a] + 31) / 32;
        const size_t bytes = (size_t)rb->rows[a] * rb->words[a] * 4;
        CK(d.rows[a].ensure(bytes));
        rb->bits[a] = d.rows[a].as<uint32_t>();
        CK(cudaMemsetAsync(rb->bits[a], 0, bytes, s));
    }
    const uint32_t nt = d.bvh.nt;
    if (nt == 0) return cudaSuccess;
    CK(d.big_list.ensure((size_t)nt * 3 * 4));
    CK(d.big_count.ensure(4));
    CK(cudaMemsetAsync(d.big_count.p, 0, 4, s));
    const RowCtx c = make_row_ctx(g);
    // any order of the records does: every triangle toggles its own rows
    k_rows_small<<<blocks_for(nt, 256), 256, 0, s>>>(rec, nt, c, rb->bits[0], rb->bits[1], rb->bits[2],
                                                     d.big_list.as<uint32_t>(), d.big_count.as<uint32_t>());
    k_rows_big<<<d.sm_count * 4, 256, 0, s>>>(rec, c, rb->bits[0], rb->bits[1], rb->bits[2],
                                              d.big_list.as<uint32_t>(), d.big_count.as<uint32_t>());
    for (int a = 0; a < 3; ++a)
        k_rows_scan<<<blocks_for(rb->rows[a], 256), 256, 0, s>>>(rb->bits[a], rb->rows[a], rb->words[a]);
    d.launches += 5;
    return cudaGetLastError();
}

static float grid_magnitude(const GridParams& g) {
    float mag = 0.0f;
    const float f[3] = {g.fx, g.fy, g.fz}, sz[3] = {g.sx, g.sy, g.sz};
    const uint32_t n[3] = {g.nx, g.ny, g.nz};
    for (int i = 0; i < 3; ++i) {
        mag = fmaxf(mag, fabsf(f[i]));
        mag = fmaxf(mag, fabsf(f[i] + (float)n[i] * sz[i]));
    }
    return mag;
}

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Raycast / unsigned grids take their seeds from finished neighbour tiles inside the distance kernel
// (k_grid_nearest_pkt); everything else runs the separate coarse pass below.
static bool grid_uses_run_kernel(const Device& d) {
    return d.packet && d.pair && d.neighbour_seeds && d.bvh.leaf_size == 1u && d.bvh.nleaf >= 2u;
}
// k_grid_nearest_pkt keeps the coarse pass for Normal (its fold depends on the visiting order and must stay
// deterministic); k_grid_nearest_run's Normal rule does not depend on the order.
bool grid_uses_neighbour_seeds(const Device& d, int mode) {
    return d.packet && d.neighbour_seeds && (mode != MODE_NORMAL || grid_uses_run_kernel(d));
}

// Coarse seeding pass(es) over the whole slab [g.x0, g.x1).
// slot selects the seed buffer (two half-slabs may be in flight); stream defaults to the device's.
cudaError_t launch_grid_seeds(Device& d, const GridParams& g, SeedLevel* out, int slot, cudaStream_t stream) {
    cudaStream_t s = stream ? stream : d.stream;
    const uint32_t sx = g.x1 - g.x0;
    const float mag = grid_magnitude(g);
    BuildStatus* st = d.status.as<BuildStatus>();
    SeedLevel L{nullptr, 0, 0, 0, 0};
    // strides 4 (and 16 with M2S_SEED_LEVELS=2); skipped for grids that are too small to profit
    if (d.seed_levels > 0 && (uint64_t)sx * g.ny * g.nz >= 4096) {
        const int nlev = d.seed_levels > 2 ? 2 : d.seed_levels;
        for (int lev = nlev; lev >= 1; --lev) {
            uint32_t stride = 1;
            for (int k = 0; k < lev; ++k) stride *= d.seed_stride;
            const uint32_t cx = cdiv(sx, stride), cy = cdiv(g.ny, stride), cz = cdiv(g.nz, stride);
            DevBuf& buf = d.seeds[(lev - 1 + slot) & 1];
            CK(buf.ensure((size_t)cx * cy * cz * 4));
            const unsigned nb = cdiv(cx, BX) * cdiv(cy, BY) * cdiv(cz, BZ);
            if (d.packet && d.seed_packet)
                k_grid_nearest_pkt<MODE_UNSIGNED, false, true><<<nb, 256, 0, s>>>(
                    d.bvh, g, mag, L, nullptr, nullptr, nullptr, buf.as<float>(), st, stride, make_uint3(cx, cy, cz),
                    nullptr);
            else
                k_grid_seed<<<nb, 256, 0, s>>>(d.bvh, g, mag, stride, cx, cy, cz, L, buf.as<uint32_t>(), st);
            d.launches++;
            L = SeedLevel{buf.as<uint32_t>(), cx, cy, cz, stride};
        }
    }
    *out = L;
    return cudaGetLastError();
}

// The distance kernel over planes [g.xa, g.xb) of the slab.
cudaError_t launch_grid_final(Device& d, const GridParams& g, const SeedLevel& L, int mode, const RowBits* rb,
                              float* d_out) {
    cudaStream_t s = d.stream;
    const uint64_t nblocks = (uint64_t)cdiv(g.xb - g.xa, BX) * cdiv(g.ny, BY) * cdiv(g.nz, BZ);
    if (nblocks == 0) return cudaSuccess;
    if (nblocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const float mag = grid_magnitude(g);
    BuildStatus* st = d.status.as<BuildStatus>();
    const unsigned nb = (unsigned)nblocks;
    if (grid_uses_run_kernel(d)) {
        // several voxels per lane. M2S_PAIR: 1 = default (V, LAYOUT) = (2, 0); 4..7 = (2,0) (2,1) (4,0) (4,1)
        const int variant = d.pair < 4 ? 4 : d.pair;
        const uint32_t V = variant >= 6 ? 4u : 2u;
        const uint32_t bzr = 4u * V * (RUN_WARPS / 4);
        const uint64_t nrun = (uint64_t)cdiv(g.xb - g.xa, BX) * cdiv(g.ny, BY) * cdiv(g.nz, bzr);
        if (nrun > 0x7fffffffull) return cudaErrorInvalidConfiguration;
        const unsigned nbr = (unsigned)nrun;
        CK(d.tile_slot.ensure((size_t)nbr * RUN_WARPS * 16 * 4));
        CK(cudaMemsetAsync(d.tile_slot.p, 0xff, (size_t)nbr * RUN_WARPS * 16 * 4, s));
        uint32_t* tile_slot = d.tile_slot.as<uint32_t>();
        const uint32_t *b0 = rb ? rb->bits[0] : nullptr, *b1 = rb ? rb->bits[1] : nullptr, *b2 = rb ? rb->bits[2] : nullptr;
        // seeds come from the brick `planes` steps back in x: far enough in dispatch order to have finished
        // (about 1.25 x the resident blocks), at most 4 steps (16 cells)
        const uint32_t plane_bricks = cdiv(g.ny, BY) * cdiv(g.nz, bzr);
        const uint32_t resident = (uint32_t)d.sm_count * (V == 4 ? RUN4_MIN_BLOCKS : RUN_SEED_BLOCKS) * 4u / RUN_WARPS;
        const uint32_t planes = std::min(4u, std::max(1u, cdiv(resident * 5u / 4u, plane_bricks)));
        CK(launch_nodes_interleave(d, mag));  // node frames in units of S = 2^k >= 4 x the largest |coordinate|
        const int sign = rb ? RUN_SIGN_RAYCAST : (mode == MODE_NORMAL ? RUN_SIGN_NORMAL : RUN_SIGN_NONE);
#define M2S_RUN(SG, VV, LL) k_grid_nearest_run<SG, VV, LL><<<nbr, 32 * RUN_WARPS, 0, s>>>(d.bvh, g, mag, b0, b1, b2, d_out, st, tile_slot, planes)
#define M2S_RUN3(VV, LL)                                            \
    do {                                                            \
        if (sign == RUN_SIGN_RAYCAST) M2S_RUN(RUN_SIGN_RAYCAST, VV, LL); \
        else if (sign == RUN_SIGN_NORMAL) M2S_RUN(RUN_SIGN_NORMAL, VV, LL); \
        else M2S_RUN(RUN_SIGN_NONE, VV, LL);                        \
    } while (0)
        switch (variant) {
            case 4: M2S_RUN3(2, 0); break;
            case 5: M2S_RUN3(2, 1); break;
            case 6: M2S_RUN3(4, 0); break;
            default: M2S_RUN3(4, 1); break;
        }
#undef M2S_RUN3
#undef M2S_RUN
    } else if (d.packet) {
        uint32_t* tile_slot = nullptr;
        if (grid_uses_neighbour_seeds(d, mode)) {
            CK(d.tile_slot.ensure((size_t)nb * 8 * 4 * 4));
            CK(cudaMemsetAsync(d.tile_slot.p, 0xff, (size_t)nb * 8 * 4 * 4, s));
            tile_slot = d.tile_slot.as<uint32_t>();
        }
        if (rb) {
            k_grid_nearest_pkt<MODE_UNSIGNED, true, false><<<nb, 256, 0, s>>>(
                d.bvh, g, mag, L, rb->bits[0], rb->bits[1], rb->bits[2], d_out, st, 1u, make_uint3(0, 0, 0), tile_slot);
        } else if (mode == MODE_NORMAL) {
            k_grid_nearest_pkt<MODE_NORMAL, false, false><<<nb, 256, 0, s>>>(
                d.bvh, g, mag, L, nullptr, nullptr, nullptr, d_out, st, 1u, make_uint3(0, 0, 0), nullptr);
        } else {
            k_grid_nearest_pkt<MODE_UNSIGNED, false, false><<<nb, 256, 0, s>>>(
                d.bvh, g, mag, L, nullptr, nullptr, nullptr, d_out, st, 1u, make_uint3(0, 0, 0), tile_slot);
        }
    } else if (rb) {
        k_grid_nearest<MODE_UNSIGNED, true><<<nb, 256, 0, s>>>(d.bvh, g, mag, L, rb->bits[0], rb->bits[1], rb->bits[2],
                                                               d_out, st);
    } else if (mode == MODE_NORMAL) {
        k_grid_nearest<MODE_NORMAL, false><<<nb, 256, 0, s>>>(d.bvh, g, mag, L, nullptr, nullptr, nullptr, d_out, st);
    } else {
        k_grid_nearest<MODE_UNSIGNED, false><<<nb, 256, 0, s>>>(d.bvh, g, mag, L, nullptr, nullptr, nullptr, d_out, st);
    }
    d.launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_nearest(Device& d, const GridParams& g, int mode, const RowBits* rb, float* d_out,
                                cudaEvent_t after_seeds) {
    SeedLevel L{};
    if (!grid_uses_neighbour_seeds(d, mode) || d.neighbour_and_coarse) CK(launch_grid_seeds(d, g, &L));
    if (after_seeds) cudaEventRecord(after_seeds, d.stream);
    return launch_grid_final(d, g, L, mode, rb, d_out);
}

// sign_rule: 0 none, 1 = +X parity, 3 = best of three axes
cudaError_t launch_points(Device& d, uint64_t nq, int mode, int sign_rule, float* d_out, cudaEvent_t after_seeds) {
    cudaStream_t s = d.stream;
    if (nq == 0) return cudaSuccess;
    const float4* q = d.q_sorted.as<float4>();
    BuildStatus* st = d.status.as<BuildStatus>();
    const uint32_t n = (uint32_t)nq;

    if (d.packet && d.pair && d.bvh.leaf_size == 1u && d.bvh.nleaf >= 2u) {
        // one query per lane by default: measured on C4, 64-query packets lose more to their wider union of
        // candidate triangles than they gain (3.79 vs 3.18 ms; k_points_pkt 3.57). M2S_PAIR=8: two per lane (A/B)
        CK(launch_nodes_interleave(d, 0.0f));  // after sort_queries: the scene bounds now include the queries
        if (after_seeds) cudaEventRecord(after_seeds, s);
        const bool one = d.pair != 8;
        const unsigned nbr = blocks_for(nq, one ? 128 : 256);
#define M2S_PRUN(MD, SG)                                                              \
    do {                                                                              \
        if (one) k_points_run<MD, SG, 1><<<nbr, 128, 0, s>>>(d.bvh, q, n, d_out, st); \
        else k_points_run<MD, SG, 2><<<nbr, 128, 0, s>>>(d.bvh, q, n, d_out, st);     \
    } while (0)
        if (mode == MODE_NORMAL) M2S_PRUN(MODE_NORMAL, 0);
        else if (mode == MODE_ARGMIN) M2S_PRUN(MODE_ARGMIN, 0);
        else if (sign_rule == 1) M2S_PRUN(MODE_UNSIGNED, 1);
        else if (sign_rule == 3) M2S_PRUN(MODE_UNSIGNED, 3);
        else M2S_PRUN(MODE_UNSIGNED, 0);
#undef M2S_PRUN
        d.launches++;
        return cudaGetLastError();
    }
    if (d.packet) {
        if (after_seeds) cudaEventRecord(after_seeds, s);
        const unsigned nbp = blocks_for(nq, 256);
        if (mode == MODE_NORMAL) k_points_pkt<MODE_NORMAL, 0><<<nbp, 256, 0, s>>>(d.bvh, q, n, d_out, st);
        else if (mode == MODE_ARGMIN) k_points_pkt<MODE_ARGMIN, 0><<<nbp, 256, 0, s>>>(d.bvh, q, n, d_out, st);
        else if (sign_rule == 1) k_points_pkt<MODE_UNSIGNED, 1><<<nbp, 256, 0, s>>>(d.bvh, q, n, d_out, st);
        else if (sign_rule == 3) k_points_pkt<MODE_UNSIGNED, 3><<<nbp, 256, 0, s>>>(d.bvh, q, n, d_out, st);
        else k_points_pkt<MODE_UNSIGNED, 0><<<nbp, 256, 0, s>>>(d.bvh, q, n, d_out, st);
        d.launches++;
        return cudaGetLastError();
    }
    const uint32_t* parent = nullptr;
    uint32_t parent_count = 0;
    if (d.seed_levels > 0 && n >= 4096) {
        const int nlev = d.seed_levels > 2 ? 2 : d.seed_levels;
        for (int lev = nlev; lev >= 1; --lev) {
            uint32_t stride = 1;
            for (int k = 0; k < lev; ++k) stride *= POINT_SEED_STRIDE;
            const uint32_t count = cdiv(n, stride);
            DevBuf& buf = d.seeds[lev - 1];
            CK(buf.ensure((size_t)count * 4));
            k_points_seed<<<blocks_for(count, 256), 256, 0, s>>>(d.bvh, q, n, stride, count, parent, parent_count,
                                                                 buf.as<uint32_t>(), st);
            d.launches++;
            parent = buf.as<uint32_t>();
            parent_count = count;
        }
    }
    if (after_seeds) cudaEventRecord(after_seeds, s);
    const unsigned nb = blocks_for(nq, 256);
    if (mode == MODE_NORMAL) k_points<MODE_NORMAL, 0><<<nb, 256, 0, s>>>(d.bvh, q, n, parent, parent_count, d_out, st);
    else if (mode == MODE_ARGMIN) k_points<MODE_ARGMIN, 0><<<nb, 256, 0, s>>>(d.bvh, q, n, parent, parent_count, d_out, st);
    else if (sign_rule == 1) k_points<MODE_UNSIGNED, 1><<<nb, 256, 0, s>>>(d.bvh, q, n, parent, parent_count, d_out, st);
    else if (sign_rule == 3) k_points<MODE_UNSIGNED, 3><<<nb, 256, 0, s>>>(d.bvh, q, n, parent, parent_count, d_out, st);
    else k_points<MODE_UNSIGNED, 0><<<nb, 256, 0, s>>>(d.bvh, q, n, parent, parent_count, d_out, st);
    d.launches++;
    return cudaGetLastError();
}

cudaError_t launch_fill(Device& d, float* d_out, uint64_t n, float value) {
    if (n == 0) return cudaSuccess;
    const unsigned nb = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)d.sm_count * 16);
    k_fill<<<nb, 256, 0, d.stream>>>(d_out, n, value);
    d.launches++;
    return cudaGetLastError();
}

}  // namespace m2s
