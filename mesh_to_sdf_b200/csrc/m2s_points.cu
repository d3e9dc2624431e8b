// Scattered query points (generate_sdf): the exact nearest-triangle search of 32 Morton-sorted queries per warp
// (k_points_run) followed, for the Raycast sign rules, by ONE packet walk of the box tree that counts the hits of
// the +X / +Y / +Z rays of all 32 queries together.
//
// Replaces the per-query closures of generate_sdf_{default,bvh,rtree,rtree_bvh}
// (mesh_to_sdf/src/generate/generic/default.rs:28-73, bvh.rs:77-143, rtree.rs:114-124, rtree_bvh.rs:124-172)
// including bvh_ext.rs:59-169 (nearest candidates), the rstar nearest_neighbor search and the three
// Vec-allocating bvh.traverse ray walks per query.
#include "m2s_search.cuh"

namespace m2s {
namespace {

// ---------------------------------------------------------------------------------------------------
// Ray parities. The reference decides a hit with geo.rs:165-216 alone; bvh.traverse (bvh.rs:119,
// rtree_bvh.rs:149) is a conservative filter over the padded triangle boxes (geo.rs:4-22). Here the 32 queries of
// a warp walk the box tree (Bvh::boxes: two padded child boxes per node) as ONE packet for all three axis rays at
// once: a child is entered if any lane's +X, +Y or +Z ray touches its box; the stack is the warp's shared-memory
// stack of the nearest search (no per-thread local stack), node and triangle loads are warp-uniform. The walk
// needs no ordering: every overlapped leaf is tested exactly once. At a leaf the edge vectors and origin offsets
// are shared by the three axis tests.
// ---------------------------------------------------------------------------------------------------

// which of the rays o + t e_axis (t >= 0 side only matters through the far face) touch the padded box [lo, hi]:
// bit 0 = +X, bit 1 = +Y, bit 2 = +Z (the slab test of the bvh crate restricted to axis-aligned rays)
template <int NAX>
__device__ __forceinline__ uint32_t ray_box_mask(const f3 o, const float4 lo, const float4 hi) {
    const bool inx = o.x >= lo.x && o.x <= hi.x, iny = o.y >= lo.y && o.y <= hi.y, inz = o.z >= lo.z && o.z <= hi.z;
    uint32_t m = (iny && inz && o.x <= hi.x) ? 1u : 0u;
    if (NAX == 3) {
        m |= (inz && inx && o.y <= hi.y) ? 2u : 0u;
        m |= (inx && iny && o.z <= hi.z) ? 4u : 0u;
    }
    return m;
}

// geo.rs:165-216 for the axes in `mask` at once; returns the mask of axes whose ray hits (Some(t), t > 0).
// Per axis the expressions are those of ray_aligned<AXIS> (m2s_geom.cuh), un-fused, in the reference's order.
__device__ __forceinline__ uint32_t ray_aligned_axes(const f3 o, const f3 v0, const f3 v1, const f3 v2, uint32_t mask) {
    const f3 e01 = v_sub(v1, v0), e12 = v_sub(v2, v1), e20 = v_sub(v0, v2);
    const f3 p0 = v_sub(o, v0), p1 = v_sub(o, v1), p2 = v_sub(o, v2);
    uint32_t hit = 0u;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        if (!(mask & (1u << axis))) continue;
        // axis rotation X:(x;y,z) Y:(y;z,x) Z:(z;x,y)   (geo.rs:181-195)
        auto gx = [&](const f3& v) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); };
        auto gy = [&](const f3& v) { return axis == 0 ? v.y : (axis == 1 ? v.z : v.x); };
        auto gz = [&](const f3& v) { return axis == 0 ? v.z : (axis == 1 ? v.x : v.y); };
        const float w0 = fsub(fmul(gz(p1), gy(e12)), fmul(gy(p1), gz(e12)));
        const float w1 = fsub(fmul(gz(p2), gy(e20)), fmul(gy(p2), gz(e20)));
        const float w2 = fsub(fmul(gz(p0), gy(e01)), fmul(gy(p0), gz(e01)));
        if ((w0 < 0.0f && w1 < 0.0f && w2 < 0.0f) || (w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) {
            const float num = fadd(fadd(fmul(w0, gx(p0)), fmul(w2, gx(p2))), fmul(w1, gx(p1)));
            const float t = fdiv(-num, fadd(fadd(w0, w1), w2));
            if (t > 0.0f) hit |= 1u << axis;
        }
    }
    return hit;
}

// All 32 lanes call it together. Returns this lane's parity bits (bit a = parity of the hits of the ray along
// axis a). NAX = 1: only +X (default.rs:34-38); NAX = 3: all three (bvh.rs:106-134, rtree_bvh.rs:136-164).
template <int NAX>
__device__ __forceinline__ uint32_t ray_parities_packet(const Bvh& bvh, const f3 o, const bool valid, uint2* stack,
                                                        int* overflow) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t parity = 0u;
    int sp = 0;
    uint32_t cur = 0u;  // the root: always an internal node
    __syncwarp();
    for (;;) {
        const float4* nd = bvh.boxes + BOX_F4 * (size_t)cur;  // warp-uniform address
        const float4 n0 = ldg4(nd), n1 = ldg4(nd + 1), n2 = ldg4(nd + 2), n3 = ldg4(nd + 3);
        const uint32_t ml = valid ? ray_box_mask<NAX>(o, n0, n1) : 0u;
        const uint32_t mr = valid ? ray_box_mask<NAX>(o, n2, n3) : 0u;
        unsigned bl = __ballot_sync(full, ml != 0u), br = __ballot_sync(full, mr != 0u);
        const uint32_t lref = __float_as_uint(n0.w), rref = __float_as_uint(n2.w);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const uint32_t ref = side ? rref : lref;
            if (!(ref & LEAF_BIT)) continue;
            const unsigned want = side ? br : bl;
            const uint32_t m = side ? mr : ml;
            if (want) {
                const uint32_t j = ref & LEAF_INDEX_MASK;
                const float4 r0 = ldg4(bvh.rec + 3 * (size_t)j);
                const float4 r1 = ldg4(bvh.rec + 3 * (size_t)j + 1);
                const float4 r2 = ldg4(bvh.rec + 3 * (size_t)j + 2);
                if (m) {
                    const f3 a = {r0.x, r0.y, r0.z}, bb = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
                    parity ^= ray_aligned_axes(o, a, bb, c, m);
                }
            }
            if (side) br = 0u; else bl = 0u;
        }
        if (bl && br) {
            if (sp < PKT_STACK) {
                if (lane == 0) stack[sp].x = rref;
                ++sp;
                __syncwarp();
            } else {
                *overflow = 1;
            }
            cur = lref;
        } else if (bl) {
            cur = lref;
        } else if (br) {
            cur = rref;
        } else {
            if (sp == 0) break;
            cur = stack[--sp].x;
            __syncwarp();  // every lane has read the entry before lane 0 may overwrite the slot
        }
    }
    return parity;
}

// ---------------------------------------------------------------------------------------------------
// Ray parities through the ray bins (RayBins, m2s_internal.h): per axis the query looks up the cell of its two
// in-plane coordinates and tests the triangles listed there - exactly the triangles whose padded box its ray can
// touch (plus the few big ones) - with the same box filter and the same geo.rs:165-216 arithmetic as the packet walk
// above, so both give the same parities. One lane per query; no tree, no stack, ~5 candidates per axis on a mesh
// whose triangles are about a cell wide.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ord2f_b(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

template <int AXIS>
__device__ __forceinline__ uint32_t ray_hit_padded(const Bvh& bvh, uint32_t t, const f3 o) {
    const float4 r0 = ldg4(bvh.rec + 3 * (size_t)t), r1 = ldg4(bvh.rec + 3 * (size_t)t + 1),
                 r2 = ldg4(bvh.rec + 3 * (size_t)t + 2);
    const f3 a = {r0.x, r0.y, r0.z}, b = {r0.w, r1.x, r1.y}, c = {r1.z, r1.w, r2.x};
    // the padded box of geo.rs:4-22, as the refit wrote it into the box tree (fsub / fadd of 1e-4 on min / max)
    const float EPS = 0.0001f;
    const float4 lo = make_float4(fsub(fminf(a.x, fminf(b.x, c.x)), EPS), fsub(fminf(a.y, fminf(b.y, c.y)), EPS),
                                  fsub(fminf(a.z, fminf(b.z, c.z)), EPS), 0.f);
    const float4 hi = make_float4(fadd(fmaxf(a.x, fmaxf(b.x, c.x)), EPS), fadd(fmaxf(a.y, fmaxf(b.y, c.y)), EPS),
                                  fadd(fmaxf(a.z, fmaxf(b.z, c.z)), EPS), 0.f);
    if (!(ray_box_mask<3>(o, lo, hi) & (1u << AXIS))) return 0u;
    float t_hit;
    return ray_aligned<AXIS>(o, a, b, c, &t_hit) ? (1u << AXIS) : 0u;
}

template <int AXIS>
__device__ __forceinline__ uint32_t ray_parity_bins_axis(const Bvh& bvh, const f3 o, const bool valid) {
    const RayBins& B = bvh.bins;
    constexpr int IY = (AXIS + 1) % 3, IZ = (AXIS + 2) % 3;
    const float oc[3] = {o.x, o.y, o.z};
    const float lo_y = ord2f_b(B.mesh->lo[IY]), hi_y = ord2f_b(B.mesh->hi[IY]);
    const float lo_z = ord2f_b(B.mesh->lo[IZ]), hi_z = ord2f_b(B.mesh->hi[IZ]);
    uint32_t parity = 0u;
    // outside the mesh's (padded) bounds in the projection: the ray touches no box
    const bool in = valid && oc[IY] >= lo_y && oc[IY] <= hi_y && oc[IZ] >= lo_z && oc[IZ] <= hi_z;
    uint32_t beg = 0u, end = 0u;
    if (in) {
        const float inv_y = hi_y > lo_y ? (float)B.R / (hi_y - lo_y) : 0.0f;
        const float inv_z = hi_z > lo_z ? (float)B.R / (hi_z - lo_z) : 0.0f;
        const float cy = floorf((oc[IY] - lo_y) * inv_y), cz = floorf((oc[IZ] - lo_z) * inv_z);
        const uint32_t j = !(cy > 0.0f) ? 0u : (cy >= (float)B.R ? B.R - 1u : (uint32_t)cy);
        const uint32_t k = !(cz > 0.0f) ? 0u : (cz >= (float)B.R ? B.R - 1u : (uint32_t)cz);
        const uint32_t* off = B.offsets + (size_t)AXIS * B.R * B.R + (size_t)k * B.R + j;
        beg = off[0];
        end = off[1];
    }
    for (uint32_t i = beg; i < end; ++i) parity ^= ray_hit_padded<AXIS>(bvh, B.items[i], o);
    const uint32_t nbig = B.meta[AXIS];
    if (in)
        for (uint32_t i = 0; i < nbig; ++i) parity ^= ray_hit_padded<AXIS>(bvh, B.big[AXIS * RAYBIN_MAX_BIG + i], o);
    return parity;
}

template <int NAX>
__device__ __forceinline__ uint32_t ray_parities_bins(const Bvh& bvh, const f3 o, const bool valid) {
    uint32_t parity = ray_parity_bins_axis<0>(bvh, o, valid);
    if (NAX == 3) {
        parity |= ray_parity_bins_axis<1>(bvh, o, valid);
        parity |= ray_parity_bins_axis<2>(bvh, o, valid);
    }
    return parity;
}

// the sign rules' parities: through the bins when the mesh fits their budgets, else the packet walk of the box tree
template <int NAX>
__device__ __forceinline__ uint32_t ray_parities(const Bvh& bvh, const f3 o, const bool valid, uint2* stack,
                                                 int* overflow) {
    if (bvh.bins.meta != nullptr && bvh.bins.meta[4] != 0u) return ray_parities_bins<NAX>(bvh, o, valid);
    return ray_parities_packet<NAX>(bvh, o, valid, stack, overflow);
}

// ---------------------------------------------------------------------------------------------------
// Run kernel for scattered queries: the packet walk of k_grid_nearest_run (m2s_grid.cu) for 32 consecutive
// Morton-sorted queries per warp - both children of a node per packed-fp32 instruction (interleaved, pre-scaled
// nodes, FADD.SAT excess), one warp-shared queue of (triangle, query) items for the exact arithmetic. The queries
// of a packet are independent points (no lattice step), one per lane (two per lane measured slower: a 64-query
// packet's union of candidates grows faster than its overhead shrinks).
//   MODE_UNSIGNED: min |d| (+ the ray-parity sign rules)
//   MODE_ARGMIN:   signed distance of THE nearest triangle, ties -> lowest original index (rtree.rs:116-123):
//                  the packed word carries (d2, original index, sign) so the winner's sign comes with it
//   MODE_NORMAL:   the order-independent restatement of the compare_distances fold (see k_grid_nearest_run)
// SIGN: 0 = value already signed / unsigned  1 = +X parity (default.rs:65-72)
//       3 = best of the 3 axes (bvh.rs:137-141, rtree_bvh.rs:167-171)
// ---------------------------------------------------------------------------------------------------
#ifndef PRUN_MIN_BLOCKS
#define PRUN_MIN_BLOCKS 5
#endif
#ifndef PRUN_BLOCK_WARPS
#define PRUN_BLOCK_WARPS 1  // packets per thread block: a block keeps its resources until its slowest warp is done
#endif

template <int MODE, int SIGN>
__global__ void __launch_bounds__(32 * PRUN_BLOCK_WARPS, PRUN_MIN_BLOCKS * 4 / PRUN_BLOCK_WARPS)
k_points_run(const Bvh bvh, const float4* __restrict__ q_sorted, uint32_t nq, float* __restrict__ out,
             BuildStatus* __restrict__ st) {
    constexpr int QCAP = 32 + 2 * 32;
    constexpr bool NORMAL = MODE == MODE_NORMAL, ARGMIN = MODE == MODE_ARGMIN;
    __shared__ uint2 s_stack[PRUN_BLOCK_WARPS][PKT_STACK];
    __shared__ uint2 s_queue[PRUN_BLOCK_WARPS][QCAP];            // (triangle slot | degen, owner lane)
    __shared__ unsigned long long s_best[PRUN_BLOCK_WARPS][32];  // per owner: (d2 bits << 32) | payload (see pack below)
    __shared__ uint32_t s_pos[PRUN_BLOCK_WARPS][NORMAL ? 32 : 1];
    const unsigned full = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2* const stack = s_stack[warp];
    uint2* const queue = s_queue[warp];
    unsigned long long* const best = s_best[warp];
    uint32_t* const pos = s_pos[warp];

    const uint32_t base = (blockIdx.x * (uint32_t)PRUN_BLOCK_WARPS + warp) * 32u;
    if (base >= nq) return;  // warp-uniform
    const uint32_t idx = base + lane;
    const bool valid = idx < nq;
    const float4 q = valid ? q_sorted[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    const f3 p = {q.x, q.y, q.z};
    const uint32_t orig = __float_as_uint(q.w);

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
    float best2 = INFINITY, bnd, pos2 = INFINITY;
    uint32_t pay = 0u;
    bool nan = false;

    // start: a greedy descent per query; its triangle gives the search a finite radius
    if (valid) {
        const uint32_t j = greedy_leaf(bvh, p);
        const bool degen = (bvh.tri_id[j] & TRI_DEGEN_BIT) != 0u;
        bool neg = false;
        best2 = exact_d2_sign<NORMAL || ARGMIN>(bvh, j, degen, p, &neg);
        pay = payload(j, neg);
        if (NORMAL && !neg) pos2 = best2;
        if (NORMAL) nan |= !(best2 == best2);
    }
    bnd = valid ? bound_of(best2) : -1.0f;
    auto warp_max_b = [&]() { return __uint_as_float(__reduce_max_sync(full, __float_as_uint(fmaxf(bnd, 0.0f)))); };
    float max_b = warp_max_b();

    int qn = 0, sp = 0;
    int overflow = 0;
    [[maybe_unused]] uint32_t n_nodes = 0, n_leaves = 0;
    auto enqueue = [&](const bool w, uint32_t item) {
        const unsigned m = __ballot_sync(full, w);
        if (w) queue[qn + __popc(m & lt_mask)] = make_uint2(item, lane);
        qn += __popc(m);
    };
    auto flush = [&](bool everything) {
        const int nb = everything ? (qn + 31) >> 5 : qn >> 5;
        if (nb == 0) return;
        best[lane] = pack_best(best2, pay);
        if (NORMAL) pos[lane] = __float_as_uint(pos2);
        __syncwarp();
        for (int b = 0; b < nb; ++b) {
            const int i = b * 32 + (int)lane;
            const bool act = i < qn;
            const uint2 it = act ? queue[i] : make_uint2(0u, lane);
            const int ow = (int)(it.y & 31u);
            const f3 po = {__shfl_sync(full, p.x, ow), __shfl_sync(full, p.y, ow), __shfl_sync(full, p.z, ow)};
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
        const unsigned long long v = best[lane];
        if (NORMAL) pos2 = __uint_as_float(pos[lane]);
        __syncwarp();
        if ((int)lane < rem) queue[lane] = keep;
        qn = rem;
        const float n2 = __uint_as_float((unsigned)(v >> 32));
        if (n2 < best2) bnd = bound_of(n2);
        best2 = n2;
        pay = (uint32_t)v;
        __syncwarp();
        max_b = warp_max_b();
    };

    uint32_t cur = 0u;  // the root: always an internal node
    for (;;) {
        if (qn >= 32) flush(false);  // here, where the loop-carried state merges anyway
        PKT_COUNT(n_nodes);
        const float4* nd = bvh.nodes_il + NODE_F4 * (size_t)cur;  // warp-uniform address
        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
        const float4 q4 = ldg4(nd + 4), q5 = ldg4(nd + 5), q6 = ldg4(nd + 6), q7 = ldg4(nd + 7);
        const uint32_t lref = __float_as_uint(q1.z), rref = __float_as_uint(q1.w);
#ifdef M2S_PREFETCH
        if (!(lref & LEAF_BIT)) asm volatile("prefetch.global.L1 [%0];" ::"l"(bvh.nodes_il + NODE_F4 * (size_t)lref));
        if (!(rref & LEAF_BIT)) asm volatile("prefetch.global.L1 [%0];" ::"l"(bvh.nodes_il + NODE_F4 * (size_t)rref));
#endif
        const float2 m1 = make_float2(-1.0f, -1.0f);
        const float2 eu = f2hi(q3), ev = f2hi(q5), ew = f2hi(q7);
        const float2 dx = __ffma2_rn(f2lo(q0), m1, make_float2(p.x, p.x));
        const float2 dy = __ffma2_rn(f2hi(q0), m1, make_float2(p.y, p.y));
        const float2 dz = __ffma2_rn(f2lo(q1), m1, make_float2(p.z, p.z));
        const float2 tu = __ffma2_rn(dz, f2lo(q3), __ffma2_rn(dy, f2hi(q2), __fmul2_rn(dx, f2lo(q2))));
        const float2 tv = __ffma2_rn(dz, f2lo(q5), __ffma2_rn(dy, f2hi(q4), __fmul2_rn(dx, f2lo(q4))));
        const float2 tw = __ffma2_rn(dz, f2lo(q7), __ffma2_rn(dy, f2hi(q6), __fmul2_rn(dx, f2lo(q6))));
        const float2 dd = sumsq2(excess2(tu, eu), excess2(tv, ev), excess2(tw, ew));  // (left child, right child)
        const bool wl = dd.x <= bnd, wr = dd.y <= bnd;
        unsigned bl = __ballot_sync(full, wl), br = __ballot_sync(full, wr);
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
            const float kl = wl ? dd.x : INFINITY, kr = wr ? dd.y : INFINITY;
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

    float d = __fsqrt_rn(best2);
    if (ARGMIN) {
        if (pay & 1u) d = -d;  // sign of THE nearest triangle (rtree.rs:118-123)
    } else if (NORMAL) {
        if (pay & RUN_NEG_BIT) {
            const float dp = __fsqrt_rn(pos2);
            d = approx_eq_abs(dp, d) ? dp : -d;
        }
    }
    if (SIGN == 1) {
        if (ray_parities<1>(bvh, p, valid, stack, &overflow) & 1u) d = -d;
    } else if (SIGN == 3) {
        if (__popc(ray_parities<3>(bvh, p, valid, stack, &overflow)) > 1) d = -d;
    }
    if (valid) out[orig] = d;
    if (NORMAL && __any_sync(full, nan) && lane == 0) atomicExch(&st->nan_distance, 1);
    if (overflow) atomicExch(&st->stack_overflow, 1);
#ifdef M2S_STATS_BUILD
    if (bvh.stats && lane == 0) {
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
        atomicAdd(bvh.stats + 2, 1ull);
    }
#endif
}

}  // namespace

#define CK(x)                               \
    do {                                    \
        cudaError_t e__ = (x);              \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// sign_rule: 0 none, 1 = +X parity, 3 = best of three axes
cudaError_t launch_points(Device& d, MeshDev& m, uint64_t nq, int mode, int sign_rule, float* d_out) {
    cudaStream_t s = d.stream;
    if (nq == 0) return cudaSuccess;
    const float4* q = d.q_sorted.as<float4>();
    BuildStatus* st = d.call_status.as<BuildStatus>();
    const uint32_t n = (uint32_t)nq;
    // after sort_queries: the call's scene bounds now include the queries, which only the device knows
    CK(launch_nodes_interleave(d, m, 0.0f, true));
    if (sign_rule != 0 && !d.no_ray_bins) CK(launch_ray_bins(d, m));
    const unsigned nbr = (unsigned)((nq + 32 * PRUN_BLOCK_WARPS - 1) / (32 * PRUN_BLOCK_WARPS));
    Bvh bvh = m.bvh;
    if (d.no_ray_bins) bvh.bins = RayBins{};
#ifdef M2S_STATS_BUILD
    bvh.stats = d.want_stats ? d.stats.as<unsigned long long>() : nullptr;
#endif
    if (mode == MODE_NORMAL) k_points_run<MODE_NORMAL, 0><<<nbr, 32 * PRUN_BLOCK_WARPS, 0, s>>>(bvh, q, n, d_out, st);
    else if (mode == MODE_ARGMIN) k_points_run<MODE_ARGMIN, 0><<<nbr, 32 * PRUN_BLOCK_WARPS, 0, s>>>(bvh, q, n, d_out, st);
    else if (sign_rule == 1) k_points_run<MODE_UNSIGNED, 1><<<nbr, 32 * PRUN_BLOCK_WARPS, 0, s>>>(bvh, q, n, d_out, st);
    else if (sign_rule == 3) k_points_run<MODE_UNSIGNED, 3><<<nbr, 32 * PRUN_BLOCK_WARPS, 0, s>>>(bvh, q, n, d_out, st);
    else k_points_run<MODE_UNSIGNED, 0><<<nbr, 32 * PRUN_BLOCK_WARPS, 0, s>>>(bvh, q, n, d_out, st);
    d.launches++;
    return cudaGetLastError();
}

}  // namespace m2s
