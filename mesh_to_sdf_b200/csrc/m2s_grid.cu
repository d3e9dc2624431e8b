// Grid kernels of libm2s.so: the exact nearest-triangle search of every voxel of a slab (k_grid_nearest_run) with
// the sign applied in its epilogue, and the Raycast row parities (k_rows_*).
//
// Replaces, for every voxel at once:
//   - the splat + heap propagation of generate_grid_sdf (mesh_to_sdf/src/generate/grid.rs:383-558),
//   - compute_raycasts / generate_raycasts (generate/grid.rs:568-684).
// The leaf arithmetic (m2s_geom.cuh) is bit-identical to src/geo.rs; the tree only prunes, with a conservative
// slack, so |d| equals the brute-force minimum of generic/default.rs bit for bit.
#include <algorithm>
#include <type_traits>

#include "m2s_search.cuh"

namespace m2s {
namespace {

// ---------------------------------------------------------------------------------------------------
// Run kernel. One warp = 32 lanes x a RUN of V = 2 or 4 consecutive voxels along z (warp tile 2 x 4 x 4V cells, or
// 4 x 4 x 2V with LAYOUT 1; grid_run_length picks per grid / mesh shape) walks the LBVH ONCE as a packet: the stack lives in shared memory, node loads are warp-uniform (one L1 wavefront instead of up
// to 32), control flow is uniform, and every voxel still prunes with its own radius, so each result is the same
// exact minimum. A child is entered if any voxel needs it; the child most lanes are nearer to goes first; a stack
// entry is dropped when its warp-min bound exceeds the warp-max radius.
//
// Blackwell-specific arithmetic (packed fp32, FFMA2 / FMUL2, sm_100a):
//   * the node is stored with its children interleaved (Bvh::nodes_il): the low half of every packed
//     operation is the left child, the high half the right child;
//   * voxel i of a lane differs from voxel 0 by the constant step z_i - z_0 along one grid axis, so its
//     projections are one FFMA2 each: (p_i - c).u = (p_0 - c).u + (z_i - z_0) u.z;
//   * frames and extents are pre-scaled (k_nodes_interleave) so that max(|t| - e, 0) is one FADD.SAT.
// Only the pruning bounds are computed this way (plain fp32, conservative: the extents carry the slack).
//
// Leaves hold one triangle and are never pushed: the voxels that want a leaf child queue (triangle, voxel) items at
// the parent, in a warp-shared queue in shared memory that is drained 32 at a time, any lane working for any voxel
// of the tile - the owner's position arrives by shuffle (its z recomputed with the owner's own Grid::get_cell_center
// arithmetic), the minimum returns through a 64-bit shared-memory atomicMin on (d2 bits, slot) - so the expensive
// un-fused reference arithmetic runs with ~all lanes busy. The per-voxel results (d2, slot) live only in that shared
// memory; registers keep what every node visit needs, the pruning bound of the lane's own voxels.
//
// Seeds: every search starts from ONE known-near triangle so that its radius is tight from the first node on: each
// tile publishes the nearest-triangle slots of its x-far voxels (tile_slot, __stcg) and a tile starts from the
// entries of the brick `seed_planes` steps back in x (same y and z run; __ldcg). Bricks are dispatched x-major, so
// that brick has normally finished; a straggler's entry is taken from the brick twice as far back, and a tile
// without any (first brick planes of a launch) runs one greedy descent for its middle voxel. The seed only
// initialises the radius - the result is the same exact minimum either way, so this benign race cannot change a
// bit of the output (compute-sanitizer racecheck does not see it: global memory, by design; profiles/).
//
// SIGN: RUN_SIGN_NONE / RUN_SIGN_RAYCAST search min |d| (Raycast reads the row parities in the epilogue).
// RUN_SIGN_NORMAL restates the compare_distances fold (lib.rs:242-259) without its dependence on the visiting
// order: per voxel it keeps the nearest triangle (a positive one wins an exact tie) AND the nearest positive
// triangle; the result is the positive one if it is approximately equal (2 ulps / 1e-6) to the nearest, else the
// nearest with its own sign. Everything inside the near-tie window of the radius stays alive. Values can differ
// from a triangle-order fold by the width of that window (compare_distances is not transitive); signs agree.
// ---------------------------------------------------------------------------------------------------
constexpr int BX = (int)GRID_BRICK_X, BY = 8;  // a brick = 4 warp tiles = 4 x 8 x 4V cells
#ifndef RUN_MIN_BLOCKS
#define RUN_MIN_BLOCKS 7
#endif
#ifndef RUN_SEED_BLOCKS
#define RUN_SEED_BLOCKS 5  // resident blocks per SM assumed when choosing how far back the seeds come from
#endif
constexpr int RUN_WARPS = 4;  // warp tiles per brick
// One warp per thread block (a block of four warps held its registers and shared memory until its slowest warp was
// done: -3 .. -4 % kernel time, -7 .. -11 % for the scattered-query kernel). On grids large enough to keep the device
// busy with blocks twice as long (lane layout 0), a block computes TILES = 2 warp tiles one after the other along x -
// the two x-halves of its half brick - and the second one starts from the first one's results: seeds from 1 - 2 cells
// away, handed over in registers (C3: -4 % kernel time; the global seeds of a 256^3 grid come from 8 - 16 cells away,
// because that is how thick the band of bricks in flight is).
#ifndef RUN_TILES_X
#define RUN_TILES_X 2
#endif
#ifndef RUN_TILES_MIN_WAVES
#define RUN_TILES_MIN_WAVES 8  // two tiles per block only from this many waves of resident blocks on
#endif
static_assert(RUN_TILES_X == 1 || RUN_TILES_X == 2, "a half brick has two x-halves");
#ifndef RUN_FLUSH_AT
#define RUN_FLUSH_AT 32  // queued (triangle, voxel) items that trigger the exact stage
#endif

enum : int { RUN_SIGN_NONE = 0, RUN_SIGN_RAYCAST = 1, RUN_SIGN_NORMAL = 2 };
#ifdef M2S_STATS_ITEMS
#define PKT_COUNT_LEAF(x)
#else
#define PKT_COUNT_LEAF(x) PKT_COUNT(x)
#endif

template <int SIGN, int V, int LAYOUT, int TILES>
__global__ void __launch_bounds__(32, RUN_MIN_BLOCKS * 4)
k_grid_nearest_run(const Bvh bvh, const GridParams g, const float grid_mag, const uint32_t* __restrict__ px,
                   const uint32_t* __restrict__ py, const uint32_t* __restrict__ pz, float* __restrict__ out,
                   BuildStatus* __restrict__ st, uint32_t* tile_slot, const uint32_t seed_planes,
                   const Progress progress) {
    constexpr int NV = 32 * V;         // voxels per tile
    constexpr int QCAP = RUN_FLUSH_AT + 2 * NV;  // < RUN_FLUSH_AT items left over + at most 2 leaves x NV voxels appended by one node
    constexpr uint32_t BZR = 4u * V;   // brick extent in z
    constexpr bool NORMAL = SIGN == RUN_SIGN_NORMAL;
    static_assert(TILES == 1 || (TILES == 2 && LAYOUT == 0), "tiles of a block: one, or the two x-halves (layout 0)");
    constexpr uint32_t BLOCKS_PER_BRICK = RUN_WARPS / TILES;
    __shared__ uint2 stack[PKT_STACK];
    __shared__ uint2 queue[QCAP];             // exact items: (triangle slot | degen, owner voxel = i * 32 + lane)
    __shared__ unsigned long long best[NV];   // per owner voxel: (d2 bits << 32) | [negative bit] | slot
    __shared__ uint32_t pos[NORMAL ? NV : 1];  // NORMAL: d2 bits of the nearest positive triangle
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x;
    // BLOCKS_PER_BRICK consecutive blocks share a brick
    const uint32_t brick = blockIdx.x / BLOCKS_PER_BRICK, sub = blockIdx.x % BLOCKS_PER_BRICK;
    const unsigned lt_mask = (1u << lane) - 1u;

    // bricks numbered z fastest, x slowest: consecutive blocks share tree nodes in L1 / L2
    const uint32_t nby = (g.ny + BY - 1) / BY, nbz = (g.nz + BZR - 1) / BZR;
    uint32_t bid = brick;
    const uint32_t bz = bid % nbz;
    bid /= nbz;
    const uint32_t by = bid % nby, bx = bid / nby;
    // every BLOCK reports the completion of its stores; the last block of a brick plane publishes the plane's flag
    // in mapped host memory (the host copies finished planes while the kernel runs, m2s_api.cu): one thread releases
    // the block's stores at device scope - cumulative over what the barrier made visible to it - and bumps the plane's
    // counter; the block that completes the plane fences at system scope before it publishes the flag.
    auto signal_done = [&]() {
        if (progress.count == nullptr) return;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const uint32_t done = atomicAdd(progress.count + bx, 1u) + 1u;
            if (done == nby * nbz * BLOCKS_PER_BRICK) {
                __threadfence_system();
                progress.flag[bx] = progress.epoch;
            }
        }
    };
    const float mag = fmaxf(scene_magnitude(st), grid_mag);
    const float eps = 4.0e-6f * mag;
    const float inv_s = pair_inv_scale(mag), inv_s2 = inv_s * inv_s;
    int overflow = 0;
    bool nan = false;
    uint32_t carry = 0xffffffffu;  // nearest-triangle slot handed over by the previous tile of this block
#pragma unroll 1
    for (int tile = 0; tile < TILES; ++tile) {
    const uint32_t warp = TILES == 2 ? (uint32_t)tile * 2u + sub : (TILES == 4 ? (uint32_t)tile : sub);  // (wx, wy) / (wy, wz)
    // first voxel of this lane's run (x relative to the slab start); the 4 warp tiles of a brick cover 4 x 8 x 4V cells
    //   LAYOUT 0: warp tile 2 x 4 x 4V - warps (wx, wy), lanes (lx:2, ly:4, run:4): compact where cells are thin in z
    //   LAYOUT 1: warp tile 4 x 4 x 2V - warps (wy, wz), lanes (lx:4, ly:4, run:2): compact where cells are cubic
    const uint32_t xr = bx * BX + (LAYOUT == 0 ? ((warp >> 1) & 1u) * 2u + (lane >> 4) : (lane >> 3));
    const uint32_t y = by * BY + (warp & 1u) * 4u + (LAYOUT == 0 ? ((lane >> 2) & 3u) : ((lane >> 1) & 3u));
    const uint32_t z0 = bz * BZR + (LAYOUT == 0 ? (lane & 3u) * V : (warp >> 1) * 2u * V + (lane & 1u) * V);
    const uint32_t x = g.x0 + xr;
    const bool valid_xy = x < g.x1 && y < g.ny;
    bool valid[V];
#pragma unroll
    for (int i = 0; i < V; ++i) valid[i] = valid_xy && z0 + i < g.nz;

    if (!__any_sync(full, valid[0])) {  // warp-uniform (voxel 0 is the first of the run to be valid)
        carry = 0xffffffffu;
        continue;
    }

    const f3 p0 = {cell_center(g.fx, g.sx, x), cell_center(g.fy, g.sy, y), cell_center(g.fz, g.sz, z0)};
    float step[V];  // z_i - z_0 (step[0] unused)
#pragma unroll
    for (int i = 1; i < V; ++i) step[i] = cell_center(g.fz, g.sz, z0 + i) - p0.z;

    // (dist + slack)^2, rounded up a little: the squared search radius in scene units
    auto radius2_of = [&](float d2) {
        const float dist = sqrt_approx(d2);
        float r = dist + eps;
        if (NORMAL) r += fmaxf(1.0e-6f, dist * 2.4e-7f) * 1.5f;  // the near-tie window of compare_distances
        return r * r * 1.000001f;
    };
    // the same in the squared units of the scaled nodes
    auto bound_of = [&](float d2) { return radius2_of(d2) * inv_s2; };
    // Per-voxel results live in shared memory (best / pos): any lane improves any voxel of the tile with atomicMin.
    // Registers only keep what every node visit needs: the pruning bound of the lane's own voxels.
    float bnd[V];

    // Seed: the nearest triangle of the voxel with the same (y, z run) on the x-far face of the brick
    // `seed_planes` steps back in x, published by the warp that computed it.
    uint32_t nseed = 0xffffffffu;
    const uint32_t back = seed_planes * nby * nbz;  // dispatch distance of that brick
    // the x-far voxels of a brick: LAYOUT 0 lanes 16..31 of the warps with wx = 1, LAYOUT 1 lanes 24..31 of every warp
    const uint32_t src_warp = LAYOUT == 0 ? (warp | 2u) : warp, src_idx = LAYOUT == 0 ? (lane & 15u) : (lane & 7u);
    const bool publishes = LAYOUT == 0 ? (lane >= 16u && (warp & 2u)) : ((lane >> 3) == 3u);
    if (tile > 0) {
        nseed = carry;  // the tile this block has just finished: its x-far voxel with the same (y, z run)
    } else if (tile_slot && brick >= back) {
        nseed = __ldcg(tile_slot + ((size_t)(brick - back) * RUN_WARPS + src_warp) * 16u + src_idx);
        // a straggler: the brick twice as far back has certainly finished (still a good radius)
        if (nseed == 0xffffffffu && brick >= 2u * back)
            nseed = __ldcg(tile_slot + ((size_t)(brick - 2u * back) * RUN_WARPS + src_warp) * 16u + src_idx);
    }
#if defined(M2S_STATS_BUILD) && !defined(M2S_STATS_HEAVY)
    if (tile_slot && bvh.stats && lane == 0 && nseed == 0xffffffffu) atomicAdd(bvh.stats + 3, 1ull);
#endif
    if (__any_sync(full, nseed >= bvh.nt)) {
        // no neighbour result (first brick planes of a launch): one greedy descent for a voxel in the middle of
        // the tile, the same on every lane (uniform loads, no divergence); its triangle seeds the lanes without one
        // (an exact uniform search for that voxel instead measured no better: profiles/r2d, r2e)
        const f3 pc = {__shfl_sync(full, p0.x, 13), __shfl_sync(full, p0.y, 13), __shfl_sync(full, p0.z, 13)};
        const uint32_t gl = greedy_leaf(bvh, pc);
        if (nseed >= bvh.nt) nseed = gl;
    }
    {
        const bool degen = (bvh.tri_id[nseed] & TRI_DEGEN_BIT) != 0u;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float d2 = INFINITY, p2 = INFINITY;
            uint32_t sl = 0u;
            if (valid[i]) {
                const f3 pi = {p0.x, p0.y, i == 0 ? p0.z : cell_center(g.fz, g.sz, z0 + i)};
                bool neg = false;
                d2 = exact_d2_sign<NORMAL>(bvh, nseed, degen, pi, &neg);
                sl = nseed | (NORMAL && neg ? RUN_NEG_BIT : 0u);
                if (NORMAL && !neg) p2 = d2;
                if (NORMAL) nan |= !(d2 == d2);
            }
            best[lane + 32u * i] = pack_best(d2, sl);
            if (NORMAL) pos[lane + 32u * i] = __float_as_uint(p2);
            // voxels outside the grid never want a child or a triangle
            bnd[i] = valid[i] ? bound_of(d2) : -1.0f;
        }
    }
    __syncwarp();
    auto warp_max_b = [&]() {
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) m = fmaxf(m, bnd[i]);
        return __uint_as_float(__reduce_max_sync(full, __float_as_uint(m)));
    };
    float max_b = warp_max_b();

    int qn = 0, sp = 0;  // warp-uniform
    [[maybe_unused]] uint32_t n_nodes = 0, n_leaves = 0;

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
#ifdef M2S_STATS_ITEMS  // development: the "leaves" counter counts exact (triangle, voxel) evaluations instead
        n_leaves += (uint32_t)min(nb * 32, qn);
#endif
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
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (valid[i]) bnd[i] = bound_of(__uint_as_float((unsigned)(best[lane + 32u * i] >> 32)));
        __syncwarp();
        if ((int)lane < rem) queue[lane] = keep;
        qn = rem;
        __syncwarp();
        max_b = warp_max_b();
    };
    // squared lower bounds of this lane's voxels against the two children of a node: low half of every packed
    // operation = left child, high half = right child
    auto pair_bounds = [&](const float4* nd, float2 (&dd)[V], uint32_t& lref, uint32_t& rref) {
        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
        const float4 q4 = ldg4(nd + 4), q5 = ldg4(nd + 5), q6 = ldg4(nd + 6), q7 = ldg4(nd + 7);
        lref = __float_as_uint(q1.z);
        rref = __float_as_uint(q1.w);
        const float2 m1 = make_float2(-1.0f, -1.0f);
        const float2 dx = __ffma2_rn(f2lo(q0), m1, make_float2(p0.x, p0.x));
        const float2 dy = __ffma2_rn(f2hi(q0), m1, make_float2(p0.y, p0.y));
        const float2 dz = __ffma2_rn(f2lo(q1), m1, make_float2(p0.z, p0.z));
        const float2 tu = __ffma2_rn(dz, f2lo(q3), __ffma2_rn(dy, f2hi(q2), __fmul2_rn(dx, f2lo(q2))));
        const float2 tv = __ffma2_rn(dz, f2lo(q5), __ffma2_rn(dy, f2hi(q4), __fmul2_rn(dx, f2lo(q4))));
        const float2 tw = __ffma2_rn(dz, f2lo(q7), __ffma2_rn(dy, f2hi(q6), __fmul2_rn(dx, f2lo(q6))));
        const float2 eu = f2hi(q3), ev = f2hi(q5), ew = f2hi(q7);
        dd[0] = sumsq2(excess2(tu, eu), excess2(tv, ev), excess2(tw, ew));
#pragma unroll
        for (int i = 1; i < V; ++i) {
            const float2 s2 = make_float2(step[i], step[i]);
            dd[i] = sumsq2(excess2(__ffma2_rn(s2, f2lo(q3), tu), eu), excess2(__ffma2_rn(s2, f2lo(q5), tv), ev),
                           excess2(__ffma2_rn(s2, f2lo(q7), tw), ew));
        }
    };
    uint32_t cur = 0u;  // the root: always an internal node; leaves are consumed at their parent
    for (;;) {
        if (qn >= RUN_FLUSH_AT) flush(false);  // here, where the loop-carried state merges anyway
        PKT_COUNT(n_nodes);
        float2 dd[V];  // squared lower bounds of voxel i: (left child, right child)
        uint32_t lref, rref;
        pair_bounds(bvh.nodes_il + NODE_F4 * (size_t)cur, dd, lref, rref);  // warp-uniform address
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
        if ((lref | rref) & LEAF_BIT) {
            if (lref & LEAF_BIT) {
                if (bl) {
                    enqueue(wl, (lref & LEAF_INDEX_MASK) | ((lref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT_LEAF(n_leaves);
                }
                bl = 0u;
            }
            if (rref & LEAF_BIT) {
                if (br) {
                    enqueue(wr, (rref & LEAF_INDEX_MASK) | ((rref & LEAF_DEGEN_BIT) ? TRI_DEGEN_BIT : 0u));
                    PKT_COUNT_LEAF(n_leaves);
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
            if (sp < PKT_STACK) {  // depth of a Karras tree over 48-bit keys + index tie-break bits < PKT_STACK
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

    float best2[V];
    uint32_t slot[V];  // NORMAL: | RUN_NEG_BIT if that triangle sees the voxel from behind
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const unsigned long long v = best[lane + 32u * i];
        best2[i] = __uint_as_float((unsigned)(v >> 32));
        slot[i] = (uint32_t)v;
    }

    // publish the x-far voxels' nearest triangles for the bricks further in x
    if (tile_slot && publishes) {
        // the middle voxel of the run (the first one where the run is cut by the grid's end)
        if (valid[0])
            __stcg(tile_slot + ((size_t)brick * RUN_WARPS + warp) * 16u + src_idx,
                   (valid[V / 2] ? slot[V / 2] : slot[0]) & ~(NORMAL ? RUN_NEG_BIT : 0u));
    }
    if (TILES > 1) {
        // ... and hand them to the next tile of this block: lane L continues from lane 16 + (L & 15), the voxel of
        // the x-far half with the same (y, z run), 1 or 2 cells away
        const uint32_t mine =
            valid[0] ? ((valid[V / 2] ? slot[V / 2] : slot[0]) & ~(NORMAL ? RUN_NEG_BIT : 0u)) : 0xffffffffu;
        carry = __shfl_sync(full, mine, 16 + (int)(lane & 15u));
    }

    // sqrt is monotone: min sqrt = sqrt min
    float res[V];
#pragma unroll
    for (int i = 0; i < V; ++i) res[i] = __fsqrt_rn(best2[i]);
    if (NORMAL) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (slot[i] & RUN_NEG_BIT) {
                // lib.rs:242-254: an approximately equal positive distance beats the negative one
                const float dp = __fsqrt_rn(__uint_as_float(pos[lane + 32u * (NORMAL ? i : 0)]));
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
    if (V == 4 && valid[V - 1] && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[V > 2 ? 2 : 0], res[V > 3 ? 3 : 0]);
    } else if (V == 2 && valid[1] && (reinterpret_cast<uintptr_t>(o) & 7u) == 0) {
        *reinterpret_cast<float2*>(o) = make_float2(res[0], res[1]);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (valid[i]) o[i] = res[i];
    }
#ifdef M2S_STATS_BUILD
    if (bvh.stats && lane == 0) {
#ifdef M2S_STATS_HEAVY  // development: how heavy is the tail? node visits in / number of tiles above M2S_STATS_HEAVY visits, max
        if (n_nodes > M2S_STATS_HEAVY) {
            atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
            atomicAdd(bvh.stats + 1, 1ull);
        }
        atomicMax(bvh.stats + 3, (unsigned long long)n_nodes);
#else
        atomicAdd(bvh.stats + 0, (unsigned long long)n_nodes);
        atomicAdd(bvh.stats + 1, (unsigned long long)n_leaves);
#endif
        atomicAdd(bvh.stats + 2, 1ull);
    }
#endif
    __syncwarp();  // the shared arrays are reused by the next tile
    }  // tiles of this block
    signal_done();
    if (overflow) atomicExch(&st->stack_overflow, 1);
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

__global__ void __launch_bounds__(256) k_fill(float* __restrict__ out, uint64_t n, float v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = v;
}

// after a fill of a host-visible destination: every plane is complete
__global__ void k_flags_all(volatile uint32_t* flag, uint32_t planes, uint32_t epoch) {
    __threadfence_system();
    for (uint32_t i = threadIdx.x; i < planes; i += blockDim.x) flag[i] = epoch;
}

}  // namespace

#define CK(x)                               \
    do {                                    \
        cudaError_t e__ = (x);              \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

static RowCtx make_row_ctx(const GridParams& g) {
    RowCtx c;
    c.first[0] = g.fx; c.first[1] = g.fy; c.first[2] = g.fz;
    c.size[0] = g.sx; c.size[1] = g.sy; c.size[2] = g.sz;
    c.n[0] = g.nx; c.n[1] = g.ny; c.n[2] = g.nz;
    c.x0 = g.x0; c.x1 = g.x1;
    return c;
}

// Row parity bitmaps for the slab [g.x0, g.x1) (all X rows; the Y and Z rows of the slab's planes).
cudaError_t launch_grid_rows(Device& d, const float4* rec, uint32_t nt, const GridParams& g, RowBits* rb,
                             cudaStream_t s) {
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
    if (nt == 0) return cudaSuccess;
    CK(d.big_list.ensure((size_t)nt * 3 * 4));
    CK(d.big_count.ensure(4));
    CK(cudaMemsetAsync(d.big_count.p, 0, 4, s));
    const RowCtx c = make_row_ctx(g);
    k_rows_small<<<blocks_for(nt, 256), 256, 0, s>>>(rec, nt, c, rb->bits[0], rb->bits[1], rb->bits[2],
                                                     d.big_list.as<uint32_t>(), d.big_count.as<uint32_t>());
    k_rows_big<<<d.sm_count * 4, 256, 0, s>>>(rec, c, rb->bits[0], rb->bits[1], rb->bits[2],
                                              d.big_list.as<uint32_t>(), d.big_count.as<uint32_t>());
    for (int a = 0; a < 3; ++a)
        k_rows_scan<<<blocks_for(rb->rows[a], 256), 256, 0, s>>>(rb->bits[a], rb->rows[a], rb->words[a]);
    d.launches += 5;
    return cudaGetLastError();
}

float grid_magnitude(const GridParams& g) {
    float mag = 0.0f;
    const float f[3] = {g.fx, g.fy, g.fz}, sz[3] = {g.sx, g.sy, g.sz};
    const uint32_t n[3] = {g.nx, g.ny, g.nz};
    for (int i = 0; i < 3; ++i) {
        mag = fmaxf(mag, fabsf(f[i]));
        mag = fmaxf(mag, fabsf(f[i] + (float)n[i] * sz[i]));
    }
    return mag;
}

uint32_t grid_brick_planes(const GridParams& g) { return cdiv(g.x1 - g.x0, (uint32_t)BX); }

// The distance kernel over the slab [g.x0, g.x1). progress (optional): completion flags per brick plane.
template <int V, int LAYOUT>
static cudaError_t launch_grid_nearest_v(Device& d, MeshDev& m, const GridParams& g, int mode, const RowBits* rb,
                                         float* d_out, const Progress* progress) {
    cudaStream_t s = d.stream;
    constexpr uint32_t BZR = 4u * V;
    const uint64_t nrun = (uint64_t)cdiv(g.x1 - g.x0, BX) * cdiv(g.ny, BY) * cdiv(g.nz, BZR);
    if (nrun == 0) return cudaSuccess;
    if (nrun * RUN_WARPS > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const unsigned nbr = (unsigned)nrun;
    const float mag = grid_magnitude(g);
    BuildStatus* st = d.call_status.as<BuildStatus>();
    CK(d.tile_slot.ensure((size_t)nbr * RUN_WARPS * 16 * 4));
    CK(cudaMemsetAsync(d.tile_slot.p, 0xff, (size_t)nbr * RUN_WARPS * 16 * 4, s));
    uint32_t* tile_slot = d.tile_slot.as<uint32_t>();
    const uint32_t *b0 = rb ? rb->bits[0] : nullptr, *b1 = rb ? rb->bits[1] : nullptr, *b2 = rb ? rb->bits[2] : nullptr;
    // seeds come from the brick `planes` steps back in x: far enough in dispatch order to have finished
    // (about 1.25 x the resident blocks), at most 4 steps (16 cells)
    const uint32_t plane_bricks = cdiv(g.ny, BY) * cdiv(g.nz, BZR);
    const uint32_t resident = (uint32_t)d.sm_count * RUN_SEED_BLOCKS;
    const uint32_t planes = std::min(4u, std::max(1u, cdiv(resident * 5u / 4u, plane_bricks)));
    CK(launch_nodes_interleave(d, m, mag, false));  // node frames in units of S = 2^k >= 4 x the largest |coordinate|
    Progress pr{nullptr, nullptr, 0u};
    if (progress) pr = *progress;
    Bvh bvh = m.bvh;
#ifdef M2S_STATS_BUILD
    bvh.stats = d.want_stats ? d.stats.as<unsigned long long>() : nullptr;
#endif
    auto launch = [&](auto tiles) {
        constexpr int T = decltype(tiles)::value;
        const unsigned blocks = nbr * (RUN_WARPS / T);
        if (rb)
            k_grid_nearest_run<RUN_SIGN_RAYCAST, V, LAYOUT, T><<<blocks, 32, 0, s>>>(bvh, g, mag, b0, b1, b2, d_out, st,
                                                                                  tile_slot, planes, pr);
        else if (mode == MODE_NORMAL)
            k_grid_nearest_run<RUN_SIGN_NORMAL, V, LAYOUT, T><<<blocks, 32, 0, s>>>(bvh, g, mag, b0, b1, b2, d_out, st,
                                                                                 tile_slot, planes, pr);
        else
            k_grid_nearest_run<RUN_SIGN_NONE, V, LAYOUT, T><<<blocks, 32, 0, s>>>(bvh, g, mag, b0, b1, b2, d_out, st,
                                                                               tile_slot, planes, pr);
    };
    // blocks of two tiles only where there are enough of them: on a small grid the longer blocks lengthen the tail
    const uint64_t resident_blocks = (uint64_t)d.sm_count * RUN_MIN_BLOCKS * 4;
    if (LAYOUT == 0 && RUN_TILES_X == 2 && nrun * 2 >= (uint64_t)RUN_TILES_MIN_WAVES * resident_blocks)
        launch(std::integral_constant<int, LAYOUT == 0 ? 2 : 1>{});
    else
        launch(std::integral_constant<int, 1>{});
    d.launches++;
    return cudaGetLastError();
}

// Voxels per lane (the run along z). 4 amortises a node visit over 128 voxels instead of 64 but needs more registers
// (24 instead of 28 resident warps) and a tile twice as long in z. Measured over mesh sizes, grid sizes and cell shapes
// (profiles/r2m_run_length.md): it wins on large grids that are fine relative to the mesh and whose cells are thin in
// z - 0.65x .. 0.97x the kernel time at 256^3 for 5k .. 100k triangles, 0.91x on C5 - and loses 3 .. 20 % on cubic
// cells, on small grids and where the mesh is as fine as the grid. Return value: V (+ 16 for lane layout 1).
uint32_t grid_run_length(const Device& d, const MeshDev& m, const GridParams& g) {
    if (d.run_v != 0u) return d.run_v;  // M2S_OPT_RUN_LENGTH: 2, 4 (layout 0), 18, 20 (layout 1)
    const double cells = (double)g.nx * g.ny * g.nz;
    const double sx = fabs((double)g.sx), sy = fabs((double)g.sy), sz = fabs((double)g.sz);
    const bool big_grid = cells >= 8.0e6;
    const bool fine_grid = cells >= 64.0 * (double)m.nt;                   // voxels per triangle
    const bool thin_z = 16.0 * sz <= 1.5 * std::max(2.0 * sx, 4.0 * sy);   // the 2 x 4 x 16 tile stays compact
    if (big_grid && fine_grid && thin_z) return 4u;
    // cells about as wide as deep and a mesh about as fine as the grid: the 4 x 4 x 4 warp tile of layout 1 is the
    // compact one (0.94 .. 0.96x at 100k .. 1M triangles on cubic cells; slower everywhere else)
    if (!thin_z && !fine_grid) return 18u;
    return 2u;
}

// The distance kernel over the slab [g.x0, g.x1). progress (optional): completion flags per brick plane.
cudaError_t launch_grid_nearest(Device& d, MeshDev& m, const GridParams& g, int mode, const RowBits* rb, float* d_out,
                                const Progress* progress) {
    switch (grid_run_length(d, m, g)) {
        case 4u: return launch_grid_nearest_v<4, 0>(d, m, g, mode, rb, d_out, progress);
        case 18u: return launch_grid_nearest_v<2, 1>(d, m, g, mode, rb, d_out, progress);
        case 20u: return launch_grid_nearest_v<4, 1>(d, m, g, mode, rb, d_out, progress);
        default: return launch_grid_nearest_v<2, 0>(d, m, g, mode, rb, d_out, progress);
    }
}

cudaError_t launch_fill(Device& d, float* d_out, uint64_t n, float value, const Progress* progress, uint32_t planes) {
    if (n == 0) return cudaSuccess;
    const unsigned nb = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)d.sm_count * 16);
    k_fill<<<nb, 256, 0, d.stream>>>(d_out, n, value);
    d.launches++;
    if (progress && progress->flag) {
        k_flags_all<<<1, 256, 0, d.stream>>>(progress->flag, planes, progress->epoch);
        d.launches++;
    }
    return cudaGetLastError();
}

}  // namespace m2s
