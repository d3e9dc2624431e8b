// LBVH build on the GPU: triangle records -> 48-bit Morton keys -> radix sort -> Karras hierarchy over
// single-triangle leaves -> bottom-up refit -> oriented boxes per child slot.
//
// Replaces the reference's per-call acceleration-structure builds: bvh::Bvh::build_par
// (mesh_to_sdf/src/generate/grid.rs:95-111, generic/bvh.rs:62-74, generic/rtree_bvh.rs:108-116) and
// rstar::RTree::bulk_load (generic/rtree.rs:111, generic/rtree_bvh.rs:118). Leaf boxes are the
// reference's padded triangle boxes (geo::triangle_bounding_box, src/geo.rs:4-22).
#include <algorithm>

#include <cuda/atomic>

#include "m2s_geom.cuh"
#include "m2s_internal.h"
#include "m2s_sort.cuh"

namespace m2s {

cudaError_t DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) {
        cudaError_t e = cudaFree(p);
        p = nullptr;
        cap = 0;
        if (e != cudaSuccess) return e;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        return e;
    }
    cap = want;
    return cudaSuccess;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

cudaError_t PinBuf::ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) {
        cudaError_t e = cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        if (e != cudaSuccess) return e;
    }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {
        p = nullptr;
        return e;
    }
    cap = want;
    return cudaSuccess;
}
void PinBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

void MeshDev::release() {
    DevBuf* bufs[] = {&rec_sorted, &tri_id_sorted, &nodes, &nodes_il, &boxes, &status, &node_range,
                      &bin_offsets, &bin_cursor, &bin_items, &bin_big, &bin_meta};
    for (DevBuf* b : bufs) b->release();
    bvh = Bvh{};
    nv = nt = 0;
    nodes_il_mag = -1.0f;
    bins_built = false;
}

namespace {

// order-preserving float <-> int mapping for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float scene_mag(const BuildStatus* __restrict__ st) {
    float m = 0.0f;
    for (int i = 0; i < 3; ++i) {
        const float lo = ord2f(st->lo[i]), hi = ord2f(st->hi[i]);
        if (lo <= hi) m = fmaxf(m, fmaxf(fabsf(lo), fabsf(hi)));
    }
    return m;
}

// The mesh's status block: error flags of the mesh (bad index, non-finite vertex), degenerate count, bounds.
__global__ void k_mesh_status_init(BuildStatus* st) {
    if (threadIdx.x == 0) {
        st->bad_index = 0;
        st->nonfinite = 0;
        st->stack_overflow = 0;
        st->nan_distance = 0;
        st->n_degenerate = 0;
        for (int i = 0; i < 3; ++i) {
            st->lo[i] = f2ord(INFINITY);
            st->hi[i] = f2ord(-INFINITY);
        }
        for (int i = 0; i < 5; ++i) st->pad[i] = 0;
    }
}

// The call's status block starts from the mesh's (bounds, mesh errors); the sticky error flags of earlier device
// calls survive until the host has read them (clear_errors).
__global__ void k_call_status_init(BuildStatus* st, const BuildStatus* mesh, int clear_errors) {
    if (threadIdx.x == 0) {
        const int bad = mesh ? mesh->bad_index : 0, nonf = mesh ? mesh->nonfinite : 0;
        if (clear_errors) {
            st->bad_index = bad;
            st->nonfinite = nonf;
            st->stack_overflow = 0;
            st->nan_distance = 0;
        } else {
            st->bad_index |= bad;
            st->nonfinite |= nonf;
        }
        st->n_degenerate = mesh ? mesh->n_degenerate : 0;
        for (int i = 0; i < 3; ++i) {
            st->lo[i] = mesh ? mesh->lo[i] : f2ord(INFINITY);
            st->hi[i] = mesh ? mesh->hi[i] : f2ord(-INFINITY);
        }
        for (int i = 0; i < 5; ++i) st->pad[i] = 0;
    }
}

// K1: gather vertices by index, write the 48-byte record (original order), the padded AABB
// (geo.rs:4-22) and reduce the scene bounds. One thread per triangle; float4 stores are coalesced.
__global__ void __launch_bounds__(256)
k_tri_setup(const float* __restrict__ verts, uint32_t nv, const uint32_t* __restrict__ tris, uint32_t nt,
            float4* __restrict__ rec, float4* __restrict__ tri_lo, float4* __restrict__ tri_hi,
            BuildStatus* __restrict__ st) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (t < nt) {
        uint32_t i0 = tris[3 * t], i1 = tris[3 * t + 1], i2 = tris[3 * t + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) {
            atomicExch(&st->bad_index, 1);
            i0 = i1 = i2 = 0;  // keep going with harmless data; the host reports M2S_EINDEX
        }
        const f3 a = {verts[3 * i0], verts[3 * i0 + 1], verts[3 * i0 + 2]};
        const f3 b = {verts[3 * i1], verts[3 * i1 + 1], verts[3 * i1 + 2]};
        const f3 c = {verts[3 * i2], verts[3 * i2 + 1], verts[3 * i2 + 2]};
        const f3 n = v_cross(v_sub(b, a), v_sub(c, a));
        rec[3 * t + 0] = make_float4(a.x, a.y, a.z, b.x);
        rec[3 * t + 1] = make_float4(b.y, b.z, c.x, c.y);
        rec[3 * t + 2] = make_float4(c.z, n.x, n.y, n.z);
        const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(b.x) && isfinite(b.y) &&
                         isfinite(b.z) && isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
        if (!fin) atomicExch(&st->nonfinite, 1);
        const bool degen = v_eq(a, b) || v_eq(b, c) || v_eq(a, c);
        if (degen) atomicAdd(&st->n_degenerate, 1);
        const float EPS = 0.0001f;  // geo.rs:5
        lo[0] = fsub(fminf(a.x, fminf(b.x, c.x)), EPS);
        lo[1] = fsub(fminf(a.y, fminf(b.y, c.y)), EPS);
        lo[2] = fsub(fminf(a.z, fminf(b.z, c.z)), EPS);
        hi[0] = fadd(fmaxf(a.x, fmaxf(b.x, c.x)), EPS);
        hi[1] = fadd(fmaxf(a.y, fmaxf(b.y, c.y)), EPS);
        hi[2] = fadd(fmaxf(a.z, fmaxf(b.z, c.z)), EPS);
        tri_lo[t] = make_float4(lo[0], lo[1], lo[2], degen ? 1.0f : 0.0f);
        tri_hi[t] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        if (!fin) {
            for (int i = 0; i < 3; ++i) {
                lo[i] = INFINITY;
                hi[i] = -INFINITY;
            }
        }
    }
    // warp reduce, one atomic per warp and axis
    for (int i = 0; i < 3; ++i) {
        float l = lo[i], h = hi[i];
        for (int o = 16; o; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&st->lo[i], f2ord(l));
            atomicMax(&st->hi[i], f2ord(h));
        }
    }
}

constexpr int MORTON_BITS = 48;

__device__ __forceinline__ uint64_t spread21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__device__ __forceinline__ uint64_t morton63(float x, float y, float z, const BuildStatus* st) {
    const float lo[3] = {ord2f(st->lo[0]), ord2f(st->lo[1]), ord2f(st->lo[2])};
    const float hi[3] = {ord2f(st->hi[0]), ord2f(st->hi[1]), ord2f(st->hi[2])};
    const float p[3] = {x, y, z};
    uint32_t q[3];
    for (int i = 0; i < 3; ++i) {
        float ext = hi[i] - lo[i];
        float u = ext > 0.0f ? (p[i] - lo[i]) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        if (!(u == u)) u = 0.0f;
        q[i] = min((uint32_t)(u * 2097152.0f), 2097151u);
    }
    // 16 bits per axis are kept (65 536 cells per axis; equal keys are ordered by index, see delta()): 48-bit keys
    // sort in 6 radix passes instead of 8
    return ((spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2])) >> (63 - MORTON_BITS);
}

// K2: Morton key of the centre of the padded box.
__global__ void __launch_bounds__(256)
k_tri_morton(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi, uint32_t nt,
             const BuildStatus* __restrict__ st, uint64_t* __restrict__ keys) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const float4 l = tri_lo[t], h = tri_hi[t];
    keys[t] = morton63(0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z), st);
}

__global__ void __launch_bounds__(256)
k_point_morton(const float* __restrict__ q, uint32_t nq, const BuildStatus* __restrict__ st,
               uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    keys[i] = morton63(q[3 * i], q[3 * i + 1], q[3 * i + 2], st);
}

// query bounds (so that query Morton codes use a box that contains the queries) + finite check
__global__ void __launch_bounds__(256)
k_point_bounds(const float* __restrict__ q, uint32_t nq, BuildStatus* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < nq) {
        const float x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            lo[0] = hi[0] = x;
            lo[1] = hi[1] = y;
            lo[2] = hi[2] = z;
        } else {
            atomicExch(&st->nonfinite, 1);
        }
    }
    for (int k = 0; k < 3; ++k) {
        float l = lo[k], h = hi[k];
        for (int o = 16; o; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&st->lo[k], f2ord(l));
            atomicMax(&st->hi[k], f2ord(h));
        }
    }
}

__global__ void __launch_bounds__(256)
k_point_gather(const float* __restrict__ q, const uint32_t* __restrict__ perm, uint32_t nq,
               float4* __restrict__ q_sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint32_t s = perm[i];
    q_sorted[i] = make_float4(q[3 * s], q[3 * s + 1], q[3 * s + 2], __uint_as_float(s));
}

// ---------------------------------------------------------------------------------------------------
// Oriented bounds. With axis-aligned boxes alone a far-field query visits hundreds of nodes: the box of
// a tilted, nearly flat patch is loose by about a third of its size TOWARDS the query (first order)
// while the distance to neighbouring patches only grows like x^2 / 2D. An oriented box whose first
// axis is the patch's mean normal is tight to within the patch's sag in that direction, and fits an
// elongated patch laterally (a disc / cylinder does not: measured 99 -> see DESIGN.md). The distance
// from p to anything inside {c + a u + b v + g w : |a|<=eu, |b|<=ev, |g|<=ew} is at least
// hypot(max(0,|(p-c).u|-eu), max(0,|(p-c).v|-ev), max(0,|(p-c).w|-ew)) for orthonormal (u, v, w).
// ---------------------------------------------------------------------------------------------------
struct Frame {
    float3 u, v, w;
};

__device__ __forceinline__ float3 f3sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float f3dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 f3cross(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ bool f3normalize(float3& a) {
    const float l2 = f3dot(a, a);
    if (!(l2 > 1e-30f) || !isfinite(l2)) return false;
    const float inv = rsqrtf(l2);
    a = make_float3(a.x * inv, a.y * inv, a.z * inv);
    return true;
}

// Orthonormal frame whose first axis is n (normalised here); false if n is degenerate.
__device__ __forceinline__ bool frame_from_normal(float3 n, Frame* F) {
    if (!f3normalize(n)) return false;
    const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    const float3 e = ax <= ay && ax <= az ? make_float3(1.f, 0.f, 0.f)
                                          : (ay <= az ? make_float3(0.f, 1.f, 0.f) : make_float3(0.f, 0.f, 1.f));
    float3 t1 = f3cross(n, e);
    if (!f3normalize(t1)) return false;
    float3 t2 = f3cross(n, t1);
    f3normalize(t2);
    F->u = n; F->v = t1; F->w = t2;
    return true;
}

// Rotate (v, w) about u to the principal axes of the 2-D covariance (saa, sab, sbb already centred).
__device__ __forceinline__ void frame_align(Frame* F, float caa, float cab, float cbb) {
    const float theta = 0.5f * atan2f(2.0f * cab, caa - cbb);
    float sn, cs;
    sincosf(theta, &sn, &cs);
    float3 v = make_float3(cs * F->v.x + sn * F->w.x, cs * F->v.y + sn * F->w.y, cs * F->v.z + sn * F->w.z);
    f3normalize(v);
    float3 w = f3cross(F->u, v);
    f3normalize(w);
    F->v = f3cross(w, F->u);  // re-orthogonalise
    f3normalize(F->v);
    F->w = w;
}

struct Extent {
    float lo[3], hi[3];
    __device__ __forceinline__ void reset() {
        for (int i = 0; i < 3; ++i) { lo[i] = INFINITY; hi[i] = -INFINITY; }
    }
    __device__ __forceinline__ void add(const Frame& F, float3 d) {
        const float a = f3dot(d, F.u), b = f3dot(d, F.v), g = f3dot(d, F.w);
        lo[0] = fminf(lo[0], a); hi[0] = fmaxf(hi[0], a);
        lo[1] = fminf(lo[1], b); hi[1] = fmaxf(hi[1], b);
        lo[2] = fminf(lo[2], g); hi[2] = fmaxf(hi[2], g);
    }
};

// Writes the 4 x float4 oriented-box record: (centre.xyz, w0) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew).
// Extents are inflated for the rounding of this computation and of the query-side evaluation.
__device__ __forceinline__ void write_obb(float4* out, float3 origin, const Frame& F, const Extent& E, float mag,
                                          float w0) {
    const float ma = 0.5f * (E.lo[0] + E.hi[0]), mb = 0.5f * (E.lo[1] + E.hi[1]), mg = 0.5f * (E.lo[2] + E.hi[2]);
    const float3 c = make_float3(origin.x + ma * F.u.x + mb * F.v.x + mg * F.w.x,
                                 origin.y + ma * F.u.y + mb * F.v.y + mg * F.w.y,
                                 origin.z + ma * F.u.z + mb * F.v.z + mg * F.w.z);
    const float slack = 6.0e-6f * mag;
    const float eu = 0.5f * (E.hi[0] - E.lo[0]) * 1.0001f + slack;
    const float ev = 0.5f * (E.hi[1] - E.lo[1]) * 1.0001f + slack;
    const float ew = 0.5f * (E.hi[2] - E.lo[2]) * 1.0001f + slack;
    out[0] = make_float4(c.x, c.y, c.z, w0);
    out[1] = make_float4(F.u.x, F.u.y, F.u.z, eu);
    out[2] = make_float4(F.v.x, F.v.y, F.v.z, ev);
    out[3] = make_float4(F.w.x, F.w.y, F.w.z, ew);
}

// K4a: permute the records into leaf (sorted) order; per-triangle oriented box (normal, longest edge).
__global__ void __launch_bounds__(256)
k_tri_permute(const float4* __restrict__ rec, const float4* __restrict__ tri_lo,
              const uint32_t* __restrict__ order, uint32_t nt, const BuildStatus* __restrict__ st,
              float4* __restrict__ rec_sorted, float4* __restrict__ tobb, uint32_t* __restrict__ tri_id_sorted) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nt) return;
    const uint32_t t = order[j];
    const float4 r0 = rec[3 * t + 0], r1 = rec[3 * t + 1], r2 = rec[3 * t + 2];
    rec_sorted[3 * j + 0] = r0;
    rec_sorted[3 * j + 1] = r1;
    rec_sorted[3 * j + 2] = r2;
    tri_id_sorted[j] = t | (tri_lo[t].w != 0.0f ? TRI_DEGEN_BIT : 0u);
    const float3 a = make_float3(r0.x, r0.y, r0.z), b = make_float3(r0.w, r1.x, r1.y), c = make_float3(r1.z, r1.w, r2.x);
    const float mag = scene_mag(st);
    Frame F;
    Extent E;
    E.reset();
    const float3 ab = f3sub(b, a), bc = f3sub(c, b), ca = f3sub(a, c);
    bool ok = frame_from_normal(make_float3(r2.y, r2.z, r2.w), &F);
    if (ok) {
        // second axis along the longest edge
        const float lab = f3dot(ab, ab), lbc = f3dot(bc, bc), lca = f3dot(ca, ca);
        float3 e = lab >= lbc && lab >= lca ? ab : (lbc >= lca ? bc : ca);
        // project on the plane, normalise, rebuild w
        const float k = f3dot(e, F.u);
        e = make_float3(e.x - k * F.u.x, e.y - k * F.u.y, e.z - k * F.u.z);
        if (f3normalize(e)) {
            F.v = e;
            F.w = f3cross(F.u, F.v);
            f3normalize(F.w);
        }
    } else {
        // degenerate triangle: axis-aligned frame (still a valid box around the three points)
        F.u = make_float3(1.f, 0.f, 0.f); F.v = make_float3(0.f, 1.f, 0.f); F.w = make_float3(0.f, 0.f, 1.f);
    }
    E.add(F, make_float3(0.f, 0.f, 0.f));
    E.add(F, ab);
    E.add(F, f3sub(c, a));
    write_obb(tobb + 4 * (size_t)j, a, F, E, mag, 0.0f);
}

// K4d: search node of every internal node: per child either the padded box (for big / strongly curved
// subtrees) or an oriented box fitted to the child's contiguous leaf-order triangle range. A group of G lanes
// per child slot, three strided passes: normal sum -> lateral covariance -> extents. Most slots are tiny (half of
// all internal children have two or three triangles), so the first launch gives every slot 8 lanes and leaves the
// slots with more than SMALL_SLOT_TRIS triangles to a second, warp-per-slot launch over a compacted list.
#ifndef M2S_OBB_MAX_TRIS
#define M2S_OBB_MAX_TRIS 512  // one warp fits a slot serially: 4096 left a 128-iteration tail (build 0.45 -> 0.37 ms, same walk)
#endif
constexpr uint32_t OBB_MAX_TRIS = M2S_OBB_MAX_TRIS;
constexpr uint32_t SMALL_SLOT_TRIS = 32;

// fits child slot `slot` with the G lanes of the calling group (G = 8 or 32; all G lanes call it together)
template <int G>
__device__ __forceinline__ void fit_child_slot(const uint32_t slot, const uint32_t gl, const unsigned gmask, const uint32_t b,
                                               const uint32_t e, const float4 c0, const float4 c1, const uint32_t ref,
                                               const float4* __restrict__ rec_sorted, float4* __restrict__ ch,
                                               const BuildStatus* __restrict__ st, const float obb_bias) {
    const float3 origin = make_float3(0.5f * (c0.x + c1.x), 0.5f * (c0.y + c1.y), 0.5f * (c0.z + c1.z));
    bool use_obb = (e - b) <= OBB_MAX_TRIS;
    Frame F;
    Extent E;
    if (use_obb) {
        float3 ns = make_float3(0.f, 0.f, 0.f);
        for (uint32_t j = b + gl; j < e; j += G) {
            const float4 r2 = rec_sorted[3 * (size_t)j + 2];
            ns.x += r2.y; ns.y += r2.z; ns.z += r2.w;
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            ns.x += __shfl_xor_sync(gmask, ns.x, o);
            ns.y += __shfl_xor_sync(gmask, ns.y, o);
            ns.z += __shfl_xor_sync(gmask, ns.z, o);
        }
        use_obb = frame_from_normal(ns, &F);  // identical on all lanes (xor-butterfly sums are bitwise equal)
    }
    if (use_obb) {
        // lateral covariance of the vertices in the (v, w) plane
        float sa = 0.f, sb = 0.f, saa = 0.f, sab = 0.f, sbb = 0.f, cnt = 0.f;
        for (uint32_t j = b + gl; j < e; j += G) {
            const float4 r0 = rec_sorted[3 * (size_t)j], r1 = rec_sorted[3 * (size_t)j + 1],
                         r2 = rec_sorted[3 * (size_t)j + 2];
            const float3 p3[3] = {make_float3(r0.x, r0.y, r0.z), make_float3(r0.w, r1.x, r1.y),
                                  make_float3(r1.z, r1.w, r2.x)};
            for (int k = 0; k < 3; ++k) {
                const float3 d = f3sub(p3[k], origin);
                const float pa = f3dot(d, F.v), pb = f3dot(d, F.w);
                sa += pa; sb += pb; saa += pa * pa; sab += pa * pb; sbb += pb * pb; cnt += 1.0f;
            }
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            sa += __shfl_xor_sync(gmask, sa, o); sb += __shfl_xor_sync(gmask, sb, o);
            saa += __shfl_xor_sync(gmask, saa, o); sab += __shfl_xor_sync(gmask, sab, o);
            sbb += __shfl_xor_sync(gmask, sbb, o); cnt += __shfl_xor_sync(gmask, cnt, o);
        }
        const float inv = 1.0f / fmaxf(cnt, 1.0f);
        const float ma = sa * inv, mb = sb * inv;
        frame_align(&F, saa * inv - ma * ma, sab * inv - ma * mb, sbb * inv - mb * mb);
        E.reset();
        for (uint32_t j = b + gl; j < e; j += G) {
            const float4 r0 = rec_sorted[3 * (size_t)j], r1 = rec_sorted[3 * (size_t)j + 1],
                         r2 = rec_sorted[3 * (size_t)j + 2];
            E.add(F, f3sub(make_float3(r0.x, r0.y, r0.z), origin));
            E.add(F, f3sub(make_float3(r0.w, r1.x, r1.y), origin));
            E.add(F, f3sub(make_float3(r1.z, r1.w, r2.x), origin));
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1)
            for (int k = 0; k < 3; ++k) {
                E.lo[k] = fminf(E.lo[k], __shfl_xor_sync(gmask, E.lo[k], o));
                E.hi[k] = fmaxf(E.hi[k], __shfl_xor_sync(gmask, E.hi[k], o));
            }
        // keep the oriented box only where it is the smaller volume (top-level, curved subtrees are
        // better served by the axis-aligned box, which is also cheaper to test)
        const float vo = (E.hi[0] - E.lo[0] + 1e-4f) * (E.hi[1] - E.lo[1] + 1e-4f) * (E.hi[2] - E.lo[2] + 1e-4f);
        const float va = (c1.x - c0.x) * (c1.y - c0.y) * (c1.z - c0.z);
        use_obb = vo > 0.0f && isfinite(vo) && vo <= obb_bias * va;  // NaN / overflowed fits fall back to the padded box
    }
    if (gl == 0) {
        if (use_obb) {
            write_obb(ch, origin, F, E, scene_mag(st), __uint_as_float(ref));
        } else {
            // the padded axis-aligned box in the same format (identity frame): one code path, no
            // per-child branch in the traversal
            const float slack = 6.0e-6f * scene_mag(st);  // rounding of the centre and of the query side
            ch[0] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(ref));
            ch[1] = make_float4(1.f, 0.f, 0.f, fmaxf(c1.x - origin.x, origin.x - c0.x) + slack);
            ch[2] = make_float4(0.f, 1.f, 0.f, fmaxf(c1.y - origin.y, origin.y - c0.y) + slack);
            ch[3] = make_float4(0.f, 0.f, 1.f, fmaxf(c1.z - origin.z, origin.z - c0.z) + slack);
        }
    }
    (void)slot;
}

// first launch: 8 lanes per child slot; leaves copy their triangle's box, small subtrees are fitted here, the
// others are listed for k_search_nodes_big
__global__ void __launch_bounds__(256)
k_search_nodes(const float4* __restrict__ rec_sorted, const float4* __restrict__ tobb, uint32_t nt,
               int nleaf, const float4* __restrict__ boxes, float4* __restrict__ nodes,
               const uint2* __restrict__ node_range, const BuildStatus* __restrict__ st, float obb_bias,
               uint32_t* __restrict__ big_list, uint32_t* __restrict__ big_count) {
    constexpr int G = 8;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t slot = tid / G, gl = threadIdx.x & (G - 1);
    const unsigned gmask = 0xffu << ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
    if (slot >= 2u * (uint32_t)(nleaf - 1)) return;  // whole groups leave together (256 % G == 0)
    const float4* bx = boxes + BOX_F4 * (size_t)(slot >> 1) + 2 * (slot & 1u);
    float4* ch = nodes + NODE_F4 * (size_t)(slot >> 1) + CHILD_F4 * (slot & 1u);
    const float4 c0 = bx[0], c1 = bx[1];
    const uint32_t ref = __float_as_uint(c0.w);
    if (ref & LEAF_BIT) {
        // single-triangle leaf: its oriented box was already fitted by k_tri_permute
        if (gl < 4) {
            float4 v = tobb[4 * (size_t)(ref & LEAF_INDEX_MASK) + gl];
            if (gl == 0) v.w = __uint_as_float(ref);
            ch[gl] = v;
        }
        return;
    }
    const uint2 r = node_range[ref];
    const uint32_t b = r.x, e = min(nt, r.y + 1);
    if (e - b > SMALL_SLOT_TRIS) {
        if (gl == 0) big_list[atomicAdd(big_count, 1u)] = slot;
        return;
    }
    fit_child_slot<G>(slot, gl, gmask, b, e, c0, c1, ref, rec_sorted, ch, st, obb_bias);
}

// second launch: one warp per listed slot (subtrees of more than SMALL_SLOT_TRIS triangles)
__global__ void __launch_bounds__(256)
k_search_nodes_big(const float4* __restrict__ rec_sorted, uint32_t nt, const float4* __restrict__ boxes,
                   float4* __restrict__ nodes, const uint2* __restrict__ node_range, const BuildStatus* __restrict__ st,
                   float obb_bias, const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_count) {
    const uint32_t n = *big_count, lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint32_t slot = big_list[i];
        const float4* bx = boxes + BOX_F4 * (size_t)(slot >> 1) + 2 * (slot & 1u);
        float4* ch = nodes + NODE_F4 * (size_t)(slot >> 1) + CHILD_F4 * (slot & 1u);
        const float4 c0 = bx[0], c1 = bx[1];
        const uint32_t ref = __float_as_uint(c0.w);
        const uint2 r = node_range[ref];
        fit_child_slot<32>(slot, lane, 0xffffffffu, r.x, min(nt, r.y + 1), c0, c1, ref, rec_sorted, ch, st, obb_bias);
    }
}

// K4e: the same search nodes with the two children interleaved component by component, so that one
// packed-fp32 instruction (FFMA2 / FMUL2 / FADD2, sm_100a) evaluates the left child in its low half and
// the right child in its high half:
//   q0 = (cL.x, cR.x, cL.y, cR.y)  q1 = (cL.z, cR.z, refL, refR)
//   q2 = (uL.x, uR.x, uL.y, uR.y)  q3 = (uL.z, uR.z, euL, euR)     q4, q5: v     q6, q7: w
// Frames and extents are multiplied by inv_s (a power of two: exact), so the projections come out in units
// of S = 1 / inv_s >= 4 x the largest scene / grid coordinate and max(|t| - e, 0) is ONE saturating add
// (FADD.SAT clamps to [0, 1]; 1 is never reached inside the scene, and clamping a lower bound from above
// keeps it a lower bound).
__global__ void __launch_bounds__(256)
k_nodes_interleave(const float4* __restrict__ nodes, uint32_t n_nodes, float4* __restrict__ il,
                   const BuildStatus* __restrict__ st, float grid_mag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const float inv_s = pair_inv_scale(fmaxf(scene_mag(st), grid_mag));
    const float4* nd = nodes + NODE_F4 * (size_t)i;
    float4* o = il + NODE_F4 * (size_t)i;
    {
        const float4 l = nd[0], r = nd[CHILD_F4];
        o[0] = make_float4(l.x, r.x, l.y, r.y);
        o[1] = make_float4(l.z, r.z, l.w, r.w);
    }
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const float4 l = nd[k], r = nd[CHILD_F4 + k];
        o[2 * k] = make_float4(l.x * inv_s, r.x * inv_s, l.y * inv_s, r.y * inv_s);
        o[2 * k + 1] = make_float4(l.z * inv_s, r.z * inv_s, l.w * inv_s, r.w * inv_s);
    }
}

// ---- ray bins (see RayBins, m2s_internal.h) -----------------------------------------------------------------------
// cell of an in-plane coordinate: monotone in v (subtract, multiply by a positive constant, floor), so
// lo <= v <= hi  =>  cell(lo) <= cell(v) <= cell(hi): a query inside a triangle's padded box is inside its cell range
__device__ __forceinline__ uint32_t raybin_cell(float v, float lo, float inv, uint32_t R) {
    const float c = floorf((v - lo) * inv);
    if (!(c > 0.0f)) return 0u;
    return c >= (float)R ? R - 1u : (uint32_t)c;
}

struct BinRange {
    uint32_t j0, j1, k0, k1;
};

// padded box of leaf-order triangle t in the projection of `axis`: cells [j0, j1] x [k0, k1]
__device__ __forceinline__ BinRange raybin_range(const float4* __restrict__ rec, uint32_t t, int axis,
                                                 const BuildStatus* __restrict__ st, uint32_t R) {
    const float4 r0 = rec[3 * (size_t)t], r1 = rec[3 * (size_t)t + 1], r2 = rec[3 * (size_t)t + 2];
    const float a[3] = {r0.x, r0.y, r0.z}, b[3] = {r0.w, r1.x, r1.y}, c[3] = {r1.z, r1.w, r2.x};
    const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
    const float EPS = 0.0001f;  // geo.rs:5,20-21
    BinRange g;
    {
        const float lo = ord2f(st->lo[iy]), hi = ord2f(st->hi[iy]);
        const float inv = hi > lo ? (float)R / (hi - lo) : 0.0f;
        g.j0 = raybin_cell(fsub(fminf(a[iy], fminf(b[iy], c[iy])), EPS), lo, inv, R);
        g.j1 = raybin_cell(fadd(fmaxf(a[iy], fmaxf(b[iy], c[iy])), EPS), lo, inv, R);
    }
    {
        const float lo = ord2f(st->lo[iz]), hi = ord2f(st->hi[iz]);
        const float inv = hi > lo ? (float)R / (hi - lo) : 0.0f;
        g.k0 = raybin_cell(fsub(fminf(a[iz], fminf(b[iz], c[iz])), EPS), lo, inv, R);
        g.k1 = raybin_cell(fadd(fmaxf(a[iz], fmaxf(b[iz], c[iz])), EPS), lo, inv, R);
    }
    return g;
}

// pass 0: count the items of every cell, list the big triangles; pass 1: scatter the items
template <int PASS>
__global__ void __launch_bounds__(256)
k_raybins(const float4* __restrict__ rec, uint32_t nt, const BuildStatus* __restrict__ st, uint32_t R,
          uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ items,
          uint32_t* __restrict__ big, uint32_t* __restrict__ meta, uint32_t capacity) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    if (PASS == 1 && meta[4] == 0u) return;  // over budget: the queries walk the box tree instead
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        const BinRange g = raybin_range(rec, t, axis, st, R);
        const uint32_t ncell = (g.j1 - g.j0 + 1u) * (g.k1 - g.k0 + 1u);
        if (ncell > RAYBIN_BIG_CELLS) {
            if (PASS == 0) {
                const uint32_t i = atomicAdd(meta + axis, 1u);
                if (i < RAYBIN_MAX_BIG) big[axis * RAYBIN_MAX_BIG + i] = t;
            }
            continue;
        }
        uint32_t* cnt = counts + (size_t)axis * R * R;
        for (uint32_t k = g.k0; k <= g.k1; ++k)
            for (uint32_t j = g.j0; j <= g.j1; ++j) {
                const uint32_t cell = k * R + j;
                const uint32_t i = atomicAdd(cnt + cell, 1u);
                if (PASS == 1) {
                    const uint32_t at = offsets[(size_t)axis * R * R + cell] + i;
                    if (at < capacity) items[at] = t;
                }
            }
        if (PASS == 0) atomicAdd(meta + 3, ncell);
    }
}

// after the count pass: do the totals fit the budgets?
__global__ void k_raybins_verdict(uint32_t* meta, uint32_t capacity) {
    if (threadIdx.x == 0)
        meta[4] = (meta[0] <= RAYBIN_MAX_BIG && meta[1] <= RAYBIN_MAX_BIG && meta[2] <= RAYBIN_MAX_BIG &&
                   meta[3] <= capacity) ? 1u : 0u;
}

// Karras 2012 delta over the leaf keys (leaf l = sorted triangle l; equal keys are ordered by index).
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int nleaf, int i, int j) {
    if (j < 0 || j >= nleaf) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// K4b: one thread per internal node: children + parent links.
__global__ void __launch_bounds__(256)
k_hierarchy(const uint64_t* __restrict__ keys, int nleaf, float4* __restrict__ nodes,
            uint32_t* __restrict__ leaf_parent, uint32_t* __restrict__ node_parent,
            uint2* __restrict__ node_range) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nleaf - 1) return;
    const int d = (delta(keys, nleaf, i, i + 1) - delta(keys, nleaf, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, nleaf, i, i - d);
    int lmax = 2;
    while (delta(keys, nleaf, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, nleaf, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, nleaf, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, nleaf, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    uint32_t left, right;
    if (lo == gamma) {
        left = LEAF_BIT | (uint32_t)gamma;
        leaf_parent[gamma] = ((uint32_t)i << 1);
    } else {
        left = (uint32_t)gamma;
        node_parent[gamma] = ((uint32_t)i << 1);
    }
    if (hi == gamma + 1) {
        right = LEAF_BIT | (uint32_t)(gamma + 1);
        leaf_parent[gamma + 1] = ((uint32_t)i << 1) | 1u;
    } else {
        right = (uint32_t)(gamma + 1);
        node_parent[gamma + 1] = ((uint32_t)i << 1) | 1u;
    }
    // boxes are filled by the refit; store the refs now (degenerate bits are OR-ed in by the refit).
    nodes[BOX_F4 * (size_t)i + 0].w = __uint_as_float(left);
    nodes[BOX_F4 * (size_t)i + 2].w = __uint_as_float(right);
    node_range[i] = make_uint2((uint32_t)lo, (uint32_t)hi);  // leaves covered by node i (inclusive)
    if (i == 0) node_parent[0] = 0xffffffffu;
}

// K4c: the padded boxes of both children of every internal node (Bvh::boxes). Two passes.
//
// Pass 1 (k_refit_windows): a node of a Karras tree covers a contiguous range of leaves, so its box is a range
// union over the leaf boxes. A block loads the boxes of a window of REFIT_WINDOW consecutive leaves into shared
// memory, builds a sparse table over them (table[k][j] = union of leaves j .. j + 2^k - 1) and every internal node
// whose whole range lies inside the window gets both child boxes from two table lookups each: no atomics, no
// dependency between nodes, ~85 % of all nodes.
// Pass 2 (k_refit_climb): the nodes whose range straddles a window (the ancestors of the window boundaries) are
// refitted bottom-up as before - one thread per element of the frontier below them (a leaf, or a pass-1 node whose
// parent is not a pass-1 node), the second thread to reach a node continues upwards - over ~8x fewer nodes.
constexpr int REFIT_WINDOW = 256, REFIT_LEVELS = 9;  // 2^(REFIT_LEVELS - 1) == REFIT_WINDOW

__device__ __forceinline__ bool refit_in_window(uint2 range) {
    return (range.x / (uint32_t)REFIT_WINDOW) == (range.y / (uint32_t)REFIT_WINDOW);
}

struct RefitBox {
    float lo[3], hi[3];
};

__global__ void __launch_bounds__(REFIT_WINDOW)
k_refit_windows(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi,
                const uint32_t* __restrict__ order, int nleaf, float4* __restrict__ nodes,
                const uint2* __restrict__ node_range) {
    extern __shared__ float s_tab[];  // [REFIT_LEVELS][REFIT_WINDOW][6]
    __shared__ unsigned char s_degen[REFIT_WINDOW];
    const int w0 = blockIdx.x * REFIT_WINDOW, j = threadIdx.x;
    auto tab = [&](int k, int i) { return s_tab + ((size_t)k * REFIT_WINDOW + i) * 6; };
    {
        float* e = tab(0, j);
        const int l = w0 + j;
        if (l < nleaf) {
            const uint32_t t = order[l];
            const float4 a = tri_lo[t], b = tri_hi[t];
            e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = b.x; e[4] = b.y; e[5] = b.z;
            s_degen[j] = a.w != 0.0f ? 1 : 0;
        } else {
            e[0] = e[1] = e[2] = INFINITY;
            e[3] = e[4] = e[5] = -INFINITY;
            s_degen[j] = 0;
        }
    }
    __syncthreads();
    for (int k = 1; k < REFIT_LEVELS; ++k) {
        const int half = 1 << (k - 1);
        if (j + (1 << k) <= REFIT_WINDOW) {
            const float* a = tab(k - 1, j);
            const float* b = tab(k - 1, j + half);
            float* e = tab(k, j);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                e[c] = fminf(a[c], b[c]);
                e[3 + c] = fmaxf(a[3 + c], b[3 + c]);
            }
        }
        __syncthreads();
    }
    const int p = w0 + j;  // internal node p
    if (p >= nleaf - 1) return;
    const uint2 rg = node_range[p];
    if (!refit_in_window(rg)) return;  // pass 2
    float4* nd = nodes + BOX_F4 * (size_t)p;
    uint32_t lref = __float_as_uint(nd[0].w), rref = __float_as_uint(nd[2].w);
    const int gamma = (int)(lref & LEAF_INDEX_MASK);  // the split: left child covers [lo, gamma], right [gamma + 1, hi]
    auto range_box = [&](int a, int b, RefitBox* o) {   // leaves a .. b (inclusive), both inside the window
        const int len = b - a + 1;
        const int k = 31 - __clz(len);
        const float* u = tab(k, a - w0);
        const float* v = tab(k, b - (1 << k) + 1 - w0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o->lo[c] = fminf(u[c], v[c]);
            o->hi[c] = fmaxf(u[3 + c], v[3 + c]);
        }
    };
    RefitBox L, R;
    range_box((int)rg.x, gamma, &L);
    range_box(gamma + 1, (int)rg.y, &R);
    if ((lref & LEAF_BIT) && s_degen[gamma - w0]) lref |= LEAF_DEGEN_BIT;
    if ((rref & LEAF_BIT) && s_degen[gamma + 1 - w0]) rref |= LEAF_DEGEN_BIT;
    nd[0] = make_float4(L.lo[0], L.lo[1], L.lo[2], __uint_as_float(lref));
    nd[1] = make_float4(L.hi[0], L.hi[1], L.hi[2], 0.0f);
    nd[2] = make_float4(R.lo[0], R.lo[1], R.lo[2], __uint_as_float(rref));
    nd[3] = make_float4(R.hi[0], R.hi[1], R.hi[2], 0.0f);
}

// Pass 2. Thread e < nleaf: leaf e; thread e >= nleaf: internal node e - nleaf. Starts only where the parent was left
// to this pass.
__global__ void __launch_bounds__(256)
k_refit_climb(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi,
              const uint32_t* __restrict__ order, int nleaf, float4* nodes,
              const uint32_t* __restrict__ leaf_parent, const uint32_t* __restrict__ node_parent,
              const uint2* __restrict__ node_range, uint32_t* __restrict__ node_flag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * nleaf - 1) return;
    float lo[3], hi[3];
    uint32_t link;
    bool degen = false;
    if (e < nleaf) {
        link = leaf_parent[e];
        if (refit_in_window(node_range[link >> 1])) return;  // the parent's boxes came from pass 1
        const uint32_t t = order[e];
        const float4 tl = tri_lo[t], th = tri_hi[t];
        lo[0] = tl.x; lo[1] = tl.y; lo[2] = tl.z;
        hi[0] = th.x; hi[1] = th.y; hi[2] = th.z;
        degen = tl.w != 0.0f;
    } else {
        const int c = e - nleaf;
        link = node_parent[c];
        if (link == 0xffffffffu) return;                      // the root has no parent
        if (!refit_in_window(node_range[c])) return;          // a pass-2 node: reached by the climb, not a starter
        if (refit_in_window(node_range[link >> 1])) return;   // parent done in pass 1
        const float4* nd = nodes + BOX_F4 * (size_t)c;        // written by pass 1 (an earlier kernel)
        const float4 a0 = nd[0], a1 = nd[1], b0 = nd[2], b1 = nd[3];
        lo[0] = fminf(a0.x, b0.x); lo[1] = fminf(a0.y, b0.y); lo[2] = fminf(a0.z, b0.z);
        hi[0] = fmaxf(a1.x, b1.x); hi[1] = fmaxf(a1.y, b1.y); hi[2] = fmaxf(a1.z, b1.z);
    }
    bool first_level = true;
    for (;;) {
        const uint32_t p = link >> 1, side = link & 1u;
        const uint32_t up = node_parent[p];  // issued before the wait on the counter: off the critical path
        float4* nd = nodes + BOX_F4 * (size_t)p;
        // write my box into my side of the parent (the child ref lives in .w of the side's first
        // float4, each written only by its own side)
        float4* mine = nd + 2 * side;
        uint32_t ref = __float_as_uint(mine[0].w);
        if (first_level && degen) ref |= LEAF_DEGEN_BIT;
        __stcg(mine, make_float4(lo[0], lo[1], lo[2], __uint_as_float(ref)));
        __stcg(mine + 1, make_float4(hi[0], hi[1], hi[2], 0.0f));
        first_level = false;
        // one acq_rel read-modify-write instead of fence + atomic + fence: releases my box, and for the second
        // arrival acquires the sibling's (read below from L2, where the sibling's release put it)
        cuda::atomic_ref<uint32_t, cuda::thread_scope_device> flag(node_flag[p]);
        if (flag.fetch_add(1u, cuda::std::memory_order_acq_rel) == 0u) return;  // sibling not there yet
        // both children present: union with the sibling's box and go up
        const float4* sib = nd + 2 * (side ^ 1u);
        const float4 s0 = __ldcg(sib), s1 = __ldcg(sib + 1);
        lo[0] = fminf(lo[0], s0.x); lo[1] = fminf(lo[1], s0.y); lo[2] = fminf(lo[2], s0.z);
        hi[0] = fmaxf(hi[0], s1.x); hi[1] = fmaxf(hi[1], s1.y); hi[2] = fmaxf(hi[2], s1.z);
        link = up;
        if (link == 0xffffffffu) return;  // root done
    }
}

// A mesh of ONE triangle: node 0 = (the triangle, an unreachable copy of it), so that every tree has an internal
// root and the query kernels need no special case. The copy's search box has extents of -1e30 (every bound
// saturates), its ray box is empty (lo = +inf, hi = -inf).
__global__ void k_single_root(const float4* __restrict__ tobb, const float4* __restrict__ tri_lo,
                              const float4* __restrict__ tri_hi, float4* __restrict__ nodes, float4* __restrict__ boxes,
                              uint2* __restrict__ node_range) {
    if (threadIdx.x != 0) return;
    const bool degen = tri_lo[0].w != 0.0f;
    const uint32_t ref = LEAF_BIT | (degen ? LEAF_DEGEN_BIT : 0u);
    float4 c = tobb[0];
    c.w = __uint_as_float(ref);
    nodes[0] = c;
    nodes[1] = tobb[1];
    nodes[2] = tobb[2];
    nodes[3] = tobb[3];
    nodes[4] = c;
    nodes[5] = make_float4(1.f, 0.f, 0.f, -1.0e30f);
    nodes[6] = make_float4(0.f, 1.f, 0.f, -1.0e30f);
    nodes[7] = make_float4(0.f, 0.f, 1.f, -1.0e30f);
    const float4 l = tri_lo[0], h = tri_hi[0];
    boxes[0] = make_float4(l.x, l.y, l.z, __uint_as_float(ref));
    boxes[1] = make_float4(h.x, h.y, h.z, 0.0f);
    boxes[2] = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(ref));
    boxes[3] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.0f);
    node_range[0] = make_uint2(0u, 0u);
}

}  // namespace

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

#define CK(x)                          \
    do {                               \
        cudaError_t e__ = (x);         \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// The mesh's own status block (bounds + mesh errors), written by the build.
cudaError_t launch_mesh_status_reset(Device& d, MeshDev& m) {
    CK(m.status.ensure(sizeof(BuildStatus)));
    k_mesh_status_init<<<1, 32, 0, d.stream>>>(m.status.as<BuildStatus>());
    d.launches++;
    return cudaGetLastError();
}

// The call's status block = the mesh's block (or empty bounds for an empty mesh) + this call's flags.
cudaError_t launch_call_status_init(Device& d, const MeshDev& m, bool clear_errors) {
    const bool fresh = d.call_status.p == nullptr;
    CK(d.call_status.ensure(sizeof(BuildStatus)));
    k_call_status_init<<<1, 32, 0, d.stream>>>(d.call_status.as<BuildStatus>(),
                                               m.nt ? m.status.as<BuildStatus>() : nullptr, (clear_errors || fresh) ? 1 : 0);
    d.launches++;
    return cudaGetLastError();
}

// Builds records + LBVH for (d_verts, d_tris) into m, on d.stream.
cudaError_t launch_build(Device& d, MeshDev& m, const float* d_verts, uint64_t nv, const uint32_t* d_tris, uint64_t nt,
                         cudaEvent_t after_records) {
    cudaStream_t s = d.stream;
    m.bvh = Bvh{};
    m.nv = nv;
    m.nt = nt;
    m.nodes_il_mag = -1.0f;
    m.bins_built = false;
    if (nt == 0) return cudaSuccess;
    CK(launch_mesh_status_reset(d, m));
    BuildStatus* st = m.status.as<BuildStatus>();

    const uint32_t nleaf = (uint32_t)nt;
    const size_t n_nodes = nleaf > 1 ? nleaf - 1 : 1;
    CK(d.rec_orig.ensure(nt * 48));
    CK(m.rec_sorted.ensure(nt * 48));
    CK(d.tri_lo.ensure(nt * 16));
    CK(d.tri_hi.ensure(nt * 16));
    CK(d.keys_in.ensure(nt * 8));
    CK(d.keys_out.ensure(nt * 8));
    CK(d.vals_in.ensure(nt * 4));
    CK(d.vals_out.ensure(nt * 4));
    CK(m.tri_id_sorted.ensure(nt * 4));
    CK(m.nodes.ensure(n_nodes * NODE_F4 * 16));
    CK(m.nodes_il.ensure(n_nodes * NODE_F4 * 16));
    CK(m.boxes.ensure(n_nodes * BOX_F4 * 16));
    CK(m.node_range.ensure((size_t)nleaf * 8));
    CK(d.tobb.ensure(nt * 64));
    CK(d.leaf_parent.ensure((size_t)nleaf * 4));
    CK(d.node_parent.ensure((size_t)nleaf * 4));
    CK(d.node_flag.ensure((size_t)nleaf * 4));

    const unsigned bs = 256;
    k_tri_setup<<<blocks_for(nt, bs), bs, 0, s>>>(d_verts, (uint32_t)nv, d_tris, (uint32_t)nt,
                                                  d.rec_orig.as<float4>(), d.tri_lo.as<float4>(),
                                                  d.tri_hi.as<float4>(), st);
    if (after_records) CK(cudaEventRecord(after_records, s));
    k_tri_morton<<<blocks_for(nt, bs), bs, 0, s>>>(d.tri_lo.as<float4>(), d.tri_hi.as<float4>(), (uint32_t)nt, st,
                                                   d.keys_in.as<uint64_t>());
    d.launches += 2;
    // Morton order of the triangles: pass 0 reads keys_in and takes the positions as payloads (m2s_sort.cuh)
    uint64_t* const kbuf[2] = {d.keys_out.as<uint64_t>(), d.keys_in.as<uint64_t>()};
    uint32_t* const vbuf[2] = {d.vals_in.as<uint32_t>(), d.vals_out.as<uint32_t>()};
    constexpr int SORTED = (MORTON_BITS / 8 - 1) & 1;  // buffer of the last pass
    static_assert(MORTON_BITS % 8 == 0, "whole radix passes");
    CK(d.sort_tmp.ensure(radix_sort_scratch_bytes(nt)));
    CK(radix_sort_pairs(s, sort_detail::PtrSrc<uint64_t>{d.keys_in.as<uint64_t>()}, kbuf, vbuf, nt, MORTON_BITS,
                        d.sort_tmp.p, true, &d.launches));
    const uint64_t* keys_sorted = kbuf[SORTED];
    const uint32_t* order = vbuf[SORTED];
    k_tri_permute<<<blocks_for(nt, bs), bs, 0, s>>>(d.rec_orig.as<float4>(), d.tri_lo.as<float4>(),
                                                    order, (uint32_t)nt, st,
                                                    m.rec_sorted.as<float4>(), d.tobb.as<float4>(),
                                                    m.tri_id_sorted.as<uint32_t>());
    d.launches++;
    if (nleaf > 1) {
        CK(cudaMemsetAsync(d.node_flag.p, 0, (size_t)nleaf * 4, s));
        k_hierarchy<<<blocks_for(nleaf - 1, bs), bs, 0, s>>>(keys_sorted, (int)nleaf,
                                                             m.boxes.as<float4>(), d.leaf_parent.as<uint32_t>(),
                                                             d.node_parent.as<uint32_t>(), m.node_range.as<uint2>());
        // 54 KB of dynamic shared memory: opt-in per device (the attribute belongs to the current device's context)
        CK(cudaFuncSetAttribute(k_refit_windows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                REFIT_LEVELS * REFIT_WINDOW * 6 * 4));
        k_refit_windows<<<blocks_for(nleaf, REFIT_WINDOW), REFIT_WINDOW, REFIT_LEVELS * REFIT_WINDOW * 6 * 4, s>>>(
            d.tri_lo.as<float4>(), d.tri_hi.as<float4>(), order, (int)nleaf, m.boxes.as<float4>(),
            m.node_range.as<uint2>());
        k_refit_climb<<<blocks_for((uint64_t)2 * nleaf - 1, bs), bs, 0, s>>>(
            d.tri_lo.as<float4>(), d.tri_hi.as<float4>(), order, (int)nleaf, m.boxes.as<float4>(),
            d.leaf_parent.as<uint32_t>(), d.node_parent.as<uint32_t>(), m.node_range.as<uint2>(),
            d.node_flag.as<uint32_t>());
        d.launches++;
        // child-slot boxes: small subtrees with 8 lanes each, the rest through a compacted list with a warp each
        CK(d.slot_list.ensure((size_t)nleaf * 2 * 4));
        CK(d.slot_count.ensure(4));
        CK(cudaMemsetAsync(d.slot_count.p, 0, 4, s));
        k_search_nodes<<<blocks_for((uint64_t)2 * (nleaf - 1) * 8, bs), bs, 0, s>>>(
            m.rec_sorted.as<float4>(), d.tobb.as<float4>(), (uint32_t)nt, (int)nleaf, m.boxes.as<float4>(),
            m.nodes.as<float4>(), m.node_range.as<uint2>(), st, 1.0f, d.slot_list.as<uint32_t>(),
            d.slot_count.as<uint32_t>());
        k_search_nodes_big<<<d.sm_count * 8, bs, 0, s>>>(
            m.rec_sorted.as<float4>(), (uint32_t)nt, m.boxes.as<float4>(), m.nodes.as<float4>(),
            m.node_range.as<uint2>(), st, 1.0f, d.slot_list.as<uint32_t>(), d.slot_count.as<uint32_t>());
        d.launches++;
        d.launches += 3;
    } else {
        k_single_root<<<1, 32, 0, s>>>(d.tobb.as<float4>(), d.tri_lo.as<float4>(), d.tri_hi.as<float4>(),
                                       m.nodes.as<float4>(), m.boxes.as<float4>(), m.node_range.as<uint2>());
        d.launches++;
    }
    m.bvh.rec = m.rec_sorted.as<float4>();
    m.bvh.boxes = m.boxes.as<float4>();
    m.bvh.tri_id = m.tri_id_sorted.as<uint32_t>();
    m.bvh.nodes = m.nodes.as<float4>();
    m.bvh.nodes_il = m.nodes_il.as<float4>();
    m.bvh.nt = (uint32_t)nt;
    m.bvh.n_nodes = (uint32_t)n_nodes;
    m.bvh.node_range = m.node_range.as<uint2>();
    m.bvh.stats = nullptr;
    return cudaGetLastError();
}

// Writes Bvh::nodes_il in units of S = 2^k >= 4 x max(scene magnitude of the CALL's status block, mag_key).
// Grids pass their host-known magnitude as the key (a no-op when it is current); point calls force it, because
// the call's bounds include the queries, which only the device knows.
cudaError_t launch_nodes_interleave(Device& d, MeshDev& m, float mag_key, bool force) {
    if (m.bvh.n_nodes == 0 || (!force && m.nodes_il_mag == mag_key)) return cudaSuccess;
    k_nodes_interleave<<<blocks_for(m.bvh.n_nodes, 256), 256, 0, d.stream>>>(
        m.nodes.as<float4>(), m.bvh.n_nodes, m.nodes_il.as<float4>(), d.call_status.as<BuildStatus>(), mag_key);
    d.launches++;
    m.nodes_il_mag = force ? -1.0f : mag_key;
    return cudaGetLastError();
}

// Ray bins of a built mesh (RayBins, m2s_internal.h): count -> verdict -> exclusive scan -> scatter, all enqueued.
cudaError_t launch_ray_bins(Device& d, MeshDev& m) {
    if (m.bins_built || m.nt == 0) return cudaSuccess;
    cudaStream_t s = d.stream;
    const uint32_t nt = (uint32_t)m.nt;
    // about one triangle per cell of a projection: R = 2^round(log2(sqrt(nt))), 16 .. 2048
    uint32_t R = 16;
    while (R < 2048u && (double)R * R * 2.0 < (double)nt) R <<= 1;
    const size_t cells = (size_t)3 * R * R;
    const uint64_t cap64 = (uint64_t)RAYBIN_ITEMS_PER_TRI * 3u * nt + 1024u;
    const uint32_t capacity = (uint32_t)std::min<uint64_t>(cap64, 0xfffffff0ull);
    {
        // the bins are an accelerator, not a requirement: a mesh too large for their item array (144 bytes per triangle)
        // keeps the packet walk of the box tree
        cudaError_t e = m.bin_offsets.ensure((cells + 1) * 4);
        if (e == cudaSuccess) e = m.bin_cursor.ensure((cells + 1) * 4);
        if (e == cudaSuccess) e = m.bin_items.ensure((size_t)capacity * 4);
        if (e == cudaSuccess) e = m.bin_big.ensure((size_t)3 * RAYBIN_MAX_BIG * 4);
        if (e == cudaSuccess) e = m.bin_meta.ensure(32);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            DevBuf* bufs[] = {&m.bin_offsets, &m.bin_cursor, &m.bin_items, &m.bin_big, &m.bin_meta};
            for (DevBuf* b : bufs) b->release();
            m.bvh.bins = RayBins{};
            m.bins_built = true;  // decided: no bins for this mesh
            return cudaSuccess;
        }
        CK(e);
    }
    CK(cudaMemsetAsync(m.bin_cursor.p, 0, (cells + 1) * 4, s));
    CK(cudaMemsetAsync(m.bin_meta.p, 0, 32, s));
    const float4* rec = m.rec_sorted.as<float4>();
    const BuildStatus* st = m.status.as<BuildStatus>();
    uint32_t* cursor = m.bin_cursor.as<uint32_t>();
    uint32_t* offsets = m.bin_offsets.as<uint32_t>();
    uint32_t* meta = m.bin_meta.as<uint32_t>();
    k_raybins<0><<<blocks_for(nt, 256), 256, 0, s>>>(rec, nt, st, R, cursor, nullptr, nullptr, m.bin_big.as<uint32_t>(),
                                                     meta, capacity);
    k_raybins_verdict<<<1, 32, 0, s>>>(meta, capacity);
    CK(d.sort_tmp.ensure(exclusive_scan_scratch_bytes(cells + 1)));
    CK(exclusive_scan_u32(s, cursor, offsets, cells + 1, d.sort_tmp.p));
    CK(cudaMemsetAsync(m.bin_cursor.p, 0, (cells + 1) * 4, s));
    k_raybins<1><<<blocks_for(nt, 256), 256, 0, s>>>(rec, nt, st, R, cursor, offsets, m.bin_items.as<uint32_t>(),
                                                     m.bin_big.as<uint32_t>(), meta, capacity);
    d.launches += 4;  // two bin passes, the verdict, the scan
    m.bvh.bins = RayBins{R, offsets, m.bin_items.as<uint32_t>(), m.bin_big.as<uint32_t>(), meta, st};
    m.bins_built = true;
    return cudaGetLastError();
}

// Sorts the queries along a Morton curve (coherent packets). Produces q_sorted (xyz + original index) and adds the
// query bounds to the call's status block.
cudaError_t sort_queries(Device& d, const float* d_queries, uint64_t nq) {
    cudaStream_t s = d.stream;
    const unsigned bs = 256;
    BuildStatus* st = d.call_status.as<BuildStatus>();
    CK(d.q_sorted.ensure(nq * 16));
    CK(d.q_keys_in.ensure(nq * 8));
    CK(d.q_keys_out.ensure(nq * 8));
    CK(d.q_vals_in.ensure(nq * 4));
    CK(d.q_perm.ensure(nq * 4));
    k_point_bounds<<<blocks_for(nq, bs), bs, 0, s>>>(d_queries, (uint32_t)nq, st);
    k_point_morton<<<blocks_for(nq, bs), bs, 0, s>>>(d_queries, (uint32_t)nq, st, d.q_keys_in.as<uint64_t>());
    uint64_t* const kbuf[2] = {d.q_keys_out.as<uint64_t>(), d.q_keys_in.as<uint64_t>()};
    uint32_t* const vbuf[2] = {d.q_vals_in.as<uint32_t>(), d.q_perm.as<uint32_t>()};
    static_assert(((MORTON_BITS / 8 - 1) & 1) == 1, "the last pass must write q_perm");
    CK(d.sort_tmp.ensure(radix_sort_scratch_bytes(nq)));
    CK(radix_sort_pairs(s, sort_detail::PtrSrc<uint64_t>{d.q_keys_in.as<uint64_t>()}, kbuf, vbuf, nq, MORTON_BITS,
                        d.sort_tmp.p, false, &d.launches));
    k_point_gather<<<blocks_for(nq, bs), bs, 0, s>>>(d_queries, d.q_perm.as<uint32_t>(), (uint32_t)nq,
                                                     d.q_sorted.as<float4>());
    d.launches += 3;
    return cudaGetLastError();
}

}  // namespace m2s
