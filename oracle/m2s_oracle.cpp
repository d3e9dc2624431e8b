// ORACLE — TEST INFRASTRUCTURE ONLY (see m2s_oracle_geo.hpp header).
//
// CPU restatement of the drivers of Azkellas/mesh_to_sdf v0.4.0 for the hot path
// (generate_grid_sdf / generate_sdf). "restated C++, not rustc-built": the reference cannot be
// compiled in this environment (no rustc/cargo; bvh 0.10.0, rstar 0.12.0, rayon … are not vendored).
//
// Two families of drivers:
//   * exact     — brute force over every triangle, the semantics of
//                 src/generate/generic/default.rs:11-74 (ground truth for parity);
//   * faithful  — src/generate/grid.rs:265-684 restated step by step (splat → sorted seeds →
//                 per-thread binary heaps → 27-neighbour propagation under per-cell locks →
//                 face-aligned row raycasts, best of 3). Used as the timed "reference CPU path" and to
//                 quantify the reference's own propagation error. The bvh crate's ray traversal is
//                 only a conservative pre-filter in the reference (hits are decided by
//                 geo::ray_triangle_intersection_aligned), so a plain median-split AABB tree stands in.
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off -pthread).
#include "m2s_oracle_geo.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <queue>
#include <thread>
#include <vector>

using namespace m2s_oracle;

namespace {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
inline V3 ld3(const float* p, uint64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

int resolve_threads(int threads) {
    if (threads > 0) return threads;
    unsigned hc = std::thread::hardware_concurrency();  // rayon::current_num_threads() analogue
    return hc ? (int)hc : 1;
}

// rayon par_iter stand-in: dynamic chunks over [0,n).
template <class F>
void parallel_for(uint64_t n, int threads, uint64_t chunk, F&& f) {
    threads = resolve_threads(threads);
    if (threads == 1 || n <= chunk) {
        for (uint64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                uint64_t e = std::min(n, b + chunk);
                for (uint64_t i = b; i < e; ++i) f(i);
            }
        });
    for (auto& th : pool) th.join();
}

// Rust `f as usize` / `as isize`: saturating, NaN -> 0.
inline int64_t f2isize(float f) {
    if (f != f) return 0;
    if (f >= 9.2233720368547758e18f) return INT64_MAX;
    if (f <= -9.2233720368547758e18f) return INT64_MIN;
    return (int64_t)f;
}
inline uint64_t f2usize(float f) {
    if (f != f || f <= 0.0f) return 0;
    if (f >= 1.8446744073709552e19f) return UINT64_MAX;
    return (uint64_t)f;
}

// src/grid.rs:30-37
struct Grid {
    V3 first_cell;
    V3 cell_size;
    uint64_t n[3];
    // src/grid.rs:110-119
    V3 bbox_min() const { return v_sub(first_cell, v_fmul(cell_size, 0.5f)); }
    // src/grid.rs:122-124
    uint64_t idx(const uint64_t c[3]) const { return c[2] + c[1] * n[2] + c[0] * n[1] * n[2]; }
    // src/grid.rs:127-132
    void coords(uint64_t i, uint64_t c[3]) const {
        c[2] = i % n[2];
        c[1] = (i / n[2]) % n[1];
        c[0] = i / (n[1] * n[2]);
    }
    // src/grid.rs:135-141
    V3 center(const uint64_t c[3]) const {
        return {first_cell.x + (float)c[0] * cell_size.x, first_cell.y + (float)c[1] * cell_size.y,
                first_cell.z + (float)c[2] * cell_size.z};
    }
    // src/grid.rs:145-170 — returns true for SnapResult::Inside.
    bool snap(const V3& p, uint64_t out[3]) const {
        V3 d = v_sub(p, bbox_min());
        float q[3] = {d.x / cell_size.x, d.y / cell_size.y, d.z / cell_size.z};
        bool inside = true;
        for (int i = 0; i < 3; ++i) {
            int64_t c = f2isize(std::floor(q[i]));
            int64_t hi = (int64_t)n[i] - 1;
            int64_t r = c < 0 ? 0 : (c > hi ? hi : c);
            if (r != c) inside = false;
            out[i] = (uint64_t)r;
        }
        return inside;
    }
    uint64_t total() const { return n[0] * n[1] * n[2]; }
};

Grid make_grid(const float* first, const float* size, const uint64_t* count) {
    Grid g;
    g.first_cell = {first[0], first[1], first[2]};
    g.cell_size = {size[0], size[1], size[2]};
    g.n[0] = count[0];
    g.n[1] = count[1];
    g.n[2] = count[2];
    return g;
}

// ---------------------------------------------------------------------------------------------
// A plain AABB tree over the padded triangle boxes (geo.rs:4-22). Stand-in for the conservative
// filters of the bvh / rstar crates; never decides a result by itself.
// ---------------------------------------------------------------------------------------------
struct Tree {
    struct Node {
        V3 mn, mx;
        int32_t left, right;   // children, or -1
        int32_t first, count;  // leaf range into `order`
    };
    std::vector<Node> nodes;
    std::vector<uint32_t> order;
    std::vector<V3> bmn, bmx;

    void build(const float* verts, const uint32_t* tris, uint64_t nt) {
        bmn.resize(nt);
        bmx.resize(nt);
        order.resize(nt);
        std::vector<V3> cen(nt);
        for (uint64_t t = 0; t < nt; ++t) {
            V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
            triangle_bounding_box(a, b, c, &bmn[t], &bmx[t]);
            cen[t] = v_fmul(v_add(bmn[t], bmx[t]), 0.5f);
            order[t] = (uint32_t)t;
        }
        nodes.clear();
        nodes.reserve(2 * nt + 1);
        if (nt) build_rec(cen, 0, (int32_t)nt);
    }
    int32_t build_rec(const std::vector<V3>& cen, int32_t b, int32_t e) {
        int32_t id = (int32_t)nodes.size();
        nodes.push_back({});
        V3 mn = {INFINITY, INFINITY, INFINITY}, mx = {-INFINITY, -INFINITY, -INFINITY};
        V3 cmn = mn, cmx = mx;
        for (int32_t i = b; i < e; ++i) {
            uint32_t t = order[i];
            mn = {std::fmin(mn.x, bmn[t].x), std::fmin(mn.y, bmn[t].y), std::fmin(mn.z, bmn[t].z)};
            mx = {std::fmax(mx.x, bmx[t].x), std::fmax(mx.y, bmx[t].y), std::fmax(mx.z, bmx[t].z)};
            cmn = {std::fmin(cmn.x, cen[t].x), std::fmin(cmn.y, cen[t].y), std::fmin(cmn.z, cen[t].z)};
            cmx = {std::fmax(cmx.x, cen[t].x), std::fmax(cmx.y, cen[t].y), std::fmax(cmx.z, cen[t].z)};
        }
        nodes[id].mn = mn;
        nodes[id].mx = mx;
        nodes[id].left = nodes[id].right = -1;
        nodes[id].first = b;
        nodes[id].count = e - b;
        if (e - b <= 4) return id;
        V3 ext = v_sub(cmx, cmn);
        int ax = (ext.x >= ext.y && ext.x >= ext.z) ? 0 : (ext.y >= ext.z ? 1 : 2);
        int32_t mid = (b + e) / 2;
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e,
                         [&](uint32_t p, uint32_t q) { return v_get(cen[p], ax) < v_get(cen[q], ax); });
        int32_t l = build_rec(cen, b, mid);
        int32_t r = build_rec(cen, mid, e);
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }

    // All triangles whose padded box is pierced by the ray o + t*e_axis, t >= 0 (superset of what
    // bvh.traverse returns; exact hits are decided by the caller).
    template <class F>
    void ray_candidates(const V3& o, int axis, F&& f) const {
        if (nodes.empty()) return;
        const int iy = (axis + 1) % 3, iz = (axis + 2) % 3;
        const float oy = v_get(o, iy), oz = v_get(o, iz), ox = v_get(o, axis);
        int32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (oy < v_get(nd.mn, iy) || oy > v_get(nd.mx, iy) || oz < v_get(nd.mn, iz) ||
                oz > v_get(nd.mx, iz) || ox > v_get(nd.mx, axis))
                continue;
            if (nd.left < 0) {
                for (int32_t i = 0; i < nd.count; ++i) {
                    uint32_t t = order[nd.first + i];
                    if (oy < v_get(bmn[t], iy) || oy > v_get(bmx[t], iy) || oz < v_get(bmn[t], iz) ||
                        oz > v_get(bmx[t], iz) || ox > v_get(bmx[t], axis))
                        continue;
                    f(t);
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
    }

    static float box_dist2(const V3& p, const V3& mn, const V3& mx) {
        float dx = std::fmax(std::fmax(mn.x - p.x, p.x - mx.x), 0.0f);
        float dy = std::fmax(std::fmax(mn.y - p.y, p.y - mx.y), 0.0f);
        float dz = std::fmax(std::fmax(mn.z - p.z, p.z - mx.z), 0.0f);
        return dx * dx + dy * dy + dz * dz;
    }

    // Exact arg-min of geo::point_triangle_distance2 (the rstar nearest_neighbor contract,
    // rtree.rs:64-77,116). Ties -> smallest triangle index. Pruning is conservative (slack on the
    // box distance) so the result equals the brute-force minimum bit for bit.
    uint32_t nearest(const float* verts, const uint32_t* tris, const V3& p, float* best_d2_out) const {
        float best = INFINITY;
        uint32_t best_t = 0;
        int32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (box_dist2(p, nd.mn, nd.mx) * 0.9999f > best) continue;
            if (nd.left < 0) {
                for (int32_t i = 0; i < nd.count; ++i) {
                    uint32_t t = order[nd.first + i];
                    V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]),
                       c = ld3(verts, tris[3 * t + 2]);
                    float d2 = point_triangle_distance2(p, a, b, c);
                    if (d2 < best || (d2 == best && t < best_t)) {
                        best = d2;
                        best_t = t;
                    }
                }
            } else {
                const Node& l = nodes[nd.left];
                const Node& r = nodes[nd.right];
                float dl = box_dist2(p, l.mn, l.mx), dr = box_dist2(p, r.mn, r.mx);
                if (dl < dr) {
                    stack[sp++] = nd.right;
                    stack[sp++] = nd.left;
                } else {
                    stack[sp++] = nd.left;
                    stack[sp++] = nd.right;
                }
            }
        }
        *best_d2_out = best;
        return best_t;
    }
};

// ---------------------------------------------------------------------------------------------
// exact per-query evaluation — default.rs:28-73 / bvh.rs:79-142 / rtree.rs:114-124 /
// rtree_bvh.rs:124-172, all with "every triangle is a candidate".
// accel: 0=None 1=Bvh 2=Rtree 3=RtreeBvh (lib.rs:224-239); sign: 0=Raycast 1=Normal (lib.rs:204-216)
// ---------------------------------------------------------------------------------------------
float exact_query(const float* verts, const uint32_t* tris, uint64_t nt, const V3& q, int accel,
                  int sign, bool* nan) {
    if (accel == 2) {  // Rtree: sign of THE nearest triangle (rtree.rs:116-123). Tie -> first index.
        float best = INFINITY;
        uint64_t bt = 0;
        for (uint64_t t = 0; t < nt; ++t) {
            float d2 = point_triangle_distance2(q, ld3(verts, tris[3 * t]), ld3(verts, tris[3 * t + 1]),
                                                ld3(verts, tris[3 * t + 2]));
            if (d2 < best) {
                best = d2;
                bt = t;
            }
        }
        return point_triangle_signed_distance(q, ld3(verts, tris[3 * bt]), ld3(verts, tris[3 * bt + 1]),
                                              ld3(verts, tris[3 * bt + 2]));
    }
    const bool raycast = (accel == 3) || (sign == 0);
    if (!raycast) {
        // Normal: fold under compare_distances, triangle order (default.rs:52-60; bvh.rs:82-94 is the
        // mirrored test `compare(min, d) == Greater`, same outcome).
        float m = std::numeric_limits<float>::max();
        for (uint64_t t = 0; t < nt; ++t) {
            float d = point_triangle_signed_distance(q, ld3(verts, tris[3 * t]), ld3(verts, tris[3 * t + 1]),
                                                     ld3(verts, tris[3 * t + 2]));
            if (compare_distances(d, m, nan) < 0) m = d;
        }
        return m;
    }
    float m = std::numeric_limits<float>::max();
    uint32_t cnt[3] = {0, 0, 0};
    const int naxes = (accel == 0) ? 1 : 3;  // None(Raycast): +X only (default.rs:34-38)
    for (uint64_t t = 0; t < nt; ++t) {
        V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
        m = std::fmin(m, point_triangle_distance(q, a, b, c));  // f32::min (default.rs:48, bvh.rs:103)
        for (int ax = 0; ax < naxes; ++ax) {
            float tt;
            if (ray_triangle_intersection_aligned(q, a, b, c, ax, &tt)) cnt[ax]++;
        }
    }
    if (accel == 0) return (cnt[0] % 2 == 0) ? m : -m;  // default.rs:65-72
    int insides = (cnt[0] & 1) + (cnt[1] & 1) + (cnt[2] & 1);  // bvh.rs:130-141, rtree_bvh.rs:160-171
    return insides > 1 ? -m : m;
}

// grid.rs:601-617 for one row: cells 0..=k are incremented for a hit at t.
inline uint64_t row_hit_last_cell(float t, float cs, uint64_t n_axis) {
    float cc = t / cs;
    uint64_t k = f2usize(std::floor(cc));
    return std::min(k, n_axis - 1);
}

// Exact value of one grid cell: |d| = min over all triangles; sign per generate/grid.rs semantics.
float exact_grid_cell(const float* verts, const uint32_t* tris, uint64_t nt, const Grid& g,
                      const uint64_t cell[3], int sign, bool* nan) {
    V3 p = g.center(cell);
    if (sign == 1) {
        float m = std::numeric_limits<float>::max();
        for (uint64_t t = 0; t < nt; ++t) {
            float d = point_triangle_signed_distance(p, ld3(verts, tris[3 * t]), ld3(verts, tris[3 * t + 1]),
                                                     ld3(verts, tris[3 * t + 2]));
            if (compare_distances(d, m, nan) < 0) m = d;
        }
        return m;
    }
    float m = std::numeric_limits<float>::max();
    uint32_t cnt[3] = {0, 0, 0};
    V3 start[3];
    for (int ax = 0; ax < 3; ++ax) {
        uint64_t sc[3] = {cell[0], cell[1], cell[2]};
        sc[ax] = 0;  // generate_raycasts: rays start on the face cell (grid.rs:648-684)
        start[ax] = g.center(sc);
    }
    for (uint64_t t = 0; t < nt; ++t) {
        V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
        m = std::fmin(m, point_triangle_distance(p, a, b, c));
        for (int ax = 0; ax < 3; ++ax) {
            float tt;
            if (ray_triangle_intersection_aligned(start[ax], a, b, c, ax, &tt)) {
                uint64_t k = row_hit_last_cell(tt, v_get(g.cell_size, ax), g.n[ax]);
                if (cell[ax] <= k) cnt[ax]++;
            }
        }
    }
    int odd = (cnt[0] & 1) + (cnt[1] & 1) + (cnt[2] & 1);  // grid.rs:633-638
    return odd >= 2 ? -m : m;
}

// ---------------------------------------------------------------------------------------------
// faithful generate_grid_sdf (generate/grid.rs)
// ---------------------------------------------------------------------------------------------
struct SpinLock {
    std::atomic<uint8_t> f{0};
    void lock() {
        while (f.exchange(1, std::memory_order_acquire)) {
            while (f.load(std::memory_order_relaxed)) {
            }
        }
    }
    void unlock() { f.store(0, std::memory_order_release); }
};

struct State {  // grid.rs:17-25
    float distance;
    uint32_t cell[3];
    uint32_t tri;  // triangle id; the reference stores the index triple and uses it as last tie-break
};

struct Faithful {
    const float* verts;
    const uint32_t* tris;
    uint64_t nt;
    Grid g;
    int sign;
    int threads;
    std::atomic<bool> nan{false};

    inline float dist(const V3& p, uint32_t t) const {
        V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
        return sign == 0 ? point_triangle_distance(p, a, b, c)  // grid.rs:439-442, 535-540
                         : point_triangle_signed_distance(p, a, b, c);
    }
    inline int cmp(float a, float b) {
        bool n = false;
        int r = compare_distances(a, b, &n);
        if (n) nan.store(true, std::memory_order_relaxed);
        return r;
    }
    // State::cmp, grid.rs:27-35 — true iff a < b (a is popped later than b by the max-heap).
    // The triangle tie-break compares index triples in the reference (:34).
    bool state_less(const State& a, const State& b) {
        int c = cmp(b.distance, a.distance);
        if (c) return c < 0;
        for (int i = 0; i < 3; ++i)
            if (a.cell[i] != b.cell[i]) return a.cell[i] < b.cell[i];
        for (int i = 0; i < 3; ++i) {
            uint32_t x = tris[3 * a.tri + i], y = tris[3 * b.tri + i];
            if (x != y) return x < y;
        }
        return false;
    }
};

}  // namespace

// =================================================================================================
// C API (ctypes)
// =================================================================================================
extern "C" {

int m2s_oracle_hardware_threads() { return resolve_threads(0); }

// ---- leaf functions, for known-answer tests -----------------------------------------------------
void m2s_oracle_triangle_bounding_box(const float* a, const float* b, const float* c, float* mn, float* mx) {
    V3 lo, hi;
    triangle_bounding_box(ld3(a, 0), ld3(b, 0), ld3(c, 0), &lo, &hi);
    mn[0] = lo.x; mn[1] = lo.y; mn[2] = lo.z;
    mx[0] = hi.x; mx[1] = hi.y; mx[2] = hi.z;
}
void m2s_oracle_closest_point_segment(const float* p, const float* a, const float* b, float* out) {
    V3 r = closest_point_segment(ld3(p, 0), ld3(a, 0), ld3(b, 0));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void m2s_oracle_closest_point_triangle(const float* p, const float* a, const float* b, const float* c,
                                       float* out) {
    V3 r = closest_point_triangle(ld3(p, 0), ld3(a, 0), ld3(b, 0), ld3(c, 0));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float m2s_oracle_point_triangle_distance(const float* p, const float* a, const float* b, const float* c) {
    return point_triangle_distance(ld3(p, 0), ld3(a, 0), ld3(b, 0), ld3(c, 0));
}
float m2s_oracle_point_triangle_distance2(const float* p, const float* a, const float* b, const float* c) {
    return point_triangle_distance2(ld3(p, 0), ld3(a, 0), ld3(b, 0), ld3(c, 0));
}
float m2s_oracle_point_triangle_signed_distance(const float* p, const float* a, const float* b,
                                                const float* c) {
    return point_triangle_signed_distance(ld3(p, 0), ld3(a, 0), ld3(b, 0), ld3(c, 0));
}
// returns 1 and writes *t when the reference returns Some(t)
int m2s_oracle_ray_triangle_intersection_aligned(const float* o, const float* v0, const float* v1,
                                                 const float* v2, int axis, float* t) {
    return ray_triangle_intersection_aligned(ld3(o, 0), ld3(v0, 0), ld3(v1, 0), ld3(v2, 0), axis, t) ? 1 : 0;
}
int m2s_oracle_approx_eq_f32(float a, float b, int ulps, float eps) { return approx_eq_f32(a, b, ulps, eps); }
// -1 Less / 0 Equal / 1 Greater / 2 = the reference would panic (NaN)
int m2s_oracle_compare_distances(float a, float b) {
    bool nan = false;
    int r = compare_distances(a, b, &nan);
    return nan ? 2 : r;
}

// ---- Grid (src/grid.rs) ---------------------------------------------------------------------------
// grid.rs:59-74
void m2s_oracle_grid_from_bounding_box(const float* bmin, const float* bmax, const uint64_t* count,
                                       float* first_cell, float* cell_size) {
    for (int i = 0; i < 3; ++i) {
        float cs = (bmax[i] - bmin[i]) / (float)count[i];
        cell_size[i] = cs;
        first_cell[i] = bmin[i] + cs * 0.5f;
    }
}
uint64_t m2s_oracle_grid_cell_idx(const uint64_t* count, const uint64_t* cell) {
    Grid g{};
    g.n[0] = count[0]; g.n[1] = count[1]; g.n[2] = count[2];
    return g.idx(cell);
}
void m2s_oracle_grid_cell_coords(const uint64_t* count, uint64_t idx, uint64_t* cell) {
    Grid g{};
    g.n[0] = count[0]; g.n[1] = count[1]; g.n[2] = count[2];
    g.coords(idx, cell);
}
void m2s_oracle_grid_cell_center(const float* first, const float* size, const uint64_t* count,
                                 const uint64_t* cell, float* out) {
    Grid g = make_grid(first, size, count);
    V3 c = g.center(cell);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
// grid.rs:82-88
void m2s_oracle_grid_last_cell(const float* first, const float* size, const uint64_t* count, float* out) {
    for (int i = 0; i < 3; ++i) out[i] = first[i] + (float)count[i] * size[i];
}
// grid.rs:110-119
void m2s_oracle_grid_bounding_box(const float* first, const float* size, const uint64_t* count,
                                  float* mn, float* mx) {
    for (int i = 0; i < 3; ++i) {
        mn[i] = first[i] - size[i] * 0.5f;
        mx[i] = mn[i] + (float)count[i] * size[i];
    }
}
// returns 1 for Inside, 0 for Outside
int m2s_oracle_grid_snap(const float* first, const float* size, const uint64_t* count, const float* p,
                         uint64_t* cell) {
    Grid g = make_grid(first, size, count);
    return g.snap(ld3(p, 0), cell) ? 1 : 0;
}

// ---- Topology::get_triangles (lib.rs:175-193) -----------------------------------------------------
// kind: 0 = TriangleList, 1 = TriangleStrip. indices == NULL means `None` (0..nv). Returns the
// number of triangles; writes 3*count u32 into out when out != NULL.
uint64_t m2s_oracle_expand_topology(int kind, const uint32_t* indices, uint64_t n_indices, uint64_t nv,
                                    uint32_t* out) {
    uint64_t n = indices ? n_indices : nv;
    auto at = [&](uint64_t i) -> uint32_t { return indices ? indices[i] : (uint32_t)i; };
    uint64_t cnt = 0;
    if (kind == 0) {  // itertools tuples(): trailing partial tuple dropped
        cnt = n / 3;
        if (out)
            for (uint64_t t = 0; t < cnt; ++t)
                for (int k = 0; k < 3; ++k) out[3 * t + k] = at(3 * t + k);
    } else {  // tuple_windows(): no winding flip
        cnt = n >= 3 ? n - 2 : 0;
        if (out)
            for (uint64_t t = 0; t < cnt; ++t)
                for (int k = 0; k < 3; ++k) out[3 * t + k] = at(t + k);
    }
    return cnt;
}

// ---- exact drivers ---------------------------------------------------------------------------------
// Returns 0 ok, 2 index out of bounds (the reference panics on slice OOB), 3 NaN distance panic.
static int check_indices(const uint32_t* tris, uint64_t nt, uint64_t nv) {
    for (uint64_t i = 0; i < 3 * nt; ++i)
        if (tris[i] >= nv) return 2;
    return 0;
}

// generate_sdf (lib.rs:291-311) with every triangle as candidate. For RtreeBvh and nt == 0 the
// reference returns an empty Vec (rtree_bvh.rs:104-106): this function then writes nothing and
// returns 7; Rtree on an empty mesh panics (rtree.rs:117) -> 7 as well.
int m2s_oracle_generate_sdf(const float* verts, uint64_t nv, const uint32_t* tris, uint64_t nt,
                            const float* queries, uint64_t nq, int accel, int sign, int threads, float* out) {
    if (int e = check_indices(tris, nt, nv)) return e;
    if (nt == 0 && (accel == 2 || accel == 3)) return 7;
    std::atomic<bool> nan{false};
    parallel_for(nq, threads, 16, [&](uint64_t i) {
        bool n = false;
        out[i] = exact_query(verts, tris, nt, ld3(queries, i), accel, sign, &n);
        if (n) nan.store(true);
    });
    return nan.load() ? 3 : 0;
}

// Exact grid SDF at selected cells (cell_idx == NULL -> all cells, n_cells ignored).
int m2s_oracle_grid_cells_exact(const float* verts, uint64_t nv, const uint32_t* tris, uint64_t nt,
                                const float* first, const float* size, const uint64_t* count, int sign,
                                const uint64_t* cell_idx, uint64_t n_cells, int threads, float* out) {
    if (int e = check_indices(tris, nt, nv)) return e;
    Grid g = make_grid(first, size, count);
    uint64_t n = cell_idx ? n_cells : g.total();
    std::atomic<bool> nan{false};
    parallel_for(n, threads, 8, [&](uint64_t i) {
        uint64_t c[3];
        g.coords(cell_idx ? cell_idx[i] : i, c);
        bool nn = false;
        out[i] = exact_grid_cell(verts, tris, nt, g, c, sign, &nn);
        if (nn) nan.store(true);
    });
    return nan.load() ? 3 : 0;
}

// ---- faithful generate_grid_sdf (generate/grid.rs:265-378) -------------------------------------------
// phase_ms: [precompute+preheap+heap, propagate, raycast]; steps: [init steps, propagation pops, ray candidates]
int m2s_oracle_generate_grid_sdf_faithful(const float* verts, uint64_t nv, const uint32_t* tris, uint64_t nt,
                                          const float* first, const float* size, const uint64_t* count,
                                          int sign, int threads, float* out, double* phase_ms,
                                          uint64_t* steps_out) {
    if (int e = check_indices(tris, nt, nv)) return e;
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    threads = resolve_threads(threads);
    Faithful F;
    F.verts = verts; F.tris = tris; F.nt = nt; F.g = make_grid(first, size, count);
    F.sign = sign; F.threads = threads;
    const Grid& g = F.g;
    const uint64_t cells = g.total();
    const float FMAX = std::numeric_limits<float>::max();

    // Precomputations::new (grid.rs:79-166): bvh build overlaps the allocations.
    Tree tree;
    std::thread bvh_thread;
    if (sign == 0) bvh_thread = std::thread([&]() { tree.build(verts, tris, nt); });
    std::unique_ptr<std::atomic<float>[]> pre_d(new std::atomic<float>[cells]);  // preheap distance (:118-124)
    for (uint64_t i = 0; i < cells; ++i) pre_d[i].store(FMAX, std::memory_order_relaxed);
    std::vector<uint32_t> pre_t(cells, 0);            // preheap triangle
    std::unique_ptr<SpinLock[]> locks(new SpinLock[cells]);  // per-cell RwLock stand-in
    std::vector<float> distances(cells, FMAX);        // (:137-143)

    std::atomic<uint64_t> steps{0};

    // generate_preheap (grid.rs:383-457): parallel over triangles.
    parallel_for(nt, threads, 64, [&](uint64_t t) {
        V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
        V3 bmn, bmx;
        triangle_bounding_box(a, b, c, &bmn, &bmx);
        uint64_t mnc[3], mxc[3];
        g.snap(bmn, mnc);
        g.snap(bmx, mxc);
        V3 mnf = g.center(mnc);
        for (int i = 0; i < 3; ++i)  // :411-416
            if (mnc[i] > 0 && v_get(mnf, i) > v_get(bmn, i)) mnc[i] -= 1;
        V3 mxf = g.center(mxc);
        for (int i = 0; i < 3; ++i)  // :417-426
            if (mxc[i] < g.n[i] - 1 && v_get(mxf, i) < v_get(bmx, i)) mxc[i] += 1;
        uint64_t local_steps = 0;
        for (uint64_t x = mnc[0]; x <= mxc[0]; ++x)
            for (uint64_t y = mnc[1]; y <= mxc[1]; ++y)
                for (uint64_t z = mnc[2]; z <= mxc[2]; ++z) {
                    uint64_t cell[3] = {x, y, z};
                    uint64_t ci = g.idx(cell);
                    float d = F.dist(g.center(cell), (uint32_t)t);
                    // cheap check first, then lock and re-check (:445-454)
                    float stored = pre_d[ci].load(std::memory_order_relaxed);
                    if (F.cmp(d, stored) < 0) {
                        locks[ci].lock();
                        if (F.cmp(d, pre_d[ci].load(std::memory_order_relaxed)) < 0) {
                            ++local_steps;
                            pre_d[ci].store(d, std::memory_order_relaxed);
                            pre_t[ci] = (uint32_t)t;
                        }
                        locks[ci].unlock();
                    }
                }
        steps.fetch_add(local_steps, std::memory_order_relaxed);
    });

    // generate_heap (grid.rs:464-490): serial scan + sort.
    std::vector<State> heap;
    for (uint64_t ci = 0; ci < cells; ++ci) {
        float pd = pre_d[ci].load(std::memory_order_relaxed);
        if (pd < FMAX) {
            uint64_t c[3];
            g.coords(ci, c);
            distances[ci] = pd;
            heap.push_back({pd, {(uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2]}, pre_t[ci]});
        }
    }
    // sorted_unstable() under State::cmp. stable_sort is memory-safe for a non-strict-weak order.
    std::stable_sort(heap.begin(), heap.end(), [&](const State& a, const State& b) { return F.state_less(a, b); });
    pre_d.reset();
    std::vector<uint32_t>().swap(pre_t);
    auto t1 = clk::now();
    if (steps_out) steps_out[0] = steps.exchange(0);

    // split round-robin into `threads` heaps, one OS thread each (grid.rs:318-339).
    {
        std::vector<std::vector<State>> heaps(threads);
        for (auto& h : heaps) h.reserve(heap.size() / threads + 1);
        for (size_t i = 0; i < heap.size(); ++i) heaps[i % threads].push_back(heap[i]);
        std::vector<State>().swap(heap);
        auto less = [&F](const State& a, const State& b) { return F.state_less(a, b); };
        std::vector<std::thread> pool;
        for (int th = threads - 1; th >= 0; --th) {  // heaps.pop()
            pool.emplace_back([&, th]() {
                std::priority_queue<State, std::vector<State>, decltype(less)> pq(less, std::move(heaps[th]));
                uint64_t local = 0;
                // propagate_heap (grid.rs:495-558)
                while (!pq.empty()) {
                    State s = pq.top();
                    pq.pop();
                    ++local;
                    for (int dx = -1; dx <= 1; ++dx)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dz = -1; dz <= 1; ++dz) {
                                int64_t x = (int64_t)s.cell[0] + dx, y = (int64_t)s.cell[1] + dy,
                                        z = (int64_t)s.cell[2] + dz;
                                if (x < 0 || y < 0 || z < 0 || x >= (int64_t)g.n[0] || y >= (int64_t)g.n[1] ||
                                    z >= (int64_t)g.n[2])
                                    continue;
                                uint64_t nc[3] = {(uint64_t)x, (uint64_t)y, (uint64_t)z};
                                V3 p = g.center(nc);
                                uint64_t ni = g.idx(nc);
                                float d = F.dist(p, s.tri);
                                locks[ni].lock();  // write lock taken unconditionally (:542)
                                if (F.cmp(d, distances[ni]) < 0) {
                                    distances[ni] = d;
                                    locks[ni].unlock();
                                    pq.push({d, {(uint32_t)x, (uint32_t)y, (uint32_t)z}, s.tri});
                                } else {
                                    locks[ni].unlock();
                                }
                            }
                }
                steps.fetch_add(local, std::memory_order_relaxed);
            });
        }
        for (auto& th : pool) th.join();
    }
    auto t2 = clk::now();
    if (steps_out) steps_out[1] = steps.exchange(0);

    // unwrap locks (:350)
    for (uint64_t i = 0; i < cells; ++i) out[i] = distances[i];

    uint64_t ray_cands = 0;
    if (sign == 0) {
        bvh_thread.join();
        // compute_raycasts (grid.rs:568-642)
        std::unique_ptr<std::atomic<uint32_t>[]> inter(new std::atomic<uint32_t>[cells * 3]);
        for (uint64_t i = 0; i < cells * 3; ++i) inter[i].store(0, std::memory_order_relaxed);
        const uint64_t gx = g.n[0], gy = g.n[1], gz = g.n[2];
        const uint64_t nrays = gy * gz + gx * gz + gx * gy;  // generate_raycasts (:648-684)
        std::atomic<uint64_t> cands{0};
        parallel_for(nrays, threads, 64, [&](uint64_t r) {
            int axis;
            uint64_t sc[3];
            if (r < gy * gz) { axis = 0; sc[0] = 0; sc[1] = r / gz; sc[2] = r % gz; }
            else if (r < gy * gz + gx * gz) { uint64_t q = r - gy * gz; axis = 1; sc[0] = q / gz; sc[1] = 0; sc[2] = q % gz; }
            else { uint64_t q = r - gy * gz - gx * gz; axis = 2; sc[0] = q / gy; sc[1] = q % gy; sc[2] = 0; }
            V3 o = g.center(sc);
            uint64_t local = 0;
            tree.ray_candidates(o, axis, [&](uint32_t t) {
                ++local;
                V3 a = ld3(verts, tris[3 * t]), b = ld3(verts, tris[3 * t + 1]), c = ld3(verts, tris[3 * t + 2]);
                float tt;
                if (ray_triangle_intersection_aligned(o, a, b, c, axis, &tt)) {
                    uint64_t k = row_hit_last_cell(tt, v_get(g.cell_size, axis), g.n[axis]);
                    uint64_t cell[3] = {sc[0], sc[1], sc[2]};
                    for (uint64_t i = 0; i <= k; ++i) {  // :612-617
                        cell[axis] = i;
                        inter[g.idx(cell) * 3 + axis].fetch_add(1, std::memory_order_relaxed);
                    }
                }
            });
            cands.fetch_add(local, std::memory_order_relaxed);
        });
        ray_cands = cands.load();
        for (uint64_t i = 0; i < cells; ++i) {  // serial parity pass (:622-639)
            int odd = (inter[3 * i].load(std::memory_order_relaxed) & 1) +
                      (inter[3 * i + 1].load(std::memory_order_relaxed) & 1) +
                      (inter[3 * i + 2].load(std::memory_order_relaxed) & 1);
            if (odd >= 2) out[i] = -out[i];
        }
    }
    auto t3 = clk::now();
    if (steps_out) steps_out[2] = ray_cands;
    if (phase_ms) {
        phase_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        phase_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
        phase_ms[2] = std::chrono::duration<double, std::milli>(t3 - t2).count();
    }
    return F.nan.load() ? 3 : 0;
}

// ---- tree-accelerated generate_sdf, the timed CPU stand-in for Rtree / RtreeBvh / Bvh --------------
// Same results as m2s_oracle_generate_sdf for accel 2 and 3 and for Raycast sign; the trees only
// filter candidates (rtree.rs:116, rtree_bvh.rs:126-164, bvh.rs:79-134). Normal-sign Bvh is not
// offered here (its near-tie fold needs every near-equal candidate; use the exact driver).
int m2s_oracle_generate_sdf_tree(const float* verts, uint64_t nv, const uint32_t* tris, uint64_t nt,
                                 const float* queries, uint64_t nq, int accel, int sign, int threads,
                                 float* out, double* phase_ms) {
    if (int e = check_indices(tris, nt, nv)) return e;
    if (nt == 0) return 7;
    if (!(accel == 2 || accel == 3 || (accel == 1 && sign == 0))) return 1;
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    Tree tree;
    tree.build(verts, tris, nt);
    auto t1 = clk::now();
    parallel_for(nq, threads, 256, [&](uint64_t i) {
        V3 q = ld3(queries, i);
        float d2;
        uint32_t bt = tree.nearest(verts, tris, q, &d2);
        V3 a = ld3(verts, tris[3 * bt]), b = ld3(verts, tris[3 * bt + 1]), c = ld3(verts, tris[3 * bt + 2]);
        if (accel == 2) {
            out[i] = point_triangle_signed_distance(q, a, b, c);
            return;
        }
        float d = point_triangle_distance(q, a, b, c);
        int insides = 0;
        for (int ax = 0; ax < 3; ++ax) {
            uint32_t cnt = 0;
            tree.ray_candidates(q, ax, [&](uint32_t t) {
                float tt;
                if (ray_triangle_intersection_aligned(q, ld3(verts, tris[3 * t]), ld3(verts, tris[3 * t + 1]),
                                                      ld3(verts, tris[3 * t + 2]), ax, &tt))
                    ++cnt;
            });
            insides += cnt & 1;
        }
        out[i] = insides > 1 ? -d : d;
    });
    auto t2 = clk::now();
    if (phase_ms) {
        phase_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        phase_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    }
    return 0;
}

}  // extern "C"
