// mesh_to_sdf.hpp — header-only C++17 host facade over the C ABI of m2s.h, mirroring the public API of the Rust
// crate Azkellas/mesh_to_sdf 0.4.0 (mesh_to_sdf/src/lib.rs:146-311, src/grid.rs:30-170): same names, argument
// meaning and error behaviour. The reference's toolchain (Rust) is absent from this build environment, so this is
// the compiled-language host side that is actually built and tested here; the Rust facade with the identical
// surface is in rust/mesh_to_sdf.
//
//   Rust                                   C++ (namespace mesh_to_sdf)
//   generate_sdf(&v, Topology, &q, accel)  generate_sdf(v, topology, q, accel)         -> Distances (a std::vector<float>
//                                                                                        that is not zero-filled first)
//   generate_grid_sdf(&v, Topology, &g, s) generate_grid_sdf(v, topology, grid, sign)  -> Distances
//   panic!                                 throws mesh_to_sdf::Panic
//   trait Point                            any type with .x .y .z floats or operator[] (see point_traits)
// Extensions that have no counterpart in the crate (they remove copies the GPU path would otherwise pay):
//   PinnedVec + generate_grid_sdf_into     the result lands in page-locked memory, written by the kernel itself
//   Mesh                                   upload + LBVH once, then any number of grids / query sets
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "m2s.h"

namespace mesh_to_sdf {

struct Panic : std::runtime_error {
    int status;
    Panic(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

namespace detail {
// a + n * b and a - b * c with the product rounded before the sum: rustc never contracts to FMA, a host compiler
// may (GCC's default -ffp-contract=fast on aarch64), and Grid must stay bit-identical to src/grid.rs
inline float mul_then_add(float a, float n, float b) {
    volatile float prod = n * b;
    return a + prod;
}
inline float mul_then_sub(float a, float b, float c) {
    volatile float prod = b * c;
    return a - prod;
}
}  // namespace detail

// ---- Point (src/point.rs:21-62): constructor + x/y/z ------------------------------------------------------------
template <class V, class = void>
struct point_traits {  // default: indexable (std::array<float,3>, float[3]-likes)
    static V make(float x, float y, float z) { return V{x, y, z}; }
    static float x(const V& v) { return v[0]; }
    static float y(const V& v) { return v[1]; }
    static float z(const V& v) { return v[2]; }
};
template <class V>
struct point_traits<V, std::void_t<decltype(std::declval<V>().x), decltype(std::declval<V>().z)>> {  // .x .y .z
    static V make(float x, float y, float z) { return V{x, y, z}; }
    static float x(const V& v) { return v.x; }
    static float y(const V& v) { return v.y; }
    static float z(const V& v) { return v.z; }
};

// ---- enums (declaration order of lib.rs:204-239) ----------------------------------------------------------------
enum class SignMethod { Raycast = 0, Normal = 1 };

struct AccelerationMethod {
    enum Kind { None = 0, Bvh = 1, Rtree = 2, RtreeBvh = 3 } kind = RtreeBvh;  // #[default] RtreeBvh
    SignMethod sign = SignMethod::Raycast;
    static AccelerationMethod none(SignMethod s) { return {None, s}; }
    static AccelerationMethod bvh(SignMethod s) { return {Bvh, s}; }
    static AccelerationMethod rtree() { return {Rtree, SignMethod::Raycast}; }
    static AccelerationMethod rtree_bvh() { return {RtreeBvh, SignMethod::Raycast}; }
};

// ---- Topology (lib.rs:151-193) ------------------------------------------------------------------------------------
template <class I>
struct Topology {
    enum Kind { TriangleList = 0, TriangleStrip = 1 } kind;
    const I* indices;  // nullptr == None: 0..vertices.len()
    size_t count;
    static Topology triangle_list(const std::vector<I>& idx) { return {TriangleList, idx.data(), idx.size()}; }
    static Topology triangle_strip(const std::vector<I>& idx) { return {TriangleStrip, idx.data(), idx.size()}; }
    static Topology triangle_list() { return {TriangleList, nullptr, 0}; }
    static Topology triangle_strip() { return {TriangleStrip, nullptr, 0}; }
    std::vector<uint32_t> get_triangles(size_t n_vertices) const {
        static_assert(sizeof(I) == 2 || sizeof(I) == 4, "I must be u16 or u32");
        const uint64_t nt = m2s_expand_topology(kind, indices, (int)sizeof(I), count, n_vertices, nullptr);
        std::vector<uint32_t> out(nt * 3);
        if (nt) m2s_expand_topology(kind, indices, (int)sizeof(I), count, n_vertices, out.data());
        return out;
    }
};

// ---- Grid / SnapResult (src/grid.rs) ------------------------------------------------------------------------------
struct SnapResult {
    bool inside;
    std::array<size_t, 3> cell;
    bool operator==(const SnapResult& o) const { return inside == o.inside && cell == o.cell; }
};

template <class V>
class Grid {
    using T = point_traits<V>;
    V first_cell_, cell_size_;
    std::array<size_t, 3> cell_count_;

  public:
    Grid(V first_cell, V cell_size, std::array<size_t, 3> cell_count)
        : first_cell_(first_cell), cell_size_(cell_size), cell_count_(cell_count) {}
    static Grid from_bounding_box(const V& bbox_min, const V& bbox_max, std::array<size_t, 3> cell_count) {
        const float mn[3] = {T::x(bbox_min), T::y(bbox_min), T::z(bbox_min)};
        const float mx[3] = {T::x(bbox_max), T::y(bbox_max), T::z(bbox_max)};
        const uint64_t cc[3] = {cell_count[0], cell_count[1], cell_count[2]};
        float first[3], size[3];
        m2s_grid_from_bounding_box(mn, mx, cc, first, size);
        return Grid(T::make(first[0], first[1], first[2]), T::make(size[0], size[1], size[2]), cell_count);
    }
    V get_first_cell() const { return first_cell_; }
    V get_cell_size() const { return cell_size_; }
    std::array<size_t, 3> get_cell_count() const { return cell_count_; }
    size_t get_total_cell_count() const { return cell_count_[0] * cell_count_[1] * cell_count_[2]; }
    V get_last_cell() const {
        return T::make(detail::mul_then_add(T::x(first_cell_), (float)cell_count_[0], T::x(cell_size_)),
                       detail::mul_then_add(T::y(first_cell_), (float)cell_count_[1], T::y(cell_size_)),
                       detail::mul_then_add(T::z(first_cell_), (float)cell_count_[2], T::z(cell_size_)));
    }
    std::pair<V, V> get_bounding_box() const {
        const float lo[3] = {detail::mul_then_sub(T::x(first_cell_), T::x(cell_size_), 0.5f),
                             detail::mul_then_sub(T::y(first_cell_), T::y(cell_size_), 0.5f),
                             detail::mul_then_sub(T::z(first_cell_), T::z(cell_size_), 0.5f)};
        return {T::make(lo[0], lo[1], lo[2]),
                T::make(detail::mul_then_add(lo[0], (float)cell_count_[0], T::x(cell_size_)),
                        detail::mul_then_add(lo[1], (float)cell_count_[1], T::y(cell_size_)),
                        detail::mul_then_add(lo[2], (float)cell_count_[2], T::z(cell_size_)))};
    }
    size_t get_cell_idx(const std::array<size_t, 3>& c) const {
        return c[2] + cell_count_[2] * (c[1] + cell_count_[1] * c[0]);
    }
    std::array<size_t, 3> get_cell_integer_coordinates(size_t idx) const {
        return {idx / (cell_count_[1] * cell_count_[2]), (idx / cell_count_[2]) % cell_count_[1], idx % cell_count_[2]};
    }
    V get_cell_center(const std::array<size_t, 3>& c) const {
        return T::make(detail::mul_then_add(T::x(first_cell_), (float)c[0], T::x(cell_size_)),
                       detail::mul_then_add(T::y(first_cell_), (float)c[1], T::y(cell_size_)),
                       detail::mul_then_add(T::z(first_cell_), (float)c[2], T::z(cell_size_)));
    }
    SnapResult snap_point_to_grid(const V& p) const {
        const V lo = get_bounding_box().first;
        const float q[3] = {(T::x(p) - T::x(lo)) / T::x(cell_size_), (T::y(p) - T::y(lo)) / T::y(cell_size_),
                            (T::z(p) - T::z(lo)) / T::z(cell_size_)};
        SnapResult r{true, {0, 0, 0}};
        for (int i = 0; i < 3; ++i) {
            const float f = std::floor(q[i]);
            long long raw = f != f ? 0 : (f >= 9.2e18f ? std::numeric_limits<long long>::max()
                                                       : (f <= -9.2e18f ? std::numeric_limits<long long>::min() : (long long)f));
            long long hi = (long long)cell_count_[i] - 1;
            long long c = raw < 0 ? 0 : (raw > hi ? hi : raw);
            if (c != raw) r.inside = false;
            r.cell[i] = (size_t)c;
        }
        return r;
    }
};

// ---- process-global context (the Rust facade keeps a OnceLock<Mutex<..>> the same way) --------------------------
namespace detail {
inline m2s_ctx* context() {
    static m2s_ctx* ctx = [] {
        std::vector<int> devices;
        if (const char* e = std::getenv("M2S_DEVICES")) {
            std::string s(e);
            size_t pos = 0;
            while (pos < s.size()) {
                size_t next = s.find(',', pos);
                if (next == std::string::npos) next = s.size();
                if (next > pos) devices.push_back(std::atoi(s.substr(pos, next - pos).c_str()));
                pos = next + 1;
            }
        }
        m2s_ctx* c = nullptr;
        const m2s_status rc = m2s_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &c);
        if (rc != M2S_OK) throw Panic(rc, "mesh_to_sdf: no usable CUDA device; libm2s has no CPU fallback");
        return c;
    }();
    return ctx;
}
// one facade-wide lock around a call and the read of its error text: the free functions stay callable from any thread
// and a panic always carries the message of its own call
inline std::mutex& call_mutex() {
    static std::mutex mu;
    return mu;
}
inline void check(m2s_ctx* c, m2s_status rc) {
    if (rc == M2S_OK) return;
    char buf[512];
    buf[0] = '\0';
    m2s_last_error_copy(c, buf, sizeof buf);
    const std::string msg = buf;
    if (rc == M2S_ENAN) throw Panic(rc, "NaN distance (" + msg + ")");           // lib.rs:257
    if (rc == M2S_EINDEX) throw Panic(rc, "index out of bounds (" + msg + ")");  // slice index panic
    if (rc == M2S_EEMPTY) throw Panic(rc, "called `Option::unwrap()` on a `None` value (" + msg + ")");  // rtree.rs:117
    throw Panic(rc, "mesh_to_sdf backend error: " + msg);
}
// std::vector<float>(n) zero-fills its n floats on the calling thread: for a 256^3 grid that is 64 MiB of first-touch
// page faults - several times the GPU's whole call - spent on values libm2s overwrites one and all. The results of this
// facade are therefore vectors whose allocator default-initialises (no fill), which is what the reference's
// `vec![0.0; n]` costs too (alloc_zeroed hands out untouched pages): the pages are first touched by the library's copy
// threads, in parallel and while the GPU still computes.
template <class T>
struct default_init_allocator : std::allocator<T> {
    default_init_allocator() = default;
    template <class U>
    default_init_allocator(const default_init_allocator<U>&) noexcept {}
    template <class U>
    struct rebind {
        using other = default_init_allocator<U>;
    };
    template <class U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};

template <class V>
std::vector<float> pack(const std::vector<V>& v) {
    std::vector<float> out;
    out.reserve(v.size() * 3);
    for (const V& p : v) {
        out.push_back(point_traits<V>::x(p));
        out.push_back(point_traits<V>::y(p));
        out.push_back(point_traits<V>::z(p));
    }
    return out;
}
}  // namespace detail

// What generate_sdf / generate_grid_sdf return: a std::vector of f32 in every respect but one - resizing it does not
// zero-fill (see detail::default_init_allocator).
using Distances = std::vector<float, detail::default_init_allocator<float>>;

// generate_sdf(vertices, indices, query_points, acceleration_method) -> Vec<f32>       (lib.rs:291-311)
template <class V, class I>
Distances generate_sdf(const std::vector<V>& vertices, Topology<I> indices, const std::vector<V>& query_points,
                                AccelerationMethod method = {}) {
    const std::vector<uint32_t> tris = indices.get_triangles(vertices.size());
    if (tris.empty() && method.kind == AccelerationMethod::RtreeBvh) return {};  // rtree_bvh.rs:104-106
    const std::vector<float> v = detail::pack(vertices), q = detail::pack(query_points);
    Distances out(query_points.size());
    m2s_ctx* c = detail::context();
    std::lock_guard<std::mutex> lock(detail::call_mutex());
    detail::check(c, m2s_generate_sdf(c, v.data(), vertices.size(), tris.data(), tris.size() / 3, q.data(),
                                      query_points.size(), (int)method.kind, (int)method.sign, out.data()));
    return out;
}

// generate_grid_sdf(vertices, indices, grid, sign_method) -> Vec<f32>                   (generate/grid.rs:265-378)
template <class V, class I>
Distances generate_grid_sdf(const std::vector<V>& vertices, Topology<I> indices, const Grid<V>& grid,
                            SignMethod sign_method = SignMethod::Raycast) {
    using T = point_traits<V>;
    const std::vector<uint32_t> tris = indices.get_triangles(vertices.size());
    const std::vector<float> v = detail::pack(vertices);
    const V f = grid.get_first_cell(), s = grid.get_cell_size();
    const float first[3] = {T::x(f), T::y(f), T::z(f)}, size[3] = {T::x(s), T::y(s), T::z(s)};
    const auto n = grid.get_cell_count();
    const uint64_t count[3] = {n[0], n[1], n[2]};
    // a pageable Vec like the reference's (generate/grid.rs:376): libm2s fills it from its pinned ring with host
    // threads while the kernel is still running (m2s_timings.host_path == M2S_PATH_PIPELINED for >= 4 MiB)
    Distances out(grid.get_total_cell_count());
    m2s_ctx* c = detail::context();
    std::lock_guard<std::mutex> lock(detail::call_mutex());
    detail::check(c, m2s_generate_grid_sdf(c, v.data(), vertices.size(), tris.data(), tris.size() / 3, first, size, count,
                                           (int)sign_method, out.data()));
    return out;
}

// ---- extensions: page-locked results and mesh handles -----------------------------------------------------------

// A Vec<f32>-like buffer in page-locked, mapped host memory (m2s_host_alloc). generate_grid_sdf_into writes it in
// place: the distance kernel's own stores cross PCIe while it computes, so there is no staging buffer and no copy.
class PinnedVec {
    float* p_ = nullptr;
    size_t n_ = 0;

  public:
    PinnedVec() = default;
    explicit PinnedVec(size_t n) { resize(n); }
    PinnedVec(const PinnedVec&) = delete;
    PinnedVec& operator=(const PinnedVec&) = delete;
    PinnedVec(PinnedVec&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    PinnedVec& operator=(PinnedVec&& o) noexcept {
        if (this != &o) { m2s_host_free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~PinnedVec() { m2s_host_free(p_); }
    void resize(size_t n) {  // contents are not preserved
        if (n == n_) return;
        m2s_host_free(p_);
        p_ = nullptr;
        n_ = 0;
        void* q = nullptr;
        if (m2s_host_alloc(n * sizeof(float), &q) != M2S_OK) throw Panic(M2S_ECUDA, "mesh_to_sdf: page-locked allocation failed");
        p_ = static_cast<float*>(q);
        n_ = n;
    }
    float* data() { return p_; }
    const float* data() const { return p_; }
    size_t size() const { return n_; }
    float& operator[](size_t i) { return p_[i]; }
    const float& operator[](size_t i) const { return p_[i]; }
    const float* begin() const { return p_; }
    const float* end() const { return p_ + n_; }
};

// generate_grid_sdf into a caller-owned PinnedVec (resized to the grid). Same values as generate_grid_sdf.
template <class V, class I>
void generate_grid_sdf_into(const std::vector<V>& vertices, Topology<I> indices, const Grid<V>& grid, SignMethod sign_method,
                            PinnedVec& out) {
    using T = point_traits<V>;
    const std::vector<uint32_t> tris = indices.get_triangles(vertices.size());
    const std::vector<float> v = detail::pack(vertices);
    const V f = grid.get_first_cell(), s = grid.get_cell_size();
    const float first[3] = {T::x(f), T::y(f), T::z(f)}, size[3] = {T::x(s), T::y(s), T::z(s)};
    const auto n = grid.get_cell_count();
    const uint64_t count[3] = {n[0], n[1], n[2]};
    out.resize(grid.get_total_cell_count());
    m2s_ctx* c = detail::context();
    std::lock_guard<std::mutex> lock(detail::call_mutex());
    detail::check(c, m2s_generate_grid_sdf(c, v.data(), vertices.size(), tris.data(), tris.size() / 3, first, size, count,
                                           (int)sign_method, out.data()));
}

// How the result of the last grid call reached the host (m2s_host_path_taken) and its phase timings.
inline m2s_timings last_timings() {
    m2s_timings t{};
    m2s_last_timings(detail::context(), &t);
    return t;
}

// A mesh uploaded once with its LBVH (m2s_mesh): the reference rebuilds its trees per call; its viewer regenerates
// the grid of ONE mesh on every parameter change (mesh_to_sdf_client/src/sdf_program.rs:679-721).
template <class V>
class Mesh {
    m2s_mesh* h_ = nullptr;

  public:
    template <class I>
    Mesh(const std::vector<V>& vertices, Topology<I> indices) {
        const std::vector<uint32_t> tris = indices.get_triangles(vertices.size());
        const std::vector<float> v = detail::pack(vertices);
        m2s_ctx* c = detail::context();
        std::lock_guard<std::mutex> lock(detail::call_mutex());
        detail::check(c, m2s_mesh_create(c, v.data(), vertices.size(), tris.data(), tris.size() / 3, &h_));
    }
    Mesh(const Mesh&) = delete;
    Mesh& operator=(const Mesh&) = delete;
    ~Mesh() { m2s_mesh_destroy(h_); }
    Distances generate_grid_sdf(const Grid<V>& grid, SignMethod sign_method = SignMethod::Raycast) const {
        using T = point_traits<V>;
        const V f = grid.get_first_cell(), s = grid.get_cell_size();
        const float first[3] = {T::x(f), T::y(f), T::z(f)}, size[3] = {T::x(s), T::y(s), T::z(s)};
        const auto n = grid.get_cell_count();
        const uint64_t count[3] = {n[0], n[1], n[2]};
        Distances out(grid.get_total_cell_count());
        m2s_ctx* c = detail::context();
        std::lock_guard<std::mutex> lock(detail::call_mutex());
        detail::check(c, m2s_mesh_grid_sdf(c, h_, first, size, count, (int)sign_method, 0, n[0], out.data()));
        return out;
    }
    Distances generate_sdf(const std::vector<V>& query_points, AccelerationMethod method = {}) const {
        const std::vector<float> q = detail::pack(query_points);
        Distances out(query_points.size());
        m2s_ctx* c = detail::context();
        std::lock_guard<std::mutex> lock(detail::call_mutex());
        const m2s_status rc = m2s_mesh_sdf(c, h_, q.data(), query_points.size(), (int)method.kind, (int)method.sign, out.data());
        if (rc == M2S_EEMPTY && method.kind == AccelerationMethod::RtreeBvh) return {};  // rtree_bvh.rs:104-106
        detail::check(c, rc);
        return out;
    }
};

// ---- post-passes on a finished grid: what the reference's in-repo caller does next ------------------------------
// (not part of the crate's API: mesh_to_sdf_client/src/sdf.rs:62-68, :123 and shaders/draw_raymarching.wgsl:118-200)

struct GridOrder {
    std::vector<uint32_t> ordered_indices;  // (0..n).sorted_by(|i, j| data[i].total_cmp(&data[j]))
    float min, max;                         // data.iter().copied().minmax()
};
template <class A>
GridOrder grid_order(const std::vector<float, A>& sdf) {
    GridOrder r{std::vector<uint32_t>(sdf.size()), 0.0f, 0.0f};
    float mm[2] = {0.0f, 0.0f};
    m2s_ctx* c = detail::context();
    std::lock_guard<std::mutex> lock(detail::call_mutex());
    detail::check(c, m2s_grid_order(c, sdf.data(), sdf.size(), r.ordered_indices.data(), mm));
    r.min = mm[0];
    r.max = mm[1];
    return r;
}

enum class SampleMode { Snap = 0, Trilinear = 1, Tetrahedral = 2 };  // raymarch_mode of draw_raymarching.wgsl

// sdf_grid(position, iso) at every point: 100.0 outside [first_cell, last_cell], else the snapped / interpolated
// grid value minus iso — the "distance from any point with interpolation" of the TODO at src/grid.rs:172.
template <class V, class A>
Distances sample_grid_sdf(const std::vector<float, A>& sdf, const Grid<V>& grid, const std::vector<V>& points,
                          SampleMode mode = SampleMode::Trilinear, float iso = 0.0f) {
    using T = point_traits<V>;
    if (sdf.size() != grid.get_total_cell_count()) throw Panic(M2S_EINVAL, "sdf length does not match the grid");
    const V f = grid.get_first_cell(), s = grid.get_cell_size();
    const float first[3] = {T::x(f), T::y(f), T::z(f)}, size[3] = {T::x(s), T::y(s), T::z(s)};
    const auto n = grid.get_cell_count();
    const uint64_t count[3] = {n[0], n[1], n[2]};
    const std::vector<float> p = detail::pack(points);
    Distances out(points.size());
    m2s_ctx* c = detail::context();
    std::lock_guard<std::mutex> lock(detail::call_mutex());
    detail::check(c, m2s_sample_grid_sdf(c, sdf.data(), first, size, count, p.data(), points.size(), (int)mode, iso,
                                         out.data()));
    return out;
}

}  // namespace mesh_to_sdf
