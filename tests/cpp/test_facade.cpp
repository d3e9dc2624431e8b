// The reference's doc-tests and cross-implementation unit tests, re-expressed against the C++ facade
// (include/mesh_to_sdf.hpp) so they read like the crate's own tests. Built by tests/test_gpu_cpp_facade.py with
// g++ and linked against libm2s.so; needs a GPU.
#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "mesh_to_sdf.hpp"

using namespace mesh_to_sdf;
using V3 = std::array<float, 3>;
struct Vec3 { float x, y, z; };  // a user type with fields, like glam::Vec3

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static void doc_generate_sdf() {  // lib.rs:13-31
    std::vector<V3> vertices = {{0.5f, 1.5f, 0.5f}, {1.f, 2.f, 3.f}, {1.f, 3.f, 7.f}};
    std::vector<uint32_t> indices = {0, 1, 2};
    std::vector<V3> query_points = {{0.5f, 0.5f, 0.5f}};
    auto sdf = generate_sdf(vertices, Topology<uint32_t>::triangle_list(indices), query_points, AccelerationMethod::rtree_bvh());
    CHECK(sdf.size() == 1 && sdf[0] == 1.0f);
    for (auto m : {AccelerationMethod::none(SignMethod::Raycast), AccelerationMethod::none(SignMethod::Normal),
                   AccelerationMethod::bvh(SignMethod::Raycast), AccelerationMethod::bvh(SignMethod::Normal), AccelerationMethod::rtree()})
        CHECK(generate_sdf(vertices, Topology<uint32_t>::triangle_list(indices), query_points, m)[0] == 1.0f);
}

static void doc_generate_grid_sdf() {  // lib.rs:34-58, generate/grid.rs:205-231
    std::vector<V3> vertices = {{0.5f, 1.5f, 0.5f}, {1.f, 2.f, 3.f}, {1.f, 3.f, 7.f}};
    std::vector<uint32_t> indices = {0, 1, 2};
    auto grid = Grid<V3>::from_bounding_box({0.f, 0.f, 0.f}, {10.f, 10.f, 10.f}, {10, 10, 10});
    auto sdf = generate_grid_sdf(vertices, Topology<uint32_t>::triangle_list(indices), grid, SignMethod::Raycast);
    CHECK(sdf.size() == 1000 && sdf[0] == 1.0f);
    for (size_t x = 0; x < 10; ++x)
        for (size_t y = 0; y < 10; ++y)
            for (size_t z = 0; z < 10; ++z) {
                size_t i = grid.get_cell_idx({x, y, z});
                CHECK((grid.get_cell_integer_coordinates(i) == std::array<size_t, 3>{x, y, z}));
                CHECK(std::isfinite(sdf[i]));
            }
}

static void test_generate_grid() {  // generate/grid.rs:692-724: grid == generate_sdf(None(Raycast)), assert_eq!
    std::vector<Vec3> vertices = {{0.f, 1.f, 0.f}, {1.f, 2.f, 3.f}, {1.f, 3.f, 4.f}, {2.f, 0.f, 0.f}};
    std::vector<uint16_t> indices = {0, 1, 2, 1, 2, 3};
    auto grid = Grid<Vec3>::from_bounding_box({0.f, 0.f, 0.f}, {5.f, 5.f, 5.f}, {5, 5, 5});
    std::vector<Vec3> query_points;
    for (size_t x = 0; x < 5; ++x)
        for (size_t y = 0; y < 5; ++y)
            for (size_t z = 0; z < 5; ++z) query_points.push_back(grid.get_cell_center({x, y, z}));
    auto sdf = generate_sdf(vertices, Topology<uint16_t>::triangle_list(indices), query_points, AccelerationMethod::none(SignMethod::Raycast));
    auto grid_sdf = generate_grid_sdf(vertices, Topology<uint16_t>::triangle_list(indices), grid, SignMethod::Raycast);
    CHECK(sdf.size() == grid_sdf.size());
    for (size_t i = 0; i < sdf.size(); ++i) CHECK(sdf[i] == grid_sdf[i]);
}

static void test_grid_type() {  // grid.rs:180-297
    auto g = Grid<V3>::from_bounding_box({-1.f, 0.f, 1.f}, {0.f, 2.f, 5.f}, {2, 2, 2});
    CHECK((g.get_first_cell() == V3{-0.75f, 0.5f, 2.f}) && (g.get_cell_size() == V3{0.5f, 1.f, 2.f}));
    CHECK((g.get_bounding_box().first == V3{-1.f, 0.f, 1.f}) && (g.get_bounding_box().second == V3{0.f, 2.f, 5.f}));
    Grid<V3> h({0.f, 1.f, 2.f}, {1.f, 2.f, 3.f}, {10, 20, 30});
    CHECK((h.get_last_cell() == V3{10.f, 41.f, 92.f}));
    auto s = Grid<V3>::from_bounding_box({0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}, {2, 2, 2});
    CHECK((s.snap_point_to_grid({0.4f, 0.8f, 0.1f}) == SnapResult{true, {0, 1, 0}}));
    CHECK((s.snap_point_to_grid({-0.5f, 0.8f, 0.8f}) == SnapResult{false, {0, 1, 1}}));
    CHECK((s.snap_point_to_grid({0.8f, 1.5f, 0.8f}) == SnapResult{false, {1, 1, 1}}));
    auto t = Grid<V3>::from_bounding_box({0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}, {2, 3, 4});
    CHECK(t.get_cell_idx({0, 1, 1}) == 5 && t.get_cell_idx({1, 0, 0}) == 12 && t.get_cell_idx({1, 1, 1}) == 17);
}

static void test_panics() {
    std::vector<V3> vertices = {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}};
    std::vector<V3> q = {{0.f, 0.f, 1.f}};
    std::vector<uint32_t> bad = {0, 1, 9};
    bool threw = false;
    try { generate_sdf(vertices, Topology<uint32_t>::triangle_list(bad), q); } catch (const Panic& p) { threw = p.status == M2S_EINDEX; }
    CHECK(threw);
    std::vector<V3> none;
    std::vector<uint32_t> empty;
    CHECK(generate_sdf(none, Topology<uint32_t>::triangle_list(empty), q, AccelerationMethod::rtree_bvh()).empty());  // rtree_bvh.rs:104-106
    threw = false;
    try { generate_sdf(none, Topology<uint32_t>::triangle_list(empty), q, AccelerationMethod::rtree()); } catch (const Panic& p) { threw = p.status == M2S_EEMPTY; }
    CHECK(threw);
    CHECK(generate_sdf(vertices, Topology<uint32_t>::triangle_list(), q)[0] == 1.0f);  // TriangleList(None), default method
}

// the caller-side post-passes (mesh_to_sdf_client/src/sdf.rs:62-68, :123; draw_raymarching.wgsl:118-200)
static void test_post_passes() {
    std::vector<V3> vertices = {{0.5f, 1.5f, 0.5f}, {1.f, 2.f, 3.f}, {1.f, 3.f, 7.f}};
    std::vector<uint32_t> indices = {0, 1, 2};
    auto grid = Grid<V3>::from_bounding_box({0.f, 0.f, 0.f}, {10.f, 10.f, 10.f}, {6, 5, 4});
    const auto sdf = generate_grid_sdf(vertices, Topology<uint32_t>::triangle_list(indices), grid, SignMethod::Raycast);
    const GridOrder o = grid_order(sdf);
    CHECK(o.ordered_indices.size() == sdf.size());
    for (size_t i = 1; i < sdf.size(); ++i) {
        const float a = sdf[o.ordered_indices[i - 1]], b = sdf[o.ordered_indices[i]];
        CHECK(a < b || (a == b && o.ordered_indices[i - 1] < o.ordered_indices[i]));  // sorted, stable
    }
    CHECK(o.min == sdf[o.ordered_indices.front()] && o.max == sdf[o.ordered_indices.back()]);
    // every mode returns the cell's own value at a cell centre, 100 outside the grid
    std::vector<V3> pts = {grid.get_cell_center({2, 3, 1}), {-5.f, 0.f, 0.f}};
    for (SampleMode m : {SampleMode::Snap, SampleMode::Trilinear, SampleMode::Tetrahedral}) {
        const auto s = sample_grid_sdf(sdf, grid, pts, m);
        CHECK(std::fabs(s[0] - sdf[grid.get_cell_idx({2, 3, 1})]) <= 1e-6f && s[1] == 100.0f);
    }
}

// How results reach the host: a pageable Vec of >= 4 MiB is filled through the pinned ring while the kernel runs,
// a PinnedVec is written in place, a Mesh handle pays neither upload nor build. All three give the same bits.
static void test_host_paths_and_handles() {
    std::vector<V3> vertices;
    std::vector<uint32_t> indices;
    const int nu = 48, nv = 32;  // a torus, outward wound
    for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nv; ++j) {
            const float u = 6.2831853f * i / nu, v = 6.2831853f * j / nv;
            vertices.push_back({(1.f + 0.35f * std::cos(v)) * std::cos(u), (1.f + 0.35f * std::cos(v)) * std::sin(u), 0.35f * std::sin(v)});
        }
    for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nv; ++j) {
            const uint32_t a = i * nv + j, b = ((i + 1) % nu) * nv + j, c = ((i + 1) % nu) * nv + (j + 1) % nv, d = i * nv + (j + 1) % nv;
            for (uint32_t k : {a, b, c, a, c, d}) indices.push_back(k);
        }
    auto grid = Grid<V3>::from_bounding_box({-1.7f, -1.6f, -0.6f}, {1.75f, 1.65f, 0.66f}, {112, 104, 96});  // 4.3 MiB
    const auto topo = Topology<uint32_t>::triangle_list(indices);
    const auto sdf = generate_grid_sdf(vertices, topo, grid, SignMethod::Raycast);
    CHECK(last_timings().host_path == M2S_PATH_PIPELINED);  // the drop-in Vec<f32> path
    // the result is a std::vector<float> that was never zero-filled by the calling thread (its pages are first touched
    // by the library's copy threads); it converts to a plain vector by range
    static_assert(std::is_same<std::remove_const_t<decltype(sdf)>, Distances>::value, "facade result type");
    const std::vector<float> plain(sdf.begin(), sdf.end());
    CHECK(plain.size() == grid.get_total_cell_count() && plain.back() == sdf.back());
    PinnedVec pinned;
    generate_grid_sdf_into(vertices, topo, grid, SignMethod::Raycast, pinned);
    CHECK(last_timings().host_path == M2S_PATH_ZEROCOPY && pinned.size() == sdf.size());
    CHECK(std::memcmp(pinned.data(), sdf.data(), sdf.size() * sizeof(float)) == 0);
    Mesh<V3> mesh(vertices, topo);
    const auto again = mesh.generate_grid_sdf(grid, SignMethod::Raycast);
    const m2s_timings t = last_timings();
    CHECK(t.build_ms < 0.01f && t.h2d_ms < 0.01f && t.host_path == M2S_PATH_PIPELINED);
    CHECK(std::memcmp(again.data(), sdf.data(), sdf.size() * sizeof(float)) == 0);
    size_t inside = 0;
    for (float d : sdf) inside += d < 0.f;
    CHECK(inside > sdf.size() / 20 && inside < sdf.size() / 3);
    // generic positions: axis rays from (1, 0, 0) would run along the mesh's u = 0 seam and graze its vertices
    std::vector<V3> q = {{0.97f, 0.11f, 0.04f}, {0.03f, 0.02f, 0.01f}, {3.f, 0.1f, 0.05f}};
    const auto d = mesh.generate_sdf(q);
    const auto e = generate_sdf(vertices, topo, q);
    CHECK(d.size() == 3 && d[0] < 0.f && d[1] > 0.f && d[2] > 0.f);
    for (size_t i = 0; i < 3; ++i) CHECK(d[i] == e[i]);
}

int main() {
    test_host_paths_and_handles();
    test_post_passes();
    doc_generate_sdf();
    doc_generate_grid_sdf();
    test_generate_grid();
    test_grid_type();
    test_panics();
    std::printf("cpp facade: all tests passed\n");
    return 0;
}
