// The reference's serde tests (mesh_to_sdf/src/serde.rs:229-372) against include/mesh_to_sdf_serde.hpp.
// usage: test_serde <dir with sdf_generic_v1.bin and sdf_grid_v1.bin> <scratch dir>
#include <cstdio>
#include <fstream>
#include <iterator>

#include "mesh_to_sdf_serde.hpp"

using namespace mesh_to_sdf;
using V = std::array<float, 3>;
struct Vec3 { float x, y, z; };  // a cgmath::Vector3-like point

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static std::vector<uint8_t> slurp(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string golden = argv[1], tmp = argv[2];
    // test_serde / test_backward_compatibility_serde_generic_v1
    serde::Generic<Vec3> gen{{{1, 2, 3}, {6, 5, 4}}, {1.0f, 3.0f}};
    const auto gen_bytes = serde::serialize<Vec3>(gen);
    CHECK(gen_bytes == slurp(golden + "/sdf_generic_v1.bin"));
    {
        auto de = serde::read_from_file<Vec3>(golden + "/sdf_generic_v1.bin");
        auto* g = std::get_if<serde::Generic<Vec3>>(&de);
        CHECK(g && g->query_points.size() == 2 && g->query_points[1].x == 6 && g->query_points[1].z == 4);
        CHECK(g && g->distances == std::vector<float>({1.0f, 3.0f}));
    }
    // test_serde_grid / test_backward_compatibility_serde_grid_v1
    Grid<V> grid({1, 2, 3}, {4, 5, 6}, {7, 8, 9});
    std::vector<float> dist(grid.get_total_cell_count());
    for (size_t i = 0; i < dist.size(); ++i) dist[i] = (float)i;
    serde::GridSdf<V> gs{grid, dist};
    CHECK(serde::serialize<V>(gs) == slurp(golden + "/sdf_grid_v1.bin"));
    {
        auto de = serde::read_from_file<V>(golden + "/sdf_grid_v1.bin");
        auto* g = std::get_if<serde::GridSdf<V>>(&de);
        CHECK(g && g->grid.get_first_cell() == V({1, 2, 3}) && g->grid.get_cell_size() == V({4, 5, 6}));
        CHECK(g && g->grid.get_cell_count() == (std::array<size_t, 3>{7, 8, 9}) && g->distances == dist);
    }
    // test_serde_file + header widths (array16 / array32, u8 / u16 / u32 counts)
    {
        serde::Generic<V> big;
        for (int i = 0; i < 70000; ++i) { big.query_points.push_back({(float)i, 0.5f, -1.0f}); big.distances.push_back(i * 0.25f); }
        serde::save_to_file<V>(big, tmp + "/sdf.bin");
        auto de = serde::read_from_file<V>(tmp + "/sdf.bin");
        auto* g = std::get_if<serde::Generic<V>>(&de);
        CHECK(g && g->query_points == big.query_points && g->distances == big.distances);
        serde::GridSdf<V> cnt{Grid<V>({0, 0, 0}, {1, 1, 1}, {200, 70000, 5}), {}};
        auto de2 = serde::deserialize<V>(serde::serialize<V>(cnt));
        CHECK(std::get<serde::GridSdf<V>>(de2).grid.get_cell_count() == (std::array<size_t, 3>{200, 70000, 5}));
    }
    // malformed input -> DeserializationFailed; missing file -> IoError
    for (std::vector<uint8_t> bad : {std::vector<uint8_t>{}, {0x93, 1, 2, 3}, {0x81, 0xa2, 'V', '2', 0x81, 0xa4, 'G', 'r', 'i', 'd', 0x92, 0x90, 0x90},
                                     std::vector<uint8_t>(gen_bytes.begin(), gen_bytes.end() - 3)}) {
        try { serde::deserialize<V>(bad); CHECK(false); }
        catch (const serde::SerdeError& e) { CHECK(e.kind == serde::SerdeError::DeserializationFailed); }
    }
    try { serde::read_from_file<V>(tmp + "/missing.bin"); CHECK(false); }
    catch (const serde::SerdeError& e) { CHECK(e.kind == serde::SerdeError::IoError); }
    if (!fails) std::printf("all tests passed\n");
    return fails ? 1 : 0;
}
