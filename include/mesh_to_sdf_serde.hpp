// mesh_to_sdf_serde.hpp — the `mesh_to_sdf::serde` module of the reference (mesh_to_sdf/src/serde.rs) for the C++
// facade: format V1 = rmp_serde::to_vec(&SerializeVersion::V1(sdf)) (serde.rs:156-160), i.e. MessagePack with enum
// variants as one-entry maps keyed by name, structs as arrays in field order, f32 as 0xca + big-endian bits, usize
// in the shortest unsigned form and the shortest array headers:
//     {"V1": {"Generic": [[[x, y, z], ...], [d, ...]]}}                         serde.rs:84-95
//     {"V1": {"Grid": [[[fx, fy, fz], [sx, sy, sz], [nx, ny, nz]], [d, ...]]}}  serde.rs:97-106, src/grid.rs:30-37
// Host-only, header-only, no dependency besides mesh_to_sdf.hpp. Pinned byte for byte by the reference's fixtures
// (mesh_to_sdf/tests/sdf_generic_v1.bin, sdf_grid_v1.bin; tests/cpp/test_serde.cpp).
#pragma once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

#include "mesh_to_sdf.hpp"

namespace mesh_to_sdf {
namespace serde {

// serde.rs:43-52
struct SerdeError : std::runtime_error {
    enum Kind { SerializationFailed, DeserializationFailed, IoError } kind;
    SerdeError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

template <class V>
struct Generic {  // SerializeGeneric / DeserializeGeneric, serde.rs:84-95, :122-131
    std::vector<V> query_points;
    std::vector<float> distances;
};
template <class V>
struct GridSdf {  // SerializeGrid / DeserializeGrid, serde.rs:97-106, :133-142
    Grid<V> grid;
    std::vector<float> distances;
};
template <class V>
using Sdf = std::variant<Generic<V>, GridSdf<V>>;  // SerializeSdf / DeserializeSdf

namespace detail {
inline void put_be(std::vector<uint8_t>& o, uint64_t v, int bytes) {
    for (int i = bytes - 1; i >= 0; --i) o.push_back((uint8_t)(v >> (8 * i)));
}
inline void put_array(std::vector<uint8_t>& o, size_t n) {
    if (n < 16) o.push_back((uint8_t)(0x90 | n));
    else if (n < (1u << 16)) { o.push_back(0xdc); put_be(o, n, 2); }
    else if (n <= 0xffffffffull) { o.push_back(0xdd); put_be(o, n, 4); }
    else throw SerdeError(SerdeError::SerializationFailed, "sequence longer than 2^32-1");
}
inline void put_uint(std::vector<uint8_t>& o, uint64_t v) {
    if (v < 128) o.push_back((uint8_t)v);
    else if (v < (1u << 8)) { o.push_back(0xcc); put_be(o, v, 1); }
    else if (v < (1u << 16)) { o.push_back(0xcd); put_be(o, v, 2); }
    else if (v <= 0xffffffffull) { o.push_back(0xce); put_be(o, v, 4); }
    else { o.push_back(0xcf); put_be(o, v, 8); }
}
inline void put_f32(std::vector<uint8_t>& o, float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    o.push_back(0xca);
    put_be(o, b, 4);
}
inline void put_str(std::vector<uint8_t>& o, const char* s) {
    const size_t n = std::strlen(s);  // variant names: always < 32 bytes
    o.push_back((uint8_t)(0xa0 | n));
    o.insert(o.end(), s, s + n);
}
template <class V>
void put_point(std::vector<uint8_t>& o, const V& p) {
    using T = point_traits<V>;
    o.push_back(0x93);
    put_f32(o, T::x(p)); put_f32(o, T::y(p)); put_f32(o, T::z(p));
}
inline void put_f32_seq(std::vector<uint8_t>& o, const std::vector<float>& a) {
    put_array(o, a.size());
    for (float f : a) put_f32(o, f);
}

struct Reader {
    const uint8_t* p;
    size_t n, i = 0;
    [[noreturn]] static void bad(const char* m) { throw SerdeError(SerdeError::DeserializationFailed, m); }
    uint8_t byte() { if (i >= n) bad("truncated input"); return p[i++]; }
    uint64_t be(int bytes) { uint64_t v = 0; for (int k = 0; k < bytes; ++k) v = (v << 8) | byte(); return v; }
    std::string map1_key() {
        if (byte() != 0x81) bad("expected a one-entry map (enum variant)");
        const uint8_t t = byte();
        size_t len = 0;
        if ((t & 0xe0) == 0xa0) len = t & 0x1f;
        else if (t == 0xd9) len = byte();
        else bad("expected a string key");
        std::string s;
        for (size_t k = 0; k < len; ++k) s.push_back((char)byte());
        return s;
    }
    size_t array() {
        const uint8_t t = byte();
        if ((t & 0xf0) == 0x90) return t & 0x0f;
        if (t == 0xdc) return (size_t)be(2);
        if (t == 0xdd) return (size_t)be(4);
        bad("expected an array");
    }
    void expect_array(size_t k) { if (array() != k) bad("unexpected array length"); }
    uint64_t uint() {
        const uint8_t t = byte();
        if (t < 0x80) return t;
        if (t == 0xcc) return be(1);
        if (t == 0xcd) return be(2);
        if (t == 0xce) return be(4);
        if (t == 0xcf) return be(8);
        bad("expected an unsigned integer");
    }
    float f32() {
        if (byte() != 0xca) bad("expected an f32");
        const uint32_t b = (uint32_t)be(4);
        float f;
        std::memcpy(&f, &b, 4);
        return f;
    }
    template <class V>
    V point() {
        expect_array(3);
        const float x = f32(), y = f32(), z = f32();
        return point_traits<V>::make(x, y, z);
    }
    std::vector<float> f32_seq() {
        const size_t k = array();
        if (k > (n - i) / 5) bad("truncated input");
        std::vector<float> a(k);
        for (auto& f : a) f = f32();
        return a;
    }
};
}  // namespace detail

// serde.rs:156-160: always the latest version (V1)
template <class V>
std::vector<uint8_t> serialize(const Sdf<V>& sdf) {
    using namespace detail;
    std::vector<uint8_t> o;
    o.push_back(0x81);
    put_str(o, "V1");
    o.push_back(0x81);
    if (const auto* g = std::get_if<Generic<V>>(&sdf)) {
        put_str(o, "Generic");
        o.push_back(0x92);
        put_array(o, g->query_points.size());
        for (const V& p : g->query_points) put_point(o, p);
        put_f32_seq(o, g->distances);
    } else {
        const auto& gr = std::get<GridSdf<V>>(sdf);
        put_str(o, "Grid");
        o.push_back(0x92);
        o.push_back(0x93);
        put_point(o, gr.grid.get_first_cell());
        put_point(o, gr.grid.get_cell_size());
        o.push_back(0x93);
        for (size_t c : gr.grid.get_cell_count()) put_uint(o, c);
        put_f32_seq(o, gr.distances);
    }
    return o;
}

// serde.rs:162-170
template <class V>
Sdf<V> deserialize(const uint8_t* data, size_t n) {
    detail::Reader r{data, n};
    if (r.map1_key() != "V1") detail::Reader::bad("unknown format version");
    const std::string kind = r.map1_key();
    r.expect_array(2);
    Sdf<V> out = Generic<V>{};
    if (kind == "Generic") {
        Generic<V> g;
        const size_t k = r.array();
        if (k > (n - r.i) / 16) detail::Reader::bad("truncated input");
        g.query_points.reserve(k);
        for (size_t j = 0; j < k; ++j) g.query_points.push_back(r.point<V>());
        g.distances = r.f32_seq();
        out = std::move(g);
    } else if (kind == "Grid") {
        r.expect_array(3);
        const V first = r.point<V>(), size = r.point<V>();
        r.expect_array(3);
        std::array<size_t, 3> cc;
        for (auto& c : cc) c = (size_t)r.uint();
        GridSdf<V> g{Grid<V>(first, size, cc), r.f32_seq()};
        out = std::move(g);
    } else {
        detail::Reader::bad("unknown variant");
    }
    if (r.i != n) detail::Reader::bad("trailing bytes");
    return out;
}
template <class V>
Sdf<V> deserialize(const std::vector<uint8_t>& data) { return deserialize<V>(data.data(), data.size()); }

// serde.rs:187-193
template <class V>
void save_to_file(const Sdf<V>& sdf, const std::string& path) {
    const auto bytes = serialize(sdf);
    std::ofstream f(path, std::ios::binary);
    if (!f || !f.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size()))
        throw SerdeError(SerdeError::IoError, "cannot write " + path);
}

// serde.rs:217-221
template <class V>
Sdf<V> read_from_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw SerdeError(SerdeError::IoError, "cannot read " + path);
    const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return deserialize<V>(bytes);
}

}  // namespace serde
}  // namespace mesh_to_sdf
