"""mesh_to_sdf::serde format V1 (mesh_to_sdf/src/serde.rs): the reference's own fixtures are the golden vectors
(tests/golden/sdf_*_v1.bin are byte copies of mesh_to_sdf/tests/sdf_*_v1.bin, written by the reference at V1)."""
import os

import numpy as np
import pytest

import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import serde


def _generic():
    # serde.rs:231-240 / :315-324
    return serde.Generic(np.array([[1., 2., 3.], [6., 5., 4.]], np.float32), np.array([1.0, 3.0], np.float32))


def _grid():
    # serde.rs:258-262 / :349-353: Grid::new([1,2,3],[4,5,6],[7,8,9]), distances = 0..504 as f32
    g = m2s.Grid([1., 2., 3.], [4., 5., 6.], [7, 8, 9])
    return serde.GridSdf(g, np.arange(7 * 8 * 9, dtype=np.float32))


def test_encoder_reproduces_the_reference_fixtures_byte_for_byte(golden_dir):
    assert serde.serialize(_generic()) == open(os.path.join(golden_dir, "sdf_generic_v1.bin"), "rb").read()
    assert serde.serialize(_grid()) == open(os.path.join(golden_dir, "sdf_grid_v1.bin"), "rb").read()


def test_backward_compatibility_v1(golden_dir):
    # serde.rs:313-372
    de = serde.read_from_file(os.path.join(golden_dir, "sdf_generic_v1.bin"))
    assert isinstance(de, serde.Generic)
    assert np.array_equal(de.query_points, _generic().query_points) and np.array_equal(de.distances, [1.0, 3.0])
    de = serde.read_from_file(os.path.join(golden_dir, "sdf_grid_v1.bin"))
    assert isinstance(de, serde.GridSdf)
    want = _grid()
    assert np.array_equal(de.grid.first_cell, want.grid.first_cell) and np.array_equal(de.grid.cell_size, want.grid.cell_size)
    assert list(de.grid.cell_count) == [7, 8, 9]
    assert np.array_equal(de.distances, want.distances)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 65535, 65536, 200_000])
def test_round_trip_and_array_headers(tmp_path, n):
    # serde.rs:229-311 (test_serde, test_serde_grid, test_serde_file) across the fixarray / array16 / array32 bounds
    rng = np.random.default_rng(n)
    gen = serde.Generic(rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal(n).astype(np.float32))
    path = tmp_path / "sdf.bin"
    serde.save_to_file(gen, path)
    de = serde.read_from_file(path)
    assert np.array_equal(de.query_points.view(np.uint32), gen.query_points.view(np.uint32))
    assert np.array_equal(de.distances.view(np.uint32), gen.distances.view(np.uint32))
    grid = serde.GridSdf(m2s.Grid([0., 0., 0.], [1., 1., 1.], [n, 1, 300]), np.zeros(0, np.float32))
    de = serde.deserialize(serde.serialize(grid))
    assert list(de.grid.cell_count) == [n, 1, 300]  # shortest unsigned encodings: fixint, u8, u16, u32


def test_agrees_with_an_independent_messagepack_decoder():
    msgpack = pytest.importorskip("msgpack")
    doc = msgpack.unpackb(serde.serialize(_grid()))
    assert list(doc) == ["V1"] and list(doc["V1"]) == ["Grid"]
    (first, size, count), dist = doc["V1"]["Grid"]
    assert first == [1.0, 2.0, 3.0] and size == [4.0, 5.0, 6.0] and count == [7, 8, 9] and len(dist) == 504
    # and the other way: a document written by the independent encoder (float32 mode) is read back
    blob = msgpack.packb({"V1": {"Generic": [[[1.0, 2.0, 3.0]], [0.5]]}}, use_single_float=True)
    de = serde.deserialize(blob)
    assert np.array_equal(de.query_points, [[1., 2., 3.]]) and np.array_equal(de.distances, [0.5])


@pytest.mark.parametrize("blob", [b"", b"\x81\xa2V2\x81\xa4Grid\x92\x90\x90", b"\x81\xa2V1\x81\xa3Foo\x92\x90\x90",
                                  b"\x81\xa2V1\x81\xa7Generic\x92\x91\x93\xca\x00\x00", b"\x93\x01\x02\x03"])
def test_malformed_documents_fail(blob):
    with pytest.raises(serde.SerdeError):
        serde.deserialize(blob)


def test_io_error(tmp_path):
    with pytest.raises(serde.SerdeError):
        serde.read_from_file(tmp_path / "missing.bin")


def test_cpp_serde_header_against_the_fixtures(tmp_path, golden_dir):
    # include/mesh_to_sdf_serde.hpp (the C++ facade's serde module): tests/cpp/test_serde.cpp restates serde.rs:229-372.
    # Host-only code: it links libm2s.so (Grid::from_bounding_box lives there) but makes no compute call.
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "mesh_to_sdf_b200")
    exe = str(tmp_path / "test_serde")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "cpp", "test_serde.cpp"), "-o", exe, "-L", libdir, "-l:libm2s.so",
           f"-Wl,-rpath,{libdir}", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, golden_dir, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all tests passed" in r.stdout, r.stdout + r.stderr
