"""Post-passes on a finished grid: render order + iso limits (mesh_to_sdf_client/src/sdf.rs:
62-68, :123) and sdf_grid() sampling (shaders/draw_raymarching.wgsl:118-200).

CPU part: the numpy oracle against hand-made known answers of the std / itertools semantics it restates.
GPU part (``-m gpu``): the CUDA kernels through the C ABI against that oracle, bit for bit."""
import numpy as np
import pytest

from mesh_to_sdf_b200 import synth
from oracle import post


# ---- oracle pinned by known answers (CPU) ------------------------------------------------------------------
def test_total_order_known_answers():
    # core::f32::total_cmp: -NaN < -inf < -1 < -0 < +0 < 1 < +inf < NaN
    neg_nan = np.array([0xffc00000], np.uint32).view(np.float32)[0]
    a = np.array([np.nan, np.inf, 1.0, 0.0, -0.0, -1.0, -np.inf, neg_nan], np.float32)
    assert list(post.grid_order(a)) == [7, 6, 5, 4, 3, 2, 1, 0]


def test_order_is_stable_and_minmax_first_last():
    a = np.array([2.0, -1.0, 2.0, -1.0, 0.0, -0.0, 5.0, 5.0], np.float32)
    assert list(post.grid_order(a)) == [1, 3, 5, 4, 0, 2, 6, 7]  # equal keys keep index order; -0 before +0
    lo, hi = post.minmax(a)
    assert lo == -1.0 and hi == 5.0
    # PartialOrd: -0.0 == +0.0, so the first minimum / last maximum decides which zero comes back
    z = np.array([0.0, -0.0], np.float32)
    lo, hi = post.minmax(z)
    assert not np.signbit(lo) and np.signbit(hi)


def _affine_grid(count, first, size, coef):
    x, y, z = np.meshgrid(*[np.arange(c) for c in count], indexing="ij")
    pos = [first[k] + v * size[k] for k, v in enumerate((x, y, z))]
    return (coef[0] * pos[0] + coef[1] * pos[1] + coef[2] * pos[2] + coef[3]).astype(np.float32).reshape(-1)


def test_sample_oracle_properties():
    count, first, size = [6, 5, 7], np.array([0.5, -1.0, 2.0], np.float32), np.array([0.25, 0.5, 0.125], np.float32)
    rng = np.random.default_rng(7)
    sdf = rng.standard_normal(int(np.prod(count))).astype(np.float32)
    # at cell centres every mode returns the cell's own value; iso is subtracted
    ijk = np.array([[i, j, k] for i in range(6) for j in range(5) for k in range(7)])
    centres = (first + ijk * size).astype(np.float32)
    for mode in (post.SNAP, post.TRILINEAR, post.TETRAHEDRAL):
        got = post.sample_grid(sdf, first, size, count, centres, mode, iso=0.25)
        assert np.allclose(got, sdf - 0.25, atol=1e-6), mode
    # outside [first, first + count * size] -> 100 (draw_raymarching.wgsl:121-123; Grid::get_last_cell quirk)
    out = post.sample_grid(sdf, first, size, count, [[0.4, 0.0, 2.5], [0.5 + 6 * 0.25 + 0.01, 0.0, 2.5]], post.TRILINEAR)
    assert list(out) == [100.0, 100.0]
    # affine fields are reproduced by both interpolations inside the dual grid
    lin = _affine_grid(count, first, size, (0.5, -2.0, 3.0, 1.0))
    p = (first + rng.uniform(0, 1, (200, 3)) * (np.array(count) - 1) * size).astype(np.float32)
    want = 0.5 * p[:, 0] - 2.0 * p[:, 1] + 3.0 * p[:, 2] + 1.0
    for mode in (post.TRILINEAR, post.TETRAHEDRAL):
        assert np.allclose(post.sample_grid(lin, first, size, count, p, mode), want, atol=2e-5), mode
    # snap: a point anywhere inside a cell returns that cell
    q = (first + np.array([2, 3, 4]) * size + np.array([0.4, -0.4, 0.3]) * size).astype(np.float32)
    assert post.sample_grid(sdf, first, size, count, [q], post.SNAP)[0] == sdf[4 + 3 * 7 + 2 * 35]


# ---- CUDA kernels against the oracle (GPU) -----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 1000, 4096, 4097, 12_289, 262_144 + 17, 2_000_003])
def test_gpu_grid_order_matches_oracle(m2s, n):
    # sizes around the 4096-pair tiles of the radix sort (csrc/m2s_sort.cuh): one tile, a full tile, a tile with a
    # single pair, many tiles with look-back
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype(np.float32)
    a[rng.integers(0, n, n // 3 + 1)] = np.float32(0.5)  # many ties -> stability matters
    if n > 8:
        a[3], a[5] = 0.0, -0.0
    if n > 64:  # every class f32::total_cmp orders: NaNs of both signs, infinities, denormals
        a[7:13] = np.array([0x7fc00000, 0xffc00000, 0x7f800000, 0xff800000, 0x00000001, 0x80000001], np.uint32).view(np.float32)
        a[40:60] *= np.float32(1e-30)  # spread over many exponents
    with m2s.Context() as c:
        order, (lo, hi) = c.grid_order(a)
    assert np.array_equal(order, post.grid_order(a))
    if not np.isnan(a).any():  # itertools::minmax over PartialOrd has no defined answer with NaNs
        wlo, whi = post.minmax(a)
        assert lo.tobytes() == wlo.tobytes() and hi.tobytes() == whi.tobytes()
    b = np.where(np.isnan(a), np.float32(0.25), a)  # the same array without NaNs: iso limits as well
    with m2s.Context() as c:
        order, (lo, hi) = c.grid_order(b)
    assert np.array_equal(order, post.grid_order(b))
    wlo, whi = post.minmax(b)
    assert lo.tobytes() == wlo.tobytes() and hi.tobytes() == whi.tobytes()


@pytest.mark.gpu
def test_gpu_grid_order_of_generated_sdf_on_device(m2s):
    torch = pytest.importorskip("torch")
    verts, tris = synth.bumpy_torus(32, 20)
    mn, mx = synth.padded_grid_box(verts)
    grid = m2s.Grid.from_bounding_box(mn, mx, [33, 30, 21])
    n = 33 * 30 * 21
    dv = torch.from_numpy(verts).cuda()
    dt = torch.from_numpy(tris.view(np.int32)).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    mm = torch.empty(2, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with m2s.Context() as c:
        c.grid_sdf_device(dv.data_ptr(), len(verts), dt.data_ptr(), len(tris), grid, 0, 0, 33, out.data_ptr())
        c.grid_order_device(out.data_ptr(), n, order.data_ptr(), mm.data_ptr())  # same stream: no host round trip
        c.synchronize()
    sdf = out.cpu().numpy()
    assert np.array_equal(order.cpu().numpy().view(np.uint32), post.grid_order(sdf))
    assert tuple(mm.cpu().numpy()) == (sdf.min(), sdf.max())
    assert np.all(np.diff(sdf[order.cpu().numpy()]) >= 0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gpu_sample_matches_oracle_bit_for_bit(m2s, mode):
    verts, tris = synth.bumpy_torus(24, 16)
    mn, mx = synth.padded_grid_box(verts)
    grid = m2s.Grid.from_bounding_box(mn, mx, [20, 17, 23])
    rng = np.random.default_rng(mode)
    with m2s.Context() as c:
        sdf = c.grid_sdf(verts, tris, grid, 0)
        # points inside, on the border cells (clamped fetches) and outside the box
        lo = grid.first_cell - 2 * grid.cell_size
        hi = grid.first_cell + (np.array(grid.cell_count) + 2) * grid.cell_size
        p = (lo + rng.uniform(0, 1, (20000, 3)) * (hi - lo)).astype(np.float32)
        ijk = rng.integers(0, [20, 17, 23], (500, 3))
        p = np.concatenate([p, (grid.first_cell + ijk * grid.cell_size).astype(np.float32)])  # exact cell centres
        got = c.sample_grid_sdf(sdf, grid, p, mode, iso=0.03)
    want = post.sample_grid(sdf, grid.first_cell, grid.cell_size, grid.cell_count, p, mode, iso=0.03)
    assert (want == 100.0).any() and (want != 100.0).any()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_sample_interpolates_the_distance_field(m2s, oracle):
    # end to end: trilinear samples of the generated grid stay within one cell diagonal's worth of curvature
    # error of the exact distance at the same points (1-Lipschitz field, h = cell diagonal)
    verts, tris = synth.bumpy_torus(48, 30)
    mn, mx = synth.padded_grid_box(verts)
    grid = m2s.Grid.from_bounding_box(mn, mx, [64, 64, 48])
    rng = np.random.default_rng(3)
    inner_lo = grid.first_cell
    inner_hi = grid.first_cell + (np.array(grid.cell_count) - 1) * grid.cell_size
    p = (inner_lo + rng.uniform(0, 1, (4000, 3)) * (inner_hi - inner_lo)).astype(np.float32)
    with m2s.Context() as c:
        sdf = c.grid_sdf(verts, tris, grid, 0)
        tri = c.sample_grid_sdf(sdf, grid, p, 1)
        tet = c.sample_grid_sdf(sdf, grid, p, 2)
    exact = oracle.generate_sdf(verts, tris, p, oracle.ACCEL_BVH, oracle.RAYCAST)
    h = float(np.linalg.norm(grid.cell_size))
    assert np.max(np.abs(tri - exact)) <= h and np.max(np.abs(tet - exact)) <= h
    assert np.median(np.abs(tri - exact)) <= 0.05 * h


# ---- hand-derived known answers for sdf_grid() (draw_raymarching.wgsl:118-200, :92-99, :585-650) -----------------
# A 2x2x2 grid with first_cell = (0,0,0), cell_size = (1,1,1): cell_index == position, so the fractions are the
# coordinates themselves. Corner values are powers of two and the fractions multiples of 1/4: every product and sum
# below is exact in f32, so the expected numbers are worked out by hand from the shader's formulas, not by running
# either implementation. d[x][y][z]:
_KAT_D = {(0, 0, 0): 1.0, (1, 0, 0): 2.0, (0, 1, 0): 4.0, (0, 0, 1): 8.0,
          (1, 1, 0): 16.0, (1, 0, 1): 32.0, (0, 1, 1): 64.0, (1, 1, 1): 128.0}
# (position, expected tetrahedral value): one point strictly inside each of the six tetrahedra of :596-648
_KAT_TETRA = [
    # fG>=fB>=fR: (1-g) d000 + (g-b) d010 + (b-r) d011 + r d111
    ((0.25, 0.75, 0.5), 0.25 * 1 + 0.25 * 4 + 0.25 * 64 + 0.25 * 128),      # 49.25
    # fB>fR>fG: (1-b) d000 + (b-r) d001 + (r-g) d101 + g d111
    ((0.5, 0.25, 0.75), 0.25 * 1 + 0.25 * 8 + 0.25 * 32 + 0.25 * 128),      # 42.25
    # fB>fG>=fR: (1-b) d000 + (b-g) d001 + (g-r) d011 + r d111
    ((0.25, 0.5, 0.75), 0.25 * 1 + 0.25 * 8 + 0.25 * 64 + 0.25 * 128),      # 50.25
    # fR>=fG>fB: (1-r) d000 + (r-g) d100 + (g-b) d110 + b d111
    ((0.75, 0.5, 0.25), 0.25 * 1 + 0.25 * 2 + 0.25 * 16 + 0.25 * 128),      # 36.75
    # fG>fR>=fB: (1-g) d000 + (g-r) d010 + (r-b) d110 + b d111
    ((0.5, 0.75, 0.25), 0.25 * 1 + 0.25 * 4 + 0.25 * 16 + 0.25 * 128),      # 37.25
    # fR>=fB>=fG: (1-r) d000 + (r-b) d100 + (b-g) d101 + g d111
    ((0.75, 0.25, 0.5), 0.25 * 1 + 0.25 * 2 + 0.25 * 32 + 0.25 * 128),      # 40.75
    # the diagonal r = g = b = 1/2 takes the LAST matching branch (fR>=fB>=fG): 0.5 d000 + 0.5 d111
    ((0.5, 0.5, 0.5), 0.5 * 1 + 0.5 * 128),                                  # 64.5
]


def _kat_grid():
    sdf = np.zeros(8, np.float32)
    for (x, y, z), v in _KAT_D.items():
        sdf[z + 2 * y + 4 * x] = v  # Grid::get_cell_idx
    return sdf, np.zeros(3, np.float32), np.ones(3, np.float32), [2, 2, 2]


def _kat_cases():
    pts, want = [], {0: [], 1: [], 2: []}
    for p, tet in _KAT_TETRA:
        x, y, z = p
        pts.append(p)
        # trilinear (:155-166): x first, then y, then z
        c00 = _KAT_D[0, 0, 0] * (1 - x) + _KAT_D[1, 0, 0] * x
        c01 = _KAT_D[0, 0, 1] * (1 - x) + _KAT_D[1, 0, 1] * x
        c10 = _KAT_D[0, 1, 0] * (1 - x) + _KAT_D[1, 1, 0] * x
        c11 = _KAT_D[0, 1, 1] * (1 - x) + _KAT_D[1, 1, 1] * x
        want[1].append((c00 * (1 - y) + c10 * y) * (1 - z) + (c01 * (1 - y) + c11 * y) * z)
        want[2].append(tet)
        # snap (:126-134): floor((p - (first - size/2)) / size) = floor(p + 0.5)
        want[0].append(_KAT_D[int(x + 0.5), int(y + 0.5), int(z + 0.5)])
    # hand-checked spot values of the table above
    assert want[2][0] == 49.25 and want[2][3] == 36.75 and want[2][6] == 64.5
    assert want[1][6] == (1 + 2 + 4 + 8 + 16 + 32 + 64 + 128) / 8.0  # the centre is the mean of the corners
    # outside [first_cell, get_last_cell()] = [0, 2]^3: 100.0 whatever the mode (:120-122); the upper border itself is
    # inside and clamps to the last cell (:92-99): position (2, 2, 2) -> idx (2,2,2), fractions 0 -> d111
    pts += [(-0.25, 0.5, 0.5), (0.5, 2.25, 0.5), (2.0, 2.0, 2.0), (1.5, 0.0, 0.0)]
    for m in (0, 1, 2):
        want[m] += [100.0, 100.0, 128.0]
    # (1.5, 0, 0): idx (1,0,0) r = (0.5, 0, 0); x+1 clamps to the last cell: every x-neighbour is d1yz
    want[0].append(_KAT_D[1, 0, 0])
    want[1].append(_KAT_D[1, 0, 0])
    want[2].append(0.5 * _KAT_D[1, 0, 0] + 0.5 * _KAT_D[1, 0, 0])  # fR>=fB>=fG: (1-r) d100 + (r-b) d(2->1)00
    return np.asarray(pts, np.float32), {m: np.asarray(v, np.float32) for m, v in want.items()}


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sample_known_answers_oracle(mode):
    sdf, first, size, count = _kat_grid()
    pts, want = _kat_cases()
    got = post.sample_grid(sdf, first, size, count, pts, mode)
    assert np.array_equal(got, want[mode]), (mode, got, want[mode])
    # iso is subtracted from every fetched value, not from the interpolated result (:98): affine weights sum to 1
    got_iso = post.sample_grid(sdf, first, size, count, pts[:7], mode, iso=0.5)
    assert np.array_equal(got_iso, want[mode][:7] - np.float32(0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sample_known_answers_gpu(m2s, mode):
    sdf, first, size, count = _kat_grid()
    pts, want = _kat_cases()
    grid = m2s.Grid(first, size, count)
    got = m2s.default_context().sample_grid_sdf(sdf, grid, pts, mode)
    assert np.array_equal(got, want[mode]), (mode, got, want[mode])
