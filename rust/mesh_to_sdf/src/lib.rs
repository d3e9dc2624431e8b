//! mesh_to_sdf — the public surface of Azkellas/mesh_to_sdf 0.4.0 (`generate_sdf`, `generate_grid_sdf`, `Grid`,
//! `SnapResult`, `Topology`, `SignMethod`, `AccelerationMethod`, `Point`) as a thin facade over `libm2s.so`, the
//! B200-native CUDA implementation of the hot path. There is no CPU fallback: without a CUDA device the first call
//! panics. Signatures stay infallible like the reference's; every non-OK status becomes a panic.
mod ffi;
mod grid;
mod point;
#[cfg(feature = "serde")]
pub mod serde;

pub use grid::{Grid, SnapResult};
pub use point::Point;

use std::sync::{Mutex, OnceLock};

/// How indices are stored. `None` means `0..vertices.len()`.
#[derive(Copy, Clone)]
pub enum Topology<'a, I: Into<u32>> {
    TriangleList(Option<&'a [I]>),
    TriangleStrip(Option<&'a [I]>),
}

#[derive(Debug, Clone, Copy, Default, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub enum SignMethod {
    #[default]
    Raycast,
    Normal,
}

#[derive(Default, Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub enum AccelerationMethod {
    None(SignMethod),
    Bvh(SignMethod),
    Rtree,
    #[default]
    RtreeBvh,
}

struct Ctx(*mut ffi::m2s_ctx);
unsafe impl Send for Ctx {}

fn ctx() -> &'static Mutex<Ctx> {
    static CTX: OnceLock<Mutex<Ctx>> = OnceLock::new();
    CTX.get_or_init(|| {
        // M2S_DEVICES="0,1,..." shards grids by x-slabs and queries by ranges over several GPUs of one box.
        let devices: Vec<i32> = std::env::var("M2S_DEVICES")
            .map(|s| s.split(',').filter_map(|t| t.trim().parse().ok()).collect())
            .unwrap_or_default();
        let mut raw = core::ptr::null_mut();
        let rc = unsafe { ffi::m2s_create(if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() }, devices.len() as i32, &mut raw) };
        assert!(rc == ffi::M2S_OK, "mesh_to_sdf: no usable CUDA device (libm2s status {rc}); there is no CPU fallback");
        Mutex::new(Ctx(raw))
    })
}

fn check(c: &Ctx, rc: i32) {
    if rc != ffi::M2S_OK {
        // copied while the context is still locked by this call (m2s_last_error_copy)
        let mut buf = [0 as core::ffi::c_char; 512];
        unsafe { ffi::m2s_last_error_copy(c.0, buf.as_mut_ptr(), buf.len()) };
        let msg = unsafe { std::ffi::CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned();
        match rc {
            ffi::M2S_ENAN => panic!("NaN distance ({msg})"),
            ffi::M2S_EINDEX => panic!("index out of bounds ({msg})"),
            ffi::M2S_EEMPTY => panic!("called `Option::unwrap()` on a `None` value (empty mesh: {msg})"),
            _ => panic!("mesh_to_sdf backend error {rc}: {msg}"),
        }
    }
}

fn pack<V: Point>(v: &[V]) -> Vec<f32> {
    let mut out = Vec::with_capacity(v.len() * 3);
    for p in v {
        out.extend_from_slice(&[p.x(), p.y(), p.z()]);
    }
    out
}

fn triangles<I: Copy + Into<u32>>(nv: usize, topology: Topology<'_, I>) -> Vec<u32> {
    let (kind, idx): (i32, Option<Vec<u32>>) = match topology {
        Topology::TriangleList(i) => (0, i.map(|s| s.iter().map(|x| (*x).into()).collect())),
        Topology::TriangleStrip(i) => (1, i.map(|s| s.iter().map(|x| (*x).into()).collect())),
    };
    let (ptr, n) = idx.as_ref().map_or((core::ptr::null(), 0), |v| (v.as_ptr().cast(), v.len() as u64));
    let count = unsafe { ffi::m2s_expand_topology(kind, ptr, 4, n, nv as u64, core::ptr::null_mut()) } as usize;
    let mut out = vec![0u32; count * 3];
    unsafe { ffi::m2s_expand_topology(kind, ptr, 4, n, nv as u64, out.as_mut_ptr()) };
    out
}

/// Signed distance from every query point to the mesh, in query order.
pub fn generate_sdf<V, I>(vertices: &[V], indices: Topology<I>, query_points: &[V], acceleration_method: AccelerationMethod) -> Vec<f32>
where
    V: Point + 'static,
    I: Copy + Into<u32> + Sync + Send,
{
    let tris = triangles(vertices.len(), indices);
    if tris.is_empty() && acceleration_method == AccelerationMethod::RtreeBvh {
        return vec![]; // what the reference returns for an empty mesh on this path
    }
    let (accel, sign) = match acceleration_method {
        AccelerationMethod::None(s) => (0, s as i32),
        AccelerationMethod::Bvh(s) => (1, s as i32),
        AccelerationMethod::Rtree => (2, 0),
        AccelerationMethod::RtreeBvh => (3, 0),
    };
    let (v, q) = (pack(vertices), pack(query_points));
    let mut out = vec![0f32; query_points.len()];
    let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
    let rc = unsafe {
        ffi::m2s_generate_sdf(c.0, v.as_ptr(), vertices.len() as u64, tris.as_ptr(), (tris.len() / 3) as u64, q.as_ptr(),
                              query_points.len() as u64, accel, sign, out.as_mut_ptr())
    };
    check(&c, rc);
    out
}

/// Signed distance at every cell centre of `grid`, flat in `Grid::get_cell_idx` order.
pub fn generate_grid_sdf<V, I>(vertices: &[V], indices: Topology<I>, grid: &Grid<V>, sign_method: SignMethod) -> Vec<f32>
where
    V: Point + 'static,
    I: Copy + Into<u32> + Sync + Send,
{
    let tris = triangles(vertices.len(), indices);
    let v = pack(vertices);
    let (f, s, n) = (grid.get_first_cell(), grid.get_cell_size(), grid.get_cell_count());
    let (first, size) = ([f.x(), f.y(), f.z()], [s.x(), s.y(), s.z()]);
    let count = [n[0] as u64, n[1] as u64, n[2] as u64];
    let mut out = vec![0f32; grid.get_total_cell_count()];
    let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
    let rc = unsafe {
        ffi::m2s_generate_grid_sdf(c.0, v.as_ptr(), vertices.len() as u64, tris.as_ptr(), (tris.len() / 3) as u64,
                                   first.as_ptr(), size.as_ptr(), count.as_ptr(), sign_method as i32, out.as_mut_ptr())
    };
    check(&c, rc);
    out
}

/// A `Vec<f32>`-like buffer in page-locked, mapped host memory (`m2s_host_alloc`). `generate_grid_sdf_into` has the
/// distance kernel write it in place over PCIe: no staging buffer, no copy of the result. Derefs to `[f32]`.
pub struct PinnedVec {
    ptr: *mut f32,
    len: usize,
}
unsafe impl Send for PinnedVec {}
impl PinnedVec {
    pub fn new(len: usize) -> Self {
        let mut raw = core::ptr::null_mut();
        let rc = unsafe { ffi::m2s_host_alloc(len.max(1) * 4, &mut raw) };
        assert!(rc == ffi::M2S_OK, "mesh_to_sdf: m2s_host_alloc failed ({rc})");
        Self { ptr: raw.cast(), len }
    }
}
impl Drop for PinnedVec {
    fn drop(&mut self) {
        unsafe { ffi::m2s_host_free(self.ptr.cast()) }
    }
}
impl core::ops::Deref for PinnedVec {
    type Target = [f32];
    fn deref(&self) -> &[f32] {
        unsafe { core::slice::from_raw_parts(self.ptr, self.len) }
    }
}
impl core::ops::DerefMut for PinnedVec {
    fn deref_mut(&mut self) -> &mut [f32] {
        unsafe { core::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

/// `generate_grid_sdf` into a caller-owned destination of `grid.get_total_cell_count()` floats. A `PinnedVec` is
/// written in place by the kernel; any other slice is filled through the library's pinned ring while the kernel runs.
pub fn generate_grid_sdf_into<V, I>(vertices: &[V], indices: Topology<I>, grid: &Grid<V>, sign_method: SignMethod, out: &mut [f32])
where
    V: Point + 'static,
    I: Copy + Into<u32> + Sync + Send,
{
    assert_eq!(out.len(), grid.get_total_cell_count(), "destination length does not match the grid");
    let tris = triangles(vertices.len(), indices);
    let v = pack(vertices);
    let (f, s, n) = (grid.get_first_cell(), grid.get_cell_size(), grid.get_cell_count());
    let (first, size) = ([f.x(), f.y(), f.z()], [s.x(), s.y(), s.z()]);
    let count = [n[0] as u64, n[1] as u64, n[2] as u64];
    let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
    let rc = unsafe {
        ffi::m2s_generate_grid_sdf(c.0, v.as_ptr(), vertices.len() as u64, tris.as_ptr(), (tris.len() / 3) as u64,
                                   first.as_ptr(), size.as_ptr(), count.as_ptr(), sign_method as i32, out.as_mut_ptr())
    };
    check(&c, rc);
}

/// A mesh uploaded once with its LBVH kept on the GPU(s) (`m2s_mesh_create`): repeated `generate_*` calls on it pay
/// neither the upload nor the build — the pattern of the reference's viewer, which regenerates the grid of one mesh
/// on every parameter change (mesh_to_sdf_client/src/sdf_program.rs:679-721).
pub struct Mesh {
    raw: *mut ffi::m2s_mesh,
    n_triangles: usize,
}
unsafe impl Send for Mesh {}
impl Mesh {
    pub fn new<V: Point, I: Copy + Into<u32>>(vertices: &[V], indices: Topology<I>) -> Self {
        let tris = triangles(vertices.len(), indices);
        let v = pack(vertices);
        let mut raw = core::ptr::null_mut();
        let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
        let rc = unsafe { ffi::m2s_mesh_create(c.0, v.as_ptr(), vertices.len() as u64, tris.as_ptr(), (tris.len() / 3) as u64, &mut raw) };
        check(&c, rc);
        Self { raw, n_triangles: tris.len() / 3 }
    }

    pub fn generate_grid_sdf<V: Point>(&self, grid: &Grid<V>, sign_method: SignMethod) -> Vec<f32> {
        let (f, s, n) = (grid.get_first_cell(), grid.get_cell_size(), grid.get_cell_count());
        let (first, size) = ([f.x(), f.y(), f.z()], [s.x(), s.y(), s.z()]);
        let count = [n[0] as u64, n[1] as u64, n[2] as u64];
        let mut out = vec![0f32; grid.get_total_cell_count()];
        let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
        let rc = unsafe {
            ffi::m2s_mesh_grid_sdf(c.0, self.raw, first.as_ptr(), size.as_ptr(), count.as_ptr(), sign_method as i32, 0, count[0],
                                   out.as_mut_ptr())
        };
        check(&c, rc);
        out
    }

    pub fn generate_sdf<V: Point>(&self, query_points: &[V], acceleration_method: AccelerationMethod) -> Vec<f32> {
        if self.n_triangles == 0 && acceleration_method == AccelerationMethod::RtreeBvh {
            return vec![];
        }
        let (accel, sign) = match acceleration_method {
            AccelerationMethod::None(s) => (0, s as i32),
            AccelerationMethod::Bvh(s) => (1, s as i32),
            AccelerationMethod::Rtree => (2, 0),
            AccelerationMethod::RtreeBvh => (3, 0),
        };
        let q = pack(query_points);
        let mut out = vec![0f32; query_points.len()];
        let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
        let rc = unsafe { ffi::m2s_mesh_sdf(c.0, self.raw, q.as_ptr(), query_points.len() as u64, accel, sign, out.as_mut_ptr()) };
        check(&c, rc);
        out
    }
}
impl Drop for Mesh {
    fn drop(&mut self) {
        unsafe { ffi::m2s_mesh_destroy(self.raw) }
    }
}

/// What the reference's in-repo caller does with the grid next (mesh_to_sdf_client/src/sdf.rs:62-68, :123):
/// `(0..n).sorted_by(|i, j| data[*i].total_cmp(&data[*j]))` and `data.iter().copied().minmax()`, on the GPU.
pub fn grid_order(sdf: &[f32]) -> (Vec<u32>, (f32, f32)) {
    let mut order = vec![0u32; sdf.len()];
    let mut mm = [0f32; 2];
    let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
    let rc = unsafe { ffi::m2s_grid_order(c.0, sdf.as_ptr(), sdf.len() as u64, order.as_mut_ptr(), mm.as_mut_ptr()) };
    check(&c, rc);
    (order, (mm[0], mm[1]))
}

/// `raymarch_mode` of mesh_to_sdf_client/shaders/draw_raymarching.wgsl.
#[derive(Debug, Clone, Copy, PartialEq, Eq, Default)]
pub enum SampleMode {
    Snap = 0,
    #[default]
    Trilinear = 1,
    Tetrahedral = 2,
}

/// Distance at arbitrary points by interpolating a grid SDF — the TODO of the reference's `src/grid.rs:172`, with
/// the semantics of `sdf_grid()` in the reference's raymarching shader (100.0 outside the grid).
pub fn sample_grid_sdf<V: Point>(sdf: &[f32], grid: &Grid<V>, points: &[V], mode: SampleMode, iso: f32) -> Vec<f32> {
    assert_eq!(sdf.len(), grid.get_total_cell_count(), "sdf length does not match the grid");
    let p = pack(points);
    let (f, s, n) = (grid.get_first_cell(), grid.get_cell_size(), grid.get_cell_count());
    let (first, size) = ([f.x(), f.y(), f.z()], [s.x(), s.y(), s.z()]);
    let count = [n[0] as u64, n[1] as u64, n[2] as u64];
    let mut out = vec![0f32; points.len()];
    let c = ctx().lock().unwrap_or_else(|e| e.into_inner());
    let rc = unsafe {
        ffi::m2s_sample_grid_sdf(c.0, sdf.as_ptr(), first.as_ptr(), size.as_ptr(), count.as_ptr(), p.as_ptr(),
                                 points.len() as u64, mode as i32, iso, out.as_mut_ptr())
    };
    check(&c, rc);
    out
}

#[cfg(test)]
mod tests {
    use super::*;

    // the reference's doc-tests, unchanged in meaning
    #[test]
    fn doc_generate_sdf() {
        let vertices: Vec<[f32; 3]> = vec![[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]];
        let indices: Vec<u32> = vec![0, 1, 2];
        let sdf = generate_sdf(&vertices, Topology::TriangleList(Some(&indices)), &[[0.5, 0.5, 0.5]], AccelerationMethod::RtreeBvh);
        assert_eq!(sdf, vec![1.0]);
    }

    #[test]
    fn doc_generate_grid_sdf() {
        let vertices: Vec<[f32; 3]> = vec![[0.5, 1.5, 0.5], [1., 2., 3.], [1., 3., 7.]];
        let indices: Vec<u32> = vec![0, 1, 2];
        let grid = Grid::from_bounding_box(&[0., 0., 0.], &[10., 10., 10.], [10, 10, 10]);
        let sdf = generate_grid_sdf(&vertices, Topology::TriangleList(Some(&indices)), &grid, SignMethod::Raycast);
        assert_eq!(sdf.len(), 1000);
        assert_eq!(sdf[0], 1.0);
    }
}
