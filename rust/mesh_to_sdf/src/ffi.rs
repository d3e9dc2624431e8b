//! Raw declarations of include/m2s.h (ABI version 2). Plain pointers and sizes only.
#![allow(non_camel_case_types, dead_code)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct m2s_ctx {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct m2s_mesh {
    _opaque: [u8; 0],
}

pub const M2S_OK: c_int = 0;
pub const M2S_EINVAL: c_int = 1;
pub const M2S_EINDEX: c_int = 2;
pub const M2S_ENAN: c_int = 3;
pub const M2S_ECUDA: c_int = 4;
pub const M2S_ENODEV: c_int = 6;
pub const M2S_EEMPTY: c_int = 7;

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct m2s_timings {
    pub h2d_ms: f32,
    pub build_ms: f32,
    pub sign_ms: f32,
    pub dist_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
    pub seed_ms: f32,
    pub host_path: c_int,
}

extern "C" {
    pub fn m2s_abi_version() -> c_int;
    pub fn m2s_create(devices: *const c_int, n_devices: c_int, out: *mut *mut m2s_ctx) -> c_int;
    pub fn m2s_create_on_stream(device: c_int, cuda_stream: *mut c_void, out: *mut *mut m2s_ctx) -> c_int;
    pub fn m2s_destroy(ctx: *mut m2s_ctx);
    pub fn m2s_last_error(ctx: *const m2s_ctx) -> *const c_char;
    pub fn m2s_last_error_copy(ctx: *mut m2s_ctx, buf: *mut c_char, n: usize) -> c_int;
    pub fn m2s_last_timings(ctx: *const m2s_ctx, out: *mut m2s_timings) -> c_int;
    pub fn m2s_set_option(ctx: *mut m2s_ctx, option: c_int, value: i64) -> c_int;
    pub fn m2s_mesh_create(ctx: *mut m2s_ctx, verts_xyz: *const f32, nv: u64, tri_idx: *const u32, nt: u64, out: *mut *mut m2s_mesh) -> c_int;
    pub fn m2s_mesh_destroy(mesh: *mut m2s_mesh);
    pub fn m2s_mesh_grid_sdf(
        ctx: *mut m2s_ctx, mesh: *mut m2s_mesh, first_cell: *const f32, cell_size: *const f32, cell_count: *const u64,
        sign_method: c_int, x_begin: u64, x_end: u64, out_slab: *mut f32,
    ) -> c_int;
    pub fn m2s_mesh_sdf(
        ctx: *mut m2s_ctx, mesh: *mut m2s_mesh, queries_xyz: *const f32, nq: u64, accel_method: c_int, sign_method: c_int,
        out: *mut f32,
    ) -> c_int;
    pub fn m2s_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn m2s_host_free(p: *mut c_void);
    pub fn m2s_host_register(p: *mut c_void, bytes: usize) -> c_int;
    pub fn m2s_host_unregister(p: *mut c_void) -> c_int;
    pub fn m2s_launch_count(ctx: *const m2s_ctx) -> u64;
    pub fn m2s_device_count(ctx: *const m2s_ctx) -> c_int;
    pub fn m2s_synchronize(ctx: *mut m2s_ctx) -> c_int;
    pub fn m2s_generate_grid_sdf(
        ctx: *mut m2s_ctx, verts_xyz: *const f32, nv: u64, tri_idx: *const u32, nt: u64,
        first_cell: *const f32, cell_size: *const f32, cell_count: *const u64, sign_method: c_int, out: *mut f32,
    ) -> c_int;
    pub fn m2s_generate_grid_sdf_slab(
        ctx: *mut m2s_ctx, verts_xyz: *const f32, nv: u64, tri_idx: *const u32, nt: u64,
        first_cell: *const f32, cell_size: *const f32, cell_count: *const u64, sign_method: c_int,
        x_begin: u64, x_end: u64, out_slab: *mut f32,
    ) -> c_int;
    pub fn m2s_generate_sdf(
        ctx: *mut m2s_ctx, verts_xyz: *const f32, nv: u64, tri_idx: *const u32, nt: u64,
        queries_xyz: *const f32, nq: u64, accel_method: c_int, sign_method: c_int, out: *mut f32,
    ) -> c_int;
    // post-passes on a finished grid: mesh_to_sdf_client/src/sdf.rs:62-68, :123; shaders/draw_raymarching.wgsl:118-200
    pub fn m2s_grid_order(ctx: *mut m2s_ctx, sdf: *const f32, n: u64, order: *mut u32, minmax: *mut f32) -> c_int;
    pub fn m2s_sample_grid_sdf(
        ctx: *mut m2s_ctx, sdf: *const f32, first_cell: *const f32, cell_size: *const f32, cell_count: *const u64,
        points_xyz: *const f32, np: u64, sample_mode: c_int, iso: f32, out: *mut f32,
    ) -> c_int;
    pub fn m2s_expand_topology(
        topology: c_int, indices: *const c_void, index_bytes: c_int, n_indices: u64, nv: u64, out: *mut u32,
    ) -> u64;
    pub fn m2s_grid_from_bounding_box(
        bbox_min: *const f32, bbox_max: *const f32, cell_count: *const u64, first_cell: *mut f32, cell_size: *mut f32,
    );
}
