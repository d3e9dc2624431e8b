/* m2s.h — C ABI of libm2s.so, the B200-native (sm_100a) replacement for the hot path of the Rust
 * crate Azkellas/mesh_to_sdf v0.4.0: nearest-triangle distance + Raycast/Normal sign behind
 * `generate_grid_sdf` and `generate_sdf`.
 *
 * The reference has no FFI of its own: its boundary is the crate's public Rust API
 * (mesh_to_sdf/src/lib.rs:146-148 re-exports). Each entry point below cites the reference item it
 * replaces; the Rust facade that binds them is in rust/mesh_to_sdf/ (source) and INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; nothing is retained after return;
 *   - vertices / queries are packed xyz float32 (what the facade reads through Point::x()/y()/z(),
 *     src/point.rs:46-56); triangles are already expanded u32 index triples
 *     (Topology::get_triangles, src/lib.rs:175-193 — the facade / m2s_expand_topology does that);
 *   - grid output index = z + y*nz + x*ny*nz (Grid::get_cell_idx, src/grid.rs:122-124);
 *   - every error is a status code (the reference panics; the facade maps non-OK to panic!);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call returns
 *     M2S_ENODEV / M2S_ECUDA.
 */
#ifndef M2S_H
#define M2S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2S_ABI_VERSION 2

#if defined(__GNUC__)
#define M2S_API __attribute__((visibility("default")))
#else
#define M2S_API
#endif

typedef enum m2s_status {
    M2S_OK = 0,
    M2S_EINVAL = 1, /* bad argument (null pointer, unknown enum, zero cell count …)              */
    M2S_EINDEX = 2, /* a triangle index >= nv   (reference: slice index panic)                    */
    M2S_ENAN = 3,   /* non-finite input / NaN distance (reference: panic "NaN distance" lib.rs:257)*/
    M2S_ECUDA = 4,  /* CUDA runtime error, see m2s_last_error                                     */
    M2S_ENCCL = 5,  /* reserved (the multi-GPU reassembly uses peer-mapped stores, not a collective)      */
    M2S_ENODEV = 6, /* no usable CUDA device                                                      */
    M2S_EEMPTY = 7  /* Rtree / RtreeBvh on a mesh without triangles (rtree.rs:117 panics;
                       rtree_bvh.rs:104-106 returns an empty Vec — the facade handles both)       */
} m2s_status;

/* SignMethod, src/lib.rs:204-216 (declaration order; Raycast is #[default]). */
typedef enum m2s_sign_method { M2S_SIGN_RAYCAST = 0, M2S_SIGN_NORMAL = 1 } m2s_sign_method;

/* AccelerationMethod, src/lib.rs:224-239. `sign` is only read for NONE and BVH. */
typedef enum m2s_accel_method {
    M2S_ACCEL_NONE = 0,
    M2S_ACCEL_BVH = 1,
    M2S_ACCEL_RTREE = 2,
    M2S_ACCEL_RTREE_BVH = 3 /* #[default] */
} m2s_accel_method;

/* Topology, src/lib.rs:151-167. */
typedef enum m2s_topology { M2S_TRIANGLE_LIST = 0, M2S_TRIANGLE_STRIP = 1 } m2s_topology;

typedef struct m2s_ctx m2s_ctx;   /* opaque: device(s), streams, scratch arenas, host copy threads */
typedef struct m2s_mesh m2s_mesh; /* opaque: a mesh uploaded once + its LBVH on every device of a context */

/* Phase timings of the last call, milliseconds, measured with CUDA events on the context's stream
 * (replaces the reference's log::info! phase timers, src/generate/grid.rs:303-307,342-346,369-373). */
typedef struct m2s_timings {
    float h2d_ms;   /* host -> device copies (0 for the *_device entry points)             */
    float build_ms; /* triangle records, Morton sort, LBVH hierarchy + refit (0 on a mesh handle) */
    float sign_ms;  /* raycast row toggles + scan (grid) / query sort (points) beyond the build */
    float dist_ms;  /* the final nearest-triangle kernel alone (sign applied in its epilogue) */
    float d2h_ms;   /* from the end of the kernel to the result being complete in the caller's buffer */
    float total_ms; /* first event to last event                                           */
    float seed_ms;  /* work between the sign phase and the kernel (node interleave)        */
    int host_path;  /* how the result reached a host destination: one of m2s_host_path_taken */
} m2s_timings;

/* m2s_timings.host_path */
typedef enum m2s_host_path_taken {
    M2S_PATH_DEVICE = 0,    /* device destination, nothing copied                                              */
    M2S_PATH_ZEROCOPY = 1,  /* page-locked destination written by the kernel's own stores                      */
    M2S_PATH_PIPELINED = 2, /* pageable destination filled from the library's pinned ring by host threads while
                               the kernel runs                                                                 */
    M2S_PATH_STAGED = 3,    /* device buffer + one cudaMemcpy into the destination                             */
    M2S_PATH_REGISTERED = 4 /* pageable destination page-locked for the duration of the call, then as ZEROCOPY */
} m2s_host_path_taken;

/* m2s_set_option keys */
typedef enum m2s_option {
    M2S_OPT_BUILD_MODE = 1,  /* multi-device contexts: m2s_build_mode                                           */
    M2S_OPT_HOST_PATH = 2,   /* how pageable host destinations are filled: m2s_host_path                        */
    M2S_OPT_COPY_THREADS = 3, /* host threads of the pipelined paths (1..64; default 4..8 by host size)         */
    M2S_OPT_RAY_BINS = 4,     /* generate_sdf Raycast sign rules: 1 (default) = axis-ray parities through per-axis
                                 2-D triangle bins, 0 = always through the packet walk of the box tree (same results) */
    M2S_OPT_BALANCE = 5,      /* multi-device contexts: 1 (default) = a grid call on the same grid shape as the previous
                                 one cuts its x-slabs at equal shares of that call's measured per-slab kernel time,
                                 0 = always equal-width slabs (same results either way)                              */
    M2S_OPT_RUN_LENGTH = 6    /* voxels per lane of the grid kernel: 0 (default) = chosen by mesh size, 2 or 4         */
} m2s_option;
typedef enum m2s_build_mode {
    M2S_BUILD_REPLICATED = 0, /* every device builds its own LBVH from the mesh (default: measured faster)      */
    M2S_BUILD_BROADCAST = 1   /* the first device builds, the others pull the arrays over NVLink (peer copies)  */
} m2s_build_mode;
typedef enum m2s_host_path {
    M2S_HOST_AUTO = 0,      /* pinned destination: zero-copy; pageable: pipelined                               */
    M2S_HOST_STAGED = 1,    /* always device buffer + cudaMemcpy (the round-1 behaviour for pageable memory)    */
    M2S_HOST_PIPELINED = 2, /* always through the pinned ring                                                   */
    M2S_HOST_REGISTER = 3   /* cudaHostRegister the destination per call                                        */
} m2s_host_path;

/* ---- context ------------------------------------------------------------------------------------ */

/* Create a context on `n_devices` CUDA devices (`devices == NULL || n_devices == 0` -> device 0; an ordinal may be
 * listed more than once: every entry gets its own streams, arenas and slab).
 * With more than one device, m2s_generate_grid_sdf shards the grid by slabs along x (the slowest
 * axis of get_cell_idx) and m2s_generate_sdf shards the queries by contiguous ranges. */
M2S_API m2s_status m2s_create(const int* devices, int n_devices, m2s_ctx** out);

/* Same, single device, but all work is enqueued on the caller's `cudaStream_t` (passed as void*;
 * NULL = the legacy default stream). Lets a host framework time / order the library's work with
 * its own events. */
M2S_API m2s_status m2s_create_on_stream(int device, void* cuda_stream, m2s_ctx** out);

M2S_API void m2s_destroy(m2s_ctx* ctx);

/* Human-readable description of the last non-OK status on this context ("" if none). */
M2S_API const char* m2s_last_error(const m2s_ctx* ctx);

/* Same, copied into `buf` (NUL-terminated, truncated to n) while the context is locked: safe against a
 * concurrent call on another thread overwriting the message between the failing call and the read. */
M2S_API m2s_status m2s_last_error_copy(m2s_ctx* ctx, char* buf, size_t n);

/* Timings of the last call on the context's first device / on device `index` of the context. */
M2S_API m2s_status m2s_last_timings(const m2s_ctx* ctx, m2s_timings* out);
M2S_API m2s_status m2s_last_timings_device(const m2s_ctx* ctx, int index, m2s_timings* out);

M2S_API m2s_status m2s_set_option(m2s_ctx* ctx, int option, int64_t value);

/* Number of kernels this library launched on behalf of `ctx` since creation (bench.py's gpu_launches). */
M2S_API uint64_t m2s_launch_count(const m2s_ctx* ctx);

M2S_API int m2s_abi_version(void);

/* Diagnostics (development builds compiled with -DM2S_STATS_BUILD and M2S_STATS=1 in the environment): the traversal
 * counts {internal nodes visited, leaf triangles queued, tiles, tiles without a neighbour seed} of the last call on
 * the first device; zeros in the product build. */
M2S_API m2s_status m2s_debug_stats(m2s_ctx* ctx, uint64_t out[4]);

/* Number of devices the context drives. */
M2S_API int m2s_device_count(const m2s_ctx* ctx);

/* ---- host-buffer entry points: the drop-in boundary ----------------------------------------------- */

/* generate_grid_sdf(vertices, indices, grid, sign_method) -> Vec<f32>
 *   replaces src/generate/grid.rs:265-378 (generate_preheap :383-457, generate_heap :464-490,
 *   propagate_heap :495-558, compute_raycasts :568-642).
 * first_cell / cell_size / cell_count are Grid's three fields (src/grid.rs:30-37).
 * `out` has nx*ny*nz floats. nt == 0 fills f32::MAX (what the reference's un-seeded grid returns). */
M2S_API m2s_status m2s_generate_grid_sdf(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv,
                                 const uint32_t* tri_idx, uint64_t nt, const float first_cell[3],
                                 const float cell_size[3], const uint64_t cell_count[3], int sign_method,
                                 float* out);

/* One slab of the same grid, host buffers, single-device context: cells x in [x_begin, x_end) are written
 * at out_slab[(x - x_begin)*ny*nz + y*nz + z]. This is the per-process call of a one-process-per-GPU
 * deployment (each rank owns a contiguous range of the flat Vec<f32>, src/grid.rs:122-124). */
M2S_API m2s_status m2s_generate_grid_sdf_slab(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv,
                                              const uint32_t* tri_idx, uint64_t nt, const float first_cell[3],
                                              const float cell_size[3], const uint64_t cell_count[3],
                                              int sign_method, uint64_t x_begin, uint64_t x_end, float* out_slab);

/* generate_sdf(vertices, indices, query_points, acceleration_method) -> Vec<f32>
 *   replaces src/lib.rs:291-311 and the four drivers it dispatches to:
 *   generic/default.rs:11-74 (NONE), generic/bvh.rs:52-145 + bvh_ext.rs (BVH),
 *   generic/rtree.rs:87-126 (RTREE), generic/rtree_bvh.rs:79-174 (RTREE_BVH).
 * `out` has nq floats, in query order. nt == 0: NONE/BVH fill f32::MAX per query; RTREE and
 * RTREE_BVH return M2S_EEMPTY without touching `out`. */
M2S_API m2s_status m2s_generate_sdf(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                            uint64_t nt, const float* queries_xyz, uint64_t nq, int accel_method,
                            int sign_method, float* out);

/* ---- device-buffer entry points ---------------------------------------------------------------------
 * Same semantics, but every pointer is a device pointer on the context's FIRST device and all work is
 * only ENQUEUED: the call returns without synchronising. On a multi-device context (m2s_create(devices, n),
 * peer access between the devices required) the other devices pull the mesh over NVLink, compute their x-slabs
 * (query ranges) and store them STRAIGHT INTO d_out on the first device through peer-mapped pointers - the
 * reassembly of the flat Vec<f32> (src/grid.rs:122-124) is fused into the distance kernel's epilogue, there is no
 * gather step; the first device's stream then waits for the others, so stream order on it covers the whole grid. Data errors
 * (M2S_EINDEX, M2S_ENAN) are detected on the device and reported by the next m2s_synchronize().
 * The grid variant computes the slab
 * x in [x_begin, x_end) of the full grid and writes it at out_slab[(x - x_begin)*ny*nz + y*nz + z]:
 * this is the per-rank call of the multi-GPU path (one process per GPU, slabs along x). */
M2S_API m2s_status m2s_generate_grid_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv,
                                        const uint32_t* d_tri_idx, uint64_t nt, const float first_cell[3],
                                        const float cell_size[3], const uint64_t cell_count[3],
                                        int sign_method, uint64_t x_begin, uint64_t x_end, float* d_out_slab);

M2S_API m2s_status m2s_generate_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv,
                                   const uint32_t* d_tri_idx, uint64_t nt, const float* d_queries_xyz,
                                   uint64_t nq, int accel_method, int sign_method, float* d_out);

/* ---- mesh handles: upload + build once, query many times ---------------------------------------------------
 * The reference rebuilds its acceleration structures on every call (generate/grid.rs:95-111 and every generic driver), and so
 * do the one-shot entry points above. Its in-repo caller regenerates the grid of ONE mesh whenever a grid parameter
 * changes (mesh_to_sdf_client/src/sdf_program.rs:679-721): a handle keeps the mesh and its LBVH on every device of
 * the context, so those calls pay neither the upload nor the build (m2s_timings.h2d_ms == build_ms == 0). */
M2S_API m2s_status m2s_mesh_create(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                                   uint64_t nt, m2s_mesh** out);
/* device pointers on the context's first device; enqueue only (data errors surface at the next synchronize) */
M2S_API m2s_status m2s_mesh_create_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv,
                                          const uint32_t* d_tri_idx, uint64_t nt, m2s_mesh** out);
M2S_API void m2s_mesh_destroy(m2s_mesh* mesh);
/* host destination, cells x in [x_begin, x_end) (whole grid: 0, cell_count[0]) */
M2S_API m2s_status m2s_mesh_grid_sdf(m2s_ctx* ctx, m2s_mesh* mesh, const float first_cell[3], const float cell_size[3],
                                     const uint64_t cell_count[3], int sign_method, uint64_t x_begin, uint64_t x_end,
                                     float* out_slab);
M2S_API m2s_status m2s_mesh_grid_sdf_device(m2s_ctx* ctx, m2s_mesh* mesh, const float first_cell[3],
                                            const float cell_size[3], const uint64_t cell_count[3], int sign_method,
                                            uint64_t x_begin, uint64_t x_end, float* d_out_slab);
M2S_API m2s_status m2s_mesh_sdf(m2s_ctx* ctx, m2s_mesh* mesh, const float* queries_xyz, uint64_t nq, int accel_method,
                                int sign_method, float* out);
M2S_API m2s_status m2s_mesh_sdf_device(m2s_ctx* ctx, m2s_mesh* mesh, const float* d_queries_xyz, uint64_t nq,
                                       int accel_method, int sign_method, float* d_out);

/* ---- memory helpers ---------------------------------------------------------------------------------------------
 * Page-locked host memory is written in place by the distance kernel (zero-copy stores over PCIe): a facade that
 * returns / fills such a buffer (`PinnedVec`, `generate_grid_sdf_into`) skips every copy of the result.
 * m2s_host_register page-locks memory the caller owns (a long-lived Vec, a shared-memory segment that several
 * processes fill slab by slab); it must be unregistered before it is freed. */
M2S_API m2s_status m2s_host_alloc(size_t bytes, void** out);
M2S_API void m2s_host_free(void* p);
M2S_API m2s_status m2s_host_register(void* p, size_t bytes);
M2S_API m2s_status m2s_host_unregister(void* p);

/* Device memory on the context's first device that other PROCESSES can map (one process per GPU): rank 0 allocates
 * the flat grid and exports it, every rank opens it and passes `base + slab offset` as d_out_slab of
 * m2s_generate_grid_sdf_device - its distance kernel then stores its slab straight into rank 0's memory over NVLink.
 * handle: 64 opaque bytes (cudaIpcMemHandle_t) to be sent to the other ranks by the host program. */
M2S_API m2s_status m2s_device_alloc(m2s_ctx* ctx, size_t bytes, void** d_out);
M2S_API m2s_status m2s_device_free(m2s_ctx* ctx, void* d_ptr);
M2S_API m2s_status m2s_ipc_export(m2s_ctx* ctx, void* d_ptr, unsigned char handle[64]);
M2S_API m2s_status m2s_ipc_open(m2s_ctx* ctx, const unsigned char handle[64], void** d_out);
M2S_API m2s_status m2s_ipc_close(m2s_ctx* ctx, void* d_ptr);

/* Waits for the context's stream(s) and returns the deferred status of the device-buffer calls
 * enqueued since the previous m2s_synchronize (M2S_OK, M2S_EINDEX, M2S_ENAN or M2S_ECUDA). */
M2S_API m2s_status m2s_synchronize(m2s_ctx* ctx);

/* ---- post-passes on a finished grid SDF (what the reference's in-repo caller runs next) ------------------
 * The reference's only caller of generate_grid_sdf (mesh_to_sdf_client/src/sdf.rs:50-123) sorts the cell
 * indices by distance, takes min / max and re-uploads the Vec to a GPU buffer that its shaders sample. These
 * entry points do the same work on the grid while it is still on the device. `*_device` variants take device
 * pointers on a single-device context and only enqueue on its stream (like the entry points above). */

/* order[i] = index of the i-th cell in ascending distance: `(0..n).sorted_by(|i, j| data[i].total_cmp(&data[j]))`
 * (stable: equal values keep index order), mesh_to_sdf_client/src/sdf.rs:65-68. minmax = {min, max} as
 * `data.iter().copied().minmax()` returns them (:123; -0.0 == +0.0 there: the first minimum / last maximum wins).
 * Either output may be NULL. n < 2^31. NaN distances: order follows total_cmp, minmax is unspecified. */
M2S_API m2s_status m2s_grid_order(m2s_ctx* ctx, const float* sdf, uint64_t n, uint32_t* order, float minmax[2]);
M2S_API m2s_status m2s_grid_order_device(m2s_ctx* ctx, const float* d_sdf, uint64_t n, uint32_t* d_order,
                                         float* d_minmax);

/* raymarch_mode of mesh_to_sdf_client/shaders/draw_raymarching.wgsl (MODE_SNAP / MODE_TRILINEAR / MODE_TETRAHEDRAL) */
typedef enum m2s_sample_mode { M2S_SAMPLE_SNAP = 0, M2S_SAMPLE_TRILINEAR = 1, M2S_SAMPLE_TETRAHEDRAL = 2 } m2s_sample_mode;

/* out[i] = sdf_grid(points[i], iso) of draw_raymarching.wgsl:118-200: 100.0 outside [first_cell,
 * Grid::get_last_cell()], else the grid value (minus iso) snapped to the containing cell / interpolated on the
 * dual grid of cell centres with border clamping (:92-99) trilinearly or over the six-tetrahedra split
 * (:585-650). This is the "distance from any point with interpolation" the reference lists as TODO
 * (mesh_to_sdf/src/grid.rs:172). Arithmetic is un-fused fp32 in the shader's evaluation order. */
M2S_API m2s_status m2s_sample_grid_sdf(m2s_ctx* ctx, const float* sdf, const float first_cell[3],
                                       const float cell_size[3], const uint64_t cell_count[3],
                                       const float* points_xyz, uint64_t np, int sample_mode, float iso, float* out);
M2S_API m2s_status m2s_sample_grid_sdf_device(m2s_ctx* ctx, const float* d_sdf, const float first_cell[3],
                                              const float cell_size[3], const uint64_t cell_count[3],
                                              const float* d_points_xyz, uint64_t np, int sample_mode, float iso,
                                              float* d_out);

/* ---- host-side helpers the facade shares with the tests --------------------------------------------- */

/* Topology::get_triangles, src/lib.rs:175-193. `indices == NULL` is `None` (0..nv). index_bytes is 2
 * (u16) or 4 (u32). Returns the triangle count; writes 3*count u32 when out != NULL.
 * TriangleList drops a trailing partial tuple; TriangleStrip windows are not winding-flipped. */
M2S_API uint64_t m2s_expand_topology(int topology, const void* indices, int index_bytes, uint64_t n_indices,
                             uint64_t nv, uint32_t* out);

/* Grid::from_bounding_box, src/grid.rs:59-74. */
M2S_API void m2s_grid_from_bounding_box(const float bbox_min[3], const float bbox_max[3],
                                const uint64_t cell_count[3], float first_cell[3], float cell_size[3]);

#ifdef __cplusplus
}
#endif
#endif /* M2S_H */
