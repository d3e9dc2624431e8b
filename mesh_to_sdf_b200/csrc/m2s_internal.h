// Internal declarations shared by the translation units of libm2s.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/m2s.h"

namespace m2s {

// ---------------------------------------------------------------------------------------------------
// Device-side layouts (all in HBM; the whole mesh + LBVH of config C3 is ~14 MB and stays L2-resident)
//
//   triangle record, 48 B = 3 x float4 (one per triangle, two copies: original order for the
//   brute-force / row kernels, leaf (Morton) order for the LBVH kernels):
//       r0 = (a.x, a.y, a.z, b.x)   r1 = (b.y, b.z, c.x, c.y)   r2 = (c.z, n.x, n.y, n.z)
//       n = cross(b - a, c - a), un-normalised, un-fused (geo.rs:60-64)
//   triangle oriented box, 64 B = 4 x float4 (leaf order): (centre.xyz, 0) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew)
//   with u = unit normal, v along the longest edge: the cheap pretest in front of the exact arithmetic
//   box node, 64 B = 4 x float4 (`boxes`): per child (min.xyz, bits(child ref)) (max.xyz, 0), padded
//   -/+1e-4 like geo.rs:18-21. Written by the refit; read by the ray walks.
//   search node, 128 B = 8 x float4 (`nodes`): per child 4 x float4 = an oriented box
//   (centre.xyz, ref) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew), either fitted to the child's triangles or,
//   where that is not smaller, the padded axis-aligned box written with the identity frame; m2s_build.cu.
//   child ref: >= 0 internal node index; < 0 leaf: bit31 set, bit30 = "leaf holds a degenerate
//   triangle" (slow path with the geo.rs:73-88 guards), bits 0..29 = leaf index. Leaf l owns the
//   sorted triangles [l*K, min((l+1)*K, nt)).
// ---------------------------------------------------------------------------------------------------
constexpr int NODE_F4 = 8;   // float4 per node
constexpr int CHILD_F4 = 4;  // float4 per child slot
constexpr int BOX_F4 = 4;    // float4 per box node
constexpr uint32_t LEAF_BIT = 0x80000000u;
constexpr uint32_t LEAF_DEGEN_BIT = 0x40000000u;
constexpr uint32_t LEAF_INDEX_MASK = 0x3fffffffu;
constexpr uint32_t TRI_DEGEN_BIT = 0x80000000u;  // in tri_id_sorted

// Written by the build kernels, read back once per call (64 B).
struct BuildStatus {
    int bad_index;      // some triangle index >= nv
    int nonfinite;      // some referenced vertex / query / grid parameter is NaN or +-inf
    int n_degenerate;   // triangles with two equal vertices
    int stack_overflow; // traversal stack overflow (cannot happen for depth <= 128; checked anyway)
                        // the four error flags above/below are sticky until launch_status_reset(clear)
    int nan_distance;   // brute-force Normal fold hit the reference's "NaN distance" panic
    int lo[3];          // scene bounds (padded boxes), order-preserving int encoding of float
    int hi[3];
    int pad[5];
};
static_assert(sizeof(BuildStatus) == 64, "BuildStatus must be 64 bytes");

struct Bvh {
    const float4* rec;        // leaf-order triangle records
    const float4* tobb;       // leaf-order triangle oriented boxes (pretest)
    const float4* boxes;      // box nodes (ray walks)
    const uint32_t* tri_id;   // leaf-order -> original triangle id (| TRI_DEGEN_BIT)
    const float4* nodes;      // internal nodes
    const float4* nodes_il;   // the same nodes, children interleaved for packed-fp32 tests (m2s_build.cu, K4e)
    uint32_t nt;              // triangles
    uint32_t nleaf;           // leaves
    uint32_t leaf_size;       // K
    uint32_t root;            // root ref (a leaf ref when nleaf == 1)
    const BuildStatus* st;    // device pointer: scene bounds (-> pruning slack) and error flags
    unsigned long long* stats; // optional traversal counters (M2S_STATS=1): nodes, leaves, searches; with
                               // -DM2S_STATS_BUILD also [8 + k]: visits of nodes spanning [2^k, 2^(k+1)) leaves
    const uint2* node_range;   // per internal node: first / last leaf (diagnostics only)
};

struct GridParams {
    float fx, fy, fz;     // Grid::first_cell
    float sx, sy, sz;     // Grid::cell_size
    uint32_t nx, ny, nz;  // Grid::cell_count
    uint32_t x0, x1;      // slab [x0, x1): origin of the output / seed indexing
    uint32_t xa, xb;      // planes [xa, xb) of the slab computed by this launch (chunked host copies)
};

#ifdef __CUDACC__
// Scale of the interleaved nodes (k_nodes_interleave / k_grid_nearest_run): 1 / S with S the power of two
// >= 4 x mag, mag = largest |coordinate| of mesh and grid. Both kernels derive it from the same device-side
// inputs, so the host never has to know the mesh bounds.
__device__ __forceinline__ float pair_inv_scale(float mag) {
    if (!(mag > 0.0f) || !isfinite(mag)) return 1.0f;
    int e;
    frexpf(4.0f * mag, &e);  // 4 mag = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, -e);
}
#endif

// Row parity bitmaps for the grid Raycast sign (generate/grid.rs:568-642). For axis A the rows are
// the nB*nC rays starting on the face cell A=0; bit i of a row = parity of the hits whose last
// incremented cell k satisfies k >= i. Layout: [word][row] per axis so neighbouring rows are adjacent.
struct RowBits {
    uint32_t* bits[3];
    uint32_t rows[3];   // rows per axis: X: ny*nz, Y: nx*nz, Z: nx*ny
    uint32_t words[3];  // ceil(n_axis / 32)
};

enum QueryMode : int {
    MODE_UNSIGNED = 0,  // min |d|                      (Raycast paths)
    MODE_NORMAL = 1,    // positive-wins near-tie rule  (lib.rs:242-259)
    MODE_ARGMIN = 2     // sign of the single nearest triangle (rtree.rs:116-123)
};

// Growable device buffer, reused across calls (no cudaMalloc on the steady-state path).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes);
    void release();
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// Everything a context owns on one device.
struct Device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;

    // mesh + LBVH
    DevBuf verts, tris;  // staging for the host entry points
    DevBuf rec_orig, rec_sorted, tri_lo, tri_hi, keys_in, keys_out, vals_in, vals_out, cub_tmp;
    DevBuf tri_id_sorted, nodes, nodes_il, leaf_parent, node_parent, node_flag, node_range, tobb, boxes, status;
    DevBuf rows[3], big_list, big_count;
    DevBuf stats;             // traversal counters, only with M2S_STATS=1
    bool want_stats = false;
    int stats_mode = 0;
    float obb_bias = 1.0f;     // oriented box kept when its volume <= obb_bias * padded box volume (M2S_OBB_BIAS)
    bool seed_packet = false;  // M2S_SEED_PACKET=0: per-lane traversal for the seed pass
    bool packet = true;        // M2S_PACKET=0 selects the per-lane traversal grid kernel
    int pair = 1;              // M2S_PAIR: 0 = one voxel per lane (k_grid_nearest_pkt); 1 = k_grid_nearest_run, V = 2 voxels
                               // per lane; 4..7 = its (V, LAYOUT) variants for A/B runs
    DevBuf tile_slot;         // per-tile nearest-triangle slots published by the distance kernel
    bool neighbour_and_coarse = false;  // experiment (M2S_NSEED=2): coarse pass as the fallback of neighbour seeds
    bool neighbour_seeds = true;  // M2S_NSEED=0: separate coarse seed pass for every grid
    DevBuf seeds[2];          // nearest-triangle slots of the coarse seeding levels
    uint32_t seed_stride = 4;  // voxels per seed block edge (M2S_SEED_STRIDE)
    int seed_levels = 1;      // 0 disables the coarse-to-fine seeding (M2S_SEED_LEVELS)
    DevBuf queries, q_sorted, q_perm, q_keys_in, q_keys_out, q_vals_in, out;
    DevBuf post_keys, post_idx, post_mm, post_in, post_pts, post_out;  // post-passes (m2s_post.cu) and their host staging
    BuildStatus* h_status = nullptr;  // pinned
    cudaEvent_t ev[8] = {};
    cudaStream_t aux_stream = nullptr;   // high priority: the second half's seed pass hides under the first half's kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_half[2] = {};
    cudaEvent_t ev_records = nullptr, ev_rows = nullptr;  // the row parities run on aux_stream beside the tree build
    bool split_halves = true;            // M2S_SPLIT=0: one seed pass + one distance launch per slab
    bool last_split = false;
    bool zero_copy = true;               // M2S_ZEROCOPY=0: stage + copy even when the host destination is pinned
    cudaStream_t copy_stream = nullptr;  // D2H of finished x-chunks overlaps the next chunk's kernel
    cudaEvent_t ev_chunk[8] = {};
    cudaEvent_t ev_copied = nullptr;

    Bvh bvh{};
    float nodes_il_mag = -1.0f;  // grid magnitude the interleaved nodes were last written for (< 0: stale)
};

// ---- launchers (each returns the CUDA error of its enqueue) ------------------------------------------
cudaError_t launch_status_reset(Device& d, bool clear_errors);
// after_records (optional) is recorded once the original-order triangle records exist (what the row kernels need)
cudaError_t launch_build(Device& d, const float* d_verts, uint64_t nv, const uint32_t* d_tris, uint64_t nt,
                         uint32_t leaf_size, cudaEvent_t after_records = nullptr);
cudaError_t launch_nodes_interleave(Device& d, float grid_mag);
cudaError_t sort_queries(Device& d, const float* d_queries, uint64_t nq);

// stream / rec default to the device's stream and the leaf-order records
cudaError_t launch_grid_rows(Device& d, const GridParams& g, RowBits* rb, cudaStream_t stream = nullptr,
                             const float4* rec = nullptr);

struct SeedLevel {
    const uint32_t* parent;  // nearest-triangle slots of the parent level (nullptr: start unbounded)
    uint32_t px, py, pz;     // parent level dims
    uint32_t pstride;        // parent level stride in voxels
};
bool grid_uses_neighbour_seeds(const Device& d, int mode);
cudaError_t launch_grid_seeds(Device& d, const GridParams& g, SeedLevel* L, int slot = 0, cudaStream_t stream = nullptr);
cudaError_t launch_grid_final(Device& d, const GridParams& g, const SeedLevel& L, int mode, const RowBits* rb,
                              float* d_out);
cudaError_t launch_grid_nearest(Device& d, const GridParams& g, int mode, const RowBits* rb, float* d_out,
                                cudaEvent_t after_seeds = nullptr);

// queries must have been Morton-sorted into d.q_sorted by sort_queries().
// sign_rule: 0 = value already signed / unsigned, 1 = +X ray parity, 3 = best of the three axes
cudaError_t launch_points(Device& d, uint64_t nq, int mode, int sign_rule, float* d_out,
                          cudaEvent_t after_seeds = nullptr);

cudaError_t launch_fill(Device& d, float* d_out, uint64_t n, float value);

// post-passes on a device-resident grid (m2s_post.cu)
cudaError_t launch_grid_order(Device& d, const float* d_sdf, uint64_t n, uint32_t* d_order, float* d_minmax);
cudaError_t launch_grid_sample(Device& d, const float* d_sdf, const GridParams& g, const float* d_points, uint64_t np,
                               int mode, float iso, float* d_out);

}  // namespace m2s

struct m2s_ctx {
    int n_devices = 0;
    m2s::Device* dev = nullptr;
    std::string last_error;
    m2s_timings timings{};
    uint32_t leaf_size = 1;
    std::mutex mu;  // a context serves one call at a time
};
