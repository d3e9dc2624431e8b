// Internal declarations shared by the translation units of libm2s.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/m2s.h"

namespace m2s {

// ---------------------------------------------------------------------------------------------------
// Device-side layouts (all in HBM; the whole mesh + LBVH of config C3 is ~26 MB and stays L2-resident)
//
//   triangle record, 48 B = 3 x float4, leaf (Morton) order:
//       r0 = (a.x, a.y, a.z, b.x)   r1 = (b.y, b.z, c.x, c.y)   r2 = (c.z, n.x, n.y, n.z)
//       n = cross(b - a, c - a), un-normalised, un-fused (geo.rs:60-64)
//   triangle oriented box, 64 B = 4 x float4 (leaf order): (centre.xyz, 0) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew)
//   with u = unit normal, v along the longest edge: becomes the child slot of the triangle's parent node
//   box node, 64 B = 4 x float4 (`boxes`): per child (min.xyz, bits(child ref)) (max.xyz, escape), padded
//   -/+1e-4 like geo.rs:18-21. Written by the refit; read by the ray walks.
//   search node, 128 B = 8 x float4 (`nodes`): per child 4 x float4 = an oriented box
//   (centre.xyz, ref) (u.xyz, eu) (v.xyz, ev) (w.xyz, ew), either fitted to the child's triangles or,
//   where that is not smaller, the padded axis-aligned box written with the identity frame; m2s_build.cu.
//   child ref: >= 0 internal node index; < 0 leaf: bit31 set, bit30 = "the triangle is degenerate"
//   (slow path with the geo.rs:73-88 guards), bits 0..29 = leaf = sorted triangle index (one triangle per leaf).
//   A mesh of ONE triangle gets a root node whose second child is an unreachable copy of the first, so every
//   tree has at least one internal node and there is a single code path.
// ---------------------------------------------------------------------------------------------------
constexpr int NODE_F4 = 8;   // float4 per node
constexpr int CHILD_F4 = 4;  // float4 per child slot
constexpr int BOX_F4 = 4;    // float4 per box node
constexpr uint32_t LEAF_BIT = 0x80000000u;
constexpr uint32_t LEAF_DEGEN_BIT = 0x40000000u;
constexpr uint32_t LEAF_INDEX_MASK = 0x3fffffffu;
constexpr uint32_t TRI_DEGEN_BIT = 0x80000000u;  // in tri_id_sorted

// Status block, 64 B. One per mesh (written by the build: error flags of the mesh + its bounds) and one per
// call (k_call_status_init copies the mesh's block, the query kernels add the query bounds and their own flags).
struct BuildStatus {
    int bad_index;      // some triangle index >= nv
    int nonfinite;      // some referenced vertex / query / grid parameter is NaN or +-inf
    int n_degenerate;   // triangles with two equal vertices
    int stack_overflow; // packet stack overflow (cannot happen for depth <= 128; checked anyway)
    int nan_distance;   // Normal fold hit the reference's "NaN distance" panic (lib.rs:257)
    int lo[3];          // scene bounds (padded boxes), order-preserving int encoding of float
    int hi[3];
    int pad[5];
};
static_assert(sizeof(BuildStatus) == 64, "BuildStatus must be 64 bytes");

// Ray bins (scattered queries, Raycast sign rules): per axis A a uniform R x R grid over the two other coordinates
// of the mesh's bounds; a cell lists the triangles whose PADDED box (geo.rs:4-22, the bvh crate's filter) overlaps it
// in that projection. The axis-aligned ray of a query only has to test the triangles of its own cell - what
// bvh.traverse returns for that ray, found without a tree walk. Triangles that cover more than RAYBIN_BIG_CELLS cells
// go to a short per-axis list every query scans. `ok` == 0 (too many items / big triangles for the budgets: a mesh of
// huge overlapping triangles) makes the query kernel fall back to the packet walk of the box tree.
constexpr uint32_t RAYBIN_BIG_CELLS = 64;
constexpr uint32_t RAYBIN_MAX_BIG = 256;       // per axis
constexpr uint32_t RAYBIN_ITEMS_PER_TRI = 12;  // item capacity per axis = this x nt
struct RayBins {
    uint32_t R;               // cells per in-plane axis (power of two)
    const uint32_t* offsets;  // [3][R*R] + 1: exclusive prefix of the cell counts (axis-major, cell = k * R + j)
    const uint32_t* items;    // leaf-order triangle slots
    const uint32_t* big;      // [3][RAYBIN_MAX_BIG]
    const uint32_t* meta;     // [0..2] big count per axis, [3] total items, [4] ok flag
    const BuildStatus* mesh;  // the mesh's bounds (lo / hi of the padded boxes)
};

struct Bvh {
    const float4* rec;        // leaf-order triangle records
    const float4* boxes;      // box nodes (ray walks)
    const uint32_t* tri_id;   // leaf-order -> original triangle id (| TRI_DEGEN_BIT)
    const float4* nodes;      // search nodes
    const float4* nodes_il;   // the same nodes, children interleaved for packed-fp32 tests (k_nodes_interleave)
    uint32_t nt;              // triangles = leaves
    uint32_t n_nodes;         // internal nodes (>= 1 when nt >= 1)
    unsigned long long* stats; // optional traversal counters (-DM2S_STATS_BUILD + M2S_STATS=1): nodes, leaves, tiles
    const uint2* node_range;   // per internal node: first / last leaf (box fitting, diagnostics)
    RayBins bins;              // valid after launch_ray_bins (point calls with a Raycast sign rule)
};

struct GridParams {
    float fx, fy, fz;     // Grid::first_cell
    float sx, sy, sz;     // Grid::cell_size
    uint32_t nx, ny, nz;  // Grid::cell_count
    uint32_t x0, x1;      // slab [x0, x1) computed by this launch; the output starts at plane x0
};

// Completion signalling of the distance kernel for host destinations that are copied while it runs: every warp
// bumps count[brick plane] after its stores; the last one publishes flag[brick plane] = epoch in mapped host memory.
struct Progress {
    uint32_t* count;            // device, one per brick plane (4 x-planes), zeroed before the launch
    volatile uint32_t* flag;    // mapped pinned host memory (device alias), one per brick plane
    uint32_t epoch;
};

#ifdef __CUDACC__
// Scale of the interleaved nodes (k_nodes_interleave / the run kernels): 1 / S with S the power of two
// >= 4 x mag, mag = largest |coordinate| of mesh and grid / queries. Both sides derive it from the same device-side
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

// Growable page-locked + mapped + portable host buffer (staging ring of the pageable-destination path).
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes);
    void release();
};

// A mesh + its LBVH on one device: everything the query kernels read. Owned either by a Device (the scratch mesh
// of the one-shot entry points, rebuilt per call) or by an m2s_mesh handle (built once, queried many times).
struct MeshDev {
    DevBuf rec_sorted, tri_id_sorted, nodes, nodes_il, boxes, status, node_range;
    DevBuf bin_offsets, bin_cursor, bin_items, bin_big, bin_meta;  // ray bins, built at the first point call that needs them
    bool bins_built = false;
    Bvh bvh{};
    uint64_t nv = 0, nt = 0;
    float nodes_il_mag = -1.0f;  // magnitude key the interleaved nodes were last written for (< 0: stale)
    void release();
};

// Everything a context owns on one device.
struct Device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    bool peer_to_first = false;   // this device can store into / load from the context's first device

    DevBuf verts, tris;           // staging for the host entry points
    // build scratch (dead after launch_build)
    DevBuf rec_orig, tobb, tri_lo, tri_hi, keys_in, keys_out, vals_in, vals_out, sort_tmp;
    DevBuf leaf_parent, node_parent, node_flag;
    DevBuf slot_list, slot_count; // child slots whose box is fitted by a whole warp (k_search_nodes_big)
    MeshDev scratch;              // the mesh of the one-shot entry points
    DevBuf call_status;           // per-call BuildStatus (mesh block + query bounds + this call's flags)
    DevBuf rows[3], big_list, big_count;
    DevBuf stats;                 // traversal counters (stats builds only)
    bool want_stats = false;
    uint32_t run_v = 0;           // M2S_OPT_RUN_LENGTH: 0 = by mesh size, 2 / 4 = voxels per lane of the grid kernel
    bool no_ray_bins = false;     // M2S_OPT_RAY_BINS = 0: ray parities always through the box tree (tests, A/B)
    DevBuf tile_slot;             // per-tile nearest-triangle slots published by the distance kernel
    DevBuf progress;              // per brick plane completion counters (pipelined host copies)
    DevBuf queries, q_sorted, q_perm, q_keys_in, q_keys_out, q_vals_in, out;
    DevBuf post_keys, post_idx, post_mm, post_in, post_pts, post_out;  // post-passes (m2s_post.cu) and their host staging
    PinBuf stage;                 // pinned staging ring: pageable destinations are filled from here by host threads
    PinBuf in_mesh, in_queries;   // pinned staging of pageable INPUTS (host threads copy chunks in, each followed by its H2D)
    PinBuf flags;                 // mapped completion flags of the pipelined path
    uint32_t epoch = 0;
    BuildStatus* h_status = nullptr;  // pinned
    cudaEvent_t ev[8] = {};       // 0 start, 1 inputs on device, 2 build done, 3 sign done, 4 kernel done, 5 end, 6 pre-kernel
    cudaStream_t aux_stream = nullptr;   // high priority: row parities beside the tree build
    cudaEvent_t ev_records = nullptr, ev_rows = nullptr;
    cudaEvent_t ev_inputs = nullptr, ev_done = nullptr, ev_built = nullptr;  // cross-device ordering (multi-device contexts)
    m2s_timings timings{};
};

// Host worker threads that copy finished plane groups from the pinned ring into pageable destinations.
class CopyPool {
public:
    explicit CopyPool(int n);
    ~CopyPool();
    void start(const std::function<void(int)>* job, int n_jobs);  // workers run (*job)(i) for i in [0, n_jobs)
    void wait();                                                   // until every job has returned
    int threads() const { return (int)workers_.size(); }
private:
    void loop();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* job_ = nullptr;
    std::atomic<int> next_{0};
    int total_ = 0, active_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

// ---- launchers (each returns the CUDA error of its enqueue) ------------------------------------------
cudaError_t launch_mesh_status_reset(Device& d, MeshDev& m);
cudaError_t launch_call_status_init(Device& d, const MeshDev& m, bool clear_errors);
// after_records (optional) is recorded once the leaf-order triangle records exist (what the row kernels need)
cudaError_t launch_build(Device& d, MeshDev& m, const float* d_verts, uint64_t nv, const uint32_t* d_tris, uint64_t nt,
                         cudaEvent_t after_records = nullptr);
cudaError_t launch_nodes_interleave(Device& d, MeshDev& m, float mag_key, bool force);
cudaError_t sort_queries(Device& d, const float* d_queries, uint64_t nq);
cudaError_t launch_ray_bins(Device& d, MeshDev& m);  // no-op when the mesh already has them

// rec: triangle records in any order (every triangle toggles its own rows)
cudaError_t launch_grid_rows(Device& d, const float4* rec, uint32_t nt, const GridParams& g, RowBits* rb, cudaStream_t stream);
cudaError_t launch_grid_nearest(Device& d, MeshDev& m, const GridParams& g, int mode, const RowBits* rb, float* d_out,
                                const Progress* progress);
uint32_t grid_brick_planes(const GridParams& g);  // number of completion flags a launch over g publishes
float grid_magnitude(const GridParams& g);       // largest |coordinate| of the grid's box: key of the node interleave
constexpr uint32_t GRID_BRICK_X = 4;               // x-planes per brick plane

// queries must have been Morton-sorted into d.q_sorted by sort_queries().
// sign_rule: 0 = value already signed / unsigned, 1 = +X ray parity, 3 = best of the three axes
cudaError_t launch_points(Device& d, MeshDev& m, uint64_t nq, int mode, int sign_rule, float* d_out);

cudaError_t launch_fill(Device& d, float* d_out, uint64_t n, float value, const Progress* progress, uint32_t planes);

// post-passes on a device-resident grid (m2s_post.cu)
cudaError_t launch_grid_order(Device& d, const float* d_sdf, uint64_t n, uint32_t* d_order, float* d_minmax);
cudaError_t launch_grid_sample(Device& d, const float* d_sdf, const GridParams& g, const float* d_points, uint64_t np,
                               int mode, float iso, float* d_out);

}  // namespace m2s

struct m2s_mesh {
    m2s_ctx* owner = nullptr;
    m2s::MeshDev* dev = nullptr;  // one per device of the owning context
    uint64_t nv = 0, nt = 0;
};

// Slab cuts of the last multi-device grid call and, once its timings are in, the cuts the next call on the same grid
// shape will use: equal-width x-slabs are uneven in cost (the slabs through the middle of a mesh do more work), so
// the cuts move to equal shares of the measured per-slab kernel time. Results do not depend on the cuts.
struct SlabBalance {
    uint64_t xa = 0, xb = 0, ny = 0, nz = 0, nt = 0;
    int nd = 0;
    int kind = 0;                // 0: device destination, 1: host destination (their kernel times differ: PCIe stores)
    std::vector<uint64_t> cuts;  // nd + 1 entries, cuts[0] = xa, cuts[nd] = xb
    bool valid = false;          // cuts hold a measured split for the key above
    bool pending = false;        // a call with these cuts is in flight / finished and its timings were not used yet
};

struct m2s_ctx {
    int n_devices = 0;
    m2s::Device* dev = nullptr;
    std::string last_error;
    int build_mode = M2S_BUILD_REPLICATED;
    int host_path = M2S_HOST_AUTO;
    int copy_threads = 4;
    bool balance_slabs = true;     // M2S_OPT_BALANCE
    SlabBalance balance;
    m2s::CopyPool* pool = nullptr;
    std::mutex mu;  // a context serves one call at a time
};
