// C ABI of libm2s.so (include/m2s.h): contexts, argument checks, mesh handles, host <-> device staging, slab
// sharding over the context's devices with the reassembly fused into the distance kernel's stores (peer-mapped
// device memory / page-locked host memory), the pipelined copy into pageable destinations, deferred status.
// No CPU fallback: without a CUDA device m2s_create fails and nothing else can be called.
//
// Boundary replaced: the public free functions of the Rust crate, mesh_to_sdf/src/lib.rs:291-311
// (generate_sdf) and src/generate/grid.rs:265-378 (generate_grid_sdf); their infallible signatures
// panic where this ABI returns a status (lib.rs:257 "NaN distance", slice index panics, rtree.rs:117).
#include <sys/mman.h>

#include <algorithm>
#include <cerrno>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "m2s_internal.h"

using namespace m2s;

// ---- host copy threads ---------------------------------------------------------------------------------------
namespace m2s {

CopyPool::CopyPool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lock(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
}

void CopyPool::start(const std::function<void(int)>* job, int n_jobs) {
    {
        std::lock_guard<std::mutex> lock(mu_);
        job_ = job;
        next_.store(0);
        total_ = n_jobs;
        active_ = (int)workers_.size();
        ++generation_;
    }
    cv_.notify_all();
}

void CopyPool::wait() {
    std::unique_lock<std::mutex> lock(mu_);
    done_cv_.wait(lock, [this] { return active_ == 0; });
    job_ = nullptr;
}

void CopyPool::loop() {
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)>* job;
        int total;
        {
            std::unique_lock<std::mutex> lock(mu_);
            cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            job = job_;
            total = total_;
        }
        for (int i = next_.fetch_add(1); i < total; i = next_.fetch_add(1)) (*job)(i);
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (--active_ == 0) done_cv_.notify_all();
        }
    }
}

}  // namespace m2s

namespace {

// restores the caller's current device on every exit path
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

m2s_status fail(m2s_ctx* ctx, m2s_status s, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return s;
}

m2s_status cuda_fail(m2s_ctx* ctx, cudaError_t e, const char* where) {
    cudaGetLastError();
    return fail(ctx, M2S_ECUDA, std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define CU(ctx, expr)                                          \
    do {                                                       \
        cudaError_t e__ = (expr);                              \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr); \
    } while (0)

bool finite3(const float* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

struct GridArgs {
    GridParams g;
    uint64_t total;
};

m2s_status check_mesh(m2s_ctx* ctx, const void* verts, uint64_t nv, const void* tris, uint64_t nt) {
    if (nt > 0 && (!tris || !verts)) return fail(ctx, M2S_EINVAL, "null vertex / index pointer");
    if (nt > 0 && nv == 0) return fail(ctx, M2S_EINDEX, "triangles reference an empty vertex array");
    if (nt >= (1ull << 30)) return fail(ctx, M2S_EINVAL, "more than 2^30 triangles");
    if (nv > 0xffffffffull) return fail(ctx, M2S_EINVAL, "more than 2^32 vertices");
    return M2S_OK;
}

m2s_status check_grid(m2s_ctx* ctx, const float first[3], const float size[3], const uint64_t count[3], int sign,
                      GridArgs* out) {
    if (!first || !size || !count) return fail(ctx, M2S_EINVAL, "null grid parameter");
    if (sign != M2S_SIGN_RAYCAST && sign != M2S_SIGN_NORMAL) return fail(ctx, M2S_EINVAL, "unknown sign method");
    for (int i = 0; i < 3; ++i)
        if (count[i] > 0x7fffffffull) return fail(ctx, M2S_EINVAL, "cell_count component above 2^31-1");
    out->total = count[0] * count[1] * count[2];
    if (count[0] * count[1] > 0xffffffffull || count[0] * count[2] > 0xffffffffull ||
        count[1] * count[2] > 0xffffffffull)
        return fail(ctx, M2S_EINVAL, "grid face with more than 2^32 cells");
    if (!finite3(first) || !finite3(size)) return fail(ctx, M2S_ENAN, "non-finite grid parameter (NaN distance)");
    out->g.fx = first[0]; out->g.fy = first[1]; out->g.fz = first[2];
    out->g.sx = size[0]; out->g.sy = size[1]; out->g.sz = size[2];
    out->g.nx = (uint32_t)count[0]; out->g.ny = (uint32_t)count[1]; out->g.nz = (uint32_t)count[2];
    out->g.x0 = 0;
    out->g.x1 = out->g.nx;
    return M2S_OK;
}

m2s_status check_queries(m2s_ctx* ctx, uint64_t nq) {
    // queries are indexed with 32 bits on the device (sort payloads, packet bases)
    if (nq >= (1ull << 31)) return fail(ctx, M2S_EINVAL, "more than 2^31-1 query points");
    return M2S_OK;
}

struct PointPlan {
    int mode;
    int sign_rule;
};

// AccelerationMethod -> (distance rule, sign rule); SURVEY.md §8b.
bool plan_points(int accel, int sign, PointPlan* p) {
    if (sign != M2S_SIGN_RAYCAST && sign != M2S_SIGN_NORMAL) return false;
    switch (accel) {
        case M2S_ACCEL_NONE:  // default.rs:44-72: Raycast = +X parity only
            *p = sign == M2S_SIGN_NORMAL ? PointPlan{MODE_NORMAL, 0} : PointPlan{MODE_UNSIGNED, 1};
            return true;
        case M2S_ACCEL_BVH:  // bvh.rs:82-141
            *p = sign == M2S_SIGN_NORMAL ? PointPlan{MODE_NORMAL, 0} : PointPlan{MODE_UNSIGNED, 3};
            return true;
        case M2S_ACCEL_RTREE:  // rtree.rs:116-123
            *p = PointPlan{MODE_ARGMIN, 0};
            return true;
        case M2S_ACCEL_RTREE_BVH:  // rtree_bvh.rs:126-171
            *p = PointPlan{MODE_UNSIGNED, 3};
            return true;
        default:
            return false;
    }
}

m2s_status status_to_code(m2s_ctx* ctx, const BuildStatus& st) {
    if (st.bad_index) return fail(ctx, M2S_EINDEX, "triangle index out of bounds");
    if (st.nonfinite) return fail(ctx, M2S_ENAN, "non-finite vertex or query coordinate");
    if (st.nan_distance) return fail(ctx, M2S_ENAN, "NaN distance");
    if (st.stack_overflow) return fail(ctx, M2S_ECUDA, "LBVH traversal stack overflow");
    return M2S_OK;
}

// GPU-side phase timings of the last call on this device (the host-side tail of the pipelined path is added by
// its caller).
void collect_timings(Device& d, int host_path) {
    m2s_timings t{};
    cudaEventElapsedTime(&t.h2d_ms, d.ev[0], d.ev[1]);
    cudaEventElapsedTime(&t.build_ms, d.ev[1], d.ev[2]);
    cudaEventElapsedTime(&t.sign_ms, d.ev[2], d.ev[3]);
    cudaEventElapsedTime(&t.seed_ms, d.ev[3], d.ev[6]);
    cudaEventElapsedTime(&t.dist_ms, d.ev[6], d.ev[4]);
    cudaEventElapsedTime(&t.d2h_ms, d.ev[4], d.ev[5]);
    cudaEventElapsedTime(&t.total_ms, d.ev[0], d.ev[5]);
    cudaGetLastError();
    t.host_path = host_path;
    d.timings = t;
}

// ---- one device's share of a call ------------------------------------------------------------------------------

// Mesh of the call on device d: `handle` (already built) or the device's scratch mesh, built here from device
// pointers that are valid on d. Records ev[2] when the tree exists.
cudaError_t enqueue_mesh(Device& d, MeshDev* handle, const float* d_verts, uint64_t nv, const uint32_t* d_tris,
                         uint64_t nt, cudaEvent_t after_records, MeshDev** out) {
    cudaError_t e = cudaSuccess;
    if (handle) {
        *out = handle;
    } else {
        *out = &d.scratch;
        e = launch_build(d, d.scratch, d_verts, nv, d_tris, nt, after_records);
    }
    cudaEventRecord(d.ev[2], d.stream);
    return e;
}

// BUILD_BROADCAST: pull the finished arrays of `src` (on device sd) instead of building them.
cudaError_t enqueue_mesh_pull(Device& d, MeshDev& m, const Device& sd, const MeshDev& src) {
    cudaError_t e;
    const uint64_t nt = src.nt;
    m.bvh = Bvh{};
    m.nv = src.nv;
    m.nt = nt;
    m.nodes_il_mag = -1.0f;
    m.bins_built = false;
    if (nt == 0) return cudaSuccess;
    const size_t n_nodes = src.bvh.n_nodes;
    struct Item { DevBuf* dst; const DevBuf* s; size_t bytes; };
    const Item items[] = {{&m.rec_sorted, &src.rec_sorted, nt * 48},
                          {&m.tri_id_sorted, &src.tri_id_sorted, nt * 4},
                          {&m.nodes, &src.nodes, n_nodes * NODE_F4 * 16},
                          {&m.boxes, &src.boxes, n_nodes * BOX_F4 * 16},
                          {&m.status, &src.status, sizeof(BuildStatus)}};
    for (const Item& it : items) {
        if ((e = it.dst->ensure(it.bytes)) != cudaSuccess) return e;
        if ((e = cudaMemcpyPeerAsync(it.dst->p, d.ordinal, it.s->p, sd.ordinal, it.bytes, d.stream)) != cudaSuccess) return e;
    }
    if ((e = m.nodes_il.ensure(n_nodes * NODE_F4 * 16)) != cudaSuccess) return e;
    m.bvh.rec = m.rec_sorted.as<float4>();
    m.bvh.boxes = m.boxes.as<float4>();
    m.bvh.tri_id = m.tri_id_sorted.as<uint32_t>();
    m.bvh.nodes = m.nodes.as<float4>();
    m.bvh.nodes_il = m.nodes_il.as<float4>();
    m.bvh.nt = (uint32_t)nt;
    m.bvh.n_nodes = (uint32_t)n_nodes;
    return cudaSuccess;
}

// Row parities + distance kernel of one slab on a built mesh. rows_rec: triangle records for the row kernels
// (original order right after k_tri_setup on the one-shot path, so that the rows run beside the rest of the build
// on the side stream; the mesh's leaf-order records on a handle).
cudaError_t enqueue_grid_query(Device& d, MeshDev& m, const float4* rows_rec, bool rows_beside_build,
                               const GridParams& g, int sign, float* d_out, bool clear_errors, const Progress* pr) {
    cudaError_t e;
    if ((e = launch_call_status_init(d, m, clear_errors)) != cudaSuccess) return e;
    const uint64_t slab_cells = (uint64_t)(g.x1 - g.x0) * g.ny * g.nz;
    if (m.nt == 0) {
        cudaEventRecord(d.ev[3], d.stream);
        cudaEventRecord(d.ev[6], d.stream);
        e = launch_fill(d, d_out, slab_cells, FLT_MAX, pr, grid_brick_planes(g));  // un-seeded cells stay f32::MAX (grid.rs:137-143)
        cudaEventRecord(d.ev[4], d.stream);
        return e;
    }
    RowBits rb{};
    const bool raycast = sign == M2S_SIGN_RAYCAST;
    if (raycast) {
        if (rows_beside_build) {
            // the row parities only need the triangle records: they run on the side stream beside the sort /
            // hierarchy / refit / box fitting (small, latency-bound launches) and join before the distance kernel
            if ((e = cudaStreamWaitEvent(d.aux_stream, d.ev_records, 0)) != cudaSuccess) return e;
            if ((e = launch_grid_rows(d, rows_rec, (uint32_t)m.nt, g, &rb, d.aux_stream)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(d.ev_rows, d.aux_stream)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(d.stream, d.ev_rows, 0)) != cudaSuccess) return e;
        } else {
            if ((e = launch_grid_rows(d, rows_rec, (uint32_t)m.nt, g, &rb, d.stream)) != cudaSuccess) return e;
        }
    }
    cudaEventRecord(d.ev[3], d.stream);
    const int mode = raycast ? MODE_UNSIGNED : MODE_NORMAL;
    // node frames in units of S = 2^k >= 4 x the largest |coordinate| (a no-op when current for this grid)
    if ((e = launch_nodes_interleave(d, m, grid_magnitude(g), false)) != cudaSuccess) return e;
    cudaEventRecord(d.ev[6], d.stream);
    e = launch_grid_nearest(d, m, g, mode, raycast ? &rb : nullptr, d_out, pr);
    cudaEventRecord(d.ev[4], d.stream);
    return e;
}

cudaError_t enqueue_points_query(Device& d, MeshDev& m, const float* d_queries, uint64_t nq, const PointPlan& plan,
                                 float* d_out, bool clear_errors) {
    cudaError_t e;
    if ((e = launch_call_status_init(d, m, clear_errors)) != cudaSuccess) return e;
    if (m.nt == 0) {
        cudaEventRecord(d.ev[3], d.stream);
        cudaEventRecord(d.ev[6], d.stream);
        e = launch_fill(d, d_out, nq, FLT_MAX, nullptr, 0);  // default.rs:54 / bvh.rs:83 fold from f32::MAX
        cudaEventRecord(d.ev[4], d.stream);
        return e;
    }
    if ((e = sort_queries(d, d_queries, nq)) != cudaSuccess) return e;
    cudaEventRecord(d.ev[3], d.stream);
    cudaEventRecord(d.ev[6], d.stream);
    e = launch_points(d, m, nq, plan.mode, plan.sign_rule, d_out);
    cudaEventRecord(d.ev[4], d.stream);
    return e;
}

// Device-side address of a page-locked, mapped host range (nullptr if any part of it is pageable). Must be called
// with the target device current.
float* pinned_device_alias(const void* host_ptr, size_t bytes) {
    if (!host_ptr || bytes == 0) return nullptr;
    cudaPointerAttributes a{}, b{};
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    // the last byte must belong to the same mapping (a buffer that starts inside a registration and ends
    // outside of it would fault the kernel)
    const char* last = static_cast<const char*>(host_ptr) + (bytes - 1);
    if (cudaPointerGetAttributes(&b, last) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (b.type != cudaMemoryTypeHost || !b.devicePointer) return nullptr;
    if (static_cast<const char*>(b.devicePointer) - static_cast<const char*>(a.devicePointer) != (ptrdiff_t)(bytes - 1))
        return nullptr;
    return static_cast<float*>(a.devicePointer);
}

CopyPool* ensure_pool(m2s_ctx* ctx) {
    if (!ctx->pool || ctx->pool->threads() != ctx->copy_threads) {
        delete ctx->pool;
        ctx->pool = new CopyPool(ctx->copy_threads);
    }
    return ctx->pool;
}

// Host -> device copies of caller memory. A cudaMemcpyAsync from pageable memory is staged by the driver through its
// own bounce buffer by the calling thread (measured 11 GB/s: 1.6 ms for the 18 MB mesh of C5); from page-locked memory
// it is one DMA. Large pageable inputs therefore go through the library's pinned staging buffer: the context's host
// threads copy 1 MiB chunks into it and each enqueues the DMA of its chunk right away, so the memcpy of one chunk runs
// beside the DMA of another. `parts` are copied back to back into `stage`.
constexpr size_t H2D_STAGE_MIN_BYTES = 4u << 20, H2D_CHUNK = 1u << 20;
struct H2DPart {
    void* dst;
    const void* src;
    size_t bytes;
};
bool host_is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
cudaError_t upload_inputs(m2s_ctx* ctx, Device& d, PinBuf& stage, const H2DPart* parts, int n_parts,
                          cudaStream_t stream) {
    size_t total = 0;
    bool pageable = false;
    for (int i = 0; i < n_parts; ++i) {
        total += parts[i].bytes;
        if (parts[i].bytes && !host_is_pinned(parts[i].src)) pageable = true;
    }
    if (!pageable || total < H2D_STAGE_MIN_BYTES || ctx->host_path == M2S_HOST_STAGED) {
        for (int i = 0; i < n_parts; ++i)
            if (parts[i].bytes) {
                cudaError_t e = cudaMemcpyAsync(parts[i].dst, parts[i].src, parts[i].bytes, cudaMemcpyHostToDevice, stream);
                if (e != cudaSuccess) return e;
            }
        return cudaSuccess;
    }
    // the previous call's DMAs out of this buffer have completed: host-buffer calls return synchronised
    cudaError_t e = stage.ensure(total);
    if (e != cudaSuccess) return e;
    struct Chunk {
        char* dst;
        const char* src;
        char* pin;
        size_t bytes;
    };
    std::vector<Chunk> chunks;
    size_t off = 0;
    for (int i = 0; i < n_parts; ++i)
        for (size_t o = 0; o < parts[i].bytes; o += H2D_CHUNK) {
            const size_t n = std::min(H2D_CHUNK, parts[i].bytes - o);
            chunks.push_back(Chunk{static_cast<char*>(parts[i].dst) + o, static_cast<const char*>(parts[i].src) + o,
                                   static_cast<char*>(stage.p) + off, n});
            off += n;
        }
    std::atomic<int> failed{0};
    const int ordinal = d.ordinal;
    std::function<void(int)> job = [&](int k) {
        const Chunk& c = chunks[(size_t)k];
        std::memcpy(c.pin, c.src, c.bytes);
        if (cudaSetDevice(ordinal) != cudaSuccess ||
            cudaMemcpyAsync(c.dst, c.pin, c.bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess)
            failed.store(1);
    };
    CopyPool* pool = ensure_pool(ctx);
    pool->start(&job, (int)chunks.size());
    pool->wait();
    if (failed.load()) {
        const cudaError_t err = cudaGetLastError();
        return err != cudaSuccess ? err : cudaErrorUnknown;
    }
    return cudaSuccess;
}

// Upload (host inputs) or fan out (device inputs on the first device) the mesh of a call to every device that takes
// part: the first device gets it from the caller, the others pull it from the first device over NVLink when peer
// access exists (one PCIe upload instead of n), else from the host again.
// in_dev[i] receives the device-local pointers.
struct MeshInputs {
    const float* verts;
    const uint32_t* tris;
};
m2s_status stage_mesh_inputs(m2s_ctx* ctx, int nd, bool host_inputs, const float* verts, uint64_t nv,
                             const uint32_t* tris, uint64_t nt, std::vector<MeshInputs>& in_dev) {
    in_dev.assign(nd, MeshInputs{nullptr, nullptr});
    if (nt == 0) return M2S_OK;
    Device& d0 = ctx->dev[0];
    CU(ctx, cudaSetDevice(d0.ordinal));
    cudaEventRecord(d0.ev[0], d0.stream);
    if (host_inputs) {
        CU(ctx, d0.verts.ensure(nv * 12));
        CU(ctx, d0.tris.ensure(nt * 12));
        const H2DPart parts[2] = {{d0.verts.p, verts, nv * 12}, {d0.tris.p, tris, nt * 12}};
        CU(ctx, upload_inputs(ctx, d0, d0.in_mesh, parts, 2, d0.stream));
        in_dev[0] = MeshInputs{d0.verts.as<float>(), d0.tris.as<uint32_t>()};
    } else {
        in_dev[0] = MeshInputs{verts, tris};
    }
    cudaEventRecord(d0.ev[1], d0.stream);
    if (nd > 1) CU(ctx, cudaEventRecord(d0.ev_inputs, d0.stream));
    for (int i = 1; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        cudaEventRecord(d.ev[0], d.stream);
        CU(ctx, d.verts.ensure(nv * 12));
        CU(ctx, d.tris.ensure(nt * 12));
        if (host_inputs && !d.peer_to_first) {
            CU(ctx, cudaMemcpyAsync(d.verts.p, verts, nv * 12, cudaMemcpyHostToDevice, d.stream));
            CU(ctx, cudaMemcpyAsync(d.tris.p, tris, nt * 12, cudaMemcpyHostToDevice, d.stream));
        } else {
            CU(ctx, cudaStreamWaitEvent(d.stream, d0.ev_inputs, 0));
            CU(ctx, cudaMemcpyPeerAsync(d.verts.p, d.ordinal, in_dev[0].verts, d0.ordinal, nv * 12, d.stream));
            CU(ctx, cudaMemcpyPeerAsync(d.tris.p, d.ordinal, in_dev[0].tris, d0.ordinal, nt * 12, d.stream));
        }
        cudaEventRecord(d.ev[1], d.stream);
        in_dev[i] = MeshInputs{d.verts.as<float>(), d.tris.as<uint32_t>()};
    }
    return M2S_OK;
}

// Builds (or pulls) the call's mesh on the first nd devices. meshes[i] receives the MeshDev to query.
m2s_status stage_meshes(m2s_ctx* ctx, int nd, m2s_mesh* handle, const std::vector<MeshInputs>& in_dev, uint64_t nv,
                        uint64_t nt, bool want_records_event, std::vector<MeshDev*>& meshes) {
    meshes.assign(nd, nullptr);
    const bool broadcast = !handle && ctx->build_mode == M2S_BUILD_BROADCAST && nd > 1 && nt > 0;
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        if (handle) {
            CU(ctx, enqueue_mesh(d, &handle->dev[i], nullptr, nv, nullptr, nt, nullptr, &meshes[i]));
        } else if (broadcast && i > 0 && d.peer_to_first) {
            Device& d0 = ctx->dev[0];
            CU(ctx, cudaStreamWaitEvent(d.stream, d0.ev_built, 0));
            CU(ctx, enqueue_mesh_pull(d, d.scratch, d0, d0.scratch));
            cudaEventRecord(d.ev[2], d.stream);
            meshes[i] = &d.scratch;
        } else {
            CU(ctx, enqueue_mesh(d, nullptr, in_dev[i].verts, nv, in_dev[i].tris, nt,
                                 want_records_event ? d.ev_records : nullptr, &meshes[i]));
            if (broadcast && i == 0) CU(ctx, cudaEventRecord(d.ev_built, d.stream));
        }
    }
    return M2S_OK;
}

void mark_call_start(m2s_ctx* ctx, int nd, bool inputs_staged) {
    // stage_mesh_inputs records ev[0] / ev[1] when there is a mesh to move; calls on a handle / an empty mesh record
    // them here
    if (inputs_staged) return;
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        cudaSetDevice(d.ordinal);
        cudaEventRecord(d.ev[0], d.stream);
        cudaEventRecord(d.ev[1], d.stream);
    }
}

// A freshly allocated destination (the Vec<f32> a facade returns) takes its page faults inside the copy threads: with
// transparent huge pages in "madvise" mode, asking for them first turns 512 faults per 2 MiB into one. Advisory only.
void advise_huge_pages(void* p, size_t bytes) {
#ifdef MADV_HUGEPAGE
    const uintptr_t huge = 2u << 20;
    const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + huge - 1) & ~(huge - 1);
    const uintptr_t b = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(huge - 1);
    if (b > a) madvise(reinterpret_cast<void*>(a), b - a, MADV_HUGEPAGE);
#else
    (void)p;
    (void)bytes;
#endif
}

// ---- slab cuts over the devices of a context ---------------------------------------------------------------------
// cuts[i] .. cuts[i + 1] = the x range of device i
std::vector<uint64_t> slab_cuts(m2s_ctx* ctx, int nd, uint64_t xa, uint64_t xb, const GridArgs& ga, uint64_t nt,
                                int kind) {
    SlabBalance& b = ctx->balance;
    const bool same = b.valid && b.nd == nd && b.xa == xa && b.xb == xb && b.ny == ga.g.ny && b.nz == ga.g.nz &&
                      b.nt == nt && b.kind == kind;
    if (!(ctx->balance_slabs && same)) {
        b.cuts.resize((size_t)nd + 1);
        for (int i = 0; i <= nd; ++i) b.cuts[(size_t)i] = xa + (xb - xa) * (uint64_t)i / (uint64_t)nd;
    }
    b.xa = xa; b.xb = xb; b.ny = ga.g.ny; b.nz = ga.g.nz; b.nt = nt; b.nd = nd; b.kind = kind;
    b.valid = false;
    b.pending = nd > 1 && ctx->balance_slabs;
    return b.cuts;
}

// Called once the timings of the call that used ctx->balance.cuts are known: the cost per x-plane is taken as
// uniform inside a slab, the new cuts sit at equal shares of the cumulative cost, on whole brick planes.
void update_slab_balance(m2s_ctx* ctx) {
    SlabBalance& b = ctx->balance;
    if (!b.pending) return;
    b.pending = false;
    const int nd = b.nd;
    std::vector<double> t((size_t)nd);
    double total = 0.0;
    for (int i = 0; i < nd; ++i) {
        const m2s_timings& tm = ctx->dev[i].timings;
        t[(size_t)i] = std::max(0.0, (double)tm.dist_ms + (double)tm.sign_ms);
        total += t[(size_t)i];
    }
    const uint64_t align = GRID_BRICK_X, span = b.xb - b.xa;
    if (!(total > 0.0) || !std::isfinite(total) || span < align * (uint64_t)nd) return;  // keep the cuts, not "valid"
    const double tmax = *std::max_element(t.begin(), t.end()), tmin = *std::min_element(t.begin(), t.end());
    if (tmax <= 1.03 * tmin) {  // balanced within the noise of the measurement: keep the cuts
        b.valid = true;
        return;
    }
    // a cut moves by a quarter of the mean slab per call at most: a kernel held up for a call by a saturated host link
    // or a busy device cannot throw the split far off
    const double max_move = 0.25 * (double)span / nd;
    std::vector<uint64_t> cuts((size_t)nd + 1);
    cuts[0] = b.xa;
    double acc = 0.0;
    int r = 0;
    for (int k = 1; k < nd; ++k) {
        const double want = total * k / nd;
        while (r < nd - 1 && acc + t[(size_t)r] < want) acc += t[(size_t)r++];
        const double x0 = (double)b.cuts[(size_t)r], x1 = (double)b.cuts[(size_t)r + 1];
        double x = x0 + (x1 - x0) * (want - acc) / std::max(t[(size_t)r], 1e-12);
        x = 0.5 * (x + (double)b.cuts[(size_t)k]);  // damped: one odd measurement moves a cut half-way at most
        x = std::min(std::max(x, (double)b.cuts[(size_t)k] - max_move), (double)b.cuts[(size_t)k] + max_move);
        uint64_t xi = b.xa + (uint64_t)std::llround(std::max(0.0, x - (double)b.xa) / (double)align) * align;
        const uint64_t lo = cuts[(size_t)k - 1] + align, hi = b.xb - align * (uint64_t)(nd - k);
        cuts[(size_t)k] = std::min(std::max(xi, lo), hi);
    }
    cuts[(size_t)nd] = b.xb;
    b.cuts = cuts;
    b.valid = true;
}

constexpr size_t PIPELINE_MIN_BYTES = 4u << 20;  // smaller slabs: one staged copy is cheaper than the flag protocol

// ---- grid call, host destination -------------------------------------------------------------------------------
// cells x in [xa, xb) are split over the context's devices and written at out[(x - xa) * ny * nz + ...].
m2s_status grid_host(m2s_ctx* ctx, m2s_mesh* handle, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                     uint64_t nt, const GridArgs& ga, int sign_method, uint64_t xa, uint64_t xb, float* out) {
    DeviceGuard guard;
    const uint64_t span = xb - xa;
    const int nd = (int)std::min<uint64_t>((uint64_t)ctx->n_devices, span);
    const uint64_t plane = (uint64_t)ga.g.ny * ga.g.nz;
    const bool raycast = sign_method == M2S_SIGN_RAYCAST;

    std::vector<MeshInputs> in_dev;
    std::vector<MeshDev*> meshes;
    const bool staged_inputs = !handle && nt > 0;
    if (staged_inputs) {
        m2s_status s = stage_mesh_inputs(ctx, nd, true, verts_xyz, nv, tri_idx, nt, in_dev);
        if (s != M2S_OK) return s;
    } else {
        in_dev.assign(nd, MeshInputs{nullptr, nullptr});
    }
    mark_call_start(ctx, nd, staged_inputs);
    {
        m2s_status s = stage_meshes(ctx, nd, handle, in_dev, nv, nt, raycast, meshes);
        if (s != M2S_OK) return s;
    }

    struct Slab {
        GridParams g;
        float* host_dst;
        uint64_t cells;
        int path;
        float* stage;         // pinned ring (host address), PIPELINED
        uint32_t planes;      // completion flags
        uint32_t epoch;
        bool registered;
    };
    std::vector<Slab> slabs(nd);
    const std::vector<uint64_t> cuts = slab_cuts(ctx, nd, xa, xb, ga, nt, 1);
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        Slab& sl = slabs[i];
        sl.g = ga.g;
        sl.g.x0 = (uint32_t)cuts[(size_t)i];
        sl.g.x1 = (uint32_t)cuts[(size_t)i + 1];
        sl.cells = (uint64_t)(sl.g.x1 - sl.g.x0) * plane;
        sl.host_dst = out + (uint64_t)(sl.g.x0 - xa) * plane;
        sl.stage = nullptr;
        sl.planes = grid_brick_planes(sl.g);
        sl.epoch = 0;
        sl.registered = false;
        CU(ctx, cudaSetDevice(d.ordinal));
        const size_t bytes = sl.cells * 4;
        float* d_out = nullptr;
        Progress pr{nullptr, nullptr, 0u};
        float* alias = ctx->host_path == M2S_HOST_STAGED || ctx->host_path == M2S_HOST_PIPELINED
                           ? nullptr : pinned_device_alias(sl.host_dst, bytes);
        if (alias) {
            // a page-locked destination is written by the distance kernel itself: its stores travel over PCIe while
            // it computes, so there is no staging buffer and no copy
            sl.path = M2S_PATH_ZEROCOPY;
            d_out = alias;
        } else if (ctx->host_path == M2S_HOST_REGISTER) {
            CU(ctx, cudaHostRegister(sl.host_dst, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
            sl.registered = true;
            void* dp = nullptr;
            CU(ctx, cudaHostGetDevicePointer(&dp, sl.host_dst, 0));
            sl.path = M2S_PATH_REGISTERED;
            d_out = static_cast<float*>(dp);
        } else if (ctx->host_path == M2S_HOST_PIPELINED || (ctx->host_path == M2S_HOST_AUTO && bytes >= PIPELINE_MIN_BYTES)) {
            // pageable destination: the kernel writes into the library's pinned ring and publishes a flag per brick
            // plane; host threads copy finished planes into the caller's memory while the kernel is still running
            sl.path = M2S_PATH_PIPELINED;
            advise_huge_pages(sl.host_dst, bytes);
            CU(ctx, d.stage.ensure(bytes));
            CU(ctx, d.flags.ensure((size_t)sl.planes * 4));
            CU(ctx, d.progress.ensure((size_t)sl.planes * 4));
            if (++d.epoch == 0u) {  // wrapped: old flags could alias the new epoch
                std::memset(d.flags.p, 0, d.flags.cap);
                d.epoch = 1u;
            }
            sl.epoch = d.epoch;
            sl.stage = static_cast<float*>(d.stage.p);
            void *dp = nullptr, *fp = nullptr;
            CU(ctx, cudaHostGetDevicePointer(&dp, d.stage.p, 0));
            CU(ctx, cudaHostGetDevicePointer(&fp, d.flags.p, 0));
            CU(ctx, cudaMemsetAsync(d.progress.p, 0, (size_t)sl.planes * 4, d.stream));
            pr = Progress{d.progress.as<uint32_t>(), static_cast<volatile uint32_t*>(fp), sl.epoch};
            d_out = static_cast<float*>(dp);
        } else {
            sl.path = M2S_PATH_STAGED;
            CU(ctx, d.out.ensure(bytes));
            d_out = d.out.as<float>();
        }
        MeshDev& m = *meshes[i];
        const bool beside = !handle && raycast && nt > 0 && m.bvh.nt > 0 && &m == &d.scratch &&
                            !(ctx->build_mode == M2S_BUILD_BROADCAST && i > 0 && d.peer_to_first);
        const float4* rows_rec = beside ? d.rec_orig.as<float4>() : m.bvh.rec;
        CU(ctx, enqueue_grid_query(d, m, rows_rec, beside, sl.g, sign_method, d_out, true, pr.count ? &pr : nullptr));
        if (sl.path != M2S_PATH_STAGED) {
            CU(ctx, cudaMemcpyAsync(d.h_status, d.call_status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
            cudaEventRecord(d.ev[5], d.stream);
        }
    }
    // every device is busy by now: a copy into pageable memory blocks the calling thread, so those come last
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        Slab& sl = slabs[i];
        if (sl.path != M2S_PATH_STAGED) continue;
        CU(ctx, cudaSetDevice(d.ordinal));
        CU(ctx, cudaMemcpyAsync(sl.host_dst, d.out.p, sl.cells * 4, cudaMemcpyDeviceToHost, d.stream));
        CU(ctx, cudaMemcpyAsync(d.h_status, d.call_status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
        cudaEventRecord(d.ev[5], d.stream);
    }

    // pipelined slabs: host threads copy plane groups as their flags flip
    struct Chunk {
        int dev;
        uint32_t plane;
    };
    std::vector<Chunk> chunks;
    uint32_t max_planes = 0;
    for (int i = 0; i < nd; ++i)
        if (slabs[i].path == M2S_PATH_PIPELINED) max_planes = std::max(max_planes, slabs[i].planes);
    for (uint32_t p = 0; p < max_planes; ++p)
        for (int i = 0; i < nd; ++i)
            if (slabs[i].path == M2S_PATH_PIPELINED && p < slabs[i].planes) chunks.push_back(Chunk{i, p});
    std::vector<std::atomic<int>> stream_done(nd);
    for (int i = 0; i < nd; ++i) stream_done[i].store(0);
    std::atomic<int> copy_failed{0};
    // A fresh destination (the Vec<f32> a facade has just allocated) takes its first-touch page faults in these
    // threads. They are taken EARLY, while the GPU is still uploading and building: the first tasks populate the
    // destination 2 MiB at a time (MADV_POPULATE_WRITE leaves present pages and their contents alone, so it cannot
    // race with a copy), the copies that follow find their pages there.
    struct Block {
        char* p;
        size_t bytes;
    };
    std::vector<Block> blocks;
#ifdef MADV_POPULATE_WRITE
    static std::atomic<int> populate_ok{1};
    if (populate_ok.load(std::memory_order_relaxed))
        for (int i = 0; i < nd; ++i) {
            if (slabs[i].path != M2S_PATH_PIPELINED) continue;
            const uintptr_t a = reinterpret_cast<uintptr_t>(slabs[i].host_dst) & ~(uintptr_t)4095;
            const uintptr_t b = (reinterpret_cast<uintptr_t>(slabs[i].host_dst) + slabs[i].cells * 4 + 4095) & ~(uintptr_t)4095;
            for (uintptr_t o = a; o < b; o += (2u << 20))
                blocks.push_back(Block{reinterpret_cast<char*>(o), (size_t)std::min<uintptr_t>(2u << 20, b - o)});
        }
#endif
    const int n_blocks = (int)blocks.size();
    std::function<void(int)> job = [&](int task) {
        if (task < n_blocks) {
#ifdef MADV_POPULATE_WRITE
            if (madvise(blocks[(size_t)task].p, blocks[(size_t)task].bytes, MADV_POPULATE_WRITE) != 0 && errno == EINVAL)
                populate_ok.store(0, std::memory_order_relaxed);  // kernel older than 5.14: faults happen in the copies
#endif
            return;
        }
        const int k = task - n_blocks;
        const Chunk c = chunks[(size_t)k];
        const Slab& sl = slabs[c.dev];
        const volatile uint32_t* flag = static_cast<const volatile uint32_t*>(ctx->dev[c.dev].flags.p) + c.plane;
        unsigned spins = 0;
        while (*flag != sl.epoch) {
            if (stream_done[c.dev].load(std::memory_order_acquire)) {
                if (*flag == sl.epoch) break;
                copy_failed.store(1);  // the stream ended without publishing this plane (a failed launch)
                return;
            }
            if (++spins > 64) std::this_thread::yield();
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        const uint64_t p0 = (uint64_t)c.plane * GRID_BRICK_X;
        const uint64_t p1 = std::min<uint64_t>(p0 + GRID_BRICK_X, sl.g.x1 - sl.g.x0);
        std::memcpy(sl.host_dst + p0 * plane, sl.stage + p0 * plane, (p1 - p0) * plane * 4);
    };
    const bool pipelined = !chunks.empty();
    if (pipelined) {
        ensure_pool(ctx)->start(&job, n_blocks + (int)chunks.size());
    }
    m2s_status result = M2S_OK;
    cudaError_t sync_error = cudaSuccess;
    std::vector<std::chrono::steady_clock::time_point> t_done(nd);
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        cudaSetDevice(d.ordinal);
        const cudaError_t e = cudaStreamSynchronize(d.stream);
        t_done[i] = std::chrono::steady_clock::now();
        stream_done[i].store(1, std::memory_order_release);
        if (e != cudaSuccess && sync_error == cudaSuccess) sync_error = e;
    }
    if (pipelined) ctx->pool->wait();
    const auto t_end = std::chrono::steady_clock::now();
    for (int i = 0; i < nd; ++i)
        if (slabs[i].registered) cudaHostUnregister(slabs[i].host_dst);
    if (sync_error != cudaSuccess) return cuda_fail(ctx, sync_error, "cudaStreamSynchronize");
    if (copy_failed.load()) return fail(ctx, M2S_ECUDA, "distance kernel ended without publishing every plane");
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        if (result == M2S_OK) result = status_to_code(ctx, *d.h_status);
        cudaSetDevice(d.ordinal);
        collect_timings(d, slabs[i].path);
        if (slabs[i].path == M2S_PATH_PIPELINED) {
            // what the caller waits for after the kernel: the copies of the last planes (host clock)
            const float tail = std::chrono::duration<float, std::milli>(t_end - t_done[i]).count();
            d.timings.d2h_ms += tail;
            d.timings.total_ms += tail;
        }
    }
    update_slab_balance(ctx);
    return result;
}

// ---- grid call, device destination on the first device ------------------------------------------------------------
m2s_status grid_device(m2s_ctx* ctx, m2s_mesh* handle, const float* d_verts, uint64_t nv, const uint32_t* d_tris,
                       uint64_t nt, const GridArgs& ga, int sign_method, uint64_t xa, uint64_t xb, float* d_out) {
    DeviceGuard guard;
    const uint64_t span = xb - xa;
    const int nd = (int)std::min<uint64_t>((uint64_t)ctx->n_devices, span);
    const uint64_t plane = (uint64_t)ga.g.ny * ga.g.nz;
    const bool raycast = sign_method == M2S_SIGN_RAYCAST;
    std::vector<MeshInputs> in_dev;
    std::vector<MeshDev*> meshes;
    const bool staged_inputs = !handle && nt > 0;
    if (staged_inputs) {
        m2s_status s = stage_mesh_inputs(ctx, nd, false, d_verts, nv, d_tris, nt, in_dev);
        if (s != M2S_OK) return s;
    } else {
        in_dev.assign(nd, MeshInputs{nullptr, nullptr});
    }
    mark_call_start(ctx, nd, staged_inputs);
    {
        m2s_status s = stage_meshes(ctx, nd, handle, in_dev, nv, nt, raycast, meshes);
        if (s != M2S_OK) return s;
    }
    const std::vector<uint64_t> cuts = slab_cuts(ctx, nd, xa, xb, ga, nt, 0);
    Device& d0 = ctx->dev[0];
    if (nd > 1) {
        // the destination may still be in use by earlier work on the first device's stream
        CU(ctx, cudaSetDevice(d0.ordinal));
        CU(ctx, cudaEventRecord(d0.ev_inputs, d0.stream));
    }
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        GridParams g = ga.g;
        g.x0 = (uint32_t)cuts[(size_t)i];
        g.x1 = (uint32_t)cuts[(size_t)i + 1];
        const uint64_t cells = (uint64_t)(g.x1 - g.x0) * plane;
        float* dst = d_out + (uint64_t)(g.x0 - xa) * plane;
        // peers store their slab straight into the first device's buffer (peer-mapped pointer, NVLink): the
        // reassembly of the flat grid is the kernel's own epilogue
        float* target = (i == 0 || d.peer_to_first) ? dst : nullptr;
        if (!target) {
            CU(ctx, d.out.ensure(cells * 4));
            target = d.out.as<float>();
        }
        if (i > 0) CU(ctx, cudaStreamWaitEvent(d.stream, d0.ev_inputs, 0));
        MeshDev& m = *meshes[i];
        const bool beside = !handle && raycast && nt > 0 && &m == &d.scratch &&
                            !(ctx->build_mode == M2S_BUILD_BROADCAST && i > 0 && d.peer_to_first);
        const float4* rows_rec = beside ? d.rec_orig.as<float4>() : m.bvh.rec;
        CU(ctx, enqueue_grid_query(d, m, rows_rec, beside, g, sign_method, target, false, nullptr));
        if (target != dst) CU(ctx, cudaMemcpyPeerAsync(dst, d0.ordinal, target, d.ordinal, cells * 4, d.stream));
        cudaEventRecord(d.ev[5], d.stream);
        d.timings.host_path = M2S_PATH_DEVICE;
        if (i > 0) CU(ctx, cudaEventRecord(d.ev_done, d.stream));
    }
    if (nd > 1) {
        // stream order on the first device now covers the whole grid
        CU(ctx, cudaSetDevice(d0.ordinal));
        for (int i = 1; i < nd; ++i) CU(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[i].ev_done, 0));
    }
    return M2S_OK;
}

// ---- point calls ------------------------------------------------------------------------------------------------
m2s_status points_call(m2s_ctx* ctx, m2s_mesh* handle, bool host_io, const float* verts, uint64_t nv,
                       const uint32_t* tris, uint64_t nt, const float* queries, uint64_t nq, const PointPlan& plan,
                       float* out) {
    DeviceGuard guard;
    const int nd = (int)std::min<uint64_t>((uint64_t)ctx->n_devices, nq);
    std::vector<MeshInputs> in_dev;
    std::vector<MeshDev*> meshes;
    const bool staged_inputs = !handle && nt > 0;
    if (staged_inputs) {
        m2s_status s = stage_mesh_inputs(ctx, nd, host_io, verts, nv, tris, nt, in_dev);
        if (s != M2S_OK) return s;
    } else {
        in_dev.assign(nd, MeshInputs{nullptr, nullptr});
    }
    mark_call_start(ctx, nd, staged_inputs);
    {
        m2s_status s = stage_meshes(ctx, nd, handle, in_dev, nv, nt, false, meshes);
        if (s != M2S_OK) return s;
    }
    Device& d0 = ctx->dev[0];
    if (!host_io && nd > 1) {
        CU(ctx, cudaSetDevice(d0.ordinal));
        CU(ctx, cudaEventRecord(d0.ev_inputs, d0.stream));
    }
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        const uint64_t q0 = nq * i / nd, q1 = nq * (i + 1) / nd, n = q1 - q0;
        CU(ctx, cudaSetDevice(d.ordinal));
        const float* dq;
        float* target;
        if (host_io) {
            CU(ctx, d.queries.ensure(n * 12));
            CU(ctx, d.out.ensure(n * 4));
            // the queries travel on the side stream while the main stream builds the tree
            const H2DPart part{d.queries.p, queries + 3 * q0, n * 12};
            CU(ctx, upload_inputs(ctx, d, d.in_queries, &part, 1, d.aux_stream));
            CU(ctx, cudaEventRecord(d.ev_rows, d.aux_stream));
            CU(ctx, cudaStreamWaitEvent(d.stream, d.ev_rows, 0));
            dq = d.queries.as<float>();
            target = d.out.as<float>();
        } else if (i == 0) {
            dq = queries;
            target = out;
        } else {
            CU(ctx, cudaStreamWaitEvent(d.stream, d0.ev_inputs, 0));
            CU(ctx, d.queries.ensure(n * 12));
            CU(ctx, cudaMemcpyPeerAsync(d.queries.p, d.ordinal, queries + 3 * q0, d0.ordinal, n * 12, d.stream));
            dq = d.queries.as<float>();
            if (d.peer_to_first) {
                target = out + q0;  // results scattered straight into the first device's buffer
            } else {
                CU(ctx, d.out.ensure(n * 4));
                target = d.out.as<float>();
            }
        }
        CU(ctx, enqueue_points_query(d, *meshes[i], dq, n, plan, target, host_io));
        if (!host_io && i > 0) {
            if (target != out + q0)
                CU(ctx, cudaMemcpyPeerAsync(out + q0, d0.ordinal, target, d.ordinal, n * 4, d.stream));
            CU(ctx, cudaEventRecord(d.ev_done, d.stream));
        }
        if (!host_io) {
            cudaEventRecord(d.ev[5], d.stream);
            d.timings.host_path = M2S_PATH_DEVICE;
        }
    }
    if (!host_io) {
        if (nd > 1) {
            CU(ctx, cudaSetDevice(d0.ordinal));
            for (int i = 1; i < nd; ++i) CU(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[i].ev_done, 0));
        }
        return M2S_OK;
    }
    // results are small (4 B per query): staged copies, issued once every device is busy
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        const uint64_t q0 = nq * i / nd, q1 = nq * (i + 1) / nd, n = q1 - q0;
        CU(ctx, cudaSetDevice(d.ordinal));
        CU(ctx, cudaMemcpyAsync(out + q0, d.out.p, n * 4, cudaMemcpyDeviceToHost, d.stream));
        CU(ctx, cudaMemcpyAsync(d.h_status, d.call_status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
        cudaEventRecord(d.ev[5], d.stream);
    }
    m2s_status result = M2S_OK;
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        CU(ctx, cudaStreamSynchronize(d.stream));
        if (result == M2S_OK) result = status_to_code(ctx, *d.h_status);
        collect_timings(d, M2S_PATH_STAGED);
    }
    return result;
}

void release_device(Device& d) {
    cudaSetDevice(d.ordinal);
    if (d.stream || !d.own_stream) cudaStreamSynchronize(d.stream);
    DevBuf* bufs[] = {&d.verts, &d.tris, &d.rec_orig, &d.tobb, &d.tri_lo, &d.tri_hi, &d.keys_in, &d.keys_out,
                      &d.vals_in, &d.vals_out, &d.sort_tmp, &d.leaf_parent, &d.node_parent, &d.node_flag, &d.slot_list, &d.slot_count,
                      &d.call_status, &d.rows[0], &d.rows[1], &d.rows[2], &d.big_list, &d.big_count, &d.stats,
                      &d.tile_slot, &d.progress, &d.queries, &d.q_sorted, &d.q_perm, &d.q_keys_in, &d.q_keys_out,
                      &d.q_vals_in, &d.out, &d.post_keys, &d.post_idx, &d.post_mm, &d.post_in, &d.post_pts,
                      &d.post_out};
    for (DevBuf* b : bufs) b->release();
    d.scratch.release();
    d.stage.release();
    d.flags.release();
    d.in_mesh.release();
    d.in_queries.release();
    if (d.h_status) cudaFreeHost(d.h_status);
    for (int k = 0; k < 8; ++k)
        if (d.ev[k]) cudaEventDestroy(d.ev[k]);
    cudaEvent_t evs[] = {d.ev_records, d.ev_rows, d.ev_inputs, d.ev_done, d.ev_built};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (d.aux_stream) cudaStreamDestroy(d.aux_stream);
    if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
}

m2s_status create_common(const int* devices, int n, void* stream, bool use_stream, m2s_ctx** out) {
    if (!out) return M2S_EINVAL;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
        cudaGetLastError();
        return M2S_ENODEV;
    }
    DeviceGuard guard;
    std::vector<int> ids;
    if (!devices || n <= 0) ids.push_back(0);
    else ids.assign(devices, devices + n);
    // An ordinal may be listed more than once: every entry is a "device" of the context with its own streams, arenas
    // and slab (several slabs in flight on one GPU; also how the multi-device paths are exercised on a one-GPU box).
    if (ids.size() > 64) return M2S_EINVAL;
    for (size_t i = 0; i < ids.size(); ++i)
        if (ids[i] < 0 || ids[i] >= visible) return M2S_ENODEV;
    m2s_ctx* ctx = new (std::nothrow) m2s_ctx();
    if (!ctx) return M2S_EINVAL;
    ctx->n_devices = (int)ids.size();
    {
        // host copy threads of the pipelined paths: a fresh destination (the Vec<f32> a facade returns) takes its page
        // faults inside them. Measured on this pool's hosts (scripts/host_fault_bench.cpp, 64 MiB, huge pages
        // advised): 4 threads fill 14.9 GB/s, 8 threads 23.8 GB/s; the C3 kernel produces 13.8 GB/s. Half of the
        // host's threads divided by the visible GPUs (one process per GPU shares the host), 4..8.
        // A context that drives several GPUs itself gets that share for each of them (up to 16 threads).
        const int hw = (int)std::thread::hardware_concurrency();
        const int share = std::max(1, hw / (2 * std::max(1, visible)));
        ctx->copy_threads = std::min(ctx->n_devices > 1 ? 16 : 8, std::max(4, share * ctx->n_devices));
    }
    ctx->dev = new (std::nothrow) Device[ids.size()];
    if (!ctx->dev) { delete ctx; return M2S_EINVAL; }
    for (size_t i = 0; i < ids.size(); ++i) {
        Device& d = ctx->dev[i];
        d.ordinal = ids[i];
        cudaError_t e = cudaSetDevice(d.ordinal);
        if (e == cudaSuccess) {
            if (use_stream) {
                d.stream = (cudaStream_t)stream;
                d.own_stream = false;
            } else {
                e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
                d.own_stream = true;
            }
        }
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.ordinal);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&d.h_status, sizeof(BuildStatus));
        for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreate(&d.ev[k]);
        if (e == cudaSuccess) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
            e = cudaStreamCreateWithPriority(&d.aux_stream, cudaStreamNonBlocking, hi);
        }
        cudaEvent_t* evs[] = {&d.ev_records, &d.ev_rows, &d.ev_inputs, &d.ev_done, &d.ev_built};
        for (cudaEvent_t* ev : evs)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            cudaGetLastError();
            m2s_destroy(ctx);
            return M2S_ECUDA;
        }
        std::memset(d.h_status, 0, sizeof(BuildStatus));
#ifdef M2S_STATS_BUILD
        if (const char* c = std::getenv("M2S_STATS")) d.want_stats = std::atoi(c) != 0;
        if (d.want_stats) {
            d.stats.ensure(64);
            cudaMemset(d.stats.p, 0, 64);
        }
#endif
    }
    // peer access between the first device and every other one: slabs are stored straight into the first
    // device's memory, the mesh is pulled from it
    for (size_t i = 1; i < ids.size(); ++i) {
        if (ids[i] == ids[0]) {  // the same GPU: its memory is simply its own
            ctx->dev[i].peer_to_first = true;
            continue;
        }
        int to0 = 0, from0 = 0;
        cudaDeviceCanAccessPeer(&to0, ids[i], ids[0]);
        cudaDeviceCanAccessPeer(&from0, ids[0], ids[i]);
        if (!to0 || !from0) continue;
        cudaSetDevice(ids[i]);
        cudaError_t e1 = cudaDeviceEnablePeerAccess(ids[0], 0);
        cudaSetDevice(ids[0]);
        cudaError_t e2 = cudaDeviceEnablePeerAccess(ids[i], 0);
        cudaGetLastError();
        const bool ok1 = e1 == cudaSuccess || e1 == cudaErrorPeerAccessAlreadyEnabled;
        const bool ok2 = e2 == cudaSuccess || e2 == cudaErrorPeerAccessAlreadyEnabled;
        ctx->dev[i].peer_to_first = ok1 && ok2;
    }
    *out = ctx;
    return M2S_OK;
}

m2s_status mesh_create_common(m2s_ctx* ctx, bool host_inputs, const float* verts, uint64_t nv, const uint32_t* tris,
                              uint64_t nt, m2s_mesh** out) {
    if (!ctx || !out) return M2S_EINVAL;
    *out = nullptr;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    m2s_status s = check_mesh(ctx, verts, nv, tris, nt);
    if (s != M2S_OK) return s;
    DeviceGuard guard;
    m2s_mesh* mesh = new (std::nothrow) m2s_mesh();
    if (!mesh) return fail(ctx, M2S_EINVAL, "out of memory");
    mesh->owner = ctx;
    mesh->nv = nv;
    mesh->nt = nt;
    mesh->dev = new (std::nothrow) MeshDev[ctx->n_devices];
    if (!mesh->dev) { delete mesh; return fail(ctx, M2S_EINVAL, "out of memory"); }
    const int nd = ctx->n_devices;
    std::vector<MeshInputs> in_dev;
    s = stage_mesh_inputs(ctx, nd, host_inputs, verts, nv, tris, nt, in_dev);
    const bool broadcast = ctx->build_mode == M2S_BUILD_BROADCAST && nd > 1 && nt > 0;
    for (int i = 0; i < nd && s == M2S_OK; ++i) {
        Device& d = ctx->dev[i];
        cudaError_t e = cudaSetDevice(d.ordinal);
        if (e == cudaSuccess) {
            if (broadcast && i > 0 && d.peer_to_first) {
                e = cudaStreamWaitEvent(d.stream, ctx->dev[0].ev_built, 0);
                if (e == cudaSuccess) e = enqueue_mesh_pull(d, mesh->dev[i], ctx->dev[0], mesh->dev[0]);
            } else {
                e = launch_build(d, mesh->dev[i], in_dev[i].verts, nv, in_dev[i].tris, nt, nullptr);
                if (e == cudaSuccess && broadcast && i == 0) e = cudaEventRecord(d.ev_built, d.stream);
            }
        }
        if (e != cudaSuccess) s = cuda_fail(ctx, e, "m2s_mesh_create");
    }
    if (s == M2S_OK && host_inputs) {
        // the host arrays may be freed after return; also surfaces the mesh's data errors right here
        for (int i = 0; i < nd && s == M2S_OK; ++i) {
            Device& d = ctx->dev[i];
            cudaSetDevice(d.ordinal);
            cudaError_t e = cudaSuccess;
            if (nt > 0)
                e = cudaMemcpyAsync(d.h_status, mesh->dev[i].status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
            if (e != cudaSuccess) s = cuda_fail(ctx, e, "m2s_mesh_create");
            else if (nt > 0) s = status_to_code(ctx, *d.h_status);
        }
    }
    if (s != M2S_OK) {
        for (int i = 0; i < nd; ++i) {
            cudaSetDevice(ctx->dev[i].ordinal);
            cudaStreamSynchronize(ctx->dev[i].stream);
            mesh->dev[i].release();
        }
        cudaGetLastError();
        delete[] mesh->dev;
        delete mesh;
        return s;
    }
    *out = mesh;
    return M2S_OK;
}

m2s_status check_handle(m2s_ctx* ctx, const m2s_mesh* mesh) {
    if (!mesh || mesh->owner != ctx) return fail(ctx, M2S_EINVAL, "mesh handle does not belong to this context");
    return M2S_OK;
}

m2s_status check_slab(m2s_ctx* ctx, const GridArgs& ga, uint64_t x_begin, uint64_t x_end) {
    if (x_begin > x_end || x_end > ga.g.nx) return fail(ctx, M2S_EINVAL, "slab outside the grid");
    return M2S_OK;
}

}  // namespace

extern "C" {

int m2s_abi_version(void) { return M2S_ABI_VERSION; }

m2s_status m2s_create(const int* devices, int n_devices, m2s_ctx** out) {
    return create_common(devices, n_devices, nullptr, false, out);
}

m2s_status m2s_create_on_stream(int device, void* cuda_stream, m2s_ctx** out) {
    return create_common(&device, 1, cuda_stream, true, out);
}

void m2s_destroy(m2s_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    delete ctx->pool;
    for (int i = 0; i < ctx->n_devices; ++i) release_device(ctx->dev[i]);
    cudaGetLastError();
    delete[] ctx->dev;
    delete ctx;
}

const char* m2s_last_error(const m2s_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

m2s_status m2s_last_error_copy(m2s_ctx* ctx, char* buf, size_t n) {
    if (!ctx || !buf || n == 0) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::strncpy(buf, ctx->last_error.c_str(), n - 1);
    buf[n - 1] = '\0';
    return M2S_OK;
}

m2s_status m2s_last_timings_device(const m2s_ctx* ctx, int index, m2s_timings* out) {
    if (!ctx || !out || index < 0 || index >= ctx->n_devices) return M2S_EINVAL;
    *out = ctx->dev[index].timings;
    return M2S_OK;
}

m2s_status m2s_last_timings(const m2s_ctx* ctx, m2s_timings* out) { return m2s_last_timings_device(ctx, 0, out); }

m2s_status m2s_set_option(m2s_ctx* ctx, int option, int64_t value) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    switch (option) {
        case M2S_OPT_BUILD_MODE:
            if (value != M2S_BUILD_REPLICATED && value != M2S_BUILD_BROADCAST) return fail(ctx, M2S_EINVAL, "unknown build mode");
            ctx->build_mode = (int)value;
            return M2S_OK;
        case M2S_OPT_HOST_PATH:
            if (value < M2S_HOST_AUTO || value > M2S_HOST_REGISTER) return fail(ctx, M2S_EINVAL, "unknown host path");
            ctx->host_path = (int)value;
            return M2S_OK;
        case M2S_OPT_COPY_THREADS:
            if (value < 1 || value > 64) return fail(ctx, M2S_EINVAL, "copy threads out of range");
            ctx->copy_threads = (int)value;
            return M2S_OK;
        case M2S_OPT_BALANCE:
            if (value != 0 && value != 1) return fail(ctx, M2S_EINVAL, "balance: 0 or 1");
            ctx->balance_slabs = value != 0;
            ctx->balance = SlabBalance{};
            return M2S_OK;
        case M2S_OPT_RUN_LENGTH:
            if (value != 0 && value != 2 && value != 4 && value != 18 && value != 20)
                return fail(ctx, M2S_EINVAL, "run length: 0, 2, 4 (+16: the 4 x 4 x 2V lane layout)");
            for (int i = 0; i < ctx->n_devices; ++i) ctx->dev[i].run_v = (uint32_t)value;
            return M2S_OK;
        case M2S_OPT_RAY_BINS:
            if (value != 0 && value != 1) return fail(ctx, M2S_EINVAL, "ray bins: 0 or 1");
            for (int i = 0; i < ctx->n_devices; ++i) ctx->dev[i].no_ray_bins = value == 0;
            return M2S_OK;
        default:
            return fail(ctx, M2S_EINVAL, "unknown option");
    }
}

uint64_t m2s_launch_count(const m2s_ctx* ctx) {
    uint64_t n = 0;
    if (ctx)
        for (int i = 0; i < ctx->n_devices; ++i) n += ctx->dev[i].launches;
    return n;
}

int m2s_device_count(const m2s_ctx* ctx) { return ctx ? ctx->n_devices : 0; }

m2s_status m2s_debug_stats(m2s_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    out[0] = out[1] = out[2] = out[3] = 0;
#ifdef M2S_STATS_BUILD
    DeviceGuard guard;
    Device& d = ctx->dev[0];
    if (!d.stats.p) return M2S_OK;
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, cudaStreamSynchronize(d.stream));
    CU(ctx, cudaMemcpy(out, d.stats.p, 32, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemset(d.stats.p, 0, 64));
#endif
    return M2S_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------------

m2s_status m2s_generate_grid_sdf(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                                 uint64_t nt, const float first_cell[3], const float cell_size[3],
                                 const uint64_t cell_count[3], int sign_method, float* out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    GridArgs ga{};
    m2s_status s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga);
    if (s != M2S_OK) return s;
    if ((s = check_mesh(ctx, verts_xyz, nv, tri_idx, nt)) != M2S_OK) return s;
    if (ga.total == 0) return M2S_OK;  // empty Vec
    if (!out) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_host(ctx, nullptr, verts_xyz, nv, tri_idx, nt, ga, sign_method, 0, ga.g.nx, out);
}

m2s_status m2s_generate_grid_sdf_slab(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                                      uint64_t nt, const float first_cell[3], const float cell_size[3],
                                      const uint64_t cell_count[3], int sign_method, uint64_t x_begin,
                                      uint64_t x_end, float* out_slab) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    GridArgs ga{};
    m2s_status s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga);
    if (s != M2S_OK) return s;
    if ((s = check_mesh(ctx, verts_xyz, nv, tri_idx, nt)) != M2S_OK) return s;
    if ((s = check_slab(ctx, ga, x_begin, x_end)) != M2S_OK) return s;
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_host(ctx, nullptr, verts_xyz, nv, tri_idx, nt, ga, sign_method, x_begin, x_end, out_slab);
}

m2s_status m2s_generate_sdf(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx,
                            uint64_t nt, const float* queries_xyz, uint64_t nq, int accel_method, int sign_method,
                            float* out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    PointPlan plan{};
    if (!plan_points(accel_method, sign_method, &plan))
        return fail(ctx, M2S_EINVAL, "unknown acceleration / sign method");
    m2s_status s = check_mesh(ctx, verts_xyz, nv, tri_idx, nt);
    if (s != M2S_OK) return s;
    if ((s = check_queries(ctx, nq)) != M2S_OK) return s;
    if (nt == 0 && (accel_method == M2S_ACCEL_RTREE || accel_method == M2S_ACCEL_RTREE_BVH))
        return fail(ctx, M2S_EEMPTY, "Rtree / RtreeBvh on a mesh without triangles");
    if (nq == 0) return M2S_OK;
    if (!queries_xyz || !out) return fail(ctx, M2S_EINVAL, "null query / output pointer");
    return points_call(ctx, nullptr, true, verts_xyz, nv, tri_idx, nt, queries_xyz, nq, plan, out);
}

// ---- device-buffer entry points ----------------------------------------------------------------------

m2s_status m2s_generate_grid_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv,
                                        const uint32_t* d_tri_idx, uint64_t nt, const float first_cell[3],
                                        const float cell_size[3], const uint64_t cell_count[3], int sign_method,
                                        uint64_t x_begin, uint64_t x_end, float* d_out_slab) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    GridArgs ga{};
    m2s_status s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga);
    if (s != M2S_OK) return s;
    if ((s = check_mesh(ctx, d_verts_xyz, nv, d_tri_idx, nt)) != M2S_OK) return s;
    if ((s = check_slab(ctx, ga, x_begin, x_end)) != M2S_OK) return s;
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!d_out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_device(ctx, nullptr, d_verts_xyz, nv, d_tri_idx, nt, ga, sign_method, x_begin, x_end, d_out_slab);
}

m2s_status m2s_generate_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv, const uint32_t* d_tri_idx,
                                   uint64_t nt, const float* d_queries_xyz, uint64_t nq, int accel_method,
                                   int sign_method, float* d_out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    PointPlan plan{};
    if (!plan_points(accel_method, sign_method, &plan))
        return fail(ctx, M2S_EINVAL, "unknown acceleration / sign method");
    m2s_status s = check_mesh(ctx, d_verts_xyz, nv, d_tri_idx, nt);
    if (s != M2S_OK) return s;
    if ((s = check_queries(ctx, nq)) != M2S_OK) return s;
    if (nt == 0 && (accel_method == M2S_ACCEL_RTREE || accel_method == M2S_ACCEL_RTREE_BVH))
        return fail(ctx, M2S_EEMPTY, "Rtree / RtreeBvh on a mesh without triangles");
    if (nq == 0) return M2S_OK;
    if (!d_queries_xyz || !d_out) return fail(ctx, M2S_EINVAL, "null query / output pointer");
    return points_call(ctx, nullptr, false, d_verts_xyz, nv, d_tri_idx, nt, d_queries_xyz, nq, plan, d_out);
}

m2s_status m2s_synchronize(m2s_ctx* ctx) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    m2s_status result = M2S_OK;
    for (int i = 0; i < ctx->n_devices; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        if (d.call_status.p) {
            CU(ctx, cudaMemcpyAsync(d.h_status, d.call_status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
            CU(ctx, cudaStreamSynchronize(d.stream));
            const BuildStatus st = *d.h_status;
            if (st.bad_index || st.nonfinite || st.stack_overflow || st.nan_distance) {
                // clear the sticky error flags for the next batch of calls
                MeshDev none;
                CU(ctx, launch_call_status_init(d, none, true));
                CU(ctx, cudaStreamSynchronize(d.stream));
            }
            if (result == M2S_OK) result = status_to_code(ctx, st);
            collect_timings(d, d.timings.host_path);
        } else {
            CU(ctx, cudaStreamSynchronize(d.stream));
        }
    }
    update_slab_balance(ctx);  // the timings of an enqueued multi-device grid call are in now
    return result;
}

// ---- mesh handles ------------------------------------------------------------------------------------------

m2s_status m2s_mesh_create(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx, uint64_t nt,
                           m2s_mesh** out) {
    return mesh_create_common(ctx, true, verts_xyz, nv, tri_idx, nt, out);
}

m2s_status m2s_mesh_create_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv, const uint32_t* d_tri_idx,
                                  uint64_t nt, m2s_mesh** out) {
    return mesh_create_common(ctx, false, d_verts_xyz, nv, d_tri_idx, nt, out);
}

void m2s_mesh_destroy(m2s_mesh* mesh) {
    if (!mesh) return;
    m2s_ctx* ctx = mesh->owner;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    for (int i = 0; i < ctx->n_devices; ++i) {
        cudaSetDevice(ctx->dev[i].ordinal);
        cudaStreamSynchronize(ctx->dev[i].stream);
        mesh->dev[i].release();
    }
    cudaGetLastError();
    delete[] mesh->dev;
    delete mesh;
}

m2s_status m2s_mesh_grid_sdf(m2s_ctx* ctx, m2s_mesh* mesh, const float first_cell[3], const float cell_size[3],
                             const uint64_t cell_count[3], int sign_method, uint64_t x_begin, uint64_t x_end,
                             float* out_slab) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    m2s_status s = check_handle(ctx, mesh);
    if (s != M2S_OK) return s;
    GridArgs ga{};
    if ((s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga)) != M2S_OK) return s;
    if ((s = check_slab(ctx, ga, x_begin, x_end)) != M2S_OK) return s;
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_host(ctx, mesh, nullptr, mesh->nv, nullptr, mesh->nt, ga, sign_method, x_begin, x_end, out_slab);
}

m2s_status m2s_mesh_grid_sdf_device(m2s_ctx* ctx, m2s_mesh* mesh, const float first_cell[3], const float cell_size[3],
                                    const uint64_t cell_count[3], int sign_method, uint64_t x_begin, uint64_t x_end,
                                    float* d_out_slab) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    m2s_status s = check_handle(ctx, mesh);
    if (s != M2S_OK) return s;
    GridArgs ga{};
    if ((s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga)) != M2S_OK) return s;
    if ((s = check_slab(ctx, ga, x_begin, x_end)) != M2S_OK) return s;
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!d_out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_device(ctx, mesh, nullptr, mesh->nv, nullptr, mesh->nt, ga, sign_method, x_begin, x_end, d_out_slab);
}

static m2s_status mesh_points(m2s_ctx* ctx, m2s_mesh* mesh, bool host_io, const float* queries, uint64_t nq,
                              int accel_method, int sign_method, float* out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    m2s_status s = check_handle(ctx, mesh);
    if (s != M2S_OK) return s;
    PointPlan plan{};
    if (!plan_points(accel_method, sign_method, &plan))
        return fail(ctx, M2S_EINVAL, "unknown acceleration / sign method");
    if ((s = check_queries(ctx, nq)) != M2S_OK) return s;
    if (mesh->nt == 0 && (accel_method == M2S_ACCEL_RTREE || accel_method == M2S_ACCEL_RTREE_BVH))
        return fail(ctx, M2S_EEMPTY, "Rtree / RtreeBvh on a mesh without triangles");
    if (nq == 0) return M2S_OK;
    if (!queries || !out) return fail(ctx, M2S_EINVAL, "null query / output pointer");
    return points_call(ctx, mesh, host_io, nullptr, mesh->nv, nullptr, mesh->nt, queries, nq, plan, out);
}

m2s_status m2s_mesh_sdf(m2s_ctx* ctx, m2s_mesh* mesh, const float* queries_xyz, uint64_t nq, int accel_method,
                        int sign_method, float* out) {
    return mesh_points(ctx, mesh, true, queries_xyz, nq, accel_method, sign_method, out);
}

m2s_status m2s_mesh_sdf_device(m2s_ctx* ctx, m2s_mesh* mesh, const float* d_queries_xyz, uint64_t nq, int accel_method,
                               int sign_method, float* d_out) {
    return mesh_points(ctx, mesh, false, d_queries_xyz, nq, accel_method, sign_method, d_out);
}

// ---- memory helpers -------------------------------------------------------------------------------------------

m2s_status m2s_host_alloc(size_t bytes, void** out) {
    if (!out) return M2S_EINVAL;
    *out = nullptr;
    if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return M2S_ECUDA;
    }
    return M2S_OK;
}

void m2s_host_free(void* p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

m2s_status m2s_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return M2S_EINVAL;
    if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return M2S_ECUDA;
    }
    return M2S_OK;
}

m2s_status m2s_host_unregister(void* p) {
    if (!p) return M2S_EINVAL;
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return M2S_ECUDA;
    }
    return M2S_OK;
}

m2s_status m2s_device_alloc(m2s_ctx* ctx, size_t bytes, void** d_out) {
    if (!ctx || !d_out) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    *d_out = nullptr;
    CU(ctx, cudaSetDevice(ctx->dev[0].ordinal));
    CU(ctx, cudaMalloc(d_out, bytes ? bytes : 1));
    return M2S_OK;
}

m2s_status m2s_device_free(m2s_ctx* ctx, void* d_ptr) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    CU(ctx, cudaSetDevice(ctx->dev[0].ordinal));
    if (d_ptr) CU(ctx, cudaFree(d_ptr));
    return M2S_OK;
}

m2s_status m2s_ipc_export(m2s_ctx* ctx, void* d_ptr, unsigned char handle[64]) {
    if (!ctx || !d_ptr || !handle) return M2S_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    CU(ctx, cudaSetDevice(ctx->dev[0].ordinal));
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, 64);
    return M2S_OK;
}

m2s_status m2s_ipc_open(m2s_ctx* ctx, const unsigned char handle[64], void** d_out) {
    if (!ctx || !handle || !d_out) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    *d_out = nullptr;
    CU(ctx, cudaSetDevice(ctx->dev[0].ordinal));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(ctx, cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
    return M2S_OK;
}

m2s_status m2s_ipc_close(m2s_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard;
    CU(ctx, cudaSetDevice(ctx->dev[0].ordinal));
    CU(ctx, cudaStreamSynchronize(ctx->dev[0].stream));
    CU(ctx, cudaIpcCloseMemHandle(d_ptr));
    return M2S_OK;
}

// ---- post-passes on a finished grid -------------------------------------------------------------------

static m2s_status post_ctx(m2s_ctx* ctx) {
    if (ctx->n_devices != 1) return fail(ctx, M2S_EINVAL, "post-pass entry points need a single-device context");
    return M2S_OK;
}

m2s_status m2s_grid_order_device(m2s_ctx* ctx, const float* d_sdf, uint64_t n, uint32_t* d_order, float* d_minmax) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    if (post_ctx(ctx) != M2S_OK) return M2S_EINVAL;
    if (n >= (1ull << 31)) return fail(ctx, M2S_EINVAL, "more than 2^31-1 cells");
    if (n == 0) return M2S_OK;
    if (!d_sdf) return fail(ctx, M2S_EINVAL, "null sdf pointer");
    DeviceGuard guard;
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, launch_grid_order(d, d_sdf, n, d_order, d_minmax));
    return M2S_OK;
}

m2s_status m2s_grid_order(m2s_ctx* ctx, const float* sdf, uint64_t n, uint32_t* order, float minmax[2]) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    if (post_ctx(ctx) != M2S_OK) return M2S_EINVAL;
    if (n >= (1ull << 31)) return fail(ctx, M2S_EINVAL, "more than 2^31-1 cells");
    if (n == 0) return M2S_OK;
    if (!sdf) return fail(ctx, M2S_EINVAL, "null sdf pointer");
    DeviceGuard guard;
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, d.post_in.ensure(n * 4));
    CU(ctx, d.post_out.ensure(n * 4 + 8));
    CU(ctx, cudaMemcpyAsync(d.post_in.p, sdf, n * 4, cudaMemcpyHostToDevice, d.stream));
    uint32_t* d_order = order ? d.post_out.as<uint32_t>() : nullptr;
    float* d_mm = minmax ? reinterpret_cast<float*>(d.post_out.as<uint32_t>() + n) : nullptr;
    CU(ctx, launch_grid_order(d, d.post_in.as<float>(), n, d_order, d_mm));
    if (order) CU(ctx, cudaMemcpyAsync(order, d_order, n * 4, cudaMemcpyDeviceToHost, d.stream));
    if (minmax) CU(ctx, cudaMemcpyAsync(minmax, d_mm, 8, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return M2S_OK;
}

static m2s_status check_sample(m2s_ctx* ctx, const float first[3], const float size[3], const uint64_t count[3],
                               int mode, uint64_t np, GridArgs* ga) {
    if (post_ctx(ctx) != M2S_OK) return M2S_EINVAL;
    if (mode != M2S_SAMPLE_SNAP && mode != M2S_SAMPLE_TRILINEAR && mode != M2S_SAMPLE_TETRAHEDRAL)
        return fail(ctx, M2S_EINVAL, "unknown sample mode");
    if (np > 0xffffffffull) return fail(ctx, M2S_EINVAL, "more than 2^32 sample points");
    m2s_status s = check_grid(ctx, first, size, count, M2S_SIGN_RAYCAST, ga);
    if (s != M2S_OK) return s;
    if (ga->total == 0 && np > 0) return fail(ctx, M2S_EINVAL, "sampling an empty grid");
    return M2S_OK;
}

m2s_status m2s_sample_grid_sdf_device(m2s_ctx* ctx, const float* d_sdf, const float first_cell[3],
                                      const float cell_size[3], const uint64_t cell_count[3],
                                      const float* d_points_xyz, uint64_t np, int sample_mode, float iso,
                                      float* d_out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    GridArgs ga{};
    m2s_status s = check_sample(ctx, first_cell, cell_size, cell_count, sample_mode, np, &ga);
    if (s != M2S_OK) return s;
    if (np == 0) return M2S_OK;
    if (!d_sdf || !d_points_xyz || !d_out) return fail(ctx, M2S_EINVAL, "null sdf / point / output pointer");
    DeviceGuard guard;
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, launch_grid_sample(d, d_sdf, ga.g, d_points_xyz, np, sample_mode, iso, d_out));
    return M2S_OK;
}

m2s_status m2s_sample_grid_sdf(m2s_ctx* ctx, const float* sdf, const float first_cell[3], const float cell_size[3],
                               const uint64_t cell_count[3], const float* points_xyz, uint64_t np, int sample_mode,
                               float iso, float* out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    GridArgs ga{};
    m2s_status s = check_sample(ctx, first_cell, cell_size, cell_count, sample_mode, np, &ga);
    if (s != M2S_OK) return s;
    if (np == 0) return M2S_OK;
    if (!sdf || !points_xyz || !out) return fail(ctx, M2S_EINVAL, "null sdf / point / output pointer");
    DeviceGuard guard;
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, d.post_in.ensure(ga.total * 4));
    CU(ctx, d.post_pts.ensure(np * 12));
    CU(ctx, d.post_out.ensure(np * 4));
    CU(ctx, cudaMemcpyAsync(d.post_in.p, sdf, ga.total * 4, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, cudaMemcpyAsync(d.post_pts.p, points_xyz, np * 12, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, launch_grid_sample(d, d.post_in.as<float>(), ga.g, d.post_pts.as<float>(), np, sample_mode, iso,
                               d.post_out.as<float>()));
    CU(ctx, cudaMemcpyAsync(out, d.post_out.p, np * 4, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return M2S_OK;
}

// ---- host-side helpers ---------------------------------------------------------------------------------

// Topology::get_triangles, src/lib.rs:175-193.
uint64_t m2s_expand_topology(int topology, const void* indices, int index_bytes, uint64_t n_indices, uint64_t nv,
                             uint32_t* out) {
    const uint64_t n = indices ? n_indices : nv;
    auto at = [&](uint64_t i) -> uint32_t {
        if (!indices) return (uint32_t)i;
        return index_bytes == 2 ? (uint32_t) static_cast<const uint16_t*>(indices)[i]
                                : static_cast<const uint32_t*>(indices)[i];
    };
    if (indices && index_bytes != 2 && index_bytes != 4) return 0;
    uint64_t count = 0;
    if (topology == M2S_TRIANGLE_LIST) {
        count = n / 3;  // itertools::tuples drops the trailing partial tuple
        if (out)
            for (uint64_t t = 0; t < count; ++t)
                for (int k = 0; k < 3; ++k) out[3 * t + k] = at(3 * t + k);
    } else if (topology == M2S_TRIANGLE_STRIP) {
        count = n >= 3 ? n - 2 : 0;  // tuple_windows, no winding flip
        if (out)
            for (uint64_t t = 0; t < count; ++t)
                for (int k = 0; k < 3; ++k) out[3 * t + k] = at(t + k);
    }
    return count;
}

// Grid::from_bounding_box, src/grid.rs:59-74: cell_size = (max - min) / count; first = min + size * 0.5.
void m2s_grid_from_bounding_box(const float bbox_min[3], const float bbox_max[3], const uint64_t cell_count[3],
                                float first_cell[3], float cell_size[3]) {
    for (int i = 0; i < 3; ++i) {
        volatile float ext = bbox_max[i] - bbox_min[i];
        volatile float cs = ext / (float)cell_count[i];
        volatile float half = cs * 0.5f;
        cell_size[i] = cs;
        first_cell[i] = bbox_min[i] + half;
    }
}

}  // extern "C"
