// C ABI of libm2s.so (include/m2s.h): contexts, argument checks, host <-> device staging, slab
// sharding over the context's devices, deferred status. No CPU fallback: without a CUDA device
// m2s_create fails and nothing else can be called.
//
// Boundary replaced: the public free functions of the Rust crate, mesh_to_sdf/src/lib.rs:291-311
// (generate_sdf) and src/generate/grid.rs:265-378 (generate_grid_sdf); their infallible signatures
// panic where this ABI returns a status (lib.rs:257 "NaN distance", slice index panics, rtree.rs:117).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "m2s_internal.h"

using namespace m2s;

namespace {

m2s_status fail(m2s_ctx* ctx, m2s_status s, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return s;
}

m2s_status cuda_fail(m2s_ctx* ctx, cudaError_t e, const char* where) {
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
    out->g.x0 = out->g.xa = 0;
    out->g.x1 = out->g.xb = out->g.nx;
    return M2S_OK;
}

// Enqueue: records + LBVH + (row parities) + nearest kernel for one slab, all on d.stream.
// host-destined slabs of at least 4 Mi cells and 64 planes are computed as two half-slabs (see enqueue_grid)
static bool grid_is_split(const Device& d, const GridParams& g, uint64_t nt, const float* host_out) {
    const uint64_t slab_cells = (uint64_t)(g.x1 - g.x0) * g.ny * g.nz;
    // device-resident output: measured slower when split (8.39 vs 8.11 ms; the seed pass is bound by the
    // latency of its longest search, not by its size) — only the host path, which hides a copy, splits
    return host_out != nullptr && nt > 0 && d.split_halves && d.seed_levels == 1 && slab_cells >= (4u << 20) && (g.x1 - g.x0) >= 64u;
}

// host_out != nullptr: every finished half is copied to host_out on the device's copy stream while the
// next half's kernel runs.
cudaError_t enqueue_grid(m2s_ctx* ctx, Device& d, const float* d_verts, uint64_t nv, const uint32_t* d_tris,
                         uint64_t nt, const GridParams& g, int sign, float* d_out, bool clear_errors,
                         bool timed, float* host_out = nullptr) {
    cudaError_t e;
    if ((e = launch_status_reset(d, clear_errors)) != cudaSuccess) return e;
    const uint64_t slab_cells = (uint64_t)(g.x1 - g.x0) * g.ny * g.nz;
    if (nt == 0) {
        d.bvh = Bvh{};
        if (timed) { cudaEventRecord(d.ev[2], d.stream); cudaEventRecord(d.ev[3], d.stream); cudaEventRecord(d.ev[6], d.stream); }
        e = launch_fill(d, d_out, slab_cells, FLT_MAX);  // un-seeded cells stay f32::MAX (grid.rs:137-143)
        if (timed) cudaEventRecord(d.ev[4], d.stream);
        return e;
    }
    RowBits rb{};
    const bool raycast = sign == M2S_SIGN_RAYCAST;
    // The row parities only need the triangle records, so they run on the side stream beside the sort / hierarchy /
    // refit / box fitting (small, latency-bound launches that leave most SMs idle) and join before the distance kernel.
    if ((e = launch_build(d, d_verts, nv, d_tris, nt, ctx->leaf_size, raycast ? d.ev_records : nullptr)) != cudaSuccess)
        return e;
    if (timed) cudaEventRecord(d.ev[2], d.stream);
    if (raycast) {
        if ((e = cudaStreamWaitEvent(d.aux_stream, d.ev_records, 0)) != cudaSuccess) return e;
        if ((e = launch_grid_rows(d, g, &rb, d.aux_stream, d.rec_orig.as<float4>())) != cudaSuccess) return e;
        if ((e = cudaEventRecord(d.ev_rows, d.aux_stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(d.stream, d.ev_rows, 0)) != cudaSuccess) return e;
    }
    if (timed) cudaEventRecord(d.ev[3], d.stream);
    const int mode = raycast ? MODE_UNSIGNED : MODE_NORMAL;
    const RowBits* rbp = raycast ? &rb : nullptr;
    const uint32_t span = g.x1 - g.x0;
    const uint64_t plane = (uint64_t)g.ny * g.nz;
    d.last_split = false;
    if (!grid_is_split(d, g, nt, host_out)) {
        e = launch_grid_nearest(d, g, mode, rbp, d_out, timed ? d.ev[6] : nullptr);
        if (timed) cudaEventRecord(d.ev[4], d.stream);
        return e;
    }
    // Big slabs run as two half-slabs. The seed pass is latency bound (few, long, divergent searches) and
    // the distance kernel issue bound, so the second half's seed pass runs on a high-priority side stream
    // underneath the first half's distance kernel. With a host destination the first half's D2H overlaps
    // the second half's kernel as well; only the last copy is exposed, hence the uneven cut.
    d.last_split = true;
    const uint32_t cut = g.x0 + (uint32_t)(span * 0.72) / 4u * 4u;
    GridParams g1 = g, g2 = g;
    g1.x1 = g1.xb = cut;
    g2.x0 = g2.xa = cut;
    float* out2 = d_out + (uint64_t)(cut - g.x0) * plane;
    SeedLevel L1{}, L2{};
    const bool coarse_pass = !grid_uses_neighbour_seeds(d, mode);
    if (coarse_pass) {
        if ((e = cudaEventRecord(d.ev_fork, d.stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(d.aux_stream, d.ev_fork, 0)) != cudaSuccess) return e;
        if ((e = launch_grid_seeds(d, g2, &L2, 1, d.aux_stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(d.ev_join, d.aux_stream)) != cudaSuccess) return e;
        if ((e = launch_grid_seeds(d, g1, &L1, 0)) != cudaSuccess) return e;
    }
    if (timed) cudaEventRecord(d.ev[6], d.stream);
    if ((e = launch_grid_final(d, g1, L1, mode, rbp, d_out)) != cudaSuccess) return e;
    if (timed) cudaEventRecord(d.ev_half[0], d.stream);
    if (host_out && (e = cudaEventRecord(d.ev_chunk[0], d.stream)) != cudaSuccess) return e;
    if (coarse_pass && (e = cudaStreamWaitEvent(d.stream, d.ev_join, 0)) != cudaSuccess) return e;
    if (timed) cudaEventRecord(d.ev_half[1], d.stream);
    if ((e = launch_grid_final(d, g2, L2, mode, rbp, out2)) != cudaSuccess) return e;
    if (timed) cudaEventRecord(d.ev[4], d.stream);
    if (!host_out) return cudaSuccess;
    // both kernels are enqueued before the copies: a D2H into pageable memory blocks the calling thread
    if ((e = cudaEventRecord(d.ev_chunk[1], d.stream)) != cudaSuccess) return e;
    const uint64_t n1 = (uint64_t)(cut - g.x0) * plane, n2 = (uint64_t)(g.x1 - cut) * plane;
    if ((e = cudaStreamWaitEvent(d.copy_stream, d.ev_chunk[0], 0)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(host_out, d_out, n1 * 4, cudaMemcpyDeviceToHost, d.copy_stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(d.copy_stream, d.ev_chunk[1], 0)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(host_out + n1, out2, n2 * 4, cudaMemcpyDeviceToHost, d.copy_stream)) != cudaSuccess) return e;
    // the caller's stream continues only after the last chunk has landed
    if ((e = cudaEventRecord(d.ev_copied, d.copy_stream)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(d.stream, d.ev_copied, 0);
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

cudaError_t enqueue_points(m2s_ctx* ctx, Device& d, const float* d_verts, uint64_t nv, const uint32_t* d_tris,
                           uint64_t nt, const float* d_queries, uint64_t nq, const PointPlan& plan, float* d_out,
                           bool clear_errors, bool timed) {
    cudaError_t e;
    if ((e = launch_status_reset(d, clear_errors)) != cudaSuccess) return e;
    if (nt == 0) {
        d.bvh = Bvh{};
        if (timed) { cudaEventRecord(d.ev[2], d.stream); cudaEventRecord(d.ev[3], d.stream); cudaEventRecord(d.ev[6], d.stream); }
        e = launch_fill(d, d_out, nq, FLT_MAX);  // default.rs:54 / bvh.rs:83 fold from f32::MAX
        if (timed) cudaEventRecord(d.ev[4], d.stream);
        return e;
    }
    if ((e = launch_build(d, d_verts, nv, d_tris, nt, ctx->leaf_size)) != cudaSuccess) return e;
    if ((e = sort_queries(d, d_queries, nq)) != cudaSuccess) return e;
    if (timed) { cudaEventRecord(d.ev[2], d.stream); cudaEventRecord(d.ev[3], d.stream); }
    e = launch_points(d, nq, plan.mode, plan.sign_rule, d_out, timed ? d.ev[6] : nullptr);
    if (timed) cudaEventRecord(d.ev[4], d.stream);
    return e;
}

m2s_status status_to_code(m2s_ctx* ctx, const BuildStatus& st) {
    if (st.bad_index) return fail(ctx, M2S_EINDEX, "triangle index out of bounds");
    if (st.nonfinite) return fail(ctx, M2S_ENAN, "non-finite vertex or query coordinate");
    if (st.nan_distance) return fail(ctx, M2S_ENAN, "NaN distance");
    if (st.stack_overflow) return fail(ctx, M2S_ECUDA, "LBVH traversal stack overflow");
    return M2S_OK;
}

void collect_timings(m2s_ctx* ctx, Device& d) {
    m2s_timings t{};
    cudaEventElapsedTime(&t.h2d_ms, d.ev[0], d.ev[1]);
    cudaEventElapsedTime(&t.build_ms, d.ev[1], d.ev[2]);
    cudaEventElapsedTime(&t.sign_ms, d.ev[2], d.ev[3]);
    cudaEventElapsedTime(&t.seed_ms, d.ev[3], d.ev[6]);
    if (d.last_split) {  // two launches of the distance kernel; the second half's seed pass ran beside the first
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, d.ev[6], d.ev_half[0]);
        cudaEventElapsedTime(&b, d.ev_half[1], d.ev[4]);
        t.dist_ms = a + b;
    } else {
        cudaEventElapsedTime(&t.dist_ms, d.ev[6], d.ev[4]);
    }
    cudaEventElapsedTime(&t.d2h_ms, d.ev[4], d.ev[5]);
    cudaEventElapsedTime(&t.total_ms, d.ev[0], d.ev[5]);
    ctx->timings = t;
}

m2s_status create_common(const int* devices, int n, void* stream, bool use_stream, m2s_ctx** out) {
    if (!out) return M2S_EINVAL;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
        cudaGetLastError();
        return M2S_ENODEV;
    }
    std::vector<int> ids;
    if (!devices || n <= 0) ids.push_back(0);
    else ids.assign(devices, devices + n);
    for (size_t i = 0; i < ids.size(); ++i) {
        if (ids[i] < 0 || ids[i] >= visible) return M2S_ENODEV;
        for (size_t j = 0; j < i; ++j)
            if (ids[j] == ids[i]) return M2S_EINVAL;
    }
    m2s_ctx* ctx = new (std::nothrow) m2s_ctx();
    if (!ctx) return M2S_EINVAL;
    ctx->n_devices = (int)ids.size();
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
        for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&d.ev_chunk[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_copied, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
            e = cudaStreamCreateWithPriority(&d.aux_stream, cudaStreamNonBlocking, hi);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_records, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_rows, cudaEventDisableTiming);
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreate(&d.ev_half[k]);
        if (const char* c = std::getenv("M2S_SPLIT")) d.split_halves = std::atoi(c) != 0;
        if (const char* c = std::getenv("M2S_NSEED")) { d.neighbour_seeds = std::atoi(c) != 0; d.neighbour_and_coarse = std::atoi(c) == 2; }
        if (e != cudaSuccess) {
            cudaGetLastError();
            m2s_destroy(ctx);
            return M2S_ECUDA;
        }
        std::memset(d.h_status, 0, sizeof(BuildStatus));
        if (const char* e = std::getenv("M2S_PACKET")) d.packet = std::atoi(e) != 0;
        if (const char* e = std::getenv("M2S_ZEROCOPY")) d.zero_copy = std::atoi(e) != 0;
        if (const char* e = std::getenv("M2S_PAIR")) d.pair = std::max(0, std::min(8, std::atoi(e)));
        if (const char* e = std::getenv("M2S_SEED_PACKET")) d.seed_packet = std::atoi(e) != 0;
        if (const char* e = std::getenv("M2S_OBB_BIAS")) d.obb_bias = (float)std::atof(e);
        if (const char* e = std::getenv("M2S_STATS")) { d.want_stats = std::atoi(e) != 0; d.stats_mode = std::atoi(e); }
        if (const char* e = std::getenv("M2S_SEED_STRIDE")) d.seed_stride = (uint32_t)std::max(2, std::min(64, std::atoi(e)));
        if (const char* e = std::getenv("M2S_SEED_LEVELS")) d.seed_levels = std::max(0, std::min(2, std::atoi(e)));
    }
    if (const char* e = std::getenv("M2S_LEAF_SIZE")) ctx->leaf_size = (uint32_t)std::max(1, std::min(32, std::atoi(e)));
    *out = ctx;
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
    for (int i = 0; i < ctx->n_devices; ++i) {
        Device& d = ctx->dev[i];
        cudaSetDevice(d.ordinal);
        if (d.stream || !d.own_stream) cudaStreamSynchronize(d.stream);
        DevBuf* bufs[] = {&d.verts, &d.tris, &d.rec_orig, &d.rec_sorted, &d.tri_lo, &d.tri_hi, &d.keys_in,
                          &d.keys_out, &d.vals_in, &d.vals_out, &d.cub_tmp, &d.tri_id_sorted, &d.nodes, &d.nodes_il,
                          &d.leaf_parent, &d.node_parent, &d.node_flag, &d.status, &d.rows[0], &d.rows[1],
                          &d.rows[2], &d.big_list, &d.big_count, &d.queries, &d.q_sorted, &d.q_perm,
                          &d.q_keys_in, &d.q_keys_out, &d.q_vals_in, &d.out, &d.seeds[0], &d.seeds[1], &d.stats, &d.node_range, &d.tobb, &d.boxes, &d.tile_slot,
                          &d.post_keys, &d.post_idx, &d.post_mm, &d.post_in, &d.post_pts, &d.post_out};
        for (DevBuf* b : bufs) b->release();
        if (d.h_status) cudaFreeHost(d.h_status);
        for (int k = 0; k < 8; ++k)
            if (d.ev[k]) cudaEventDestroy(d.ev[k]);
        for (int k = 0; k < 8; ++k)
            if (d.ev_chunk[k]) cudaEventDestroy(d.ev_chunk[k]);
        if (d.ev_copied) cudaEventDestroy(d.ev_copied);
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.ev_join) cudaEventDestroy(d.ev_join);
        if (d.ev_records) cudaEventDestroy(d.ev_records);
        if (d.ev_rows) cudaEventDestroy(d.ev_rows);
        for (int k = 0; k < 2; ++k)
            if (d.ev_half[k]) cudaEventDestroy(d.ev_half[k]);
        if (d.aux_stream) cudaStreamDestroy(d.aux_stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
    }
    cudaGetLastError();
    delete[] ctx->dev;
    delete ctx;
}

const char* m2s_last_error(const m2s_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

m2s_status m2s_last_timings(const m2s_ctx* ctx, m2s_timings* out) {
    if (!ctx || !out) return M2S_EINVAL;
    *out = ctx->timings;
    return M2S_OK;
}

uint64_t m2s_launch_count(const m2s_ctx* ctx) {
    uint64_t n = 0;
    if (ctx)
        for (int i = 0; i < ctx->n_devices; ++i) n += ctx->dev[i].launches;
    return n;
}

int m2s_device_count(const m2s_ctx* ctx) { return ctx ? ctx->n_devices : 0; }

#ifdef M2S_STATS_BUILD
// development builds only (not part of the ABI): visits per subtree size class, see Bvh::stats
M2S_API m2s_status m2s_debug_hist(m2s_ctx* ctx, uint64_t out[32]) {
    if (!ctx || !out) return M2S_EINVAL;
    std::memset(out, 0, 32 * 8);
    Device& d = ctx->dev[0];
    if (!d.stats.p) return M2S_OK;
    cudaSetDevice(d.ordinal);
    cudaStreamSynchronize(d.stream);
    return cudaMemcpy(out, (char*)d.stats.p + 64, 32 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? M2S_OK : M2S_ECUDA;
}
#endif

m2s_status m2s_debug_stats(m2s_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Device& d = ctx->dev[0];
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!d.stats.p) return M2S_OK;
    CU(ctx, cudaSetDevice(d.ordinal));
    CU(ctx, cudaStreamSynchronize(d.stream));
    CU(ctx, cudaMemcpy(out, d.stats.p, 32, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(out + 3, (char*)d.stats.p + 32, 8, cudaMemcpyDeviceToHost));  // [3] = tiles that found no neighbour seed
    return M2S_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------------

// Device-side address of a page-locked, mapped host buffer (nullptr for pageable memory). Must be called with the
// target device current.
static float* pinned_device_alias(const void* host_ptr) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return static_cast<float*>(a.devicePointer);
}

// Shared by the whole-grid and the slab entry points: cells x in [xa, xb) are split over the context's
// devices and written at out[(x - xa) * ny * nz + ...].
static m2s_status grid_host(m2s_ctx* ctx, const float* verts_xyz, uint64_t nv, const uint32_t* tri_idx, uint64_t nt,
                            const GridArgs& ga, int sign_method, uint64_t xa, uint64_t xb, float* out) {
    const uint64_t span = xb - xa;
    const int nd = (int)std::min<uint64_t>((uint64_t)ctx->n_devices, span);
    const uint64_t plane = (uint64_t)ga.g.ny * ga.g.nz;
    // enqueue on every device, then wait for all of them
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        GridParams g = ga.g;
        g.x0 = g.xa = (uint32_t)(xa + span * i / nd);
        g.x1 = g.xb = (uint32_t)(xa + span * (i + 1) / nd);
        const uint64_t cells = (uint64_t)(g.x1 - g.x0) * plane;
        CU(ctx, cudaSetDevice(d.ordinal));
        const bool timed = i == 0;
        if (timed) cudaEventRecord(d.ev[0], d.stream);
        if (nt > 0) {
            CU(ctx, d.verts.ensure(nv * 12));
            CU(ctx, d.tris.ensure(nt * 12));
            CU(ctx, cudaMemcpyAsync(d.verts.p, verts_xyz, nv * 12, cudaMemcpyHostToDevice, d.stream));
            CU(ctx, cudaMemcpyAsync(d.tris.p, tri_idx, nt * 12, cudaMemcpyHostToDevice, d.stream));
        }
        float* host_dst = out + (uint64_t)(g.x0 - xa) * plane;
        // A pinned (page-locked, mapped) destination is written by the distance kernel itself: its stores go over
        // PCIe while it computes (64 MiB in ~5 ms is a quarter of the link), measured at no cost to the kernel
        // (4.92 ms either way on C3), so no staging buffer, no D2H copy and no half-slab split are needed.
        float* zero_copy = d.zero_copy ? pinned_device_alias(host_dst) : nullptr;
        if (!zero_copy) CU(ctx, d.out.ensure(cells * 4));
        if (timed) cudaEventRecord(d.ev[1], d.stream);
        if (zero_copy) {
            CU(ctx, enqueue_grid(ctx, d, d.verts.as<float>(), nv, d.tris.as<uint32_t>(), nt, g, sign_method, zero_copy,
                                 true, timed, nullptr));
        } else {
            CU(ctx, enqueue_grid(ctx, d, d.verts.as<float>(), nv, d.tris.as<uint32_t>(), nt, g, sign_method,
                                 d.out.as<float>(), true, timed, host_dst));
            if (!grid_is_split(d, g, nt, host_dst))
                CU(ctx, cudaMemcpyAsync(host_dst, d.out.p, cells * 4, cudaMemcpyDeviceToHost, d.stream));
        }
        CU(ctx, cudaMemcpyAsync(d.h_status, d.status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
        if (timed) cudaEventRecord(d.ev[5], d.stream);
    }
    m2s_status result = M2S_OK;
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        CU(ctx, cudaStreamSynchronize(d.stream));
        if (result == M2S_OK) result = status_to_code(ctx, *d.h_status);
    }
    collect_timings(ctx, ctx->dev[0]);
    return result;
}

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
    return grid_host(ctx, verts_xyz, nv, tri_idx, nt, ga, sign_method, 0, ga.g.nx, out);
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
    if (x_begin > x_end || x_end > ga.g.nx) return fail(ctx, M2S_EINVAL, "slab outside the grid");
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    return grid_host(ctx, verts_xyz, nv, tri_idx, nt, ga, sign_method, x_begin, x_end, out_slab);
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
    if (nq > 0xffffffffull) return fail(ctx, M2S_EINVAL, "more than 2^32 query points");
    if (nt == 0 && (accel_method == M2S_ACCEL_RTREE || accel_method == M2S_ACCEL_RTREE_BVH))
        return fail(ctx, M2S_EEMPTY, "Rtree / RtreeBvh on a mesh without triangles");
    if (nq == 0) return M2S_OK;
    if (!queries_xyz || !out) return fail(ctx, M2S_EINVAL, "null query / output pointer");

    const int nd = (int)std::min<uint64_t>((uint64_t)ctx->n_devices, nq);
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        const uint64_t q0 = nq * i / nd, q1 = nq * (i + 1) / nd, n = q1 - q0;
        CU(ctx, cudaSetDevice(d.ordinal));
        const bool timed = i == 0;
        if (timed) cudaEventRecord(d.ev[0], d.stream);
        if (nt > 0) {
            CU(ctx, d.verts.ensure(nv * 12));
            CU(ctx, d.tris.ensure(nt * 12));
            CU(ctx, cudaMemcpyAsync(d.verts.p, verts_xyz, nv * 12, cudaMemcpyHostToDevice, d.stream));
            CU(ctx, cudaMemcpyAsync(d.tris.p, tri_idx, nt * 12, cudaMemcpyHostToDevice, d.stream));
        }
        CU(ctx, d.queries.ensure(n * 12));
        CU(ctx, d.out.ensure(n * 4));
        CU(ctx, cudaMemcpyAsync(d.queries.p, queries_xyz + 3 * q0, n * 12, cudaMemcpyHostToDevice, d.stream));
        if (timed) cudaEventRecord(d.ev[1], d.stream);
        CU(ctx, enqueue_points(ctx, d, d.verts.as<float>(), nv, d.tris.as<uint32_t>(), nt, d.queries.as<float>(), n,
                               plan, d.out.as<float>(), true, timed));
        CU(ctx, cudaMemcpyAsync(out + q0, d.out.p, n * 4, cudaMemcpyDeviceToHost, d.stream));
        CU(ctx, cudaMemcpyAsync(d.h_status, d.status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
        if (timed) cudaEventRecord(d.ev[5], d.stream);
    }
    m2s_status result = M2S_OK;
    for (int i = 0; i < nd; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        CU(ctx, cudaStreamSynchronize(d.stream));
        if (result == M2S_OK) result = status_to_code(ctx, *d.h_status);
    }
    collect_timings(ctx, ctx->dev[0]);
    return result;
}

// ---- device-buffer entry points ----------------------------------------------------------------------

m2s_status m2s_generate_grid_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv,
                                        const uint32_t* d_tri_idx, uint64_t nt, const float first_cell[3],
                                        const float cell_size[3], const uint64_t cell_count[3], int sign_method,
                                        uint64_t x_begin, uint64_t x_end, float* d_out_slab) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    if (ctx->n_devices != 1) return fail(ctx, M2S_EINVAL, "device-buffer entry points need a single-device context");
    GridArgs ga{};
    m2s_status s = check_grid(ctx, first_cell, cell_size, cell_count, sign_method, &ga);
    if (s != M2S_OK) return s;
    if ((s = check_mesh(ctx, d_verts_xyz, nv, d_tri_idx, nt)) != M2S_OK) return s;
    if (x_begin > x_end || x_end > ga.g.nx) return fail(ctx, M2S_EINVAL, "slab outside the grid");
    if (x_begin == x_end || ga.total == 0) return M2S_OK;
    if (!d_out_slab) return fail(ctx, M2S_EINVAL, "null output pointer");
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    GridParams g = ga.g;
    g.x0 = g.xa = (uint32_t)x_begin;
    g.x1 = g.xb = (uint32_t)x_end;
    cudaEventRecord(d.ev[0], d.stream);
    cudaEventRecord(d.ev[1], d.stream);
    CU(ctx, enqueue_grid(ctx, d, d_verts_xyz, nv, d_tri_idx, nt, g, sign_method, d_out_slab, false, true));
    cudaEventRecord(d.ev[5], d.stream);
    return M2S_OK;
}

m2s_status m2s_generate_sdf_device(m2s_ctx* ctx, const float* d_verts_xyz, uint64_t nv, const uint32_t* d_tri_idx,
                                   uint64_t nt, const float* d_queries_xyz, uint64_t nq, int accel_method,
                                   int sign_method, float* d_out) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    if (ctx->n_devices != 1) return fail(ctx, M2S_EINVAL, "device-buffer entry points need a single-device context");
    PointPlan plan{};
    if (!plan_points(accel_method, sign_method, &plan))
        return fail(ctx, M2S_EINVAL, "unknown acceleration / sign method");
    m2s_status s = check_mesh(ctx, d_verts_xyz, nv, d_tri_idx, nt);
    if (s != M2S_OK) return s;
    if (nq > 0xffffffffull) return fail(ctx, M2S_EINVAL, "more than 2^32 query points");
    if (nt == 0 && (accel_method == M2S_ACCEL_RTREE || accel_method == M2S_ACCEL_RTREE_BVH))
        return fail(ctx, M2S_EEMPTY, "Rtree / RtreeBvh on a mesh without triangles");
    if (nq == 0) return M2S_OK;
    if (!d_queries_xyz || !d_out) return fail(ctx, M2S_EINVAL, "null query / output pointer");
    Device& d = ctx->dev[0];
    CU(ctx, cudaSetDevice(d.ordinal));
    cudaEventRecord(d.ev[0], d.stream);
    cudaEventRecord(d.ev[1], d.stream);
    CU(ctx, enqueue_points(ctx, d, d_verts_xyz, nv, d_tri_idx, nt, d_queries_xyz, nq, plan, d_out, false, true));
    cudaEventRecord(d.ev[5], d.stream);
    return M2S_OK;
}

m2s_status m2s_synchronize(m2s_ctx* ctx) {
    if (!ctx) return M2S_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    m2s_status result = M2S_OK;
    for (int i = 0; i < ctx->n_devices; ++i) {
        Device& d = ctx->dev[i];
        CU(ctx, cudaSetDevice(d.ordinal));
        if (d.status.p) {
            CU(ctx, cudaMemcpyAsync(d.h_status, d.status.p, sizeof(BuildStatus), cudaMemcpyDeviceToHost, d.stream));
            // clear the sticky error flags (first five ints) for the next batch of calls
            CU(ctx, cudaStreamSynchronize(d.stream));
            const BuildStatus st = *d.h_status;
            if (st.bad_index || st.nonfinite || st.stack_overflow || st.nan_distance) {
                CU(ctx, launch_status_reset(d, true));
                CU(ctx, cudaStreamSynchronize(d.stream));
            }
            if (result == M2S_OK) result = status_to_code(ctx, st);
            if (i == 0) collect_timings(ctx, d);
        } else {
            CU(ctx, cudaStreamSynchronize(d.stream));
        }
    }
    return result;
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
