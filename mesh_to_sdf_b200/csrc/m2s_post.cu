// Post-passes on a finished grid SDF that is still on the device ("next" rows 2 and 3): what the
// reference's only in-repo caller does with the Vec<f32> right after generate_grid_sdf returns.
//
//   launch_grid_order  - mesh_to_sdf_client/src/sdf.rs:62-68: cell indices stably sorted by f32::total_cmp of
//                        their distance (the voxel pass draws a prefix of that order), and sdf.rs:123: the iso
//                        limits `data.iter().copied().minmax()`.
//   launch_grid_sample - sdf_grid() of mesh_to_sdf_client/shaders/draw_raymarching.wgsl:118-200 with the clamped
//                        fetch :92-99 and the tetrahedral weights :585-650: snap / trilinear / tetrahedral sampling
//                        of the cell-centred grid at arbitrary points (the reference's own TODO, src/grid.rs:172).
// Both are HBM-bound streaming passes (4 B read + 4..8 B written per cell / 12 B + 32 B gathered per sample).
#include <algorithm>

#include "m2s_geom.cuh"
#include "m2s_internal.h"
#include "m2s_sort.cuh"

namespace m2s {
namespace {

// f32::total_cmp as an unsigned key (the sort computes the same key on the fly: sort_detail::F32TotalOrderSrc)
__device__ __forceinline__ uint32_t total_order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// itertools::minmax compares with PartialOrd (`<`), under which -0.0 == +0.0: "if several elements are equally
// minimum, the first is returned; if several are equally maximum, the last". The packed (key, index) words
// below reproduce that: zeros share one key, ties resolve by index.
__device__ __forceinline__ uint32_t partial_order_key(float v) { return total_order_key(v == 0.0f ? 0.0f : v); }

__global__ void __launch_bounds__(256)
k_order_limits(const float* __restrict__ sdf, uint32_t n, unsigned long long* __restrict__ mm) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long pk = ((unsigned long long)partial_order_key(__ldg(sdf + i)) << 32) | i;
        lo = pk < lo ? pk : lo;
        hi = pk > hi ? pk : hi;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 0, lo);
        atomicMax(mm + 1, hi);
    }
}

// The same limits, taken on the sort's own read of the distances (its histogram kernel): the partial-order key differs
// from the total-order key the sort sees only for -0.0, whose key 0x7fffffff becomes that of +0.0.
struct LimitsObserver {
    unsigned long long* mm;
    unsigned long long lo = ~0ull, hi = 0ull;
    __device__ __forceinline__ void see(uint32_t i, uint64_t total_key) {
        const uint32_t k = (uint32_t)total_key == 0x7fffffffu ? 0x80000000u : (uint32_t)total_key;
        const unsigned long long pk = ((unsigned long long)k << 32) | i;
        lo = pk < lo ? pk : lo;
        hi = pk > hi ? pk : hi;
    }
    __device__ __forceinline__ void finish() {
        for (int o = 16; o; o >>= 1) {
            const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = l2 < lo ? l2 : lo;
            hi = h2 > hi ? h2 : hi;
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(mm + 0, lo);
            atomicMax(mm + 1, hi);
        }
    }
};

__global__ void k_order_minmax(const float* __restrict__ sdf, const unsigned long long* __restrict__ mm,
                               float* __restrict__ out) {
    out[0] = sdf[(uint32_t)mm[0]];
    out[1] = sdf[(uint32_t)mm[1]];
}

// draw_raymarching.wgsl:92-99 (`- iso` applied by the callers below, in the shader's order)
__device__ __forceinline__ float fetch_clamped(const float* __restrict__ sdf, const GridParams& g, int x, int y, int z) {
    x = min(max(x, 0), (int)g.nx - 1);
    y = min(max(y, 0), (int)g.ny - 1);
    z = min(max(z, 0), (int)g.nz - 1);
    return __ldg(sdf + ((size_t)z + (size_t)y * g.nz + (size_t)x * g.nz * g.ny));
}

__device__ __forceinline__ float lerp_wgsl(float a, float b, float f) {  // a * (1 - f) + b * f, un-fused
    return fadd(fmul(a, fsub(1.0f, f)), fmul(b, f));
}

__global__ void __launch_bounds__(256)
k_grid_sample(const float* __restrict__ sdf, const GridParams g, const float* __restrict__ pts, uint32_t np, int mode,
              float iso, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const float p[3] = {__ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2)};
    const float first[3] = {g.fx, g.fy, g.fz}, size[3] = {g.sx, g.sy, g.sz};
    const uint32_t cnt[3] = {g.nx, g.ny, g.nz};
    // :121-123: outside [start, end] -> 100.0; end = Grid::get_last_cell = first + count * size (src/grid.rs:82-88)
    bool outside = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float end = fadd(first[a], fmul((float)cnt[a], size[a]));
        outside |= p[a] < first[a] || p[a] > end;
    }
    if (outside) {
        out[i] = 100.0f;
        return;
    }
    float dist = 100.0f;
    if (mode == M2S_SAMPLE_SNAP) {
        // :129-135: cell = floor((p - (start - size / 2)) / size)
        int c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            c[a] = (int)floorf(fdiv(fsub(p[a], fsub(first[a], fmul(size[a], 0.5f))), size[a]));
        dist = fsub(fetch_clamped(sdf, g, c[0], c[1], c[2]), iso);
    } else {
        // dual grid: the cell centres are the corners of the interpolation cells (:137-176)
        int c[3];
        float f[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float ci = fdiv(fsub(p[a], first[a]), size[a]);
            const float fl = floorf(ci);
            f[a] = fsub(ci, fl);  // WGSL fract(e) = e - floor(e)
            c[a] = (int)fl;
        }
        auto at = [&](int dx, int dy, int dz) { return fsub(fetch_clamped(sdf, g, c[0] + dx, c[1] + dy, c[2] + dz), iso); };
        if (mode == M2S_SAMPLE_TRILINEAR) {
            const float x00 = lerp_wgsl(at(0, 0, 0), at(1, 0, 0), f[0]);
            const float x01 = lerp_wgsl(at(0, 0, 1), at(1, 0, 1), f[0]);
            const float x10 = lerp_wgsl(at(0, 1, 0), at(1, 1, 0), f[0]);
            const float x11 = lerp_wgsl(at(0, 1, 1), at(1, 1, 1), f[0]);
            const float xy0 = lerp_wgsl(x00, x10, f[1]);
            const float xy1 = lerp_wgsl(x01, x11, f[1]);
            dist = lerp_wgsl(xy0, xy1, f[2]);
        } else {
            // compute_tetrahedral_barycenter (:585-650): six sequential tests, a later match overrides an earlier one
            const float r = f[0], gg = f[1], b = f[2];
            float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
            int v2[3] = {0, 0, 0}, v3[3] = {0, 0, 0};
            auto set = [&](float a0, float a1, float a2, float a3, int x2, int y2, int z2, int x3, int y3, int z3) {
                w0 = a0; w1 = a1; w2 = a2; w3 = a3;
                v2[0] = x2; v2[1] = y2; v2[2] = z2;
                v3[0] = x3; v3[1] = y3; v3[2] = z3;
            };
            if (gg >= b && b >= r) set(fsub(1.f, gg), fsub(gg, b), fsub(b, r), r, 0, 1, 0, 0, 1, 1);
            if (b > r && r > gg) set(fsub(1.f, b), fsub(b, r), fsub(r, gg), gg, 0, 0, 1, 1, 0, 1);
            if (b > gg && gg >= r) set(fsub(1.f, b), fsub(b, gg), fsub(gg, r), r, 0, 0, 1, 0, 1, 1);
            if (r >= gg && gg > b) set(fsub(1.f, r), fsub(r, gg), fsub(gg, b), b, 1, 0, 0, 1, 1, 0);
            if (gg > r && r >= b) set(fsub(1.f, gg), fsub(gg, r), fsub(r, b), b, 0, 1, 0, 1, 1, 0);
            if (r >= b && b >= gg) set(fsub(1.f, r), fsub(r, b), fsub(b, gg), gg, 1, 0, 0, 1, 0, 1);
            // dot(bary, samples) = b.x s.x + b.y s.y + b.z s.z + b.w s.w, left to right
            const float s0 = at(0, 0, 0), s1 = at(v2[0], v2[1], v2[2]), s2 = at(v3[0], v3[1], v3[2]), s3 = at(1, 1, 1);
            dist = fadd(fadd(fadd(fmul(w0, s0), fmul(w1, s1)), fmul(w2, s2)), fmul(w3, s3));
        }
    }
    out[i] = dist;
}

}  // namespace

#define CK(x)                               \
    do {                                    \
        cudaError_t e__ = (x);              \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// d_order (n x u32) and / or d_minmax (2 floats) may be null. n < 2^31.
cudaError_t launch_grid_order(Device& d, const float* d_sdf, uint64_t n, uint32_t* d_order, float* d_minmax) {
    cudaStream_t s = d.stream;
    if (n == 0) return cudaSuccess;
    const uint32_t n32 = (uint32_t)n;
    unsigned long long* mm = nullptr;
    if (d_minmax) {
        CK(d.post_mm.ensure(16));
        mm = d.post_mm.as<unsigned long long>();
        CK(cudaMemsetAsync(mm, 0xff, 8, s));
        CK(cudaMemsetAsync(mm + 1, 0x00, 8, s));
    }
    if (d_order) {
        // four 8-bit passes over the total-order keys, computed from the distances on the fly in pass 0; the payloads
        // (cell indices) of the last pass land in the caller's array and its keys are not written (m2s_sort.cuh).
        // The iso limits ride along on the histogram kernel's read of the distances.
        CK(d.post_keys.ensure(n * 4 * 2));
        CK(d.post_idx.ensure(n * 4));
        CK(d.sort_tmp.ensure(radix_sort_scratch_bytes(n)));
        uint32_t* const kbuf[2] = {d.post_keys.as<uint32_t>(), d.post_keys.as<uint32_t>() + n};
        uint32_t* const vbuf[2] = {d.post_idx.as<uint32_t>(), d_order};
        const sort_detail::F32TotalOrderSrc src{d_sdf};
        if (d_minmax) CK(radix_sort_pairs(s, src, kbuf, vbuf, n, 32, d.sort_tmp.p, false, &d.launches, LimitsObserver{mm}));
        else CK(radix_sort_pairs(s, src, kbuf, vbuf, n, 32, d.sort_tmp.p, false, &d.launches));
    } else if (d_minmax) {
        const unsigned nb = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)d.sm_count * 8);
        k_order_limits<<<nb, 256, 0, s>>>(d_sdf, n32, mm);
        d.launches++;
    }
    if (d_minmax) {
        k_order_minmax<<<1, 1, 0, s>>>(d_sdf, mm, d_minmax);
        d.launches++;
    }
    return cudaGetLastError();
}

cudaError_t launch_grid_sample(Device& d, const float* d_sdf, const GridParams& g, const float* d_points, uint64_t np,
                               int mode, float iso, float* d_out) {
    if (np == 0) return cudaSuccess;
    k_grid_sample<<<(unsigned)((np + 255) / 256), 256, 0, d.stream>>>(d_sdf, g, d_points, (uint32_t)np, mode, iso, d_out);
    d.launches++;
    return cudaGetLastError();
}

}  // namespace m2s
