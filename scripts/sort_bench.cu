// Development harness for mesh_to_sdf_b200/csrc/m2s_sort.cuh: checks the hand-written radix sort / scan against
// std::stable_sort / a host prefix sum and times them beside cub::DeviceRadixSort (the library sort they replaced;
// cub is used HERE only, as the yardstick).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/sort_bench.cu -o build/sort_bench
//   build/sort_bench            (prints one line per case; exit code 1 on any mismatch)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../mesh_to_sdf_b200/csrc/m2s_sort.cuh"

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e__ = (x);                                                      \
        if (e__ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(2);                                                                \
        }                                                                           \
    } while (0)

using namespace m2s;

static int failures = 0;

template <class F>
static float time_ms(cudaStream_t s, int reps, F f) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

static uint32_t total_key(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// 48-bit keys (Morton codes): uniform or clustered
static void case_u64(cudaStream_t s, uint64_t n, int bits, int kind, int reps) {
    std::mt19937_64 rng(n * 7 + kind);
    std::vector<uint64_t> h(n);
    for (auto& k : h) {
        uint64_t r = rng();
        if (kind == 1) r = (r & 0xffff) | ((r >> 60) << 40);  // few distinct top digits, many duplicates
        if (kind == 2) r = 5;                                 // all equal
        k = bits >= 64 ? r : (r & ((1ull << bits) - 1));
    }
    std::vector<uint32_t> ref(n);
    std::iota(ref.begin(), ref.end(), 0u);
    std::stable_sort(ref.begin(), ref.end(), [&](uint32_t a, uint32_t b) { return h[a] < h[b]; });
    uint64_t *k_in, *k_a, *k_b;
    uint32_t *v_a, *v_b, *v_iota;
    CK(cudaMalloc(&k_in, n * 8 + 8));
    CK(cudaMalloc(&k_a, n * 8 + 8));
    CK(cudaMalloc(&k_b, n * 8 + 8));
    CK(cudaMalloc(&v_a, n * 4 + 4));
    CK(cudaMalloc(&v_b, n * 4 + 4));
    CK(cudaMalloc(&v_iota, n * 4 + 4));
    CK(cudaMemcpy(k_in, h.data(), n * 8, cudaMemcpyHostToDevice));
    std::vector<uint32_t> io(n);
    std::iota(io.begin(), io.end(), 0u);
    CK(cudaMemcpy(v_iota, io.data(), n * 4, cudaMemcpyHostToDevice));
    void* scratch;
    CK(cudaMalloc(&scratch, radix_sort_scratch_bytes(n)));
    uint64_t* const kbuf[2] = {k_a, k_b};
    uint32_t* const vbuf[2] = {v_a, v_b};
    const int res = (radix_sort_passes(bits) - 1) & 1;
    auto ours = [&] { CK(radix_sort_pairs(s, sort_detail::PtrSrc<uint64_t>{k_in}, kbuf, vbuf, n, bits, scratch, true)); };
    const float t_ours = time_ms(s, reps, ours);
    std::vector<uint32_t> got(n);
    std::vector<uint64_t> gotk(n);
    CK(cudaMemcpy(got.data(), vbuf[res], n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gotk.data(), kbuf[res], n * 8, cudaMemcpyDeviceToHost));
    bool ok = got == ref;
    for (uint64_t i = 0; ok && i < n; ++i) ok = gotk[i] == h[ref[i]];
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_a, v_iota, v_a, (int)n, 0, bits, s);
    void* ctmp;
    CK(cudaMalloc(&ctmp, tmp));
    const float t_cub = time_ms(s, reps, [&] { cub::DeviceRadixSort::SortPairs(ctmp, tmp, k_in, k_a, v_iota, v_a, (int)n, 0, bits, s); });
    printf("u64 n=%-10llu bits=%d kind=%d  ours %.4f ms  cub %.4f ms  %s\n", (unsigned long long)n, bits, kind, t_ours, t_cub,
           ok ? "OK" : "MISMATCH");
    failures += !ok;
    for (void* p : {(void*)k_in, (void*)k_a, (void*)k_b, (void*)v_a, (void*)v_b, (void*)v_iota, scratch, ctmp}) cudaFree(p);
}

// order of a float array under f32::total_cmp (grid render order)
static void case_f32(cudaStream_t s, uint64_t n, int kind, int reps) {
    std::mt19937 rng((unsigned)n + kind);
    std::vector<float> h(n);
    std::normal_distribution<float> nd(0.1f, 0.3f);
    for (uint64_t i = 0; i < n; ++i) {
        float v = nd(rng);
        if (kind == 1) {  // smooth field with ties, zeros of both signs, infinities, NaNs of both signs
            v = std::round(std::sin(i * 1e-3f) * 50.f) / 50.f;
            if (i % 1000 == 3) v = -0.0f;
            if (i % 1000 == 4) v = INFINITY;
            if (i % 1000 == 5) v = -INFINITY;
            if (i % 1000 == 6) v = NAN;
            if (i % 1000 == 7) v = -NAN;
        }
        if (kind == 2) {  // what the product sorts: the distance field of a shape on a cubic grid
            const uint64_t side = (uint64_t)std::llround(std::cbrt((double)n));
            const float x = (float)(i / (side * side)) / side - 0.5f, y = (float)((i / side) % side) / side - 0.5f,
                        z = (float)(i % side) / side - 0.5f;
            v = std::sqrt(x * x + y * y + z * z) - 0.3f;
        }
        h[i] = v;
    }
    std::vector<uint32_t> ref(n);
    std::iota(ref.begin(), ref.end(), 0u);
    std::stable_sort(ref.begin(), ref.end(), [&](uint32_t a, uint32_t b) { return total_key(h[a]) < total_key(h[b]); });
    float* d_sdf;
    uint32_t *keys, *idx, *order, *ck, *ci;
    CK(cudaMalloc(&d_sdf, n * 4 + 4));
    CK(cudaMalloc(&keys, n * 8 + 8));
    CK(cudaMalloc(&idx, n * 4 + 4));
    CK(cudaMalloc(&order, n * 4 + 4));
    CK(cudaMalloc(&ck, n * 8 + 8));
    CK(cudaMalloc(&ci, n * 4 + 4));
    CK(cudaMemcpy(d_sdf, h.data(), n * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> hk(n), io(n);
    for (uint64_t i = 0; i < n; ++i) hk[i] = total_key(h[i]);
    std::iota(io.begin(), io.end(), 0u);
    CK(cudaMemcpy(ck, hk.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ci, io.data(), n * 4, cudaMemcpyHostToDevice));
    void* scratch;
    CK(cudaMalloc(&scratch, radix_sort_scratch_bytes(n)));
    uint32_t* const kbuf[2] = {keys, keys + n};
    uint32_t* const vbuf[2] = {idx, order};
    const float t_ours =
        time_ms(s, reps, [&] { CK(radix_sort_pairs(s, sort_detail::F32TotalOrderSrc{d_sdf}, kbuf, vbuf, n, 32, scratch, false)); });
    std::vector<uint32_t> got(n);
    CK(cudaMemcpy(got.data(), order, n * 4, cudaMemcpyDeviceToHost));
    const bool ok = got == ref;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, ck, ck + n, ci, idx, (int)n, 0, 32, s);
    void* ctmp;
    CK(cudaMalloc(&ctmp, tmp));
    const float t_cub = time_ms(s, reps, [&] { cub::DeviceRadixSort::SortPairs(ctmp, tmp, ck, ck + n, ci, idx, (int)n, 0, 32, s); });
    const double gb = 60.0 * n / 1e9;
    printf("f32 n=%-10llu kind=%d  ours %.4f ms (%.0f GB/s on 60 B/pair)  cub %.4f ms (+ key/index setup)  %s\n",
           (unsigned long long)n, kind, t_ours, gb / (t_ours * 1e-3), t_cub, ok ? "OK" : "MISMATCH");
    failures += !ok;
    for (void* p : {(void*)d_sdf, (void*)keys, (void*)idx, (void*)order, (void*)ck, (void*)ci, scratch, ctmp}) cudaFree(p);
}

static void case_scan(cudaStream_t s, uint64_t n, int reps) {
    std::mt19937 rng((unsigned)n);
    std::vector<uint32_t> h(n), ref(n);
    for (auto& v : h) v = rng() % 13;
    uint32_t run = 0;
    for (uint64_t i = 0; i < n; ++i) {
        ref[i] = run;
        run += h[i];
    }
    uint32_t *in, *out;
    CK(cudaMalloc(&in, n * 4 + 4));
    CK(cudaMalloc(&out, n * 4 + 4));
    CK(cudaMemcpy(in, h.data(), n * 4, cudaMemcpyHostToDevice));
    void* scratch;
    CK(cudaMalloc(&scratch, exclusive_scan_scratch_bytes(n)));
    const float t_ours = time_ms(s, reps, [&] { CK(exclusive_scan_u32(s, in, out, n, scratch)); });
    std::vector<uint32_t> got(n);
    CK(cudaMemcpy(got.data(), out, n * 4, cudaMemcpyDeviceToHost));
    const bool ok = got == ref;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, s);
    void* ctmp;
    CK(cudaMalloc(&ctmp, tmp));
    const float t_cub = time_ms(s, reps, [&] { cub::DeviceScan::ExclusiveSum(ctmp, tmp, in, out, (int)n, s); });
    printf("scan n=%-10llu ours %.4f ms  cub %.4f ms  %s\n", (unsigned long long)n, t_ours, t_cub, ok ? "OK" : "MISMATCH");
    failures += !ok;
    for (void* p : {(void*)in, (void*)out, scratch, ctmp}) cudaFree(p);
}

int main(int argc, char** argv) {
    const bool big = argc > 1 && !strcmp(argv[1], "--big");
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    if (argc > 1 && !strcmp(argv[1], "--profile")) {  // one large case for ncu
        case_f32(s, 16777216ull, 2, 1);
        return failures;
    }
    if (argc > 1 && !strcmp(argv[1], "--quick")) {  // A/B of build variants
        case_u64(s, 100352, 48, 0, 20);
        case_u64(s, 1003520, 48, 0, 20);
        case_u64(s, 1003520, 48, 1, 20);
        for (int kind = 0; kind < 3; ++kind) case_f32(s, 16777216ull, kind, 20);
        return failures;
    }
    for (uint64_t n : {1ull, 2ull, 31ull, 33ull, 4095ull, 4096ull, 4097ull, 12289ull, 100352ull, 1003520ull}) {
        case_u64(s, n, 48, 0, 20);
        case_u64(s, n, 48, 1, 20);
    }
    case_u64(s, 50000, 48, 2, 20);
    case_u64(s, 300000, 64, 0, 20);
    case_u64(s, 300000, 8, 0, 20);
    case_u64(s, 300000, 16, 0, 20);
    for (uint64_t n : {1ull, 100ull, 4097ull, 262144ull, 2097152ull, 16777216ull}) {
        case_f32(s, n, 0, 20);
        case_f32(s, n, 1, 20);
        if (n >= 262144) case_f32(s, n, 2, 20);
    }
    if (big) case_f32(s, 134217728ull, 0, 5);
    for (uint64_t n : {1ull, 2047ull, 2048ull, 2049ull, 65537ull, 786433ull, 5000001ull}) case_scan(s, n, 20);
    printf(failures ? "FAILED %d\n" : "all OK\n", failures);
    return failures ? 1 : 0;
}
