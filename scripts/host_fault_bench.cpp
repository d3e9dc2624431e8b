// Micro-benchmark (host only): how fast can T threads fill a FRESH 64 MiB allocation (what a facade's Vec<f32> is)
// from a warm source buffer, with the page-fault strategies the pipelined host path could use?
//   g++ -O2 -pthread scripts/host_fault_bench.cpp -o /tmp/hfb && /tmp/hfb
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t bytes = 64u << 20, chunk = 1u << 20;
    char* src = (char*)malloc(bytes);
    memset(src, 1, bytes);
    for (int threads : {1, 2, 4, 8}) {
        for (int mode = 0; mode < 4; ++mode) {  // 0 plain, 1 hugepage advice, 2 populate per chunk, 3 hugepage + populate
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                char* dst = (char*)malloc(bytes);  // fresh mmap'd region: untouched pages
                const double t0 = now();
                if (mode == 1 || mode == 3) {
                    uintptr_t a = ((uintptr_t)dst + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1), b = ((uintptr_t)dst + bytes) & ~(uintptr_t)((2u << 20) - 1);
                    madvise((void*)a, b - a, MADV_HUGEPAGE);
                }
                std::vector<std::thread> ts;
                for (int t = 0; t < threads; ++t)
                    ts.emplace_back([&, t] {
                        for (size_t o = (size_t)t * chunk; o < bytes; o += (size_t)threads * chunk) {
                            if (mode >= 2) {
                                uintptr_t a = ((uintptr_t)dst + o) & ~(uintptr_t)4095, b = ((uintptr_t)dst + o + chunk + 4095) & ~(uintptr_t)4095;
                                madvise((void*)a, b - a, MADV_POPULATE_WRITE);
                            }
                            memcpy(dst + o, src + o, chunk);
                        }
                    });
                for (auto& th : ts) th.join();
                best = std::min(best, now() - t0);
                free(dst);
            }
            printf("threads %d mode %d (%s): %.2f ms = %.1f GB/s\n", threads, mode,
                   mode == 0 ? "plain" : mode == 1 ? "MADV_HUGEPAGE" : mode == 2 ? "POPULATE_WRITE per chunk" : "HUGEPAGE + POPULATE", best, bytes / best / 1e6);
        }
    }
    // warm destination for reference
    char* dst = (char*)malloc(bytes);
    memset(dst, 0, bytes);
    for (int threads : {1, 4, 8}) {
        const double t0 = now();
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; ++t)
            ts.emplace_back([&, t] { for (size_t o = (size_t)t * chunk; o < bytes; o += (size_t)threads * chunk) memcpy(dst + o, src + o, chunk); });
        for (auto& th : ts) th.join();
        printf("threads %d warm destination: %.2f ms\n", threads, now() - t0);
    }
    return 0;
}
