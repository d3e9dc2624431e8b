// Hand-written device sort and scan of libm2s (no library kernels on any path of the product).
//
//   radix_sort_pairs  - stable least-significant-digit radix sort of (key, u32 payload) pairs, 8 bits per pass, every
//                       pass a single sweep over the data ("onesweep": Adinets & Merrill 2022): one kernel reads a
//                       tile of 4096 pairs, ranks it, learns the tile's global digit offsets by decoupled look-back
//                       over the tiles before it (Merrill & Garland 2016) and scatters through shared memory.
//                       Users: the Morton order of the triangles (LBVH build, replaces Bvh::build_par's own sort,
//                       generic/bvh.rs:37-41), the Morton order of scattered queries, and the render order of a grid
//                       (`(0..n).sorted_by(total_cmp)`, mesh_to_sdf_client/src/sdf.rs:65-68).
//   exclusive_scan_u32 - single-pass exclusive prefix sum with the same look-back (ray-bin offsets).
//
// What a general library sort cannot know and this one uses: the payload of the first pass is always the element's
// own position (never read), the keys of the first pass may be computed on the fly from another array (the f32
// distances of a grid: no key array is ever written for pass 0), and the last pass of an index sort does not write
// its keys. HBM traffic per pair for a 32-bit index sort of n floats: 4 (histograms) + 4 + 8 | 8 + 8 | 8 + 8 | 8 + 4
// = 60 B against 4 + 8 (key / index setup) + 4 + 4 x 16 = 80 B for the same passes through a key-value library sort.
//
// Memory model notes. A tile's digit counts and its flag travel in ONE 64-bit word (flag and pass tag in the top byte),
// stored and polled with volatile accesses, so no fence is needed between "value" and "flag". Tiles take their number
// from an atomic counter, so a tile only ever waits for tiles that already run (forward progress without assumptions
// about the block scheduler). The look-back array is tagged with the pass number and reused by every pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m2s {
namespace sort_detail {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
#ifndef M2S_SORT_ITEMS
#define M2S_SORT_ITEMS 16
#endif
#ifndef M2S_SORT_MIN_BLOCKS32
#define M2S_SORT_MIN_BLOCKS32 4
#endif
constexpr int ITEMS = M2S_SORT_ITEMS;     // pairs per thread (even)
constexpr int TILE = THREADS * ITEMS;     // 4096 pairs per block
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr int LOOKBACK_WINDOW = 4;        // predecessor words polled per round trip

constexpr unsigned long long VALUE_MASK = (1ull << 56) - 1;
// top byte of a look-back word: 4 * (pass + 1) + state, state 1 = the tile's own counts, 2 = inclusive prefix
__host__ __device__ constexpr unsigned tag_of(int pass, int state) { return 4u * (unsigned)(pass + 1) + (unsigned)state; }

template <typename K>
struct PtrSrc {  // keys read from an array
    const K* p;
    __device__ __forceinline__ K operator()(uint32_t i) const { return p[i]; }
};

// f32::total_cmp as an unsigned key: -NaN < -inf < ... < -0 < +0 < ... < +inf < NaN (sdf.rs:65-68)
struct F32TotalOrderSrc {
    const float* p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const uint32_t b = __float_as_uint(__ldg(p + i));
        return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    }
};

// The histogram kernel reads every key once; an observer rides along on that read (the iso limits of a grid).
struct NoObserver {
    __device__ __forceinline__ void see(uint32_t, uint64_t) {}
    __device__ __forceinline__ void finish() {}
};

template <typename K>
__device__ __forceinline__ uint32_t digit_of(K k, int shift) { return (uint32_t)(k >> shift) & (RADIX - 1); }

// Lanes of the warp that hold the same digit. `match.any` takes time proportional to the number of distinct values in
// the warp (measured: a pass over random digits cost twice a pass over clustered ones), eight ballots do not; the
// upper digits of neighbouring cells' distances / of Morton-sorted codes are usually all equal, which `match.all` sees
// in one instruction.
#ifndef M2S_SORT_RANK_LDS
#define M2S_SORT_RANK_LDS 1
#endif
#ifndef M2S_SORT_BACKOFF
#define M2S_SORT_BACKOFF 0
#endif
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t dg) {
    int uniform;
    __match_all_sync(0xffffffffu, dg, &uniform);
    if (uniform) return 0xffffffffu;
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
        // four instructions per bit: predicate, ballot, complement where this lane's bit is clear, and
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 w;\n\t"
            "and.b32 w, %1, %2;\n\t"
            "setp.ne.u32 p, w, 0;\n\t"
            "vote.sync.ballot.b32 w, p, 0xffffffff;\n\t"
            "@!p not.b32 w, w;\n\t"
            "and.b32 %0, %0, w;\n\t}"
            : "+r"(peers)
            : "r"(dg), "r"(1u << b));
    }
    return peers;
}

__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// Digit histograms of every pass in one read of the keys: shared-memory atomics, except that a warp whose 32 keys
// share a digit (the rule for the upper digits of distances / Morton codes) adds 32 once.
template <typename K, typename Src, typename Obs>
__global__ void __launch_bounds__(THREADS)
k_sort_histograms(Src src, uint32_t n, int passes, uint32_t* __restrict__ hist /* [passes][RADIX] */, Obs obs) {
    __shared__ uint32_t sh[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += THREADS) sh[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    constexpr int UNROLL = 4;  // independent loads in flight per lane
    const uint64_t stride = (uint64_t)gridDim.x * WARPS * (32 * UNROLL);
    // a warp takes 128 consecutive keys per step and stays in the loop as a whole (match.all needs every lane)
    for (uint64_t base = ((uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * (32 * UNROLL); base < n; base += stride) {
        K k[UNROLL];
        bool valid[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + 32 * u + lane;
            valid[u] = i < n;
            k[u] = valid[u] ? src((uint32_t)i) : K(0);
            if (valid[u]) obs.see((uint32_t)i, (uint64_t)k[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            for (int p = 0; p < passes; ++p) {
                const uint32_t dg = digit_of(k[u], p * RADIX_BITS);
                int uniform;
                __match_all_sync(0xffffffffu, valid[u] ? dg : RADIX, &uniform);
                if (uniform) {
                    if (lane == 0 && valid[u]) atomicAdd(&sh[p * RADIX + dg], 32u);
                } else if (valid[u]) {
                    atomicAdd(&sh[p * RADIX + dg], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    obs.finish();
}

// counts -> exclusive prefixes, one block per pass
static __global__ void __launch_bounds__(RADIX) k_sort_digit_offsets(uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_sum[RADIX / 32];
    uint32_t* h = hist + blockIdx.x * RADIX;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t c = h[t];
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    if (lane == 31) warp_sum[w] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t k = 0; k < w; ++k) before += warp_sum[k];
    h[t] = before + inc - c;
}

// One pass over one digit. src yields the keys of this pass (an array or, in pass 0, anything computed per element);
// vals_in == nullptr means "the payload is the element's position".
template <typename K, typename Src>
__global__ void __launch_bounds__(THREADS, sizeof(K) == 4 ? M2S_SORT_MIN_BLOCKS32 : M2S_SORT_MIN_BLOCKS32 - 1)
k_sort_onesweep(Src src, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                uint32_t n, int shift, int pass, const uint32_t* __restrict__ digit_base /* [RADIX] exclusive */,
                uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ lookback) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys of the tile in local sorted order | their payloads | per warp digit counts -> local offsets of
    // (warp, digit) | per digit: global index of its first local position minus that position
    K* s_keys = reinterpret_cast<K*>(smem_raw);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);
    uint32_t (*warp_cnt)[RADIX] = reinterpret_cast<uint32_t (*)[RADIX]>(s_vals + TILE);
    uint32_t* s_base = &warp_cnt[0][0] + WARPS * RADIX;
    __shared__ uint32_t s_warp_tot[WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint32_t valid = (uint32_t)(n - tile_base < (uint64_t)TILE ? n - tile_base : (uint64_t)TILE);

    // warp-striped: item j of lane l is element  warp chunk + 32 j + l  (coalesced, 2 x 16 loads in flight per lane)
    const uint32_t first = w * (ITEMS * 32) + lane;
    K keys[ITEMS];
    uint32_t vals[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t loc = first + 32 * j;
        keys[j] = loc < valid ? src((uint32_t)(tile_base + loc)) : K(~K(0));  // padding sorts behind everything
    }
    if (vals_in) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t loc = first + 32 * j;
            vals[j] = loc < valid ? vals_in[tile_base + loc] : 0u;
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) vals[j] = (uint32_t)tile_base + first + 32 * j;
    }

    // rank inside the warp: lanes holding the same digit find each other, the first of them takes the warp's counter
    uint32_t rank2[ITEMS / 2];  // two 16-bit ranks per register
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t dg = digit_of(keys[j], shift);
        const uint32_t peers = same_digit_lanes(dg);
#if M2S_SORT_RANK_LDS
        // every lane reads its digit's counter (one LDS instruction for the converged warp), then the first lane of
        // each group stores the new count: no atomic, no shuffle
        __syncwarp();
        const uint32_t prev = warp_cnt[w][dg];
        __syncwarp();
        if ((peers & ((1u << lane) - 1)) == 0) warp_cnt[w][dg] = prev + (uint32_t)__popc(peers);
#else
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if ((int)lane == leader) prev = atomicAdd(&warp_cnt[w][dg], (uint32_t)__popc(peers));
        prev = __shfl_sync(0xffffffffu, prev, leader);
#endif
        const uint32_t r = prev + __popc(peers & ((1u << lane) - 1));
        rank2[j >> 1] = (j & 1) ? (rank2[j >> 1] | (r << 16)) : r;
    }
    __syncthreads();

    // thread t owns digit t: counts over the warps, publish, local digit starts, look-back
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) total += warp_cnt[k][tid];
    unsigned long long* my_word = lookback + (uint64_t)tile * RADIX + tid;
    st_word(my_word, ((unsigned long long)tag_of(pass, tile == 0 ? 2 : 1) << 56) | total);

    uint32_t inc = total;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    if (lane == 31) s_warp_tot[w] = inc;
    __syncthreads();
    uint32_t local_start = inc - total;
    for (uint32_t k = 0; k < w; ++k) local_start += s_warp_tot[k];
    {
        uint32_t run = local_start;
#pragma unroll
        for (int k = 0; k < WARPS; ++k) {
            const uint32_t c = warp_cnt[k][tid];
            warp_cnt[k][tid] = run;
            run += c;
        }
    }

    uint64_t before = 0;  // pairs with this digit in the tiles before this one
    if (tile > 0) {
        // Walk back over the tiles before this one until one of them carries an inclusive prefix. Tags of earlier
        // passes (and the zero of the memset) are smaller than this pass's, so "tag < AGG" means "not published yet".
        // LOOKBACK_WINDOW words are requested per round trip; they are consumed in order up to the first missing one.
        const unsigned AGG = tag_of(pass, 1), PREFIX = tag_of(pass, 2);
        const unsigned long long* wp = my_word - RADIX;  // word of the tile before
        uint32_t left = tile;                            // tiles before `wp`'s, plus one
        while (true) {
            unsigned long long wv[LOOKBACK_WINDOW];
#pragma unroll
            for (int k = 0; k < LOOKBACK_WINDOW; ++k) wv[k] = (uint32_t)k < left ? ld_word(wp - k * RADIX) : 0ull;
            uint32_t used = 0;
            bool done = false;
#pragma unroll
            for (int k = 0; k < LOOKBACK_WINDOW; ++k) {
                const unsigned tg = (unsigned)(wv[k] >> 56);
                if (!done && used == (uint32_t)k && tg >= AGG) {
                    before += wv[k] & VALUE_MASK;
                    used = k + 1;
                    done = tg == PREFIX;
                }
            }
            if (done) break;  // tile 0 always publishes a prefix, so the walk ends there at the latest
#if M2S_SORT_BACKOFF
            if (used == 0) __nanosleep(M2S_SORT_BACKOFF);
#endif
            wp -= used * RADIX;
            left -= used;
        }
        st_word(my_word, ((unsigned long long)PREFIX << 56) | (before + total));
    }
    s_base[tid] = (uint32_t)(digit_base[tid] + before) - local_start;
    __syncthreads();

    // local sorted order in shared memory, then runs of equal digits leave as coalesced stores
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t r = (j & 1) ? (rank2[j >> 1] >> 16) : (rank2[j >> 1] & 0xffffu);
        const uint32_t pos = warp_cnt[w][digit_of(keys[j], shift)] + r;
        s_keys[pos] = keys[j];
        s_vals[pos] = vals[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t p = k * THREADS + tid;
        const K key = s_keys[p];
        const uint32_t dst = s_base[digit_of(key, shift)] + p;
        if (p < valid) {
            if (keys_out) keys_out[dst] = key;
            vals_out[dst] = s_vals[p];
        }
    }
}

template <typename K>
constexpr size_t onesweep_smem_bytes() { return (size_t)TILE * (sizeof(K) + 4) + (WARPS + 1) * RADIX * 4; }

// ---- single-pass exclusive scan -------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = THREADS * SCAN_ITEMS;

// state word: bits 63..62 = 1 (tile sum) / 2 (inclusive prefix), low 62 bits value; zeroed before the launch
static __global__ void __launch_bounds__(THREADS)
k_exclusive_scan_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                     uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ state) {
    __shared__ uint32_t s_data[SCAN_TILE];
    __shared__ uint32_t s_warp_tot[WARPS];
    __shared__ uint32_t s_tile, s_before;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * SCAN_TILE;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint32_t p = k * THREADS + tid;
        s_data[p] = base + p < n ? in[base + p] : 0u;
    }
    __syncthreads();
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = s_data[tid * SCAN_ITEMS + k];
        sum += v[k];
    }
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    if (lane == 31) s_warp_tot[w] = inc;
    __syncthreads();
    uint32_t excl = inc - sum, tile_total = 0;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) {
        if ((uint32_t)k < w) excl += s_warp_tot[k];
        tile_total += s_warp_tot[k];
    }
    if (w == 0) {  // warp-wide look-back window over the 32 tiles before this one
        if (lane == 0) st_word(state + tile, ((tile == 0 ? 2ull : 1ull) << 62) | tile_total);
        uint64_t before = 0;
        if (tile > 0) {
            int64_t pt = (int64_t)tile - 1;
            while (true) {
                const int64_t p = pt - lane;
                const unsigned long long word = p >= 0 ? ld_word(state + p) : (2ull << 62);
                const unsigned flag = (unsigned)(word >> 62);
                const uint32_t not_ready = __ballot_sync(0xffffffffu, flag == 0);
                const uint32_t is_prefix = __ballot_sync(0xffffffffu, flag == 2);
                const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
                const int first_pf = is_prefix ? __ffs(is_prefix) - 1 : 32;
                const int upto = first_pf < first_nr ? first_pf + 1 : first_nr;  // lanes [0, upto) are consumed
                uint64_t part = (int)lane < upto ? (word & ((1ull << 62) - 1)) : 0ull;
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                before += part;
                if (first_pf < first_nr) break;
                pt -= upto;
            }
            if (lane == 0) st_word(state + tile, (2ull << 62) | (before + tile_total));
        }
        if (lane == 0) s_before = (uint32_t)before;
    }
    __syncthreads();
    uint32_t run = s_before + excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        s_data[tid * SCAN_ITEMS + k] = run;
        run += v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint32_t p = k * THREADS + tid;
        if (base + p < n) out[base + p] = s_data[p];
    }
}

}  // namespace sort_detail

// ---- host side ---------------------------------------------------------------------------------------------------

inline int radix_sort_passes(int key_bits) { return (key_bits + sort_detail::RADIX_BITS - 1) / sort_detail::RADIX_BITS; }

// scratch: [histograms MAX_PASSES x 256 u32][tile counters MAX_PASSES u32, padded][look-back tiles x 256 u64]
inline size_t radix_sort_scratch_bytes(uint64_t n) {
    using namespace sort_detail;
    const uint64_t tiles = (n + TILE - 1) / TILE;
    return (size_t)(MAX_PASSES * RADIX * 4 + 256 + tiles * RADIX * 8);
}

// Sorts n pairs by bits [0, key_bits) of their keys, stably. Pass 0 reads its keys from `src` (any functor) and takes
// the element positions 0..n-1 as payloads; pass i writes kbuf[i & 1] / vbuf[i & 1], so the result is in buffer
// (passes - 1) & 1. The last pass writes keys only if keys_of_result. n < 2^32; launches (optional) is bumped by
// the number of kernels enqueued.
template <typename K, typename Src, typename Obs = sort_detail::NoObserver>
cudaError_t radix_sort_pairs(cudaStream_t s, Src src, K* const kbuf[2], uint32_t* const vbuf[2], uint64_t n, int key_bits,
                             void* scratch, bool keys_of_result, uint64_t* launches = nullptr, Obs obs = Obs{}) {
    using namespace sort_detail;
    const int passes = radix_sort_passes(key_bits);
    if (n == 0 || passes == 0 || passes > MAX_PASSES) return n == 0 ? cudaSuccess : cudaErrorInvalidValue;
    const uint32_t tiles = (uint32_t)((n + TILE - 1) / TILE);
    uint32_t* hist = static_cast<uint32_t*>(scratch);
    uint32_t* counters = hist + MAX_PASSES * RADIX;
    unsigned long long* lookback = reinterpret_cast<unsigned long long*>(counters + 64);
    cudaError_t e = cudaMemsetAsync(scratch, 0, radix_sort_scratch_bytes(n), s);
    if (e != cudaSuccess) return e;
    const unsigned hist_blocks = (unsigned)(tiles < 148u * 8u ? tiles : 148u * 8u);  // a tile = 4 steps of a block
    k_sort_histograms<K, Src, Obs><<<hist_blocks, THREADS, 0, s>>>(src, (uint32_t)n, passes, hist, obs);
    k_sort_digit_offsets<<<passes, RADIX, 0, s>>>(hist);
    constexpr size_t smem = onesweep_smem_bytes<K>();
    if (smem > 48 * 1024) {  // opt-in above 48 KB; the attribute belongs to the current device's context
        e = cudaFuncSetAttribute(k_sort_onesweep<K, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_sort_onesweep<K, PtrSrc<K>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    for (int p = 0; p < passes; ++p) {
        const bool last = p == passes - 1;
        K* ko = (last && !keys_of_result) ? nullptr : kbuf[p & 1];
        if (p == 0)
            k_sort_onesweep<K, Src><<<tiles, THREADS, smem, s>>>(src, nullptr, ko, vbuf[0], (uint32_t)n, 0, 0, hist,
                                                                counters, lookback);
        else
            k_sort_onesweep<K, PtrSrc<K>><<<tiles, THREADS, smem, s>>>(PtrSrc<K>{kbuf[(p - 1) & 1]}, vbuf[(p - 1) & 1], ko,
                                                                     vbuf[p & 1], (uint32_t)n, p * RADIX_BITS, p,
                                                                     hist + p * RADIX, counters + p, lookback);
    }
    if (launches) *launches += 2 + passes;
    return cudaGetLastError();
}

inline size_t exclusive_scan_scratch_bytes(uint64_t n) {
    return (size_t)(64 + ((n + sort_detail::SCAN_TILE - 1) / sort_detail::SCAN_TILE) * 8);
}

// out[i] = in[0] + ... + in[i-1]; in and out may be the same array. scratch: [tile counter, padded][state words]
inline cudaError_t exclusive_scan_u32(cudaStream_t s, const uint32_t* in, uint32_t* out, uint64_t n, void* scratch,
                                      uint64_t* launches = nullptr) {
    using namespace sort_detail;
    if (n == 0) return cudaSuccess;
    const uint32_t tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    cudaError_t e = cudaMemsetAsync(scratch, 0, exclusive_scan_scratch_bytes(n), s);
    if (e != cudaSuccess) return e;
    uint32_t* counter = static_cast<uint32_t*>(scratch);
    k_exclusive_scan_u32<<<tiles, THREADS, 0, s>>>(in, out, (uint32_t)n, counter,
                                                   reinterpret_cast<unsigned long long*>(counter + 16));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace m2s
