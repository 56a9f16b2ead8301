// Binning: (view, tile) instance counts -> exclusive scan -> scatter of
// (float_bits(depth) << 32 | gaussian) keys into per-tile segments -> per-tile LSD radix sort.
//
// The sorted order inside a tile (ascending depth bits, ties by ascending Gaussian index) is
// exactly what upstream obtains from its global stable radix sort of (tile << 32 | depth) keys
// over Gaussian-ordered emission, so keys / tile ranges are bit-identical to
// the reference's binning buffers while the sort itself never leaves shared memory.
#include <cub/device/device_segmented_radix_sort.cuh>

#include "ps_common.cuh"

namespace ps {

// ---------------------------------------------------------------- scan of tile counts
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
k_tile_scan(int n, const uint32_t *__restrict__ count, uint32_t *__restrict__ start,
            uint32_t *__restrict__ cursor, long long *__restrict__ n_instances) {
    __shared__ uint32_t warp_sums[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kScanThreads) {
        const int i = base + tid;
        const uint32_t c = i < n ? count[i] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;  // inclusive
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const uint32_t warp_off = warp ? warp_sums[warp - 1] : 0u;
        const unsigned long long excl = carry + warp_off + (x - c);
        if (i < n) {
            // offsets beyond 2^32-1 cannot be represented; clamp (capacity check rejects the call)
            const uint32_t e = excl > 0xffffffffull ? 0xffffffffu : (uint32_t)excl;
            start[i] = e;
            cursor[i] = e;
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (tid == 0) *n_instances = (long long)carry_s;
}

// longest (view, tile) segment -> n_instances[1]; lets the host pick the sort configuration
__global__ void k_tile_max(int n, const uint32_t *__restrict__ count, long long *__restrict__ out) {
    uint32_t m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, count[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)m);
}

// ---------------------------------------------------------------- scatter
constexpr int kScatterThreads = 256;
constexpr int kScatterMaxSmemTiles = 8192;

// One CTA = 256 consecutive Gaussians of one view.  Slots are reserved per (CTA, tile) with a
// single global atomic; ranks inside the CTA come from shared-memory atomics.  (Consecutive
// Gaussians are neighbouring context pixels and land on a handful of tiles, so per-instance
// global atomics serialise on ~tiles addresses: 68 us -> a few us at configs[1].)
__global__ void __launch_bounds__(kScatterThreads)
k_scatter(Dims d, Geom geo, unsigned long long *__restrict__ keys, int use_smem) {
    extern __shared__ uint32_t s_scatter[];          // cnt[tiles], base[tiles]
    if (*geo.n_instances > d.capacity) return;       // truncated call: caller re-runs with more room
    const int vid = blockIdx.y;
    const int g = blockIdx.x * kScatterThreads + threadIdx.x;
    uint32_t *cnt = s_scatter, *base = s_scatter + d.tiles;
    uint32_t *cur = geo.tile_cursor + (size_t)vid * d.tiles;
    const size_t vg = (size_t)vid * d.P + g;
    const bool vis = g < d.P && geo.radii[vg] > 0;
    ushort4 r = make_ushort4(0, 0, 0, 0);
    unsigned long long key = 0;
    if (vis) {
        r = geo.rect[vg];
        key = ((unsigned long long)__float_as_uint(geo.depth[vg]) << 32) | (uint32_t)g;
    }
    if (!use_smem) {
        for (int ty = r.y; ty < r.w; ++ty)
            for (int tx = r.x; tx < r.z; ++tx) keys[atomicAdd(&cur[ty * d.gx + tx], 1u)] = key;
        return;
    }
    for (int i = threadIdx.x; i < d.tiles; i += kScatterThreads) cnt[i] = 0;
    __syncthreads();
    for (int ty = r.y; ty < r.w; ++ty)
        for (int tx = r.x; tx < r.z; ++tx) atomicAdd(&cnt[ty * d.gx + tx], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < d.tiles; i += kScatterThreads) {
        const uint32_t c = cnt[i];
        if (c) { base[i] = atomicAdd(&cur[i], c); cnt[i] = 0; }
    }
    __syncthreads();
    for (int ty = r.y; ty < r.w; ++ty)
        for (int tx = r.x; tx < r.z; ++tx) {
            const int t = ty * d.gx + tx;
            keys[base[t] + atomicAdd(&cnt[t], 1u)] = key;
        }
}

// ---------------------------------------------------------------- per-tile radix sort
// One CTA per (view, tile) segment.  8-bit LSD passes over the Gaussian-index bits then the
// depth bits; a pass whose digit is uniform over the segment is skipped.  Each warp owns a
// contiguous slice of the segment so the pass is stable with three barriers.
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

__device__ __forceinline__ uint32_t digit_of(unsigned long long k, int shift) {
    return (uint32_t)(k >> shift) & 255u;
}

__global__ void __launch_bounds__(kSortThreads)
k_tile_sort(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ tile_count,
            const long long *__restrict__ n_instances, long long capacity,
            unsigned long long *__restrict__ keys, unsigned long long *__restrict__ keys_alt,
            int smem_cap, int min_n, int id_bits) {
    static_assert(kSortThreads == 256, "one thread per 8-bit digit");
    extern __shared__ __align__(16) unsigned char s_raw[];
    if (*n_instances > capacity) return;
    const int seg = blockIdx.x;
    const int n = (int)tile_count[seg];
    if (n < 2 || n <= min_n) return;   // short segments were sorted by the bitonic launch
    const uint32_t s0 = tile_start[seg];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t *cnt = reinterpret_cast<uint32_t *>(s_raw);                 // [kSortWarps][256]
    uint32_t *digit_total = cnt + kSortWarps * 256;                      // [256]
    uint32_t *misc = digit_total + 256;                                  // [16]
    unsigned long long *sbuf = reinterpret_cast<unsigned long long *>(misc + 16);

    unsigned long long *a, *b;
    const bool in_smem = n <= smem_cap;
    if (in_smem) {
        a = sbuf;
        b = sbuf + smem_cap;
        for (int i = tid; i < n; i += kSortThreads) a[i] = keys[s0 + i];
    } else {
        a = keys + s0;
        b = keys_alt + s0;
    }
    __syncthreads();

    const int lo = (int)(((long long)n * warp) / kSortWarps);
    const int hi = (int)(((long long)n * (warp + 1)) / kSortWarps);

    for (int pass = 0; pass < 8; ++pass) {
        // passes 0..3 -> bits [0,32) (only those below id_bits), 4..7 -> bits [32,64)
        const int shift = pass * 8;
        if (pass < 4 && shift >= id_bits) continue;
        for (int i = tid; i < kSortWarps * 256; i += kSortThreads) cnt[i] = 0;
        if (tid == 0) misc[0] = 0;
        __syncthreads();
        for (int i = lo + lane; i < hi; i += 32) atomicAdd(&cnt[warp * 256 + digit_of(a[i], shift)], 1u);
        __syncthreads();
        // column sums: thread t owns digit t
        {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t c = cnt[w * 256 + tid];
                cnt[w * 256 + tid] = total;
                total += c;
            }
            digit_total[tid] = total;
            if (total == (uint32_t)n) misc[0] = 1;  // uniform digit: nothing to do
        }
        __syncthreads();
        if (misc[0]) { __syncthreads(); continue; }
        // exclusive scan of digit_total over 256 digits (8 warps x 32)
        {
            const uint32_t v = digit_total[tid];
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) misc[1 + warp] = x;
            __syncthreads();
            uint32_t off = 0;
            for (int w = 0; w < warp; ++w) off += misc[1 + w];
            const uint32_t excl = off + x - v;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) cnt[w * 256 + tid] += excl;
        }
        __syncthreads();
        // stable scatter of this warp's slice
        for (int base = lo; base < hi; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi;
            const unsigned long long key = valid ? a[i] : 0ull;
            const uint32_t dg = valid ? digit_of(key, shift) : (256u + (uint32_t)lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, dg);
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t off = 0;
            if (valid && lane == leader) {
                off = cnt[warp * 256 + dg];
                cnt[warp * 256 + dg] = off + (uint32_t)__popc(peers);
            }
            off = __shfl_sync(0xffffffffu, off, leader);
            if (valid) b[off + rank] = key;
            __syncwarp();
        }
        __syncthreads();
        unsigned long long *t = a; a = b; b = t;
    }
    // result is in `a`
    if (in_smem) {
        for (int i = tid; i < n; i += kSortThreads) keys[s0 + i] = a[i];
    } else if (a != keys + s0) {
        for (int i = tid; i < n; i += kSortThreads) keys[s0 + i] = a[i];
    }
}

// ---------------------------------------------------------------- per-tile bitonic sort
// For the common case (a few thousand keys per tile) a bitonic network over the full 64-bit
// key beats the radix sort by ~8x: no histograms, no atomics, ~log^2(n)/2 barrier-separated
// compare-exchange steps, and it is oblivious to ties.  One CTA per (view, tile); segments
// longer than `cap` (a power of two, chosen by the host from the previous call's longest
// segment) are left to the radix kernel.
constexpr int kBitonicThreads = 1024;

__global__ void __launch_bounds__(kBitonicThreads)
k_tile_sort_bitonic(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ tile_count,
                    const long long *__restrict__ n_instances, long long capacity,
                    unsigned long long *__restrict__ keys, int cap) {
    extern __shared__ __align__(16) unsigned long long s_keys[];
    if (*n_instances > capacity) return;
    const int seg = blockIdx.x;
    const int n = (int)tile_count[seg];
    if (n < 2 || n > cap) return;
    const uint32_t s0 = tile_start[seg];
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    for (int i = threadIdx.x; i < n_pad; i += kBitonicThreads) s_keys[i] = i < n ? keys[s0 + i] : ~0ull;
    __syncthreads();
    // Each warp owns a contiguous chunk of n_pad / 16 keys: every compare-exchange step whose
    // stride is below the chunk size stays inside one warp and needs only __syncwarp; the CTA
    // barrier is paid for the few long-stride steps (10 of 66 at 2048 keys).
    constexpr int kWarps = kBitonicThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = max(n_pad / kWarps, 2);          // power of two
    const int warps_used = n_pad / chunk;              // < kWarps only for tiny segments
    const int half = n_pad >> 1;
    for (int k = 2; k <= n_pad; k <<= 1) {
        int j = k >> 1;
        const bool had_global = j >= chunk;
        for (; j >= chunk; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += kBitonicThreads) {
                const int i = 2 * t - (t & (j - 1));      // bit j of i is clear
                const int p = i + j;
                const unsigned long long a = s_keys[i], b = s_keys[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s_keys[i] = b; s_keys[p] = a; }
            }
            __syncthreads();
        }
        if (warp < warps_used) {
            const int base = warp * chunk;
            for (; j > 0; j >>= 1) {
                for (int t = lane; t < (chunk >> 1); t += 32) {
                    const int i = base + 2 * t - (t & (j - 1));
                    const int p = i + j;
                    const unsigned long long a = s_keys[i], b = s_keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_keys[i] = b; s_keys[p] = a; }
                }
                __syncwarp();
            }
        }
        // the next k's long-stride steps (or the final copy-out) read other warps' chunks
        if (had_global || 2 * k > chunk) __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += kBitonicThreads) keys[s0 + i] = s_keys[i];
}

static size_t sort_smem_bytes(int cap) {
    return sizeof(uint32_t) * (kSortWarps * 256 + 256 + 16) + sizeof(unsigned long long) * 2 * (size_t)cap;
}

// Shared-memory capacities (keys) the host may choose from; a segment longer than the launch's
// capacity is still sorted correctly, ping-ponging through HBM.
static const int kSortCaps[] = {1024, 2048, 4096, 8192, 12288};

int launch_binning(const Dims &d, const Geom &g, unsigned long long *keys,
                   unsigned long long *keys_alt, int sort_impl, int segment_hint, cudaStream_t st) {
    const int n_seg = d.S * d.V * d.tiles;
    k_tile_scan<<<1, kScanThreads, 0, st>>>(n_seg, g.tile_count, g.tile_start, g.tile_cursor, g.n_instances);
    PS_LAUNCH_CHECK("k_tile_scan");
    PS_CUDA_CHECK(cudaMemsetAsync(g.n_instances + 1, 0, sizeof(long long), st));
    k_tile_max<<<min(148, (n_seg + 255) / 256), 256, 0, st>>>(n_seg, g.tile_count, g.n_instances);
    PS_LAUNCH_CHECK("k_tile_max");
    const int use_smem = d.tiles <= kScatterMaxSmemTiles;
    static bool scatter_attr = false;
    if (!scatter_attr) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(2 * sizeof(uint32_t) * kScatterMaxSmemTiles)));
        scatter_attr = true;
    }
    dim3 sgrid((d.P + kScatterThreads - 1) / kScatterThreads, d.S * d.V);
    k_scatter<<<sgrid, kScatterThreads, use_smem ? 2 * sizeof(uint32_t) * d.tiles : 0, st>>>(d, g, keys, use_smem);
    PS_LAUNCH_CHECK("k_scatter");
    mark(kMarkScatter, st);

    int id_bits = 1;
    while ((1ll << id_bits) < d.P) ++id_bits;

    if (sort_impl == 1) {
        // Debug path: CUB segmented sort over the same segments (used only to cross-check
        // the native sort in tests; needs n_instances <= capacity, guaranteed by the caller).
        size_t temp = 0;
        cub::DoubleBuffer<unsigned long long> db(keys, keys_alt);
        PS_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, temp, db, (int)d.capacity, n_seg,
                                                              g.tile_start, g.tile_cursor, 0, 64, st));
        void *tmp = nullptr;
        PS_CUDA_CHECK(cudaMallocAsync(&tmp, temp, st));
        PS_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortKeys(tmp, temp, db, (int)d.capacity, n_seg,
                                                              g.tile_start, g.tile_cursor, 0, 64, st));
        if (db.Current() != keys)
            PS_CUDA_CHECK(cudaMemcpyAsync(keys, db.Current(), sizeof(unsigned long long) * (size_t)d.capacity,
                                          cudaMemcpyDeviceToDevice, st));
        PS_CUDA_CHECK(cudaFreeAsync(tmp, st));
        return PS_OK;
    }

    static bool attr_set = false;
    if (!attr_set) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sort_smem_bytes(12288)));
        attr_set = true;
    }
    // Bitonic network in shared memory for segments up to 8192 keys (power-of-two capacity picked
    // from the hint, with 25 % head-room); the radix kernel then only does work for segments the
    // bitonic launch skipped (its CTAs exit at once otherwise).
    int bitonic_cap = 2048;
    if (segment_hint > 0) {
        const long long want = (long long)segment_hint + segment_hint / 4;
        while (bitonic_cap < want && bitonic_cap < 8192) bitonic_cap <<= 1;
    }
    static bool battr = false;
    if (!battr) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_tile_sort_bitonic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(8192 * sizeof(unsigned long long))));
        battr = true;
    }
    k_tile_sort_bitonic<<<n_seg, kBitonicThreads, bitonic_cap * sizeof(unsigned long long), st>>>(
        g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, bitonic_cap);
    PS_LAUNCH_CHECK("k_tile_sort_bitonic");
    const bool may_exceed = segment_hint <= 0 || (long long)segment_hint + segment_hint / 4 > bitonic_cap;
    if (may_exceed || sort_impl == 2) {
        k_tile_sort<<<n_seg, kSortThreads, sort_smem_bytes(4096), st>>>(
            g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, keys_alt, 4096, bitonic_cap, id_bits);
        PS_LAUNCH_CHECK("k_tile_sort");
    } else {
        // hint says everything fits; still guarantee correctness if the hint was stale: a tiny
        // grid re-checks and sorts any oversize segment
        k_tile_sort<<<n_seg, kSortThreads, sort_smem_bytes(1024), st>>>(
            g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, keys_alt, 1024, bitonic_cap, id_bits);
        PS_LAUNCH_CHECK("k_tile_sort");
    }
    return PS_OK;
}

}  // namespace ps
