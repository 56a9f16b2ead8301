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

// ---------------------------------------------------------------- scatter
constexpr int kScatterThreads = 256;

__global__ void __launch_bounds__(kScatterThreads)
k_scatter(Dims d, Geom geo, unsigned long long *__restrict__ keys) {
    if (*geo.n_instances > d.capacity) return;  // truncated call: caller re-runs with more room
    const size_t vg = (size_t)blockIdx.x * kScatterThreads + threadIdx.x;
    const size_t total = (size_t)d.S * d.V * d.P;
    if (vg >= total) return;
    if (geo.radii[vg] <= 0) return;
    const int vid = (int)(vg / d.P);
    const uint32_t g = (uint32_t)(vg - (size_t)vid * d.P);
    const ushort4 r = geo.rect[vg];
    const unsigned long long key = ((unsigned long long)__float_as_uint(geo.depth[vg]) << 32) | g;
    uint32_t *cur = geo.tile_cursor + (size_t)vid * d.tiles;
    for (int ty = r.y; ty < r.w; ++ty)
        for (int tx = r.x; tx < r.z; ++tx) {
            const uint32_t slot = atomicAdd(&cur[ty * d.gx + tx], 1u);
            keys[slot] = key;
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
            int smem_cap, int min_n, int max_n, int id_bits) {
    static_assert(kSortThreads == 256, "one thread per 8-bit digit");
    extern __shared__ __align__(16) unsigned char s_raw[];
    if (*n_instances > capacity) return;
    const int seg = blockIdx.x;
    const int n = (int)tile_count[seg];
    if (n < 2 || n <= min_n || n > max_n) return;  // trivial, or handled by the other launch
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

static size_t sort_smem_bytes(int cap) {
    return sizeof(uint32_t) * (kSortWarps * 256 + 256 + 16) + sizeof(unsigned long long) * 2 * (size_t)cap;
}

constexpr int kSortCapSmall = 2048;    // 32 KB of keys -> several CTAs per SM
constexpr int kSortCapLarge = 12288;   // 192 KB of keys -> one CTA per SM; beyond: global ping-pong

int launch_binning(const Dims &d, const Geom &g, unsigned long long *keys,
                   unsigned long long *keys_alt, int sort_impl, cudaStream_t st) {
    const int n_seg = d.S * d.V * d.tiles;
    k_tile_scan<<<1, kScanThreads, 0, st>>>(n_seg, g.tile_count, g.tile_start, g.tile_cursor, g.n_instances);
    PS_LAUNCH_CHECK("k_tile_scan");
    const size_t total = (size_t)d.S * d.V * d.P;
    k_scatter<<<(unsigned)((total + kScatterThreads - 1) / kScatterThreads), kScatterThreads, 0, st>>>(d, g, keys);
    PS_LAUNCH_CHECK("k_scatter");
    mark(kMarkScatter, st);

    int id_bits = 1;
    while ((1ll << id_bits) < d.P) ++id_bits;

    if (sort_impl == 1) {
        // Debug path: CUB segmented sort over the same segments (used only to cross-check
        // the native sort in tests; needs n_instances <= capacity, guaranteed by the caller).
        size_t temp = 0;
        cub::DoubleBuffer<unsigned long long> db(keys, keys_alt);
        // tile_start has n_seg entries; end offsets = start + count: build them in tile_cursor,
        // which after k_scatter already equals start + count.
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
                                           (int)sort_smem_bytes(kSortCapLarge)));
        attr_set = true;
    }
    k_tile_sort<<<n_seg, kSortThreads, sort_smem_bytes(kSortCapSmall), st>>>(
        g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, keys_alt, kSortCapSmall, 0, kSortCapSmall, id_bits);
    PS_LAUNCH_CHECK("k_tile_sort(small)");
    // second launch only does work for segments longer than kSortCapSmall
    k_tile_sort<<<n_seg, kSortThreads, sort_smem_bytes(kSortCapLarge), st>>>(
        g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, keys_alt, kSortCapLarge, kSortCapSmall, 0x7fffffff, id_bits);
    PS_LAUNCH_CHECK("k_tile_sort(large)");
    return PS_OK;
}

}  // namespace ps
