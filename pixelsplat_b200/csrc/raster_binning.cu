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
    uint32_t cmax = 0;
    for (int base = 0; base < n; base += kScanThreads) {
        const int i = base + tid;
        const uint32_t c = i < n ? count[i] : 0u;
        cmax = max(cmax, c);
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
    // longest (view, tile) segment -> n_instances[1]; lets the host pick the sort configuration
    cmax = __reduce_max_sync(0xffffffffu, cmax);
    if (lane == 0 && cmax) atomicMax(reinterpret_cast<unsigned long long *>(n_instances + 1), (unsigned long long)cmax);
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

// 8-bit LSD radix sort of one segment, ping-ponging through global memory (segments too long for
// the shared-memory sort).  `s_raw` needs sort_smem_bytes(0) bytes.
__device__ void radix8_segment(unsigned char *s_raw, int n, uint32_t s0, unsigned long long *__restrict__ keys,
                               unsigned long long *__restrict__ keys_alt, int id_bits) {
    static_assert(kSortThreads == 256, "one thread per 8-bit digit");
    const int smem_cap = 0;            // always the global ping-pong
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

constexpr int kBitonicThreads = 256;

// Bitonic network over s_keys[0, n_pad) (n_pad a power of two), executed by the whole CTA.
// Each warp owns a contiguous chunk: compare-exchange steps whose stride stays inside the chunk
// need only __syncwarp; the CTA barrier is paid for the few long-stride steps.
template <int THREADS>
__device__ void bitonic_sort_smem(unsigned long long *s_keys, int n_pad) {
    constexpr int kWarps = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = max(n_pad / kWarps, 2);
    const int warps_used = n_pad / chunk;
    const int half = n_pad >> 1;
    for (int k = 2; k <= n_pad; k <<= 1) {
        int j = k >> 1;
        const bool had_global = j >= chunk;
        for (; j >= chunk; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += THREADS) {
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
        if (had_global || 2 * k > chunk) __syncthreads();
    }
}

// ---------------------------------------------------------------- per-tile 11-bit radix sort
// The common case: a few thousand keys per (view, tile), depth bits spanning < 2^33/2^22.  Keys are
// sorted on (depth - min depth of the tile) with stable 11-bit LSD passes (3 passes for any
// 32-bit range, 2 when the range fits 22 bits), entirely in shared memory, each warp owning a
// contiguous slice (counting and ranking with match.any, no atomics).  The Gaussian-index
// tie-break of the full 64-bit order is restored afterwards: runs of equal depth (rare: exact
// float collisions) are insertion-sorted by their first thread, and a pathological segment (a run
// longer than 64) falls back to the bitonic network on the full key.  ~3x fewer instructions than
// the bitonic network alone (profiles/r01_ncu_metrics_v6.csv: 12.5 M for 256 tiles).
constexpr int kR11Threads = 256;
constexpr int kR11Warps = kR11Threads / 32;
constexpr int kR11Bins = 2048;

static size_t radix11_smem_bytes(int cap) {
    return sizeof(unsigned long long) * 2 * (size_t)cap + sizeof(uint16_t) * kR11Warps * kR11Bins + 64 * sizeof(uint32_t);
}

__global__ void __launch_bounds__(kR11Threads)
k_tile_sort_radix11(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ tile_count,
                    const long long *__restrict__ n_instances, long long capacity,
                    unsigned long long *__restrict__ keys, unsigned long long *__restrict__ keys_alt, int cap,
                    int id_bits) {
    extern __shared__ __align__(16) unsigned char s_r11[];
    if (*n_instances > capacity) return;
    const int seg = blockIdx.x;
    const int n = (int)tile_count[seg];
    if (n < 2) return;
    const uint32_t s0 = tile_start[seg];
    if (n > cap) {   // too long for shared memory: 8-bit radix through HBM (always correct, just slower)
        radix8_segment(s_r11, n, s0, keys, keys_alt, id_bits);
        return;
    }
    unsigned long long *A = reinterpret_cast<unsigned long long *>(s_r11);
    unsigned long long *B = A + cap;
    uint16_t *cnt = reinterpret_cast<uint16_t *>(B + cap);               // [warps][2048]
    uint32_t *misc = reinterpret_cast<uint32_t *>(cnt + kR11Warps * kR11Bins);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { misc[0] = 0xffffffffu; misc[1] = 0u; misc[3] = 0u; }
    __syncthreads();
    uint32_t dlo = 0xffffffffu, dhi = 0u;
    for (int i = tid; i < n; i += kR11Threads) {
        const unsigned long long k = keys[s0 + i];
        A[i] = k;
        const uint32_t dpt = (uint32_t)(k >> 32);
        dlo = min(dlo, dpt); dhi = max(dhi, dpt);
    }
    dlo = __reduce_min_sync(0xffffffffu, dlo);
    dhi = __reduce_max_sync(0xffffffffu, dhi);
    if (lane == 0) { atomicMin(&misc[0], dlo); atomicMax(&misc[1], dhi); }
    __syncthreads();
    const uint32_t dmin = misc[0], range = misc[1] - dmin;
    const int bits = range ? 32 - __clz(range) : 0;
    const int passes = (bits + 10) / 11;
    const int lo_i = (int)(((long long)n * warp) / kR11Warps);
    const int hi_i = (int)(((long long)n * (warp + 1)) / kR11Warps);
    uint16_t *my = cnt + warp * kR11Bins;

    for (int p = 0; p < passes; ++p) {
        const int shift = 11 * p;
        {   // zero this warp's counters
            uint32_t *my32 = reinterpret_cast<uint32_t *>(my);
            for (int i = lane; i < kR11Bins / 2; i += 32) my32[i] = 0u;
            __syncwarp();
        }
        for (int base = lo_i; base < hi_i; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi_i;
            const uint32_t dg = valid ? ((((uint32_t)(A[i] >> 32)) - dmin) >> shift) & (kR11Bins - 1) : (4096u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, dg);
            if (valid && lane == __ffs(peers) - 1) my[dg] = (uint16_t)(my[dg] + __popc(peers));
            __syncwarp();
        }
        __syncthreads();
        {   // column sums -> per-(warp, digit) start offsets; thread t owns digits 8t .. 8t+7.
            // All 64 counters are loaded first (independent 16-byte loads: 8 consecutive u16 per
            // warp row) so the shared-memory latency is paid once, not 64 times in a chain.
            uint32_t c[kR11Warps][8];
#pragma unroll
            for (int w = 0; w < kR11Warps; ++w) {
                const uint4 v = *reinterpret_cast<const uint4 *>(cnt + w * kR11Bins + 8 * tid);
                c[w][0] = v.x & 0xffffu; c[w][1] = v.x >> 16; c[w][2] = v.y & 0xffffu; c[w][3] = v.y >> 16;
                c[w][4] = v.z & 0xffffu; c[w][5] = v.z >> 16; c[w][6] = v.w & 0xffffu; c[w][7] = v.w >> 16;
            }
            uint32_t tot[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < kR11Warps; ++w) { const uint32_t t = c[w][j]; c[w][j] = run; run += t; }
                tot[j] = run;
                sum += run;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) misc[8 + warp] = incl;
            __syncthreads();
            uint32_t run2 = incl - sum;
            for (int w = 0; w < warp; ++w) run2 += misc[8 + w];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int w = 0; w < kR11Warps; ++w) c[w][j] += run2;
                run2 += tot[j];
            }
#pragma unroll
            for (int w = 0; w < kR11Warps; ++w) {
                uint4 v;
                v.x = c[w][0] | (c[w][1] << 16); v.y = c[w][2] | (c[w][3] << 16);
                v.z = c[w][4] | (c[w][5] << 16); v.w = c[w][6] | (c[w][7] << 16);
                *reinterpret_cast<uint4 *>(cnt + w * kR11Bins + 8 * tid) = v;
            }
        }
        __syncthreads();
        for (int base = lo_i; base < hi_i; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi_i;
            const unsigned long long key = valid ? A[i] : 0ull;
            const uint32_t dg = valid ? ((((uint32_t)(key >> 32)) - dmin) >> shift) & (kR11Bins - 1) : (4096u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, dg);
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t off = 0;
            if (valid && lane == leader) {
                off = my[dg];
                my[dg] = (uint16_t)(off + __popc(peers));
            }
            off = __shfl_sync(0xffffffffu, off, leader);
            if (valid) B[off + rank] = key;
            __syncwarp();
        }
        __syncthreads();
        unsigned long long *t = A; A = B; B = t;
    }
    // ---- restore the Gaussian-index order inside runs of identical depth
    for (int i = tid; i < n; i += kR11Threads) {
        const uint32_t dpt = (uint32_t)(A[i] >> 32);
        const bool starts = (i == 0 || (uint32_t)(A[i - 1] >> 32) != dpt) && (i + 1 < n) &&
                            (uint32_t)(A[i + 1] >> 32) == dpt;
        if (starts) {
            int e = i + 2;
            while (e < n && (uint32_t)(A[e] >> 32) == dpt) ++e;
            if (e - i > 64) {
                misc[3] = 1u;
            } else {
                for (int a = i + 1; a < e; ++a) {          // insertion sort of [i, e) on the full key
                    const unsigned long long v = A[a];
                    int b = a - 1;
                    while (b >= i && A[b] > v) { A[b + 1] = A[b]; --b; }
                    A[b + 1] = v;
                }
            }
        }
    }
    __syncthreads();
    if (misc[3]) {
        int n_pad = 2;
        while (n_pad < n) n_pad <<= 1;                      // <= cap (cap is a power of two)
        for (int i = n + tid; i < n_pad; i += kR11Threads) A[i] = ~0ull;
        __syncthreads();
        bitonic_sort_smem<kR11Threads>(A, n_pad);
    }
    for (int i = tid; i < n; i += kR11Threads) keys[s0 + i] = A[i];
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

    // In-shared-memory 11-bit radix sort for segments up to 8192 keys (power-of-two capacity picked
    // from the hint, with 25 % head-room); the 8-bit radix kernel below only does work for longer
    // segments (its CTAs exit at once otherwise).
    int bitonic_cap = 2048;
    if (segment_hint > 0) {
        const long long want = (long long)segment_hint + segment_hint / 4;
        while (bitonic_cap < want && bitonic_cap < 8192) bitonic_cap <<= 1;
    }
    static bool battr = false;
    if (!battr) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_tile_sort_radix11, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)radix11_smem_bytes(8192)));
        battr = true;
    }
    k_tile_sort_radix11<<<n_seg, kR11Threads, radix11_smem_bytes(bitonic_cap), st>>>(
        g.tile_start, g.tile_count, g.n_instances, d.capacity, keys, keys_alt, bitonic_cap, id_bits);
    PS_LAUNCH_CHECK("k_tile_sort_radix11");
    return PS_OK;
}

}  // namespace ps
