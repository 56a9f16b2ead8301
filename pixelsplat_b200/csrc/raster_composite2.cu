// Alpha compositing, forward and backward, as independent WARP TASKS (round-2 compositor).
//
// A task is one 8x4 pixel block of one (view, 16x16 tile); a warp owns a task from start to end and
// never meets another warp at a barrier.  Replaces the per-pixel blend loop of upstream's renderCUDA
// (the call the reference makes at
// /root/reference/src/model/decoder/cuda_splatting.py:113-124) without changing a per-pixel decision.
//
// Front end (shared by both directions): the warp streams the tile's depth-sorted instance list 32
// entries at a time -- key -> Gaussian id -> 16-byte cull record (screen position + half-extents of the
// box outside of which alpha < 1/255, written by k_preprocess) -- tests the box against its 8x4
// rectangle, and only for the hits gathers conic/opacity/colour into a per-warp shared-memory queue.
// The three dependent loads are software-pipelined across iterations (keys two chunks ahead, cull
// records one ahead, hit records parked one iteration later), so nothing waits on L2.
//
// Both kernels are bound by instruction issue (ncu: 65-80 % issue-active forward), so the per-hit loop is
// built to be short:
//   * a queued hit carries its log2-power as a POLYNOMIAL in the block-local pixel offsets (i, j in 0..7 x
//     0..3, lane = 8 j + i):  p2(i, j) = A + i (B + i qa + j qb) + j (C + j qc), five FFMAs per (hit, pixel)
//     with the coefficients formed once per hit by the lane that queues it.  Offsets are measured from the
//     block origin, so |A|, |B i|, ... stay O(10) for every hit that can contribute and the rounding error
//     of the expansion is ~1e-6 absolute in the exponent (the MUFU.EX2 that follows is no better).  Upstream's
//     `power > 0` rejection (unreachable for a positive-definite conic except through rounding) becomes
//     `p2 > 1e-4` so that this rounding cannot drop a pixel that sits on a Gaussian's centre;
//   * the queue is a ring consumed in ALIGNED GROUPS OF FOUR (zero-opacity records pad the last group):
//     one address computation per group, no per-entry bounds logic.
//
// Backward: per batch of queued hits
//   phase 1 (lane = pixel): walk the batch FRONT TO BACK carrying (T, S) with
//            S_i = sum_{j<=i} w_j (c_j . dL/dC),  w_j = alpha_j T_j,  and, from the forward's stored
//            pixel colour C,  Q = C . dL/dC = S_last + T_final (bg . dL/dC):
//            dL/dalpha_i = T_i (c_i . dL/dC) - (Q - S_i) / (1 - alpha_i)
//            (algebraically upstream's back-to-front recurrence).  Writes the two scalars every
//            gradient is built from, u = G dL/dalpha and w, to shared memory [entry][pixel];
//   phase 2 (lane = entry): each lane sums its entry's RAW moments sum_p u (1, i, j, i^2, i j, j^2) and
//            sum_p w dL/dC over the block's pixels straight out of shared memory (i, j are literals in the
//            unrolled loop: ~4 FFMAs per pixel), shifts them to the Gaussian's centre once, and adds the nine
//            gradients to the per-(view, Gaussian) scratch with three vector RED instructions.
//
// List SEGMENTS (d.segK = 1, 2 or 4 warps per task).  One 256x256 view is only 2048 tasks -- 14 warps per
// SM, each a long serial chain -- so when the batch is small a tile's sorted list is cut into segK runs of
// whole 32-entry chunks and each run gets its own warp:
//   forward : every warp composites its run as if nothing lay in front of it (T = 1), giving (C_k, T_k);
//             front-to-back composition is associative, (C, T) o (C_k, T_k) = (C + T C_k, T T_k), so one warp
//             folds the runs in order.  Upstream's early exit (stop once T (1 - alpha) < 1e-4) depends on
//             the true T, so a run is accepted only when T T_k stays clear of the threshold (T T_k is a lower
//             bound of every intermediate test); otherwise -- rare: the pixel saturates inside this run --
//             the run is replayed for those pixels with the true T, i.e. exactly the sequential loop.
//             The state in front of each run, (T, C), is kept per pixel for the backward.
//   backward: the forward-order prefix form needs only (T, S = C . dL/dC) in front of a run, which the
//             forward stored, so the runs are independent warps with no combination step at all.
#include <cstdlib>

#include "ps_common.cuh"

namespace ps {

namespace {

constexpr int kQ = 64;                  // hit-queue slots per warp (ring, consumed in aligned groups of 4)
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kPowerEps = 1.0e-4f;    // see the file header: rounding slack of the polynomial exponent
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_v2(float2 *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Per-warp hit queue (see the file header for the polynomial form).
struct HitQueue {
    float4 r0[kQ];   // A, B, C, opacity
    float4 r1[kQ];   // qa, qb, qc, list position (bits)
    float4 r2[kQ];   // r, g, b, Gaussian id (bits)
};

// log2 of the Gaussian falloff at block-local pixel (fi, fj)
__device__ __forceinline__ float hit_power2(const float4 &a0, const float4 &a1, float fi, float fj) {
    const float t1 = fmaf(fj, a1.y, fmaf(fi, a1.x, a0.y));     // B + i qa + j qb
    const float t2 = fmaf(fj, a1.z, a0.z);                     // C + j qc
    return fmaf(fj, t2, fmaf(fi, t1, a0.x));
}

// Registers of the cull pipeline (see file header).
struct CullPipe {
    unsigned long long key_next;   // keys of chunk c + 2
    float4 cr;                     // cull record of chunk c + 1 (this lane's entry)
    uint32_t g;                    // its Gaussian id
    // hit of chunk c waiting to be parked in the queue
    float4 h_co, h_rgb;
    float2 h_xy;
    uint32_t h_g, h_pos, h_slot;
    bool h_pending;
};

struct TaskGeom {
    int vid, pxi, pyi;
    bool inside;
    float fi, fj;                 // block-local pixel offsets of this lane (0..7, 0..3)
    float rx0, rx1, ry0, ry1;     // the block's pixel rectangle
    uint32_t start, count;        // the tile's list
    uint32_t run_begin, run_len;  // this warp's run of it (whole list when segK == 1)
    size_t gbase, pix, hw;
};

// Run k of segK over a list of `count` entries: whole 32-entry chunks, the same split in both directions.
__device__ __forceinline__ void run_bounds(uint32_t count, int segK, int k, uint32_t &begin, uint32_t &len) {
    const uint32_t chunks = (count + 31u) >> 5;
    const uint32_t per = (chunks + (uint32_t)segK - 1u) / (uint32_t)segK;
    begin = min(count, (uint32_t)k * per * 32u);
    len = min(count, (uint32_t)(k + 1) * per * 32u) - begin;
}

__device__ __forceinline__ bool task_setup(const Dims &d, const Geom &geo, long long task, int lane, TaskGeom &t) {
    const int sub = (int)(task & 7);
    const long long seg = task >> 3;
    if (seg >= (long long)d.S * d.V * d.tiles) return false;
    t.vid = (int)(seg / d.tiles);
    const int tile = (int)(seg - (long long)t.vid * d.tiles);
    const int tx = tile % d.gx, ty = tile / d.gx;
    const int wx0 = tx * kTile + (sub & 1) * 8, wy0 = ty * kTile + (sub >> 1) * 4;
    t.pxi = wx0 + (lane & 7);
    t.pyi = wy0 + (lane >> 3);
    t.inside = t.pxi < d.W && t.pyi < d.H;
    t.fi = (float)(lane & 7); t.fj = (float)(lane >> 3);
    t.rx0 = (float)wx0; t.rx1 = (float)(wx0 + 7); t.ry0 = (float)wy0; t.ry1 = (float)(wy0 + 3);
    t.start = geo.tile_start[seg];
    t.count = geo.tile_count[seg];
    t.run_begin = 0;
    t.run_len = t.count;
    t.gbase = (size_t)t.vid * d.P;
    t.hw = (size_t)d.H * d.W;
    t.pix = (size_t)t.pyi * d.W + t.pxi;
    return true;
}

// ---- cull pipeline ------------------------------------------------------------------------------
__device__ __forceinline__ void cull_prologue(CullPipe &p, const Geom &geo, const TaskGeom &t,
                                              const unsigned long long *__restrict__ keys, uint32_t n, int lane) {
    p.h_pending = false;
    p.h_slot = 0; p.h_g = 0; p.h_pos = 0;
    p.h_xy = make_float2(0.0f, 0.0f);
    p.h_co = p.h_rgb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    p.g = 0;
    p.cr = make_float4(0.0f, 0.0f, -3.0e38f, -3.0e38f);
    p.key_next = 0;
    if ((uint32_t)lane < n) {
        p.g = (uint32_t)keys[t.start + t.run_begin + lane];
        p.cr = geo.cull[t.gbase + p.g];
    }
    if (32u + (uint32_t)lane < n) p.key_next = keys[t.start + t.run_begin + 32u + lane];
}

// Parks the hit found in the previous iteration (its gathers have had a whole iteration to land): forms the
// polynomial coefficients of the hit about the block origin.  D0 = also keep (dx0, dy0) for the backward.
template <bool D0>
__device__ __forceinline__ void cull_park(CullPipe &p, HitQueue &q, float2 *d0, const TaskGeom &t) {
    if (p.h_pending) {
        const float qa = -0.5f * kLog2e * p.h_co.x, qb = -kLog2e * p.h_co.y, qc = -0.5f * kLog2e * p.h_co.z;
        const float dx0 = p.h_xy.x - t.rx0, dy0 = p.h_xy.y - t.ry0;
        const float ax = qa * dx0, cy = qc * dy0;
        const float A = fmaf(ax, dx0, fmaf(cy, dy0, qb * dx0 * dy0));
        const float B = -(2.0f * ax + qb * dy0);
        const float C = -(2.0f * cy + qb * dx0);
        q.r0[p.h_slot] = make_float4(A, B, C, p.h_co.w);
        q.r1[p.h_slot] = make_float4(qa, qb, qc, __uint_as_float(p.h_pos));
        q.r2[p.h_slot] = make_float4(p.h_rgb.x, p.h_rgb.y, p.h_rgb.z, __uint_as_float(p.h_g));
        if (D0) d0[p.h_slot] = make_float2(dx0, dy0);
    }
    p.h_pending = false;
    __syncwarp();
}

// Zero-opacity records up to the next multiple of four (they can never contribute: alpha = 0 < 1/255).
__device__ __forceinline__ uint32_t queue_pad(HitQueue &q, uint32_t tail, int lane) {
    const uint32_t padded = (tail + 3u) & ~3u;
    if ((uint32_t)lane < padded - tail) {
        const uint32_t slot = (tail + (uint32_t)lane) & (kQ - 1);
        q.r0[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        q.r1[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xffffffffu));   // position beyond any `last`
        q.r2[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncwarp();
    return padded;
}

// Tests chunk c (positions 32c .. 32c+31 of the warp's run, n = run length) and advances the pipeline.
// Returns the number of hits; they become readable in the queue after the NEXT cull_park.
__device__ __forceinline__ int cull_step(CullPipe &p, const Geom &geo, const TaskGeom &t,
                                         const unsigned long long *__restrict__ keys, uint32_t n, uint32_t c,
                                         uint32_t tail, int lane, uint2 *__restrict__ hit_out = nullptr) {
    const uint32_t pos = c * 32u + (uint32_t)lane;
    const float4 cr = p.cr;
    const bool hit = pos < n && (cr.x + cr.z >= t.rx0) && (cr.x - cr.z <= t.rx1) && (cr.y + cr.w >= t.ry0) &&
                     (cr.y - cr.w <= t.ry1);
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
        p.h_pending = true;
        const uint32_t idx = tail + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        p.h_slot = idx & (kQ - 1);
        p.h_g = p.g;
        p.h_pos = t.run_begin + pos;          // position in the TILE's list (n_contrib semantics)
        if (hit_out) hit_out[idx] = make_uint2(p.h_pos, p.g);     // the backward walks this list instead of culling
        p.h_xy = make_float2(cr.x, cr.y);
        p.h_co = geo.conic_opacity[t.gbase + p.g];
        p.h_rgb = geo.rgb[t.gbase + p.g];
    }
    // advance: cull record of chunk c+1 from the key loaded an iteration ago, key of chunk c+2
    const uint32_t pos1 = pos + 32u, pos2 = pos + 64u;
    p.cr = make_float4(0.0f, 0.0f, -3.0e38f, -3.0e38f);
    if (pos1 < n) {
        p.g = (uint32_t)p.key_next;
        p.cr = geo.cull[t.gbase + p.g];
    }
    if (pos2 < n) p.key_next = keys[t.start + t.run_begin + pos2];
    return __popc(mask);
}

}  // namespace

// ================================================================================== forward
constexpr int kFwdWarps = 4;
#ifndef PS_FWD_MIN_CTAS
#define PS_FWD_MIN_CTAS 6              // <= 80 registers: 24 resident warps per SM
#endif
constexpr float kStopT = 0.0001f;          // upstream: stop once T (1 - alpha) < 1e-4
constexpr float kStopGuard = 1.01e-4f;     // a run is folded without replay only if T T_k stays above this

struct FwdPixel {
    float T, Cr, Cg, Cb;
    uint32_t last;      // 1 + list position of the last blended entry (0 = none)
    bool done;          // no further blending for this lane (stopped, or outside the image / masked)
    bool stopped;       // the early-exit test fired
};

// Four queued hits (an aligned group) onto the lane's pixel, front to back.
__device__ __forceinline__ void fwd_blend4(const HitQueue &q, uint32_t base, const TaskGeom &t, float &T, float &Cr,
                                           float &Cg, float &Cb, uint32_t &last, bool &done, bool &stopped) {
    float pw[4], al[4];
    float3 col[4];
    uint32_t ps[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 a0 = q.r0[base + k], a1 = q.r1[base + k], a2 = q.r2[base + k];
        pw[k] = hit_power2(a0, a1, t.fi, t.fj);
        al[k] = fminf(0.99f, a0.w * fast_exp2(pw[k]));
        col[k] = make_float3(a2.x, a2.y, a2.z);
        ps[k] = __float_as_uint(a1.w);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool contrib = !done && !(pw[k] > kPowerEps) && !(al[k] < kAlphaMin);
        const float test_T = T * (1.0f - al[k]);
        const bool stop = contrib && (test_T < kStopT);
        const bool blend = contrib && !stop;
        const float w = blend ? al[k] * T : 0.0f;
        Cr = fmaf(col[k].x, w, Cr); Cg = fmaf(col[k].y, w, Cg); Cb = fmaf(col[k].z, w, Cb);
        T = blend ? test_T : T;
        last = blend ? ps[k] + 1u : last;
        done = done || stop;
        stopped = stopped || stop;
    }
}

// Front-to-back blend of the warp's run [t.run_begin, t.run_begin + t.run_len) onto the per-lane state px.
__device__ __forceinline__ uint32_t fwd_run(const Geom &geo, const TaskGeom &t, const unsigned long long *__restrict__ keys,
                                            HitQueue &q, FwdPixel &px, int lane, uint2 *__restrict__ hit_out = nullptr) {
    const uint32_t n = t.run_len;
    float T = px.T, Cr = px.Cr, Cg = px.Cg, Cb = px.Cb;
    uint32_t last = px.last;
    bool done = px.done, stopped = px.stopped;
    CullPipe p;
    cull_prologue(p, geo, t, keys, n, lane);
    uint32_t head = 0, tail = 0;
    const uint32_t nchunks = (n + 31u) >> 5;
    for (uint32_t c = 0; c <= nchunks; ++c) {          // one extra iteration drains the last parked hits
        cull_park<false>(p, q, nullptr, t);
        uint32_t avail = tail;                         // parked so far
        if (c < nchunks) tail += (uint32_t)cull_step(p, geo, t, keys, n, c, tail, lane, hit_out);
        else avail = queue_pad(q, tail, lane);
        // whole groups of four (their power / exp evaluations are independent, only the transmittance chains)
        while (avail - head >= 4u) {
            fwd_blend4(q, head & (kQ - 1), t, T, Cr, Cg, Cb, last, done, stopped);
            head += 4u;
        }
        if (__all_sync(0xffffffffu, done)) break;     // (hits beyond this point are behind every pixel's last contributor)
        __syncwarp();
    }
    __syncwarp();
    px.T = T; px.Cr = Cr; px.Cg = Cg; px.Cb = Cb; px.last = last; px.done = done; px.stopped = stopped;
    return tail;
}

template <int K>
__global__ void __launch_bounds__(kFwdWarps * 32, PS_FWD_MIN_CTAS)
k_composite_fwd2(Dims d, Geom geo, const float *__restrict__ bg_all, const unsigned long long *__restrict__ keys,
                 ImageState img, float *__restrict__ out_color, LossEpilogue loss, HitLists hl) {
    __shared__ HitQueue s_q[kFwdWarps];
    __shared__ float4 s_ct[K > 1 ? kFwdWarps : 1][32];      // a run's (Cr, Cg, Cb, T)
    __shared__ uint32_t s_last[K > 1 ? kFwdWarps : 1][32];   // its last contributor | stopped << 31
    constexpr int kTasksPerCta = kFwdWarps / K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = warp % K;
    HitQueue &q = s_q[warp];
    TaskGeom t;
    const bool valid = task_setup(d, geo, (long long)blockIdx.x * kTasksPerCta + warp / K, lane, t);
    const bool truncated = *geo.n_instances > d.capacity;
    const uint32_t count = (valid && !truncated) ? t.count : 0u;
    t.count = count;
    t.run_len = count;
    if (K > 1) run_bounds(count, K, run, t.run_begin, t.run_len);

    FwdPixel px;
    px.T = 1.0f; px.Cr = px.Cg = px.Cb = 0.0f;
    px.last = 0; px.stopped = false;
    px.done = !valid || !t.inside;
    if (valid) {
        const long long task = (long long)blockIdx.x * kTasksPerCta + warp / K;
        uint2 *hit_out = nullptr;
        if (hl.hits)   // this run's slice of the block's region (a run has at most run_len hits)
            hit_out = hl.hits + ((size_t)t.start * 8 + (size_t)(task & 7) * t.count + t.run_begin);
        const uint32_t nh = fwd_run(geo, t, keys, q, px, lane, hit_out);
        if (hl.run_hits && lane == 0) hl.run_hits[task * kMaxSegments + run] = nh;
    }

    if (K > 1) {
        s_ct[warp][lane] = make_float4(px.Cr, px.Cg, px.Cb, px.T);
        s_last[warp][lane] = px.last | (px.stopped ? 0x80000000u : 0u);
        __syncthreads();
        if (run != 0 || !valid) return;
        // fold the runs in list order; px is run 0's result, i.e. the exact sequential state after run 0
        for (int j = 1; j < K; ++j) {
            if (t.inside)   // state in front of run j: what the backward's run j starts from
                img.run_state[((size_t)t.vid * (kMaxSegments - 1) + (j - 1)) * t.hw + t.pix] =
                    make_float4(px.T, px.Cr, px.Cg, px.Cb);
            const float4 r = s_ct[warp + j][lane];
            const uint32_t rl = s_last[warp + j][lane];
            const bool replay = !px.done && ((rl >> 31) != 0u || px.T * r.w < kStopGuard);
            if (__any_sync(0xffffffffu, replay)) {
                // the pixel saturates inside (or near) run j: replay it with the true transmittance
                TaskGeom tj = t;
                run_bounds(count, K, j, tj.run_begin, tj.run_len);
                FwdPixel pj = px;
                pj.done = px.done || !replay;
                fwd_run(geo, tj, keys, q, pj, lane);
                if (replay) px = pj;
            }
            if (!replay && !px.done) {
                px.Cr = fmaf(px.T, r.x, px.Cr); px.Cg = fmaf(px.T, r.y, px.Cg); px.Cb = fmaf(px.T, r.z, px.Cb);
                px.T *= r.w;
                const uint32_t r_last = rl & 0x7fffffffu;
                px.last = r_last ? r_last : px.last;
            }
        }
    } else if (!valid) {
        return;
    }
    float sse = 0.0f, sse_clip = 0.0f;
    if (t.inside) {
        const size_t o1 = (size_t)t.vid * t.hw + t.pix;
        img.final_T[o1] = px.T;
        img.n_contrib[o1] = px.last;
        const float *bg = bg_all + 3 * t.vid;
        const float r = px.Cr + px.T * bg[0], g = px.Cg + px.T * bg[1], b = px.Cb + px.T * bg[2];
        const size_t o3 = (size_t)t.vid * 3 * t.hw + t.pix;
        if (out_color) { out_color[o3] = r; out_color[o3 + t.hw] = g; out_color[o3 + 2 * t.hw] = b; }
        img.color[o3] = r; img.color[o3 + t.hw] = g; img.color[o3 + 2 * t.hw] = b;
        if (loss.target) {
            // loss epilogue (loss_mse.py:30-31, metrics.py:11-19): squared error of this pixel, raw and clipped
            const float tr = loss.target[o3], tg = loss.target[o3 + t.hw], tb = loss.target[o3 + 2 * t.hw];
            const float er = r - tr, eg = g - tg, eb = b - tb;
            sse = er * er + eg * eg + eb * eb;
            const float cr = __saturatef(r) - __saturatef(tr), cg = __saturatef(g) - __saturatef(tg),
                        cb = __saturatef(b) - __saturatef(tb);
            sse_clip = cr * cr + cg * cg + cb * cb;
        }
    }
    if (loss.target) {
        // one pair of atomics per warp task, spread over kLossSlots addresses per view
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sse += __shfl_xor_sync(0xffffffffu, sse, o);
            sse_clip += __shfl_xor_sync(0xffffffffu, sse_clip, o);
        }
        if (lane == 0) {
            const int slot = (int)((blockIdx.x * kTasksPerCta + warp / K) & (kLossSlots - 1));
            float *dst = loss.sums + ((size_t)t.vid * 2) * kLossSlots + slot;
            atomicAdd(dst, sse);
            atomicAdd(dst + kLossSlots, sse_clip);
        }
    }
}

// ================================================================================== backward
constexpr int kBwdWarps = 4;
constexpr int kBatch = 16;              // queued hits per batch (phase 2: lanes = entry x pixel half)
static_assert(kBatch == 32 || kBatch == 16, "phase 2 maps lanes to (entry, pixel half)");
constexpr int kHalves = 32 / kBatch;    // lanes per entry in phase 2
constexpr int kRowsPerLane = 4 / kHalves;   // pixel rows (of 8) each phase-2 lane sums

struct BwdSmem {
    HitQueue q;
    float2 d0[kQ];                      // (dx0, dy0): the hit's centre relative to the block origin
    float su[kBatch][33];               // u = G dL/dalpha   [entry][pixel], +1 pad: conflict-free both ways
    float sw[kBatch][33];               // w = alpha T
    float4 dp[32];                      // dL/dC of the block's pixels (r, g, b, -)
};

struct BwdPixel {
    float dpr, dpg, dpb, Q, T, S;
    uint32_t last;
};

// Phase 1 + phase 2 for the queue entries [head, head + cnt), head a multiple of kBatch, cnt <= kBatch
// (warp-uniform; entries up to the next multiple of four exist as zero-opacity padding).
template <bool FULL>
__device__ __forceinline__ void bwd_batch(BwdSmem &sm, BwdPixel &px, const TaskGeom &t, const ViewGrads &vg,
                                          uint32_t head, int cnt, float kx, float ky, int lane) {
    // ---- phase 1: lane = pixel, entries front to back
    const uint32_t base = head & (kQ - 1);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        if (!FULL && (j & 3) == 0 && j >= cnt) break;
        const float4 a0 = sm.q.r0[base + j], a1 = sm.q.r1[base + j], a2 = sm.q.r2[base + j];
        const float p2 = hit_power2(a0, a1, t.fi, t.fj);
        const float G = fast_exp2(p2);
        const float al = fminf(0.99f, a0.w * G);
        const bool active = __float_as_uint(a1.w) < px.last && !(p2 > kPowerEps) && !(al < kAlphaMin);
        const float a = active ? al : 0.0f;
        const float Gs = active ? G : 0.0f;        // also keeps an overflowed exp2 out of 0 * inf
        const float cdp = fmaf(a2.z, px.dpb, fmaf(a2.y, px.dpg, a2.x * px.dpr));
        const float w = a * px.T;
        px.S = fmaf(w, cdp, px.S);
        const float om = 1.0f - a;
        const float dL = fmaf(px.T, cdp, -(px.Q - px.S) * fast_rcp(om));
        px.T *= om;
        sm.su[j][lane] = Gs * dL;
        sm.sw[j][lane] = w;
    }
    __syncwarp();
    // ---- phase 2: lane = (entry e, pixel half h); raw moments over the lane's rows, i / j literals
    const int e = lane & (kBatch - 1), h = lane / kBatch;
    float m00 = 0.0f, m10 = 0.0f, m01 = 0.0f, m20 = 0.0f, m11 = 0.0f, m02 = 0.0f, s_r = 0.0f, s_g = 0.0f, s_b = 0.0f;
#pragma unroll
    for (int jr = 0; jr < kRowsPerLane; ++jr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = (h * kRowsPerLane + jr) * 8 + i;
            const float u = sm.su[e][p], w = sm.sw[e][p];
            const float4 dpp = sm.dp[p];
            m00 += u;
            if (i) { m10 = fmaf(u, (float)i, m10); m20 = fmaf(u, (float)(i * i), m20); }
            if (jr) { m01 = fmaf(u, (float)jr, m01); m02 = fmaf(u, (float)(jr * jr), m02); }
            if (i && jr) m11 = fmaf(u, (float)(i * jr), m11);
            s_r = fmaf(w, dpp.x, s_r); s_g = fmaf(w, dpp.y, s_g); s_b = fmaf(w, dpp.z, s_b);
        }
    }
    // shift to the Gaussian's centre: dx = ca - i, dy = cb - jr, (ca, cb) = centre relative to this lane's first row
    const uint32_t slot = base + (uint32_t)e;
    const float2 c0 = sm.d0[slot];
    const float ca = c0.x, cb = c0.y - (float)(h * kRowsPerLane);
    float s_u = m00;
    float s_x = fmaf(ca, m00, -m10);
    float s_y = fmaf(cb, m00, -m01);
    float s_xx = fmaf(ca, fmaf(ca, m00, -2.0f * m10), m20);
    float s_xy = fmaf(ca, fmaf(cb, m00, -m01), fmaf(-cb, m10, m11));
    float s_yy = fmaf(cb, fmaf(cb, m00, -2.0f * m01), m02);
    if (kHalves == 2) {
        s_u += __shfl_xor_sync(0xffffffffu, s_u, 16); s_x += __shfl_xor_sync(0xffffffffu, s_x, 16);
        s_y += __shfl_xor_sync(0xffffffffu, s_y, 16); s_xx += __shfl_xor_sync(0xffffffffu, s_xx, 16);
        s_xy += __shfl_xor_sync(0xffffffffu, s_xy, 16); s_yy += __shfl_xor_sync(0xffffffffu, s_yy, 16);
        s_r += __shfl_xor_sync(0xffffffffu, s_r, 16); s_g += __shfl_xor_sync(0xffffffffu, s_g, 16);
        s_b += __shfl_xor_sync(0xffffffffu, s_b, 16);
    }
    const bool any = (s_u != 0.0f) | (s_x != 0.0f) | (s_y != 0.0f) | (s_xx != 0.0f) | (s_xy != 0.0f) |
                     (s_yy != 0.0f) | (s_r != 0.0f) | (s_g != 0.0f) | (s_b != 0.0f);
    if (h == 0 && e < cnt && any) {
        const float4 b0 = sm.q.r0[slot], b1 = sm.q.r1[slot], b2 = sm.q.r2[slot];
        const float o = b0.w;
        const size_t rec = t.gbase + __float_as_uint(b2.w);
        // u excludes the opacity factor: position / conic terms pick it up here, dL/dopacity does not
        const float ox = o * s_x, oy = o * s_y;
        red_add_v2(vg.d_mean2d + rec, kx * (2.0f * b1.x * ox + b1.y * oy), ky * (2.0f * b1.z * oy + b1.y * ox));
        red_add_v4(vg.d_conic + rec, -0.5f * o * s_xx, -0.5f * o * s_xy, -0.5f * o * s_yy, s_u);
        red_add_v4(vg.d_color + rec, s_r, s_g, s_b, 0.0f);
    }
    __syncwarp();
}

template <int K>
__global__ void __launch_bounds__(kBwdWarps * 32, 6)
k_composite_bwd2(Dims d, Geom geo, const float *__restrict__ bg_all, const unsigned long long *__restrict__ keys,
                 ImageState img, const float *__restrict__ d_color, ViewGrads vg, LossEpilogue loss, HitLists hl) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int kTasksPerCta = kBwdWarps / K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = warp % K;
    BwdSmem &sm = reinterpret_cast<BwdSmem *>(s_raw)[warp];
    if (*geo.n_instances > d.capacity) return;
    TaskGeom t;
    if (!task_setup(d, geo, (long long)blockIdx.x * kTasksPerCta + warp / K, lane, t)) return;

    BwdPixel px;
    px.T = 1.0f; px.S = 0.0f;
    px.last = 0; px.dpr = px.dpg = px.dpb = 0.0f; px.Q = 0.0f;
    if (t.inside) {
        const size_t o1 = (size_t)t.vid * t.hw + t.pix, o3 = (size_t)t.vid * 3 * t.hw + t.pix;
        px.last = img.n_contrib[o1];
        const float c_r = img.color[o3], c_g = img.color[o3 + t.hw], c_b = img.color[o3 + 2 * t.hw];
        if (d_color) {
            px.dpr = d_color[o3]; px.dpg = d_color[o3 + t.hw]; px.dpb = d_color[o3 + 2 * t.hw];
        } else {
            // fused loss: dL/dC = scale[view] * (C - target), never materialised as a tensor
            const float sc = loss.grad_scale[t.vid];
            px.dpr = sc * (c_r - loss.target[o3]); px.dpg = sc * (c_g - loss.target[o3 + t.hw]);
            px.dpb = sc * (c_b - loss.target[o3 + 2 * t.hw]);
        }
        px.Q = c_r * px.dpr + c_g * px.dpg + c_b * px.dpb;
    }
    sm.dp[lane] = make_float4(px.dpr, px.dpg, px.dpb, 0.0f);
    // nothing beyond the block's last contributor; the run split is the forward's (on the full count)
    const uint32_t nmax = min(t.count, __reduce_max_sync(0xffffffffu, px.last));
    if (K > 1) {
        run_bounds(t.count, K, run, t.run_begin, t.run_len);
        if (run > 0 && t.inside && t.run_begin < nmax) {
            // (T, C) in front of this run, stored by the forward: S = sum_{j < run} w_j (c_j . dL/dC) = C . dL/dC
            const float4 st = img.run_state[((size_t)t.vid * (kMaxSegments - 1) + (run - 1)) * t.hw + t.pix];
            px.T = st.x;
            px.S = st.y * px.dpr + st.z * px.dpg + st.w * px.dpb;
        }
    }
    const uint32_t run_end = min(t.run_begin + t.run_len, nmax);
    const uint32_t n = run_end > t.run_begin ? run_end - t.run_begin : 0u;
    t.run_len = n;
    const float kx = kLn2 * 0.5f * (float)d.W, ky = kLn2 * 0.5f * (float)d.H;
    (void)bg_all;

    uint32_t head = 0, tail = 0;
    if (hl.hits) {
        // ---- the forward left this run's hit list: no cull, every lane parks a hit
        const long long task = (long long)blockIdx.x * kTasksPerCta + warp / K;
        const uint2 *__restrict__ hits = hl.hits + ((size_t)t.start * 8 + (size_t)(task & 7) * t.count + t.run_begin);
        const uint32_t nh = n ? hl.run_hits[task * kMaxSegments + run] : 0u;
        uint2 hnext = make_uint2(0u, 0u);
        if ((uint32_t)lane < nh) hnext = hits[lane];
        for (uint32_t h0 = 0; h0 < nh; h0 += 32u) {
            const uint2 hcur = hnext;
            const bool live = h0 + (uint32_t)lane < nh && hcur.x < nmax;     // nothing behind the last contributor
            if (h0 + 32u + (uint32_t)lane < nh) hnext = hits[h0 + 32u + lane];
            float4 cr = make_float4(0.0f, 0.0f, 0.0f, 0.0f), co = cr, rgb = cr;
            if (live) {
                cr = geo.cull[t.gbase + hcur.y];
                co = geo.conic_opacity[t.gbase + hcur.y];
                rgb = geo.rgb[t.gbase + hcur.y];
            }
            const uint32_t m = __ballot_sync(0xffffffffu, live);
            if (live) {
                const uint32_t slot = (tail + (uint32_t)__popc(m & ((1u << lane) - 1u))) & (kQ - 1);
                const float qa = -0.5f * kLog2e * co.x, qb = -kLog2e * co.y, qc = -0.5f * kLog2e * co.z;
                const float dx0 = cr.x - t.rx0, dy0 = cr.y - t.ry0;
                const float ax = qa * dx0, cy = qc * dy0;
                sm.q.r0[slot] = make_float4(fmaf(ax, dx0, fmaf(cy, dy0, qb * dx0 * dy0)), -(2.0f * ax + qb * dy0),
                                            -(2.0f * cy + qb * dx0), co.w);
                sm.q.r1[slot] = make_float4(qa, qb, qc, __uint_as_float(hcur.x));
                sm.q.r2[slot] = make_float4(rgb.x, rgb.y, rgb.z, __uint_as_float(hcur.y));
                sm.d0[slot] = make_float2(dx0, dy0);
            }
            tail += (uint32_t)__popc(m);
            __syncwarp();
            while (tail - head >= (uint32_t)kBatch) {
                bwd_batch<true>(sm, px, t, vg, head, kBatch, kx, ky, lane);
                head += kBatch;
            }
            if (m != 0xffffffffu) break;                                      // the list is ordered by position
        }
    } else {
        CullPipe p;
        cull_prologue(p, geo, t, keys, n, lane);
        const uint32_t nchunks = (n + 31u) >> 5;
        for (uint32_t c = 0; c <= nchunks; ++c) {
            cull_park<true>(p, sm.q, sm.d0, t);
            const uint32_t avail = tail;
            if (c < nchunks) tail += (uint32_t)cull_step(p, geo, t, keys, n, c, tail, lane);
            while (avail - head >= (uint32_t)kBatch) {
                bwd_batch<true>(sm, px, t, vg, head, kBatch, kx, ky, lane);
                head += kBatch;
            }
        }
    }
    if (tail != head) {
        queue_pad(sm.q, tail, lane);
        bwd_batch<false>(sm, px, t, vg, head, (int)(tail - head), kx, ky, lane);
    }
}

// ================================================================================== launchers
// Tunables (A/B measurements, tests): initialised from the environment once, changeable through ps_set_option.
static int g_impl = 0;        // 1 = legacy CTA-per-tile compositor, 2 = warp tasks
static int g_segments = -1;   // 0 = automatic, else 1 | 2 | 4 list runs per task

int composite_impl() {
    if (g_impl == 0) {
        const char *e = getenv("PIXELSPLAT_B200_COMPOSITE");
        g_impl = (e && e[0] == '1') ? 1 : 2;
    }
    return g_impl;
}

int set_composite_option(int which, int value) {
    if (which == 0 && (value == 1 || value == 2)) { g_impl = value; return PS_OK; }
    if (which == 1 && (value == 0 || value == 1 || value == 2 || value == 4)) { g_segments = value; return PS_OK; }
    return PS_ERR_INVALID_ARGUMENT;
}

template <int K>
static int launch_fwd(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                      const ImageState &img, float *out_color, const LossEpilogue &loss, const HitLists &hl,
                      cudaStream_t st) {
    const long long tasks = (long long)d.S * d.V * d.tiles * 8;
    constexpr int per_cta = kFwdWarps / K;
    k_composite_fwd2<K><<<(unsigned)((tasks + per_cta - 1) / per_cta), kFwdWarps * 32, 0, st>>>(d, g, in.bg, keys, img, out_color, loss, hl);
    PS_LAUNCH_CHECK("k_composite_fwd2");
    return PS_OK;
}

int launch_composite_forward(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                             const ImageState &img, float *out_color, const LossEpilogue &loss, const HitLists &hl,
                             cudaStream_t st) {
    if (composite_impl() == 1) {
        if (loss.target || !out_color) { set_error("the legacy compositor has no loss epilogue"); return PS_ERR_UNSUPPORTED; }
        return launch_composite_forward_v1(d, in, g, keys, img, out_color, st);
    }
    switch (d.segK) {
        case 4: return launch_fwd<4>(d, in, g, keys, img, out_color, loss, hl, st);
        case 2: return launch_fwd<2>(d, in, g, keys, img, out_color, loss, hl, st);
        default: return launch_fwd<1>(d, in, g, keys, img, out_color, loss, hl, st);
    }
}

template <int K>
static int launch_bwd(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                      const ImageState &img, const float *d_color, const ViewGrads &vg, const LossEpilogue &loss,
                      const HitLists &hl, cudaStream_t st) {
    const long long tasks = (long long)d.S * d.V * d.tiles * 8;
    constexpr int per_cta = kBwdWarps / K;
    const size_t smem = sizeof(BwdSmem) * kBwdWarps;
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices)) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_composite_bwd2<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    k_composite_bwd2<K><<<(unsigned)((tasks + per_cta - 1) / per_cta), kBwdWarps * 32, smem, st>>>(d, g, in.bg, keys, img, d_color, vg, loss, hl);
    PS_LAUNCH_CHECK("k_composite_bwd2");
    return PS_OK;
}

int launch_composite_backward(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                              const ImageState &img, const float *d_color, const ViewGrads &vg,
                              const LossEpilogue &loss, const HitLists &hl, cudaStream_t st) {
    if (composite_impl() == 1) {
        if (!d_color) { set_error("the legacy compositor has no loss epilogue"); return PS_ERR_UNSUPPORTED; }
        return launch_composite_backward_v1(d, in, g, keys, img, d_color, vg, st);
    }
    switch (d.segK) {
        case 4: return launch_bwd<4>(d, in, g, keys, img, d_color, vg, loss, hl, st);
        case 2: return launch_bwd<2>(d, in, g, keys, img, d_color, vg, loss, hl, st);
        default: return launch_bwd<1>(d, in, g, keys, img, d_color, vg, loss, hl, st);
    }
}

// Runs per task for a batch of `tasks` warp tasks: enough warps to fill the machine (148 SMs x ~24 resident
// warps), none when the batch already does.  PIXELSPLAT_B200_SEGMENTS = 1 | 2 | 4 overrides (A/B runs).
int composite_segments(long long tasks) {
    if (g_segments < 0) {
        const char *e = getenv("PIXELSPLAT_B200_SEGMENTS");
        g_segments = (e && (e[0] == '1' || e[0] == '2' || e[0] == '4') && e[1] == 0) ? e[0] - '0' : 0;
    }
    if (composite_impl() == 1) return 1;
    if (g_segments) return g_segments;
    return tasks <= 2048 ? 4 : tasks <= 4096 ? 2 : 1;
}

// The forward's hit lists cost 64 bytes of binning state per unit of instance capacity: kept while that is at most
// 512 MB (every configuration of BASELINE.json at batch 1-2), dropped beyond (the backward then culls for itself).
// PIXELSPLAT_B200_HIT_LISTS = 0 | 1 forces it (A/B runs); the legacy compositor never uses them.
bool composite_hit_lists(long long capacity) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("PIXELSPLAT_B200_HIT_LISTS");
        forced = (e && (e[0] == '0' || e[0] == '1') && e[1] == 0) ? (e[0] - '0') : 2;
    }
    if (composite_impl() == 1) return false;
    if (forced != 2) return forced == 1;
    return capacity * 64 <= (512ll << 20);
}

}  // namespace ps
