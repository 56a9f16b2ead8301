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
// Backward: per batch of queued hits
//   phase 1 (lane = pixel): walk the batch FRONT TO BACK carrying (T, S) with
//            S_i = sum_{j<=i} w_j (c_j . dL/dC),  w_j = alpha_j T_j,  and, from the forward's stored
//            pixel colour C,  Q = C . dL/dC = S_last + T_final (bg . dL/dC):
//            dL/dalpha_i = T_i (c_i . dL/dC) - (Q - S_i) / (1 - alpha_i)
//            (algebraically upstream's back-to-front recurrence).  Writes the two scalars every
//            gradient is built from, u = G dL/dalpha and w, to shared memory [entry][pixel];
//   phase 2 (lane = entry): each lane sums its entry's nine gradient moments over the block's pixels
//            straight out of shared memory -- no shuffle reduction, no shared-memory accumulators --
//            and adds them to the per-(view, Gaussian) scratch with three vector RED instructions.
#include <cstdlib>

#include "ps_common.cuh"

namespace ps {

namespace {

constexpr int kQ = 64;                  // hit-queue slots per warp (ring)
constexpr float kAlphaMin = 1.0f / 255.0f;
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

// Per-warp hit queue.  The conic is stored pre-multiplied so that
//   power * log2(e) = qa dx^2 + qc dy^2 + qb dx dy      (one MUFU.EX2 per evaluation).
struct HitQueue {
    float4 q0[kQ];   // x, y, qa, qb
    float4 q1[kQ];   // qc, opacity, r, g
    float4 q2[kQ];   // b, gaussian id (bits), list position (bits), unused
};

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
    float px, py, rx0, rx1, ry0, ry1;
    uint32_t start, count;
    size_t gbase, pix, hw;
};

__device__ __forceinline__ bool task_setup(const Dims &d, const Geom &geo, int task, int lane, TaskGeom &t) {
    const int sub = task & 7;
    const long long seg = task >> 3;
    if (seg >= (long long)d.S * d.V * d.tiles) return false;
    t.vid = (int)(seg / d.tiles);
    const int tile = (int)(seg - (long long)t.vid * d.tiles);
    const int tx = tile % d.gx, ty = tile / d.gx;
    const int wx0 = tx * kTile + (sub & 1) * 8, wy0 = ty * kTile + (sub >> 1) * 4;
    t.pxi = wx0 + (lane & 7);
    t.pyi = wy0 + (lane >> 3);
    t.inside = t.pxi < d.W && t.pyi < d.H;
    t.px = (float)t.pxi; t.py = (float)t.pyi;
    t.rx0 = (float)wx0; t.rx1 = (float)(wx0 + 7); t.ry0 = (float)wy0; t.ry1 = (float)(wy0 + 3);
    t.start = geo.tile_start[seg];
    t.count = geo.tile_count[seg];
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
        p.g = (uint32_t)keys[t.start + lane];
        p.cr = geo.cull[t.gbase + p.g];
    }
    if (32u + (uint32_t)lane < n) p.key_next = keys[t.start + 32u + lane];
}

// Parks the hit found in the previous iteration (its gathers have had a whole iteration to land).
__device__ __forceinline__ void cull_park(CullPipe &p, HitQueue &q) {
    if (p.h_pending) {
        q.q0[p.h_slot] = make_float4(p.h_xy.x, p.h_xy.y, -0.5f * kLog2e * p.h_co.x, -kLog2e * p.h_co.y);
        q.q1[p.h_slot] = make_float4(-0.5f * kLog2e * p.h_co.z, p.h_co.w, p.h_rgb.x, p.h_rgb.y);
        q.q2[p.h_slot] = make_float4(p.h_rgb.z, __uint_as_float(p.h_g), __uint_as_float(p.h_pos), 0.0f);
    }
    p.h_pending = false;
    __syncwarp();
}

// Tests chunk c (list positions 32c .. 32c+31) and advances the pipeline.  Returns the number of hits;
// they become readable in the queue after the NEXT cull_park.
__device__ __forceinline__ int cull_step(CullPipe &p, const Geom &geo, const TaskGeom &t,
                                         const unsigned long long *__restrict__ keys, uint32_t n, uint32_t c,
                                         uint32_t tail, int lane) {
    const uint32_t pos = c * 32u + (uint32_t)lane;
    const float4 cr = p.cr;
    const bool hit = pos < n && (cr.x + cr.z >= t.rx0) && (cr.x - cr.z <= t.rx1) && (cr.y + cr.w >= t.ry0) &&
                     (cr.y - cr.w <= t.ry1);
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
        p.h_pending = true;
        p.h_slot = (tail + (uint32_t)__popc(mask & ((1u << lane) - 1u))) & (kQ - 1);
        p.h_g = p.g;
        p.h_pos = pos;
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
    if (pos2 < n) p.key_next = keys[t.start + pos2];
    return __popc(mask);
}

}  // namespace

// ================================================================================== forward
constexpr int kFwdWarps = 2;

__global__ void __launch_bounds__(kFwdWarps * 32)
k_composite_fwd2(Dims d, Geom geo, const float *__restrict__ bg_all, const unsigned long long *__restrict__ keys,
                 ImageState img, float *__restrict__ out_color) {
    __shared__ HitQueue s_q[kFwdWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HitQueue &q = s_q[warp];
    TaskGeom t;
    if (!task_setup(d, geo, blockIdx.x * kFwdWarps + warp, lane, t)) return;
    const bool truncated = *geo.n_instances > d.capacity;
    const uint32_t n = truncated ? 0u : t.count;

    float T = 1.0f, Cr = 0.0f, Cg = 0.0f, Cb = 0.0f;
    uint32_t last = 0;
    bool done = !t.inside;

    CullPipe p;
    cull_prologue(p, geo, t, keys, n, lane);
    uint32_t head = 0, tail = 0, avail = 0;
    const uint32_t nchunks = (n + 31u) >> 5;
    for (uint32_t c = 0; c <= nchunks; ++c) {          // one extra iteration drains the last parked hits
        cull_park(p, q);
        avail = tail;
        if (c < nchunks) tail += (uint32_t)cull_step(p, geo, t, keys, n, c, tail, lane);
        // blend every parked entry, four per iteration (their power / exp evaluations are independent,
        // only the transmittance update chains)
        while (head < avail) {
            float pw[4], al[4];
            float3 col[4];
            uint32_t ps[4];
            bool has[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                has[k] = head + (uint32_t)k < avail;
                const uint32_t slot = (head + (has[k] ? (uint32_t)k : 0u)) & (kQ - 1);
                const float4 a0 = q.q0[slot], a1 = q.q1[slot], a2 = q.q2[slot];
                const float dx = a0.x - t.px, dy = a0.y - t.py;
                pw[k] = a0.z * dx * dx + a1.x * dy * dy + a0.w * dx * dy;
                al[k] = fminf(0.99f, a1.y * fast_exp2(pw[k]));
                col[k] = make_float3(a1.z, a1.w, a2.x);
                ps[k] = __float_as_uint(a2.z);
            }
            head += 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool contrib = has[k] && !done && !(pw[k] > 0.0f) && !(al[k] < kAlphaMin);
                const float test_T = T * (1.0f - al[k]);
                const bool stop = contrib && (test_T < 0.0001f);
                const bool blend = contrib && !stop;
                const float w = blend ? al[k] * T : 0.0f;
                Cr += col[k].x * w; Cg += col[k].y * w; Cb += col[k].z * w;
                T = blend ? test_T : T;
                last = blend ? ps[k] + 1u : last;
                done = done || stop;
            }
        }
        head = avail;
        if (__all_sync(0xffffffffu, done)) break;
        __syncwarp();
    }
    if (t.inside) {
        const size_t o1 = (size_t)t.vid * t.hw + t.pix;
        img.final_T[o1] = T;
        img.n_contrib[o1] = last;
        const float *bg = bg_all + 3 * t.vid;
        const float r = Cr + T * bg[0], g = Cg + T * bg[1], b = Cb + T * bg[2];
        const size_t o3 = (size_t)t.vid * 3 * t.hw + t.pix;
        out_color[o3] = r; out_color[o3 + t.hw] = g; out_color[o3 + 2 * t.hw] = b;
        img.color[o3] = r; img.color[o3 + t.hw] = g; img.color[o3 + 2 * t.hw] = b;
    }
}

// ================================================================================== backward
constexpr int kBwdWarps = 2;
constexpr int kBatch = 32;              // queued hits per batch (phase 2: one lane per entry)
static_assert(kBatch == 32 || kBatch == 16, "phase 2 maps lanes to (entry, pixel half)");
constexpr int kHalves = 32 / kBatch;    // lanes per entry in phase 2
constexpr int kPixPerLane = 32 / kHalves;

struct BwdSmem {
    HitQueue q;
    float su[kBatch][33];               // u = G dL/dalpha   [entry][pixel], +1 pad: conflict-free both ways
    float sw[kBatch][33];               // w = alpha T
    float4 dp[32];                      // dL/dC of the block's pixels (r, g, b, -)
};

struct BwdPixel {
    float px, py, dpr, dpg, dpb, Q, T, S;
    uint32_t last;
};

// Phase 1 + phase 2 for the queue entries [head, head + cnt), cnt <= kBatch (warp-uniform).
template <bool FULL>
__device__ __forceinline__ void bwd_batch(BwdSmem &sm, BwdPixel &px, const TaskGeom &t, const ViewGrads &vg,
                                          uint32_t head, int cnt, float kx, float ky, int lane) {
    // ---- phase 1: lane = pixel, entries front to back
#pragma unroll 4
    for (int j = 0; j < kBatch; ++j) {
        if (!FULL && j >= cnt) break;
        const uint32_t slot = (head + (uint32_t)j) & (kQ - 1);
        const float4 a0 = sm.q.q0[slot], a1 = sm.q.q1[slot], a2 = sm.q.q2[slot];
        const float dx = a0.x - px.px, dy = a0.y - px.py;
        const float p2 = a0.z * dx * dx + a1.x * dy * dy + a0.w * dx * dy;
        const float G = fast_exp2(p2);
        const float al = fminf(0.99f, a1.y * G);
        const bool active = __float_as_uint(a2.z) < px.last && !(p2 > 0.0f) && !(al < kAlphaMin);
        const float a = active ? al : 0.0f;
        const float Gs = active ? G : 0.0f;        // also keeps an overflowed exp2 out of 0 * inf
        const float cdp = a1.z * px.dpr + a1.w * px.dpg + a2.x * px.dpb;
        const float w = a * px.T;
        px.S = fmaf(w, cdp, px.S);
        const float om = 1.0f - a;
        const float dL = px.T * cdp - (px.Q - px.S) * fast_rcp(om);
        px.T *= om;
        sm.su[j][lane] = Gs * dL;
        sm.sw[j][lane] = w;
    }
    __syncwarp();
    // ---- phase 2: lane = (entry e, pixel half h); nine moments over the lane's pixels
    const int e = lane & (kBatch - 1), h = lane / kBatch;
    const uint32_t slot = (head + (uint32_t)e) & (kQ - 1);
    const float4 b0 = sm.q.q0[slot];
    float s_u = 0.0f, s_x = 0.0f, s_y = 0.0f, s_xx = 0.0f, s_xy = 0.0f, s_yy = 0.0f, s_r = 0.0f, s_g = 0.0f, s_b = 0.0f;
#pragma unroll
    for (int k = 0; k < kPixPerLane; ++k) {
        const int p = h * kPixPerLane + k;
        const float u = sm.su[e][p], w = sm.sw[e][p];
        const float4 dpp = sm.dp[p];
        // same pixel coordinates as phase 1: px = rx0 + (k & 7) exactly (small integers)
        const float dx = (b0.x - (t.rx0 + (float)(k & 7))), dy = (b0.y - (t.ry0 + (float)(h * (kPixPerLane / 8) + (k >> 3))));
        const float ux = u * dx, uy = u * dy;
        s_u += u; s_x += ux; s_y += uy;
        s_xx = fmaf(ux, dx, s_xx); s_xy = fmaf(ux, dy, s_xy); s_yy = fmaf(uy, dy, s_yy);
        s_r = fmaf(w, dpp.x, s_r); s_g = fmaf(w, dpp.y, s_g); s_b = fmaf(w, dpp.z, s_b);
    }
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
        const float4 b1 = sm.q.q1[slot];
        const float4 b2 = sm.q.q2[slot];
        const float o = b1.y;
        const size_t rec = t.gbase + __float_as_uint(b2.y);
        // u excludes the opacity factor: position / conic terms pick it up here, dL/dopacity does not
        const float ox = o * s_x, oy = o * s_y;
        red_add_v2(vg.d_mean2d + rec, kx * (2.0f * b0.z * ox + b0.w * oy), ky * (2.0f * b1.x * oy + b0.w * ox));
        red_add_v4(vg.d_conic + rec, -0.5f * o * s_xx, -0.5f * o * s_xy, -0.5f * o * s_yy, s_u);
        red_add_v4(vg.d_color + rec, s_r, s_g, s_b, 0.0f);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kBwdWarps * 32)
k_composite_bwd2(Dims d, Geom geo, const float *__restrict__ bg_all, const unsigned long long *__restrict__ keys,
                 ImageState img, const float *__restrict__ d_color, ViewGrads vg) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BwdSmem &sm = reinterpret_cast<BwdSmem *>(s_raw)[warp];
    if (*geo.n_instances > d.capacity) return;
    TaskGeom t;
    if (!task_setup(d, geo, blockIdx.x * kBwdWarps + warp, lane, t)) return;

    BwdPixel px;
    px.px = t.px; px.py = t.py;
    px.T = 1.0f; px.S = 0.0f;
    px.last = 0; px.dpr = px.dpg = px.dpb = 0.0f; px.Q = 0.0f;
    if (t.inside) {
        const size_t o1 = (size_t)t.vid * t.hw + t.pix, o3 = (size_t)t.vid * 3 * t.hw + t.pix;
        px.last = img.n_contrib[o1];
        px.dpr = d_color[o3]; px.dpg = d_color[o3 + t.hw]; px.dpb = d_color[o3 + 2 * t.hw];
        px.Q = img.color[o3] * px.dpr + img.color[o3 + t.hw] * px.dpg + img.color[o3 + 2 * t.hw] * px.dpb;
    }
    sm.dp[lane] = make_float4(px.dpr, px.dpg, px.dpb, 0.0f);
    const uint32_t n = min(t.count, __reduce_max_sync(0xffffffffu, px.last));   // nothing beyond the block's last contributor
    const float kx = kLn2 * 0.5f * (float)d.W, ky = kLn2 * 0.5f * (float)d.H;
    (void)bg_all;

    CullPipe p;
    cull_prologue(p, geo, t, keys, n, lane);
    uint32_t head = 0, tail = 0;
    const uint32_t nchunks = (n + 31u) >> 5;
    for (uint32_t c = 0; c <= nchunks; ++c) {
        cull_park(p, sm.q);
        const uint32_t avail = tail;
        if (c < nchunks) tail += (uint32_t)cull_step(p, geo, t, keys, n, c, tail, lane);
        while (avail - head >= (uint32_t)kBatch) {
            bwd_batch<true>(sm, px, t, vg, head, kBatch, kx, ky, lane);
            head += kBatch;
        }
    }
    if (tail != head) bwd_batch<false>(sm, px, t, vg, head, (int)(tail - head), kx, ky, lane);
}

// ================================================================================== launchers
int composite_impl() {
    static int impl = 0;
    if (impl == 0) {
        const char *e = getenv("PIXELSPLAT_B200_COMPOSITE");
        impl = (e && e[0] == '1') ? 1 : 2;
    }
    return impl;
}

int launch_composite_forward(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                             const ImageState &img, float *out_color, cudaStream_t st) {
    if (composite_impl() == 1) return launch_composite_forward_v1(d, in, g, keys, img, out_color, st);
    const long long tasks = (long long)d.S * d.V * d.tiles * 8;
    k_composite_fwd2<<<(unsigned)((tasks + kFwdWarps - 1) / kFwdWarps), kFwdWarps * 32, 0, st>>>(d, g, in.bg, keys, img, out_color);
    PS_LAUNCH_CHECK("k_composite_fwd2");
    return PS_OK;
}

int launch_composite_backward(const Dims &d, const Inputs &in, const Geom &g, const unsigned long long *keys,
                              const ImageState &img, const float *d_color, const ViewGrads &vg, cudaStream_t st) {
    if (composite_impl() == 1) return launch_composite_backward_v1(d, in, g, keys, img, d_color, vg, st);
    const long long tasks = (long long)d.S * d.V * d.tiles * 8;
    const size_t smem = sizeof(BwdSmem) * kBwdWarps;
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices)) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_composite_bwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    k_composite_bwd2<<<(unsigned)((tasks + kBwdWarps - 1) / kBwdWarps), kBwdWarps * 32, smem, st>>>(d, g, in.bg, keys, img, d_color, vg);
    PS_LAUNCH_CHECK("k_composite_bwd2");
    return PS_OK;
}

}  // namespace ps
