// Alpha compositing, forward (front-to-back) and backward (back-to-front).
// One CTA per (view, 16x16 tile); each warp owns an 8x4 pixel sub-rectangle, each lane a pixel.
// The tile's sorted instance list is staged 256 entries at a time in shared memory.
//
// B200-first differences from upstream's renderCUDA, none of which
// change a per-pixel decision:
//   * warp-level culling: for every staged Gaussian one lane tests the axis-aligned bound of
//     the region where alpha >= 1/255 can hold (|d| <= sqrt(2 ln(255 o) Sigma_ii)) against the
//     warp's 8x4 rectangle; a ballot keeps only Gaussians that can touch the warp, so most
//     (pixel, Gaussian) pairs of the 3-sigma-square binning are never evaluated;
//   * backward: gradients are reduced across the warp with shuffles, combined across the
//     CTA's warps in shared memory, and flushed with ONE global atomic per value per
//     (tile, Gaussian) instead of one per (pixel, Gaussian).
#include "ps_common.cuh"

namespace ps {

constexpr int kCompThreads = 256;
constexpr int kStage = 256;
constexpr float kAlphaMin = 1.0f / 255.0f;

// Half-extents (in pixels) of the axis-aligned box outside of which this Gaussian's alpha is
// certainly < 1/255 (conservative).  Returns a negative x to mean "never contributes".
__device__ __forceinline__ float2 alpha_extent(const float4 co) {
    const float A = co.x, B = co.y, C = co.z, o = co.w;
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f) || !(o <= 3.0e38f)) return make_float2(3.0e38f, 3.0e38f);
    if (!(o * 255.0f >= 1.0f - 1e-3f)) return make_float2(-1.0f, -1.0f);  // also catches NaN/neg
    const float tau = __logf(fmaxf(o * 255.0f, 1.0f)) + 0.01f;                // q = -power <= tau
    const float inv = 2.0f * tau / det;
    return make_float2(sqrtf(inv * C) * 1.001f + 0.01f, sqrtf(inv * A) * 1.001f + 0.01f);
}

// Shared-memory staging is SoA so that both the lane-varying cull reads (8 B stride) and the
// broadcast reads of the evaluation loop are bank-conflict free.
struct StageBuf {
    float2 xy[kStage];
    float2 ext[kStage];
    float4 co[kStage];
    float4 rgb[kStage];
};

__device__ __forceinline__ void stage_entry(StageBuf &s, int slot, const Geom &geo, size_t base, uint32_t g) {
    const float4 co = geo.conic_opacity[base + g];
    s.xy[slot] = geo.xy[base + g];
    s.co[slot] = co;
    s.rgb[slot] = geo.rgb[base + g];
    s.ext[slot] = alpha_extent(co);
}

__global__ void __launch_bounds__(kCompThreads)
k_composite_fwd(Dims d, Geom geo, const float *__restrict__ bg_all,
                const unsigned long long *__restrict__ keys, float *__restrict__ final_T,
                uint32_t *__restrict__ n_contrib, float *__restrict__ out_color) {
    __shared__ StageBuf s;
    const int vid = blockIdx.y, tile = blockIdx.x;
    const int tx = tile % d.gx, ty = tile / d.gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wx0 = tx * kTile + (warp & 1) * 8, wy0 = ty * kTile + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < d.W && pyi < d.H;
    const float px = (float)pxi, py = (float)pyi;
    const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
    const bool truncated = *geo.n_instances > d.capacity;
    const size_t seg = (size_t)vid * d.tiles + tile;
    const uint32_t start = geo.tile_start[seg];
    const uint32_t count = truncated ? 0u : geo.tile_count[seg];
    const size_t gbase = (size_t)vid * d.P;

    float T = 1.0f, Cr = 0.0f, Cg = 0.0f, Cb = 0.0f;
    uint32_t last = 0;
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);

    for (uint32_t round0 = 0; round0 < count; round0 += kStage) {
        if (__syncthreads_count(warp_done ? 1 : 0) == kCompThreads) break;
        const uint32_t n_here = min((uint32_t)kStage, count - round0);
        if ((uint32_t)tid < n_here) stage_entry(s, tid, geo, gbase, (uint32_t)keys[start + round0 + tid]);
        __syncthreads();
        if (!warp_done) {
            for (uint32_t jb = 0; jb < n_here; jb += 32) {
                const uint32_t j = jb + lane;
                bool hit = false;
                if (j < n_here) {
                    const float2 c = s.xy[j], e = s.ext[j];
                    hit = (c.x + e.x >= rx0) && (c.x - e.x <= rx1) && (c.y + e.y >= ry0) && (c.y - e.y <= ry1);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int bpos = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (done) continue;
                    const float2 exy = s.xy[jb + bpos];
                    const float4 eco = s.co[jb + bpos];
                    const float dx = exy.x - px, dy = exy.y - py;
                    const float power = -0.5f * (eco.x * dx * dx + eco.z * dy * dy) - eco.y * dx * dy;
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, eco.w * __expf(power));
                    if (alpha < kAlphaMin) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float w = alpha * T;
                    const float4 ergb = s.rgb[jb + bpos];
                    Cr += ergb.x * w; Cg += ergb.y * w; Cb += ergb.z * w;
                    T = test_T;
                    last = round0 + jb + (uint32_t)bpos + 1u;
                }
                warp_done = __all_sync(0xffffffffu, done);
                if (warp_done) break;
            }
        }
        __syncthreads();
    }
    if (inside) {
        const size_t pix = (size_t)pyi * d.W + pxi;
        const size_t hw = (size_t)d.H * d.W;
        final_T[(size_t)vid * hw + pix] = T;
        n_contrib[(size_t)vid * hw + pix] = last;
        const float *bg = bg_all + 3 * vid;
        float *o = out_color + (size_t)vid * 3 * hw;
        o[pix] = Cr + T * bg[0];
        o[hw + pix] = Cg + T * bg[1];
        o[2 * hw + pix] = Cb + T * bg[2];
    }
}

// ---------------------------------------------------------------------------- backward

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kCompThreads)
k_composite_bwd(Dims d, Geom geo, const float *__restrict__ bg_all,
                const unsigned long long *__restrict__ keys, const float *__restrict__ final_T,
                const uint32_t *__restrict__ n_contrib, const float *__restrict__ d_color,
                ViewGrads vg) {
    __shared__ StageBuf s;
    __shared__ uint32_t s_g[kStage];
    __shared__ float s_acc[kStage][9];  // mean2d.xy, conic.xyz, opacity, color.rgb
    __shared__ uint32_t s_max_last;
    const int vid = blockIdx.y, tile = blockIdx.x;
    const int tx = tile % d.gx, ty = tile / d.gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wx0 = tx * kTile + (warp & 1) * 8, wy0 = ty * kTile + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < d.W && pyi < d.H;
    const float px = (float)pxi, py = (float)pyi;
    const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
    if (*geo.n_instances > d.capacity) return;
    const size_t seg = (size_t)vid * d.tiles + tile;
    const uint32_t start = geo.tile_start[seg];
    const size_t gbase = (size_t)vid * d.P;
    const size_t hw = (size_t)d.H * d.W;
    const size_t pix = (size_t)pyi * d.W + pxi;

    const float T_final = inside ? final_T[(size_t)vid * hw + pix] : 0.0f;
    const uint32_t last = inside ? n_contrib[(size_t)vid * hw + pix] : 0u;
    float dpr = 0.0f, dpg = 0.0f, dpb = 0.0f;
    if (inside) {
        const float *dc = d_color + (size_t)vid * 3 * hw;
        dpr = dc[pix]; dpg = dc[hw + pix]; dpb = dc[2 * hw + pix];
    }
    const float *bg = bg_all + 3 * vid;
    const float bg_dot = bg[0] * dpr + bg[1] * dpg + bg[2] * dpb;
    const float ddelx_dx = 0.5f * (float)d.W, ddely_dy = 0.5f * (float)d.H;

    if (tid == 0) s_max_last = 0;
    __syncthreads();
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) atomicMax(&s_max_last, warp_last);
    __syncthreads();
    const uint32_t block_last = s_max_last;   // entries at list position >= block_last are unused

    float T = T_final;
    float acc_r = 0.0f, acc_g = 0.0f, acc_b = 0.0f, last_alpha = 0.0f, lc_r = 0.0f, lc_g = 0.0f, lc_b = 0.0f;

    // walk positions block_last-1 .. 0, staged in chunks of kStage (chunk k covers the
    // positions [hi_k - n_k, hi_k) with hi_0 = block_last, highest position first)
    for (uint32_t hi = block_last; hi > 0;) {
        const uint32_t n_here = min((uint32_t)kStage, hi);
        // staged index j <-> list position  pos = hi - 1 - j
        if ((uint32_t)tid < n_here) {
            const uint32_t g = (uint32_t)keys[start + (hi - 1u - (uint32_t)tid)];
            stage_entry(s, tid, geo, gbase, g);
            s_g[tid] = g;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) s_acc[tid][k] = 0.0f;
        __syncthreads();
        for (uint32_t jb = 0; jb < n_here; jb += 32) {
            const uint32_t j = jb + lane;
            bool hit = false;
            if (j < n_here && (hi - 1u - j) < warp_last) {
                const float2 c = s.xy[j], e = s.ext[j];
                hit = (c.x + e.x >= rx0) && (c.x - e.x <= rx1) && (c.y + e.y >= ry0) && (c.y - e.y <= ry1);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int bpos = __ffs(mask) - 1;
                mask &= mask - 1;
                const uint32_t jj = jb + (uint32_t)bpos;
                const uint32_t pos = hi - 1u - jj;
                struct { float2 xy; float4 co; float4 rgb; } e;
                e.xy = s.xy[jj]; e.co = s.co[jj]; e.rgb = s.rgb[jj];
                float g_mx = 0.0f, g_my = 0.0f, g_ca = 0.0f, g_cb = 0.0f, g_cc = 0.0f, g_op = 0.0f;
                float g_r = 0.0f, g_g = 0.0f, g_b = 0.0f;
                bool active = pos < last;
                float dx = 0.0f, dy = 0.0f, G = 0.0f, alpha = 0.0f;
                if (active) {
                    dx = e.xy.x - px; dy = e.xy.y - py;
                    const float power = -0.5f * (e.co.x * dx * dx + e.co.z * dy * dy) - e.co.y * dx * dy;
                    active = !(power > 0.0f);
                    if (active) {
                        G = __expf(power);
                        alpha = fminf(0.99f, e.co.w * G);
                        active = !(alpha < kAlphaMin);
                    }
                }
                if (active) {
                    T = T / (1.0f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    acc_r = last_alpha * lc_r + (1.0f - last_alpha) * acc_r;
                    acc_g = last_alpha * lc_g + (1.0f - last_alpha) * acc_g;
                    acc_b = last_alpha * lc_b + (1.0f - last_alpha) * acc_b;
                    lc_r = e.rgb.x; lc_g = e.rgb.y; lc_b = e.rgb.z;
                    float dL_dalpha = (e.rgb.x - acc_r) * dpr + (e.rgb.y - acc_g) * dpg + (e.rgb.z - acc_b) * dpb;
                    g_r = dchannel_dcolor * dpr; g_g = dchannel_dcolor * dpg; g_b = dchannel_dcolor * dpb;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = e.co.w * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * e.co.x - gdy * e.co.y;
                    const float dG_ddely = -gdy * e.co.z - gdx * e.co.y;
                    g_mx = dL_dG * dG_ddelx * ddelx_dx;
                    g_my = dL_dG * dG_ddely * ddely_dy;
                    g_ca = -0.5f * gdx * dx * dL_dG;
                    g_cb = -0.5f * gdx * dy * dL_dG;
                    g_cc = -0.5f * gdy * dy * dL_dG;
                    g_op = G * dL_dalpha;
                }
                if (!__any_sync(0xffffffffu, active)) continue;
                g_mx = warp_sum(g_mx); g_my = warp_sum(g_my);
                g_ca = warp_sum(g_ca); g_cb = warp_sum(g_cb); g_cc = warp_sum(g_cc);
                g_op = warp_sum(g_op);
                g_r = warp_sum(g_r); g_g = warp_sum(g_g); g_b = warp_sum(g_b);
                if (lane == 0) {
                    float *a = s_acc[jj];
                    atomicAdd(a + 0, g_mx); atomicAdd(a + 1, g_my);
                    atomicAdd(a + 2, g_ca); atomicAdd(a + 3, g_cb); atomicAdd(a + 4, g_cc);
                    atomicAdd(a + 5, g_op);
                    atomicAdd(a + 6, g_r); atomicAdd(a + 7, g_g); atomicAdd(a + 8, g_b);
                }
            }
        }
        __syncthreads();
        if ((uint32_t)tid < n_here) {
            const float *a = s_acc[tid];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 9; ++k) any |= (a[k] != 0.0f);
            if (any) {
                const size_t o = gbase + s_g[tid];
                float *m = reinterpret_cast<float *>(vg.d_mean2d + o);
                float *c = reinterpret_cast<float *>(vg.d_conic + o);
                float *col = reinterpret_cast<float *>(vg.d_color + o);
                atomicAdd(m + 0, a[0]); atomicAdd(m + 1, a[1]);
                atomicAdd(c + 0, a[2]); atomicAdd(c + 1, a[3]); atomicAdd(c + 2, a[4]); atomicAdd(c + 3, a[5]);
                atomicAdd(col + 0, a[6]); atomicAdd(col + 1, a[7]); atomicAdd(col + 2, a[8]);
            }
        }
        __syncthreads();
        hi -= n_here;
    }
}

int launch_composite_forward(const Dims &d, const Inputs &in, const Geom &g,
                             const unsigned long long *keys, float *final_T, uint32_t *n_contrib,
                             float *out_color, cudaStream_t st) {
    dim3 grid(d.tiles, d.S * d.V);
    k_composite_fwd<<<grid, kCompThreads, 0, st>>>(d, g, in.bg, keys, final_T, n_contrib, out_color);
    PS_LAUNCH_CHECK("k_composite_fwd");
    return PS_OK;
}

int launch_composite_backward(const Dims &d, const Inputs &in, const Geom &g,
                              const unsigned long long *keys, const float *final_T,
                              const uint32_t *n_contrib, const float *d_color, const ViewGrads &vg,
                              cudaStream_t st) {
    dim3 grid(d.tiles, d.S * d.V);
    k_composite_bwd<<<grid, kCompThreads, 0, st>>>(d, g, in.bg, keys, final_T, n_contrib, d_color, vg);
    PS_LAUNCH_CHECK("k_composite_bwd");
    return PS_OK;
}

}  // namespace ps
