// Shared declarations of the rasterizer kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pixelsplat_b200.h"

namespace ps {

constexpr int kTile = PS_TILE;
constexpr int kTilePixels = kTile * kTile;

// Resolved device views of the workspace (host-built, passed by value to kernels).
struct Geom {
    float *depth;
    int32_t *radii;
    float2 *xy;
    float4 *conic_opacity;
    float4 *rgb;
    ushort4 *rect;
    uint8_t *clamped;
    uint32_t *tile_count;
    uint32_t *tile_start;
    uint32_t *tile_cursor;
    long long *n_instances;   // [0] instances, [1] longest segment, [2] #vis_pairs, [3] #vis_any
    uint32_t *vis_pairs;      // compact list of (view, Gaussian) flat indices that are on screen
    uint32_t *vis_any;        // compact list of (scene, Gaussian) flat indices visible in >= 1 view
    float4 *cull;             // xy + half-extents of the alpha >= 1/255 box (compositor's cull record)
};

// Per-view image-space state kept from forward to backward.
struct ImageState {
    float *final_T;           // [S*V*H*W]
    uint32_t *n_contrib;      // [S*V*H*W]
    float *color;             // [S*V*3*H*W] copy of the rendered colour (backward's forward-order prefix form)
    float4 *run_state;        // [S*V*(kMaxSegments-1)*H*W] (T, Cr, Cg, Cb) in front of list runs 1.. (segK > 1)
};

// Fused loss epilogue of the compositor: squared error against a target image summed per
// view in the forward, dL/dC = grad_scale[view] (C - target) formed inside the backward.  All null = off.
constexpr int kLossSlots = PS_LOSS_SLOTS;
struct LossEpilogue {
    const float *target;       // [S*V, 3, H, W]
    float *sums;               // forward: [S*V, 2, kLossSlots] partial sums (raw, clipped), zeroed by the caller
    const float *grad_scale;   // backward: [S*V]
};

constexpr int kMaxSegments = 4;   // list runs per warp task (raster_composite2.cu)

// The composite forward's per-(tile, block, run) hit lists, kept for the backward (binning state).  Run k of block
// `sub` of tile segment `seg` owns hits[(tile_start[seg] * 8 + sub * tile_count[seg]) + run_begin ...] (a run cannot
// have more hits than entries), its length is run_hits[(seg * 8 + sub) * kMaxSegments + k].
struct HitLists {
    uint2 *hits;              // (list position, Gaussian id)
    uint32_t *run_hits;
};

struct Dims {
    int S, V, P, M, deg, sh_layout, cov_layout, H, W, gx, gy, tiles, sh_basis, segK, hit_lists;
    long long capacity;
};

struct Inputs {
    const float *means, *cov, *opac, *sh, *view, *proj, *campos, *tanfov, *bg, *scale;
};

// Per-(view,Gaussian) gradient scratch written by the composite backward.
struct ViewGrads {
    float2 *d_mean2d;  // NDC-scaled like upstream (x * 0.5 W, y * 0.5 H)
    float4 *d_conic;   // x, y (half-weighted B), z, w = d_opacity
    float4 *d_color;   // r, g, b, unused
};

// Half-extents (pixels) of the axis-aligned box outside of which a splat's alpha is certainly
// < 1/255: alpha = o exp(-0.5 d^T Sigma^-1 d) >= 1/255  <=>  d^T Sigma^-1 d <= 2 ln(255 o), whose bounding
// box is sqrt(2 ln(255 o) Sigma_ii) -- taken from the 2D covariance itself (a, c = its diagonal), so no
// cancellation; the margins cover the compositor's approximate exp2 and the rounding of the conic.
// Large negative = never contributes; large positive = always evaluated (NaN inputs then propagate into
// the image exactly as they would upstream).
__device__ __forceinline__ float2 cull_extent(float a, float c, float o) {
    if (!(a > 0.0f) || !(c > 0.0f) || !(o <= 3.0e38f)) return make_float2(3.0e38f, 3.0e38f);
    if (!(o * 255.0f >= 1.0f - 1e-3f)) return make_float2(-3.0e38f, -3.0e38f);
    const float tau2 = 2.0f * (__logf(fmaxf(o * 255.0f, 1.0f)) + 0.01f);
    return make_float2(sqrtf(tau2 * a) * 1.001f + 0.01f, sqrtf(tau2 * c) * 1.001f + 0.01f);
}

void set_error(const char *fmt, ...);

// Optional per-stage device timing (bench.py's roofline leg): when enabled, the entry points
// record CUDA events on the launching stream between stages.
enum Mark { kMarkFwdStart = 0, kMarkPreprocess, kMarkScatter, kMarkSort, kMarkCompositeFwd,
            kMarkBwdStart, kMarkBwdZero, kMarkCompositeBwd, kMarkPreprocessBwd, kNumMarks };
void mark(int id, cudaStream_t st);
void count_launch();

// Once-per-DEVICE flags (a host process may drive several GPUs; kernel attributes and the library's
// side stream are per device).  `mask` is a caller-owned static, one bit per device ordinal.
// Thread-safe: the test-and-set happens under a process-wide mutex (the caller then sets a kernel attribute,
// which is idempotent, so a second thread racing past the flag at worst repeats it).
bool first_use_on_device(unsigned long long &mask);

#define PS_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ps::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PS_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define PS_LAUNCH_CHECK(name)                                                            \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ps::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return PS_ERR_CUDA;                                                          \
        }                                                                                \
        ps::count_launch();                                                              \
    } while (0)

// stage launchers (each returns PS_OK / PS_ERR_*)
int launch_preprocess(const Dims &d, const Inputs &in, const Geom &g, cudaStream_t st);
int launch_sh_color(const Dims &d, const Inputs &in, const Geom &g, cudaStream_t st);
int launch_binning(const Dims &d, const Geom &g, unsigned long long *keys,
                   unsigned long long *keys_alt, int sort_impl, int segment_hint, cudaStream_t st);
int launch_composite_forward(const Dims &d, const Inputs &in, const Geom &g,
                             const unsigned long long *keys, const ImageState &img,
                             float *out_color, const LossEpilogue &loss, const HitLists &hl, cudaStream_t st);
int launch_composite_backward(const Dims &d, const Inputs &in, const Geom &g,
                              const unsigned long long *keys, const ImageState &img,
                              const float *d_color, const ViewGrads &vg, const LossEpilogue &loss, const HitLists &hl,
                              cudaStream_t st);
// legacy CTA-per-tile compositor (round 1), kept selectable for A/B measurements
int launch_composite_forward_v1(const Dims &d, const Inputs &in, const Geom &g,
                                const unsigned long long *keys, const ImageState &img,
                                float *out_color, cudaStream_t st);
int launch_composite_backward_v1(const Dims &d, const Inputs &in, const Geom &g,
                                 const unsigned long long *keys, const ImageState &img,
                                 const float *d_color, const ViewGrads &vg, cudaStream_t st);
int composite_impl();   // 1 = legacy, 2 = warp-task compositor (env PIXELSPLAT_B200_COMPOSITE, default 2)
int set_composite_option(int which, int value);   // 0: impl (1 | 2), 1: segments (0 = auto | 1 | 2 | 4)
int composite_segments(long long tasks);
bool composite_hit_lists(long long capacity);      // keep the forward's hit lists for the backward?   // list runs per task (1, 2, 4) for a batch of `tasks` warp tasks
int launch_preprocess_backward(const Dims &d, const Inputs &in, const Geom &g, const ViewGrads &vg,
                               const ps_raster_grads &out, cudaStream_t st);
int launch_gradient_fill(const Dims &d, const ps_raster_grads &out, cudaStream_t st);

}  // namespace ps
