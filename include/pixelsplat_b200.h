/*
 * pixelsplat_b200.h -- C ABI of the B200-native render hot path of pixelSplat.
 *
 * Drop-in boundary.  The reference binds its rasterizer through the
 * Python extension `diff_gaussian_rasterization` (imported at
 * /root/reference/src/model/decoder/cuda_splatting.py:5-8, called at :99-124 and :192-217);
 * that extension's own C++ entry points are `_C.rasterize_gaussians` /
 * `_C.rasterize_gaussians_backward` (un-vendored dependency, requirements.txt:17).  The two
 * functions below replace exactly those two, generalised to a batch of scenes x views so that
 * `render_cuda`'s per-view Python loop (cuda_splatting.py:91-126), its two `.item()` host syncs
 * (:102-103), the SH permute copy (:75), the covariance triu gather (:123) and
 * DecoderSplattingCUDA's `repeat` of every Gaussian tensor per view
 * (decoder_splatting_cuda.py:53-56) all disappear.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; nothing here is a torch type;
 *   - all work is enqueued on `stream`; no call synchronises the device;
 *   - every function returns PS_OK (0) or a PS_ERR_* code; ps_last_error() gives the text;
 *   - matrices are 16 floats, column-major (element [4*col+row]) -- the flattened row-major
 *     transpose that cuda_splatting.py:85-87 builds;
 *   - fp32 throughout; indices uint32; sort keys uint64.
 */
#ifndef PIXELSPLAT_B200_H
#define PIXELSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PS_API __attribute__((visibility("default")))
#else
#define PS_API
#endif

#define PS_OK 0
#define PS_ERR_INVALID_ARGUMENT 1
#define PS_ERR_CUDA 2
#define PS_ERR_UNSUPPORTED 3

#define PS_TILE 16 /* upstream BLOCK_X = BLOCK_Y */

/* sh_layout */
#define PS_SH_M3 0 /* [P, M, 3]  what GaussianRasterizer.forward receives (cuda_splatting.py:75) */
#define PS_SH_3M 1 /* [P, 3, M]  pixelSplat's native Gaussians.harmonics (model/types.py:11)    */
/* SH convention (ps_raster_desc.sh_basis, ps_sh_rotation_matrices' `convention`):
 *   3DGS  z polar, Condon-Shortley sign (-1)^m, degree 1 = C1 (-y, z, -x): upstream 3DGS;
 *   E3NN  y polar, no Condon-Shortley sign, degree 1 = C1 (x, y, z): the basis of e3nn's real harmonics,
 *         in which the reference rotates its coefficients (/root/reference/src/misc/sh_rotation.py:18-22,
 *         ply_export.py:75 "our axes are swizzled for the spherical harmonics").
 *   Y_e3nn,i(x, y, z) = (-1)^m Y_3dgs,i(z, x, y). */
#define PS_SH_BASIS_3DGS 0
#define PS_SH_BASIS_E3NN 1
/* cov_layout */
#define PS_COV_TRIU6 0 /* [P, 6]  xx,xy,xz,yy,yz,zz = cov3D_precomp (cuda_splatting.py:115,123) */
#define PS_COV_3X3 1   /* [P, 3, 3]  Gaussians.covariances; only the upper triangle is read, and
                          only the upper triangle receives gradient (autograd of the triu gather) */

typedef struct ps_raster_desc {
    int32_t n_scenes;        /* S: independent Gaussian sets                                   */
    int32_t views_per_scene; /* V: cameras that share one Gaussian set                         */
    int32_t n_gaussians;     /* P per scene                                                    */
    int32_t sh_coeffs;       /* M (1..25) when colours come from SH; 0 = colors_precomp [P,3]  */
    int32_t sh_degree;       /* active degree, 0..4, (sh_degree+1)^2 <= M                      */
    int32_t sh_layout;       /* PS_SH_*                                                        */
    int32_t cov_layout;      /* PS_COV_*                                                       */
    int32_t height, width;   /* image size in pixels                                           */
    int32_t sort_impl;       /* 0 = native per-tile radix sort; 1 = CUB segmented sort (debug)  */
    int32_t sort_segment_hint; /* longest (view,tile) list expected (from a previous call's
                                  n_instances_host[1]); 0 = unknown.  Only selects the shared-memory
                                  size of the sort; any value is correct                        */
    int64_t instance_capacity; /* room, in (tile,Gaussian) instances, summed over all S*V views */
    /* appended in round 2 (struct grows at the end only) */
    int32_t sh_basis;        /* PS_SH_BASIS_*: the convention the SH coefficients are evaluated in    */
    int32_t reserved;        /* must be 0                                                      */
} ps_raster_desc;

/* Per-call inputs. Camera arrays are indexed by flat view id  vid = scene * V + view. */
typedef struct ps_raster_inputs {
    const float *means;      /* [S, P, 3]                                                      */
    const float *cov;        /* [S, P, 6] or [S, P, 3, 3] (cov_layout)                         */
    const float *opacities;  /* [S, P]                                                         */
    const float *sh;         /* [S, P, M, 3] / [S, P, 3, M] (sh_layout), or colours [S, P, 3]  */
    const float *viewmatrix; /* [S*V, 16] world->camera, column-major                          */
    const float *projmatrix; /* [S*V, 16] world->clip (view @ proj), column-major              */
    const float *campos;     /* [S*V, 3]                                                       */
    const float *tanfov;     /* [S*V, 2] tan(fov_x/2), tan(fov_y/2)                            */
    const float *background; /* [S*V, 3]                                                       */
    const float *scene_scale; /* [S*V] or NULL: means*=s, cov*=s*s before use -- the
                                 scale_invariant rescale of cuda_splatting.py:64-71, fused     */
} ps_raster_inputs;

/* Opaque state that lives from forward to backward (the analogue of upstream's geomBuffer /
 * binningBuffer / imgBuffer byte tensors).  The caller allocates; sizes from ps_raster_sizes. */
typedef struct ps_raster_state {
    void *geom;    size_t geom_bytes;
    void *binning; size_t binning_bytes;
    void *image;   size_t image_bytes;
} ps_raster_state;

typedef struct ps_raster_sizes {
    size_t geom_bytes, binning_bytes, image_bytes, backward_bytes;
} ps_raster_sizes;

/* Byte offsets of the intermediates inside the state buffers (parity tests read them;
 * "bit-exact tile/bin indices" is checked on keys / tile_start / tile_count). */
typedef struct ps_raster_layout {
    /* geom */
    size_t depth;         /* f32  [S*V*P]                                                      */
    size_t radii;         /* i32  [S*V*P]                                                      */
    size_t xy;            /* f32x2                                                             */
    size_t conic_opacity; /* f32x4                                                             */
    size_t rgb;           /* f32x4 (r,g,b,unused)                                              */
    size_t rect;          /* u16x4 (minx,miny,maxx,maxy) in tiles                              */
    size_t clamped;       /* u8   bit c set = channel c was clamped to 0                       */
    size_t tile_count;    /* u32  [S*V*tiles]                                                  */
    size_t tile_start;    /* u32  [S*V*tiles] exclusive scan, global instance offsets          */
    size_t tile_cursor;   /* u32  scratch                                                      */
    size_t n_instances;   /* i64  [4] instances needed (may exceed capacity), longest segment,
                                     #visible (view,Gaussian) pairs, #Gaussians visible in any view */
    size_t vis_pairs;     /* u32  [S*V*P] compact list of on-screen (view,Gaussian) flat indices  */
    size_t vis_any;       /* u32  [S*P]   compact list of (scene,Gaussian) visible in >= 1 view    */
    /* binning */
    size_t keys;          /* u64  [capacity]  per tile sorted (float_bits(depth)<<32 | gaussian) */
    size_t keys_alt;      /* u64  [capacity]  scratch                                          */
    /* image */
    size_t final_T;       /* f32  [S*V*H*W]                                                    */
    size_t n_contrib;     /* u32  [S*V*H*W]                                                    */
    /* appended in round 2 (struct grows at the end only) */
    size_t cull;          /* geom: f32x4 [S*V*P] (x, y, half-extent x, half-extent y) of the
                             alpha >= 1/255 box -- the compositor's cull record                */
    size_t color;         /* image: f32 [S*V*3*H*W] copy of the rendered colour (the backward's
                             forward-order prefix form needs C . dL/dC per pixel)              */
    size_t block_hits;    /* binning: u32x2 [8 * capacity] per-(tile, 8x4 block, run) hit lists (position, Gaussian)
                             the composite forward leaves for its backward; 0 when not kept (large capacity)   */
    size_t run_hits;      /* binning: u32 [S*V*tiles*8*4] their lengths                                        */
    size_t run_state;     /* image: f32x4 [S*V*3*H*W] (T, Cr, Cg, Cb) in front of list runs 1..3 when the
                             compositor cuts a tile's list into runs (small batches)            */
} ps_raster_layout;

typedef struct ps_raster_grads {
    float *d_means;     /* [S, P, 3]                                                           */
    float *d_cov;       /* same layout as cov                                                  */
    float *d_opacities; /* [S, P]                                                              */
    float *d_sh;        /* same layout as sh (or [S, P, 3] for colours)                        */
    float *d_means2d;   /* [S*V, P, 3] or NULL: upstream's screen-space gradient (x, y, 0)     */
} ps_raster_grads;

PS_API int ps_version(void);
PS_API const char *ps_last_error(void); /* thread-local, valid until the next failing call */

/* Instrumentation used by bench.py: number of kernels this library has launched so far, and
 * optional per-stage CUDA-event timing (on the launching stream) of the most recent
 * forward + backward pair: ms[7] = preprocess, count-scan + scatter, sort, composite forward,
 * gradient zero-fill, composite backward, preprocess backward. */
PS_API unsigned long long ps_launch_count(void);
PS_API void ps_timing_enable(int on);
PS_API int ps_timing_read(float *ms);

/* Process-wide tunables of the compositor, for A/B measurements and tests (defaults are automatic):
 *   "composite_impl"      2 = warp-task compositor (default), 1 = round-1 CTA-per-tile compositor;
 *   "composite_segments"  0 = automatic (default), 1 | 2 | 4 = list runs per warp task.
 * Takes effect for forwards issued afterwards (a backward uses the split its forward used only if the option is
 * unchanged in between).  Also read once from PIXELSPLAT_B200_COMPOSITE / PIXELSPLAT_B200_SEGMENTS. */
PS_API int ps_set_option(const char *name, int value);

/* Workspace sizes / layout for a descriptor. */
PS_API int ps_raster_sizes_query(const ps_raster_desc *desc, ps_raster_sizes *out);
PS_API int ps_raster_layout_query(const ps_raster_desc *desc, ps_raster_layout *out);

/*
 * Camera set-up for n_views views in one launch: the device-side restatement of
 * cuda_splatting.py:64-87 (scale-invariant rescale of the extrinsics translation and near/far,
 * get_fov, get_projection_matrix, extrinsics.inverse(), view @ proj) without the per-view
 * `.item()` host syncs of :102-103.
 *   extrinsics [n,4,4] camera-to-world row-major, intrinsics [n,3,3] normalised, near/far [n]
 *   -> viewmatrix/projmatrix [n,16] (column-major), campos [n,3], tanfov [n,2],
 *      scene_scale [n] (= 1/near when scale_invariant, else 1; feed to ps_raster_inputs).
 */
PS_API int ps_camera_setup(int32_t n_views, const float *extrinsics, const float *intrinsics,
                           const float *near_plane, const float *far_plane, int32_t scale_invariant,
                           float *viewmatrix, float *projmatrix, float *campos, float *tanfov,
                           float *scene_scale, void *stream);

/*
 * Forward: preprocess -> per-tile count/scan -> scatter -> per-tile radix sort -> composite.
 * Replaces _C.rasterize_gaussians for S*V views at once.
 *   out_color      [S*V, 3, H, W]
 *   out_radii      [S*V, P] int32 or NULL
 *   n_instances_host  pinned HOST int64[2] or NULL: receives {instance count, longest (view,tile)
 *                  list} asynchronously (valid once `stream` reaches this point).  If the count exceeds
 *                  desc->instance_capacity the binning was truncated and out_color is INVALID:
 *                  the caller must re-run with a larger capacity (pixelsplat_b200.rasterizer does).
 */
PS_API int ps_raster_forward(const ps_raster_desc *desc, const ps_raster_inputs *in,
                      const ps_raster_state *state, float *out_color, int32_t *out_radii,
                      int64_t *n_instances_host, void *stream);

/*
 * Backward: composite backward (warp-reduced, one atomic per (tile, Gaussian)) ->
 * per-Gaussian cov2D / projection / SH backward summed over the V views of a scene.
 * Replaces _C.rasterize_gaussians_backward.  `scratch` has ps_raster_sizes.backward_bytes.
 */
PS_API int ps_raster_backward(const ps_raster_desc *desc, const ps_raster_inputs *in,
                       const ps_raster_state *state, const float *d_color /* [S*V,3,H,W] */,
                       void *scratch, size_t scratch_bytes, const ps_raster_grads *grads,
                       void *stream);

/* ---- fused loss epilogue --------------------------------------------------
 * The training loss the reference applies to the render, LossMse (/root/reference/src/loss/loss_mse.py:30-31,
 * weight * mean((prediction - target)^2)), and the PSNR it logs (src/evaluation/metrics.py:11-19) need, per view,
 * sum (C - t)^2 and sum (clip(C) - clip(t))^2.  ps_raster_forward_loss accumulates both in the compositor's
 * epilogue (the image is not re-read for the loss; out_color may be NULL when nobody needs the pixels), and
 * ps_raster_backward_loss forms dL/dC = grad_scale[view] * (C - target) inside the composite backward, so no
 * gradient image exists as a tensor either.  Host side: pixelsplat_b200/loss.py. */
#define PS_LOSS_SLOTS 64
typedef struct ps_raster_loss {
    const float *target; /* [S*V, 3, H, W] */
    float *sums;         /* [S*V, 2, PS_LOSS_SLOTS]: partial sums, [.,0,.] raw, [.,1,.] clipped to [0, 1];
                            zeroed by the call; sum over the last axis for the per-view totals          */
} ps_raster_loss;

PS_API int ps_raster_forward_loss(const ps_raster_desc *desc, const ps_raster_inputs *in,
                                  const ps_raster_state *state, const ps_raster_loss *loss,
                                  float *out_color /* or NULL */, int32_t *out_radii,
                                  int64_t *n_instances_host, void *stream);
PS_API int ps_raster_backward_loss(const ps_raster_desc *desc, const ps_raster_inputs *in,
                                   const ps_raster_state *state, const float *target /* [S*V,3,H,W] */,
                                   const float *grad_scale /* [S*V]: dL/d(sum of squares) * 2 */,
                                   void *scratch, size_t scratch_bytes, const ps_raster_grads *grads,
                                   void *stream);

/* ------------------------------------------------------------------------------------------
 * Epipolar sampled cross-attention.
 * Replaces, inside EpipolarTransformer.forward
 * (/root/reference/src/model/encoder/epipolar/epipolar_transformer.py:96-142):
 *   ps_epipolar_geometry            EpipolarSampler's ray generation + project_rays
 *                                   (epipolar_sampler.py:62-88, geometry/epipolar_lines.py:157-251),
 *                                   get_depth (epipolar_lines.py:264-292, projection.py:176-230),
 *                                   depth clip + depth_to_relative_disparity
 *                                   (epipolar_transformer.py:103-119, conversions.py:17-27);
 *   ps_epipolar_attention_forward / _backward
 *                                   F.grid_sample of the samples (epipolar_sampler.py:98-111), the
 *                                   depth positional encoding (epipolar_transformer.py:120-121,
 *                                   positional_encoding.py:28-33) and Attention.forward with z != None
 *                                   (transformer/attention.py:54-70) for one transformer layer.
 * Ray index r = row * grid_w + col over the (down-scaled) feature grid; "other view" ov of view v
 * is view ov if ov < v else ov + 1 (misc/heterogeneous_pairings.py:9-24).
 */
typedef struct ps_epipolar_desc {
    int32_t batch, views;     /* b, v (v >= 2)                                                  */
    int32_t grid_h, grid_w;   /* ray / feature grid                                             */
    int32_t samples;          /* S <= 32 samples per epipolar segment                           */
    int32_t channels;         /* feature channels, must be 128                                  */
    int32_t heads;            /* 1..4                                                           */
    int32_t pe_dim;           /* 2 * num_octaves of the depth encoding (<= 32, heads*pe_dim<=96) */
} ps_epipolar_desc;

typedef struct ps_epipolar_inputs {
    const float *features;      /* [b, v, grid_h, grid_w, 128] channels-last                    */
    const float *segments;      /* [b, v, v-1, R, 4] xy_min.xy, xy_max.xy (from ps_epipolar_geometry) */
    const uint8_t *valid;       /* [b, v, v-1, R]                                               */
    const float *rel_disparity; /* [b, v, v-1, R, S]                                            */
    const float *q_feat;        /* [b*v*R, heads, 128]  scale * W_k,h^T q_h                     */
    const float *q_pe;          /* [b*v*R, heads, pe_dim]  W_d^T of the above                   */
    const float *bias;          /* [b*v*R, heads, v-1] or NULL (view-embedding score term)      */
} ps_epipolar_inputs;

/* segments/valid/rel_disparity as above; t_range [b, v, v-1, R, 2] (t_min, t_max) or NULL. */
PS_API int ps_epipolar_geometry(int32_t batch, int32_t views, int32_t grid_h, int32_t grid_w,
                                int32_t samples, const float *extrinsics /* [b,v,4,4] c2w */,
                                const float *intrinsics /* [b,v,3,3] */, const float *near_plane,
                                const float *far_plane /* [b,v] */, float *segments, uint8_t *valid,
                                float *rel_disparity, float *t_range, void *stream);

/* z [N,heads,128] = sum_s a_s f_s;  e [N,heads,pe_dim] = sum_s a_s PE(rd_s);
 * mass [N,heads,v-1] = per-other-view attention mass (or NULL);  lse [N,heads] log-sum-exp. */
PS_API int ps_epipolar_attention_forward(const ps_epipolar_desc *desc, const ps_epipolar_inputs *in,
                                         float *z, float *e, float *mass, float *lse, void *stream);

/* d_row [N,heads] = dz.z + de.e (+ dmass.mass).  dfeatures [b,v,grid_h,grid_w,128] must be
 * zero-initialised by the caller; the kernel accumulates into it atomically. */
PS_API int ps_epipolar_attention_backward(const ps_epipolar_desc *desc, const ps_epipolar_inputs *in,
                                          const float *lse, const float *dz, const float *de,
                                          const float *dmass, const float *d_row, float *dq_feat,
                                          float *dq_pe, float *dbias, float *dfeatures, void *stream);

/* ---- dense per-image self-attention (tcgen05, TF32 operands, FP32 accumulate) ---------------------
 * Replaces the z = None branch of /root/reference/src/model/transformer/attention.py:54-70 as used by
 * ImageSelfAttention (/root/reference/src/model/encoder/epipolar/image_self_attention.py:57-79):
 *   qkv  [n_images, tokens, 3 * heads * dim_head]   output of to_qkv ("b n (qkv h d)")
 *   out  [n_images, tokens, heads * dim_head]       softmax(q k^T * scale) v, "b n (h d)"
 * Supported shape: tokens == 256, dim_head == 128 (pixelSplat's ViT stage at 256x256), heads <= 16;
 * anything else returns PS_ERR_UNSUPPORTED.  debug_mode 1 writes the raw q k^T logits instead
 * (out is then [n_images, heads, 256, 256]); used by the tests only. */
PS_API int ps_self_attention_forward(int32_t n_images, int32_t tokens, int32_t heads, int32_t dim_head,
                                     const float *qkv, float scale, float *out, int32_t debug_mode,
                                     void *stream);

/* The same forward, additionally saving what the backward needs to rebuild the forward's probabilities bit for
 * bit: stats [n_images, heads, 256, 2] = (row max * scale * log2 e, 1 / row sum of the TF32-rounded numerators). */
PS_API int ps_self_attention_forward_stats(int32_t n_images, int32_t tokens, int32_t heads, int32_t dim_head,
                                           const float *qkv, float scale, float *out, float *stats, void *stream);

/* Backward (tcgen05, TF32 operands, FP32 accumulate; csrc/self_attention_tc_bwd.cu): autograd of the forward
 * above.  out / d_out [n_images, 256, heads * 128]; d_qkv has qkv's layout and is fully written. */
PS_API int ps_self_attention_backward(int32_t n_images, int32_t tokens, int32_t heads, int32_t dim_head,
                                      const float *qkv, const float *out, const float *d_out, const float *stats,
                                      float scale, float *d_qkv, void *stream);

/* ---- fused GaussianAdapter ----------------------------------------------
 * Replaces /root/reference/src/model/encoder/common/gaussian_adapter.py:48-95 (+ gaussians.py:8-44):
 * per-ray raw network outputs -> world-space Gaussians.  One camera = one (batch, view) pair; every
 * ray carries n_samples Gaussians that share its raw features and differ in depth.
 *   means [n_views, n_rays, n_samples, 3]      covariances [.., 3, 3]      harmonics [.., 3, sh_coeffs]
 *   scales [.., 3] (optional, may be NULL)      rotations [n_views, n_rays, 4] xyzw (optional)
 * Opacities pass through the reference's adapter unchanged and are not part of this call. */
typedef struct ps_adapter_desc {
    int32_t n_views;      /* b * v cameras */
    int32_t n_rays;       /* rays x surfaces per camera */
    int32_t n_samples;    /* Gaussians per ray (1..8) */
    int32_t sh_coeffs;    /* (degree + 1)^2, degree <= 4 */
    int32_t image_h, image_w;
    float scale_min, scale_max; /* GaussianAdapterCfg.gaussian_scale_min / max */
    float eps;            /* 1e-8 in the reference */
    int32_t reserved;
} ps_adapter_desc;

typedef struct ps_adapter_inputs {
    const float *extrinsics;   /* [n_views, 4, 4] camera-to-world */
    const float *intrinsics;   /* [n_views, 3, 3] normalised */
    const float *sh_rotation;  /* [n_views, sh_coeffs, sh_coeffs] block-diagonal D(c2w): c' = D c */
    const float *sh_mask;      /* [sh_coeffs] */
    const float *coordinates;  /* [n_views, n_rays, 2] */
    const float *depths;       /* [n_views, n_rays, n_samples] */
    const float *raw;          /* [n_views, n_rays, 7 + 3 sh_coeffs]: scale logits 3, quaternion xyzw 4, sh [3, sh_coeffs] */
} ps_adapter_inputs;

PS_API int ps_gaussian_adapter_forward(const ps_adapter_desc *desc, const ps_adapter_inputs *in, float *means,
                                       float *covariances, float *harmonics, float *scales, float *rotations,
                                       void *stream);

/* d_scales / d_rotations may be NULL (no gradient reached those outputs).  Outputs: d_coordinates
 * [n_views, n_rays, 2], d_depths [n_views, n_rays, n_samples], d_raw [n_views, n_rays, 7 + 3 sh_coeffs]
 * (all fully written, no zero-initialisation needed). */
PS_API int ps_gaussian_adapter_backward(const ps_adapter_desc *desc, const ps_adapter_inputs *in,
                                        const float *d_means, const float *d_covariances,
                                        const float *d_harmonics, const float *d_scales,
                                        const float *d_rotations, float *d_coordinates, float *d_depths,
                                        float *d_raw, void *stream);

/* Block-diagonal SH rotation matrices D(c2w) [n_views, sh_coeffs, sh_coeffs] for ps_adapter_inputs.sh_rotation
 * (replaces /root/reference/src/misc/sh_rotation.py:10-30, row f-3): c' = D c makes the rotated function at
 * d equal the original at R^T d, in the basis `convention` names.  PS_SH_BASIS_E3NN reproduces the reference's
 * wigner_D(l, *matrix_to_angles(R)) (degree-1 block == R); PS_SH_BASIS_3DGS is the rotation consistent with the
 * rasterizer's default basis.  fit_dirs [n_dirs, 3] are unit directions and fit_pinv [sh_coeffs, n_dirs] the
 * per-degree pseudo-inverse of the 3DGS basis sampled there (pixelsplat_b200/sh.py builds both once, in
 * float64).  extrinsics [n_views, 4, 4] camera-to-world. */
PS_API int ps_sh_rotation_matrices(int32_t n_views, int32_t sh_coeffs, int32_t n_dirs, int32_t convention,
                                   const float *extrinsics, const float *fit_dirs, const float *fit_pinv,
                                   float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PIXELSPLAT_B200_H */
