// Preprocess forward: per (scene, Gaussian) thread, looping over the scene's V views.
// Frustum cull, EWA projection, conic, 3-sigma radius, tile rectangle, SH -> RGB, and the
// per-(view, tile) instance histogram (shared-memory privatised, flushed once per block).
//
// COMPILED WITH --fmad=false (see csrc/Makefile): depth bits, radii and rectangles are then
// IEEE-exact functions of the inputs and are checked bit-for-bit against oracle/.
// Semantics: SURVEY.md A.1 / A.4 (upstream forward.cu preprocessCUDA); replaces the
// per-view call made at /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include "ps_common.cuh"
#include "raster_math.cuh"

namespace ps {

constexpr int kPreThreads = 128;

__global__ void __launch_bounds__(kPreThreads, 4)
k_preprocess(Dims d, Inputs in, Geom geo, int use_smem_hist, int row_stride) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t *s_hist = s_dyn;                                    // [V * tiles] when use_smem_hist
    const int hist_n = d.V * d.tiles;
    // [warps][32][row_stride], placed after the histogram rounded up to 16 bytes (float4 staging)
    float *s_sh = reinterpret_cast<float *>(s_dyn + (use_smem_hist ? ((hist_n + 3) & ~3) : 0));
    const int scene = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g0 = blockIdx.x * kPreThreads + warp * 32;
    const int g = g0 + lane;
    const bool live = g < d.P;
    if (use_smem_hist) {
        for (int i = threadIdx.x; i < hist_n; i += kPreThreads) s_hist[i] = 0;
        __syncthreads();
    }
    const size_t sg = (size_t)scene * d.P + (live ? g : 0);
    float mx0 = 0.0f, my0 = 0.0f, mz0 = 0.0f, opacity = 0.0f;
    if (live) {
        mx0 = in.means[3 * sg + 0]; my0 = in.means[3 * sg + 1]; mz0 = in.means[3 * sg + 2];
        opacity = in.opac[sg];
    }
    const float *covp = in.cov + sg * (d.cov_layout == PS_COV_TRIU6 ? 6 : 9);
    const int sh_n = 3 * d.M;
    float *row = s_sh + ((size_t)warp * 32 + lane) * row_stride;
    bool staged = false;

    for (int v = 0; v < d.V; ++v) {
        const int vid = scene * d.V + v;
        const size_t vg = (size_t)vid * d.P + (live ? g : 0);
        const float *__restrict__ vm = in.view + 16 * vid;
        const float *__restrict__ pm = in.proj + 16 * vid;
        const float sc = in.scale ? in.scale[vid] : 1.0f;
        const float px = in.scale ? mx0 * sc : mx0;
        const float py = in.scale ? my0 * sc : my0;
        const float pz = in.scale ? mz0 * sc : mz0;
        bool vis = live;
        float vz = 0.0f, pixx = 0.0f, pixy = 0.0f, det_inv = 0.0f;
        int r = 0, minx = 0, miny = 0, maxx = 0, maxy = 0;
        Cov2D cv;
        cv.a = cv.b = cv.c = 0.0f;
        if (vis) {
            geo.radii[vg] = 0;
            vz = vm[2] * px + vm[6] * py + vm[10] * pz + vm[14];
            vis = !(vz <= 0.2f);
        }
        if (vis) {
            const float hx = pm[0] * px + pm[4] * py + pm[8] * pz + pm[12];
            const float hy = pm[1] * px + pm[5] * py + pm[9] * pz + pm[13];
            const float hw = pm[3] * px + pm[7] * py + pm[11] * pz + pm[15];
            const float p_w = 1.0f / (hw + 0.0000001f);
            const float projx = hx * p_w, projy = hy * p_w;
            const float tanfovx = in.tanfov[2 * vid], tanfovy = in.tanfov[2 * vid + 1];
            const float focal_x = (float)d.W / (2.0f * tanfovx);
            const float focal_y = (float)d.H / (2.0f * tanfovy);
            float s6[6];
            load_cov6(covp, d.cov_layout, in.scale ? sc * sc : 1.0f, s6);
            compute_cov2d(px, py, pz, s6, vm, focal_x, focal_y, tanfovx, tanfovy, cv);
            const float det = cv.a * cv.c - cv.b * cv.b;
            vis = !(det == 0.0f);
            if (vis) {
                det_inv = 1.0f / det;
                const float mid = 0.5f * (cv.a + cv.c);
                const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
                const float lambda1 = mid + sq, lambda2 = mid - sq;
                const float my_radius = ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
                pixx = ((projx + 1.0f) * (float)d.W - 1.0f) * 0.5f;
                pixy = ((projy + 1.0f) * (float)d.H - 1.0f) * 0.5f;
                r = (int)my_radius;
                const float rf = (float)r;
                minx = min(d.gx, max(0, (int)((pixx - rf) / (float)kTile)));
                miny = min(d.gy, max(0, (int)((pixy - rf) / (float)kTile)));
                maxx = min(d.gx, max(0, (int)((pixx + rf + (float)(kTile - 1)) / (float)kTile)));
                maxy = min(d.gy, max(0, (int)((pixy + rf + (float)(kTile - 1)) / (float)kTile)));
                vis = (maxx - minx) * (maxy - miny) != 0;
            }
        }
        // the warp's SH rows are brought in once, the first time any of its Gaussians is on screen
        if (d.M > 0 && !staged && __any_sync(0xffffffffu, vis)) {
            const int rows = min(32, d.P - g0);
            stage_sh_rows(in.sh + ((size_t)scene * d.P + g0) * (size_t)sh_n, s_sh + (size_t)warp * 32 * row_stride,
                          rows, sh_n, row_stride, lane);
            __syncwarp();
            staged = true;
        }
        if (!vis) continue;

        float rgb[3];
        uint8_t clamp_bits = 0;
        if (d.M > 0) {
            const float cx = in.campos[3 * vid], cy = in.campos[3 * vid + 1], cz = in.campos[3 * vid + 2];
            const float ddx = px - cx, ddy = py - cy, ddz = pz - cz;
            const float len = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
            const float x = ddx / len, y = ddy / len, z = ddz / len;
            float acc[3] = {0.0f, 0.0f, 0.0f};
            const int M = d.M, layout = d.sh_layout;
            sh_for_each(d.deg, x, y, z, [&](int k, float Y, float, float, float) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float c = row[sh_index(layout, M, k, ch)];
                    acc[ch] = k == 0 ? Y * c : acc[ch] + Y * c;
                }
            });
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float a = acc[ch] + 0.5f;
                if (a < 0.0f) clamp_bits |= (uint8_t)(1u << ch);
                rgb[ch] = fmaxf(a, 0.0f);
            }
        } else {
            const float *__restrict__ col = in.sh + sg * 3;
            rgb[0] = col[0]; rgb[1] = col[1]; rgb[2] = col[2];
        }
        geo.depth[vg] = vz;
        geo.radii[vg] = r;
        geo.xy[vg] = make_float2(pixx, pixy);
        geo.conic_opacity[vg] = make_float4(cv.c * det_inv, -cv.b * det_inv, cv.a * det_inv, opacity);
        geo.rgb[vg] = make_float4(rgb[0], rgb[1], rgb[2], 0.0f);
        geo.rect[vg] = make_ushort4((unsigned short)minx, (unsigned short)miny,
                                    (unsigned short)maxx, (unsigned short)maxy);
        geo.clamped[vg] = clamp_bits;
        for (int ty = miny; ty < maxy; ++ty)
            for (int tx = minx; tx < maxx; ++tx) {
                const int t = ty * d.gx + tx;
                if (use_smem_hist) atomicAdd(&s_hist[v * d.tiles + t], 1u);
                else atomicAdd(&geo.tile_count[(size_t)vid * d.tiles + t], 1u);
            }
    }
    if (use_smem_hist) {
        __syncthreads();
        uint32_t *dst = geo.tile_count + (size_t)scene * d.V * d.tiles;
        for (int i = threadIdx.x; i < hist_n; i += kPreThreads) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(&dst[i], c);
        }
    }
}

int launch_preprocess(const Dims &d, const Inputs &in, const Geom &g, cudaStream_t st) {
    PS_CUDA_CHECK(cudaMemsetAsync(g.tile_count, 0, sizeof(uint32_t) * (size_t)d.S * d.V * d.tiles, st));
    const size_t hist_bytes = sizeof(uint32_t) * (size_t)d.V * d.tiles;
    const int use_smem = hist_bytes <= 32 * 1024;
    const int row_stride = d.M > 0 ? ((3 * d.M) | 1) : 1;
    const size_t sh_bytes = d.M > 0 ? sizeof(float) * kPreThreads * row_stride : 0;
    const size_t smem = (use_smem ? ((hist_bytes + 15) & ~(size_t)15) : 0) + sh_bytes;
    static bool attr = false;
    if (!attr) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr = true;
    }
    dim3 grid((d.P + kPreThreads - 1) / kPreThreads, d.S);
    k_preprocess<<<grid, kPreThreads, smem, st>>>(d, in, g, use_smem, row_stride);
    PS_LAUNCH_CHECK("k_preprocess");
    return PS_OK;
}

}  // namespace ps
