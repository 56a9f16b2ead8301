// Preprocess backward: one thread per (scene, Gaussian), summing over the scene's V views.
//   dL/dconic -> dL/dcov2D -> dL/dcov3D and dL/dmean (through the projection Jacobian),
//   dL/dmean2D -> dL/dmean (perspective divide), dL/drgb -> dL/dSH and dL/dmean (view direction).
// Semantics: SURVEY.md A.5 (upstream backward.cu computeCov2DCUDA + preprocessCUDA), including
// upstream's 1/(det^2 + 1e-7) and the clamp rule (x/y terms vanish when the +-1.3 tan(fov)
// clamp was active).  Because the thread owns the Gaussian, the per-view gradients are summed
// in registers and every output (300 B of dL/dSH at M = 25) is written exactly once per scene
// instead of once per view -- the reference writes them per view and then lets autograd sum
// the `repeat` (decoder_splatting_cuda.py:53-56).
#include "ps_common.cuh"
#include "raster_math.cuh"

namespace ps {

constexpr int kPreBwdThreads = 128;

// One thread per Gaussian that is on screen in at least one view (the compact list built by the
// forward pass), so warps are dense; Gaussians that are not listed receive zero gradients from
// launch_gradient_fill (plain memsets, issued on a side stream so that they overlap the
// compute-bound composite backward).  SH rows are staged per warp in shared memory: coefficients
// come in with coalesced row-wise loads, dL/dSH goes out the same way.
__global__ void __launch_bounds__(kPreBwdThreads, 5)
k_preprocess_bwd(Dims d, Inputs in, Geom geo, ViewGrads vgr, ps_raster_grads out, int row_stride) {
    extern __shared__ float s_dsh[];   // [warps][32][row_stride] coefficients (V == 1: reused for the gradient)
    if (*geo.n_instances > d.capacity) return;
    const long long n = geo.n_instances[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i0 = ((long long)blockIdx.x * kPreBwdThreads) + warp * 32;
    if (i0 >= n) return;
    const long long li = i0 + lane;
    const bool live = li < n;
    const uint32_t sgi = live ? geo.vis_any[li] : 0u;
    const uint32_t scene = sgi / (uint32_t)d.P, g = sgi - scene * (uint32_t)d.P;
    const size_t sg = sgi;
    const int cov_n = d.cov_layout == PS_COV_TRIU6 ? 6 : 9;
    const int sh_n = d.M > 0 ? 3 * d.M : 3;
    const int M = d.M, layout = d.sh_layout;
    float *wrows = s_dsh + (size_t)warp * 32 * row_stride;
    float *row = wrows + lane * row_stride;
    const bool in_place = M > 0 && d.V == 1;     // read each coefficient, then overwrite its slot with the gradient
    float *grow = in_place ? row : row + (size_t)kPreBwdThreads * row_stride;   // separate gradient rows when V > 1

    const int rows_valid = (int)min((long long)32, n - i0);
    // the SH rows stream into shared memory while the geometry part below runs; waited for at first use
    if (M > 0) gather_rows_async(in.sh, (unsigned long long)sg, rows_valid, sh_n, wrows, row_stride, lane);

    float mx0 = 0.0f, my0 = 0.0f, mz0 = 0.0f;
    if (live) { mx0 = in.means[3 * sg + 0]; my0 = in.means[3 * sg + 1]; mz0 = in.means[3 * sg + 2]; }
    const float *covp = in.cov + sg * cov_n;

    float dmx = 0.0f, dmy = 0.0f, dmz = 0.0f, dop = 0.0f;
    float dcov[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    float dcol[3] = {0.0f, 0.0f, 0.0f};
    bool sh_written = false, sh_ready = false;

    for (int v = 0; v < d.V; ++v) {
        const int vid = (int)scene * d.V + v;
        const size_t vg = (size_t)vid * d.P + g;
        const bool vis = live && geo.radii[vg] > 0;
        if (live && out.d_means2d && vis) {
            float *m2 = out.d_means2d + 3 * vg;
            const float2 t = vgr.d_mean2d[vg];
            m2[0] = t.x; m2[1] = t.y;
        }
        // (no `continue` for the views this Gaussian is not on screen in: the warp must stay convergent for the
        //  shared-memory hand-over of the SH rows below)
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        float4 gcol = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float sc = in.scale ? in.scale[vid] : 1.0f;
        const float px = mx0 * sc, py = my0 * sc, pz = mz0 * sc;
        if (vis) {
        const float *__restrict__ vm = in.view + 16 * vid;
        const float *__restrict__ pm = in.proj + 16 * vid;
        const float tanfovx = in.tanfov[2 * vid], tanfovy = in.tanfov[2 * vid + 1];
        const float focal_x = (float)d.W / (2.0f * tanfovx), focal_y = (float)d.H / (2.0f * tanfovy);
        float s6[6];
        load_cov6(covp, d.cov_layout, sc * sc, s6);
        Cov2D cv;
        compute_cov2d(px, py, pz, s6, vm, focal_x, focal_y, tanfovx, tanfovy, cv);

        const float2 g2 = vgr.d_mean2d[vg];
        const float4 gc = vgr.d_conic[vg];
        gcol = vgr.d_color[vg];
        dop += gc.w;

        const float a = cv.a, b = cv.b, c = cv.c;
        const float denom = a * c - b * b;
        const float denom2inv = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0.0f, dL_db = 0.0f, dL_dc = 0.0f;
        const float *m0 = cv.m0, *m1 = cv.m1;
        if (denom2inv != 0.0f) {
            dL_da = denom2inv * (-c * c * gc.x + 2.0f * b * c * gc.y + (denom - a * c) * gc.z);
            dL_dc = denom2inv * (-a * a * gc.z + 2.0f * a * b * gc.y + (denom - a * c) * gc.x);
            dL_db = denom2inv * 2.0f * (b * c * gc.x - (denom + 2.0f * b * b) * gc.y + a * b * gc.z);
            const float s2 = sc * sc;
            dcov[0] += s2 * (m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc);
            dcov[3] += s2 * (m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc);
            dcov[5] += s2 * (m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc);
            dcov[1] += s2 * (2.0f * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db + 2.0f * m1[0] * m1[1] * dL_dc);
            dcov[2] += s2 * (2.0f * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db + 2.0f * m1[0] * m1[2] * dL_dc);
            dcov[4] += s2 * (2.0f * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db + 2.0f * m1[1] * m1[2] * dL_dc);
        }
        // dL/dM (rows) from a = m0 S m0, b = m0 S m1, c = m1 S m1
        const float S[3][3] = {{s6[0], s6[1], s6[2]}, {s6[1], s6[3], s6[4]}, {s6[2], s6[4], s6[5]}};
        float dM0[3], dM1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sm0 = m0[0] * S[k][0] + m0[1] * S[k][1] + m0[2] * S[k][2];
            const float sm1 = m1[0] * S[k][0] + m1[1] * S[k][1] + m1[2] * S[k][2];
            dM0[k] = 2.0f * sm0 * dL_da + sm1 * dL_db;
            dM1[k] = 2.0f * sm1 * dL_dc + sm0 * dL_db;
        }
        const float dJ00 = vm[0] * dM0[0] + vm[4] * dM0[1] + vm[8] * dM0[2];
        const float dJ02 = vm[2] * dM0[0] + vm[6] * dM0[1] + vm[10] * dM0[2];
        const float dJ11 = vm[1] * dM1[0] + vm[5] * dM1[1] + vm[9] * dM1[2];
        const float dJ12 = vm[2] * dM1[0] + vm[6] * dM1[1] + vm[10] * dM1[2];
        const float tz = 1.0f / cv.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = cv.clamp_x ? 0.0f : -focal_x * tz2 * dJ02;
        const float dL_dty = cv.clamp_y ? 0.0f : -focal_y * tz2 * dJ12;
        const float dL_dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 +
                             (2.0f * focal_x * cv.ctx) * tz3 * dJ02 + (2.0f * focal_y * cv.cty) * tz3 * dJ12;
        gx = vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz;
        gy = vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz;
        gz = vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz;

        // screen-space mean through the perspective divide
        const float hx = pm[0] * px + pm[4] * py + pm[8] * pz + pm[12];
        const float hy = pm[1] * px + pm[5] * py + pm[9] * pz + pm[13];
        const float hw = pm[3] * px + pm[7] * py + pm[11] * pz + pm[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        gx += (pm[0] * m_w - pm[3] * mul1) * g2.x + (pm[1] * m_w - pm[3] * mul2) * g2.y;
        gy += (pm[4] * m_w - pm[7] * mul1) * g2.x + (pm[5] * m_w - pm[7] * mul2) * g2.y;
        gz += (pm[8] * m_w - pm[11] * mul1) * g2.x + (pm[9] * m_w - pm[11] * mul2) * g2.y;
        }   // vis (geometry part)

        if (M > 0 && !sh_ready) {          // warp-uniform; the rows have had the geometry math to arrive
            gather_rows_wait();
            sh_ready = true;
        }
        if (M > 0 && vis) {
            const float cx = in.campos[3 * vid], cy = in.campos[3 * vid + 1], cz = in.campos[3 * vid + 2];
            const float ddx = px - cx, ddy = py - cy, ddz = pz - cz;
            const float len2 = ddx * ddx + ddy * ddy + ddz * ddz;
            const float len = sqrtf(len2);
            const float x = ddx / len, y = ddy / len, z = ddz / len;
            const uint8_t cl = geo.clamped[vg];
            const float dl[3] = {(cl & 1) ? 0.0f : gcol.x, (cl & 2) ? 0.0f : gcol.y, (cl & 4) ? 0.0f : gcol.z};
            float dLda = 0.0f, dLdb = 0.0f, dLdc = 0.0f;
            const bool first = !sh_written;
            const float3 sa = sh_arg(d.sh_basis, x, y, z);
            const uint32_t flip = sh_flip_mask(d.sh_basis);
            sh_for_each(d.deg, sa.x, sa.y, sa.z, [&](int k, float Y, float Ya, float Yb, float Yc) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int idx = sh_index(layout, M, k, ch);
                    const float coef = row[idx];
                    const float sdl = sh_sign(flip, k, dl[ch]);      // the term's sign rides on dL/dcolour
                    const float val = Y * sdl;
                    grow[idx] = first ? val : grow[idx] + val;
                    const float cd = coef * sdl;
                    dLda += Ya * cd; dLdb += Yb * cd; dLdc += Yc * cd;
                }
            });
            const float3 dLd = sh_grad_unpermute(d.sh_basis, dLda, dLdb, dLdc);
            const float dLdx = dLd.x, dLdy = dLd.y, dLdz = dLd.z;
            if (first) {
                const int nb = (d.deg + 1) * (d.deg + 1);
                for (int k = nb; k < M; ++k)
                    for (int ch = 0; ch < 3; ++ch) grow[sh_index(layout, M, k, ch)] = 0.0f;
            }
            sh_written = true;
            const float inv3 = 1.0f / (len2 * len);
            gx += ((len2 - ddx * ddx) * dLdx - ddy * ddx * dLdy - ddz * ddx * dLdz) * inv3;
            gy += (-ddx * ddy * dLdx + (len2 - ddy * ddy) * dLdy - ddz * ddy * dLdz) * inv3;
            gz += (-ddx * ddz * dLdx - ddy * ddz * dLdy + (len2 - ddz * ddz) * dLdz) * inv3;
        } else if (vis) {
            dcol[0] += gcol.x; dcol[1] += gcol.y; dcol[2] += gcol.z;
        }
        dmx += gx * sc; dmy += gy * sc; dmz += gz * sc;
    }

    if (live) {
        out.d_means[3 * sg + 0] = dmx; out.d_means[3 * sg + 1] = dmy; out.d_means[3 * sg + 2] = dmz;
        out.d_opacities[sg] = dop;
        float *dc = out.d_cov + sg * cov_n;
        if (d.cov_layout == PS_COV_TRIU6) {
#pragma unroll
            for (int i = 0; i < 6; ++i) dc[i] = dcov[i];
        } else {
            dc[0] = dcov[0]; dc[1] = dcov[1]; dc[2] = dcov[2];
            dc[4] = dcov[3]; dc[5] = dcov[4]; dc[8] = dcov[5];      // lower triangle stays zero (pre-filled)
        }
        if (M == 0) {
            float *dsh = out.d_sh + sg * 3;
            dsh[0] = dcol[0]; dsh[1] = dcol[1]; dsh[2] = dcol[2];
        }
    }
    if (M > 0) {
        // every listed Gaussian is visible in >= 1 view, so its gradient row was written above
        __syncwarp();
        const float *gsrc = in_place ? wrows : wrows + (size_t)kPreBwdThreads * row_stride;
        scatter_rows(out.d_sh, (unsigned long long)sg, rows_valid, sh_n, gsrc, row_stride, lane);
    }
}

// Zero gradients for everything the dense kernel does not write (Gaussians that are on screen in
// no view, the lower covariance triangle, optional screen-space gradients).  Pure memsets.
int launch_gradient_fill(const Dims &d, const ps_raster_grads &out, cudaStream_t st) {
    const size_t sp = (size_t)d.S * d.P;
    PS_CUDA_CHECK(cudaMemsetAsync(out.d_means, 0, sp * 3 * sizeof(float), st));
    PS_CUDA_CHECK(cudaMemsetAsync(out.d_cov, 0, sp * (d.cov_layout == PS_COV_TRIU6 ? 6 : 9) * sizeof(float), st));
    PS_CUDA_CHECK(cudaMemsetAsync(out.d_opacities, 0, sp * sizeof(float), st));
    PS_CUDA_CHECK(cudaMemsetAsync(out.d_sh, 0, sp * (d.M > 0 ? 3 * d.M : 3) * sizeof(float), st));
    if (out.d_means2d)
        PS_CUDA_CHECK(cudaMemsetAsync(out.d_means2d, 0, (size_t)d.S * d.V * d.P * 3 * sizeof(float), st));
    return PS_OK;
}

int launch_preprocess_backward(const Dims &d, const Inputs &in, const Geom &g, const ViewGrads &vg,
                               const ps_raster_grads &out, cudaStream_t st) {
    const int row_stride = d.M > 0 ? ((3 * d.M) | 1) : 1;   // odd word count: conflict-free per-lane rows
    const size_t smem = d.M > 0 ? sizeof(float) * kPreBwdThreads * row_stride * (d.V == 1 ? 1 : 2) : 0;
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices)) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_preprocess_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    const long long sp = (long long)d.S * d.P;               // worst case; surplus warps exit at once
    k_preprocess_bwd<<<(unsigned)((sp + kPreBwdThreads - 1) / kPreBwdThreads), kPreBwdThreads, smem, st>>>(
        d, in, g, vg, out, row_stride);
    PS_LAUNCH_CHECK("k_preprocess_bwd");
    return PS_OK;
}

}  // namespace ps
