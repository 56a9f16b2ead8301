// Per-Gaussian projection math shared by the preprocess forward and backward kernels.
// Semantics: SURVEY.md Appendix A.1 / A.4 / A.5 (EWA splatting of a 3D covariance with a
// 0.3 px low-pass, 3-sigma radius, real SH up to degree 4 in the 3DGS sign convention).
//
// The expression shapes (operand order, no re-association) are deliberate: the forward kernel
// is compiled with --fmad=false so that depth bits, radii and tile rectangles are IEEE-exact
// and can be checked bit-for-bit ("bit-exact tile/bin indices").
#pragma once
#include <cuda_runtime.h>

namespace ps {

constexpr float kShC0 = 0.28209479177387814f;
constexpr float kShC1 = 0.4886025119029199f;

struct Cov2D {
    float tx, ty, tz;    // view-space mean
    float ctx, cty;      // after the +-1.3 tan(fov) clamp
    bool clamp_x, clamp_y;
    float m0[3], m1[3];  // rows of M = J * R  (2x3)
    float a, b, c;       // 2D covariance + 0.3 I
};

__device__ __forceinline__ void load_cov6(const float *cov, int layout, float scale2, float s[6]) {
    if (layout == PS_COV_TRIU6) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = cov[i];
    } else {
        s[0] = cov[0]; s[1] = cov[1]; s[2] = cov[2]; s[3] = cov[4]; s[4] = cov[5]; s[5] = cov[8];
    }
    if (scale2 != 1.0f) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = s[i] * scale2;
    }
}

__device__ __forceinline__ void compute_cov2d(float px, float py, float pz, const float s[6],
                                              const float *__restrict__ vm, float focal_x,
                                              float focal_y, float tanfovx, float tanfovy,
                                              Cov2D &o) {
    o.tx = vm[0] * px + vm[4] * py + vm[8] * pz + vm[12];
    o.ty = vm[1] * px + vm[5] * py + vm[9] * pz + vm[13];
    o.tz = vm[2] * px + vm[6] * py + vm[10] * pz + vm[14];
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = o.tx / o.tz, tytz = o.ty / o.tz;
    o.clamp_x = (txtz < -limx) || (txtz > limx);
    o.clamp_y = (tytz < -limy) || (tytz > limy);
    o.ctx = fminf(limx, fmaxf(-limx, txtz)) * o.tz;
    o.cty = fminf(limy, fmaxf(-limy, tytz)) * o.tz;
    const float tz = o.tz;
    const float j00 = focal_x / tz;
    const float j02 = -(focal_x * o.ctx) / (tz * tz);
    const float j11 = focal_y / tz;
    const float j12 = -(focal_y * o.cty) / (tz * tz);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.m0[j] = j00 * vm[4 * j + 0] + j02 * vm[4 * j + 2];
        o.m1[j] = j11 * vm[4 * j + 1] + j12 * vm[4 * j + 2];
    }
    const float sxx = s[0], sxy = s[1], sxz = s[2], syy = s[3], syz = s[4], szz = s[5];
    const float v0x = sxx * o.m0[0] + sxy * o.m0[1] + sxz * o.m0[2];
    const float v0y = sxy * o.m0[0] + syy * o.m0[1] + syz * o.m0[2];
    const float v0z = sxz * o.m0[0] + syz * o.m0[1] + szz * o.m0[2];
    const float v1x = sxx * o.m1[0] + sxy * o.m1[1] + sxz * o.m1[2];
    const float v1y = sxy * o.m1[0] + syy * o.m1[1] + syz * o.m1[2];
    const float v1z = sxz * o.m1[0] + syz * o.m1[1] + szz * o.m1[2];
    o.a = o.m0[0] * v0x + o.m0[1] * v0y + o.m0[2] * v0z + 0.3f;
    o.b = o.m0[0] * v1x + o.m0[1] * v1y + o.m0[2] * v1z;
    o.c = o.m1[0] * v1x + o.m1[1] * v1y + o.m1[2] * v1z + 0.3f;
}

// Real SH basis for a unit direction; entries >= (deg+1)^2 are left untouched.
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float b[25]) {
    b[0] = kShC0;
    if (deg < 1) return;
    b[1] = -kShC1 * y;
    b[2] = kShC1 * z;
    b[3] = -kShC1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
    b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    if (deg < 4) return;
    b[16] = 2.5033429417967046f * xy * (xx - yy);
    b[17] = -1.7701307697799304f * yz * (3.0f * xx - yy);
    b[18] = 0.9461746957575601f * xy * (7.0f * zz - 1.0f);
    b[19] = -0.6690465435572892f * yz * (7.0f * zz - 3.0f);
    b[20] = 0.10578554691520431f * (zz * (35.0f * zz - 30.0f) + 3.0f);
    b[21] = -0.6690465435572892f * xz * (7.0f * zz - 3.0f);
    b[22] = 0.47308734787878004f * (xx - yy) * (7.0f * zz - 1.0f);
    b[23] = -1.7701307697799304f * xz * (xx - 3.0f * yy);
    b[24] = 0.6258357354491761f * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
}

// d basis / d(x, y, z) with x, y, z treated as independent variables.
__device__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float dx[25],
                                              float dy[25], float dz[25]) {
#pragma unroll
    for (int i = 0; i < 25; ++i) dx[i] = dy[i] = dz[i] = 0.0f;
    if (deg < 1) return;
    dy[1] = -kShC1;
    dz[2] = kShC1;
    dx[3] = -kShC1;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    {
        const float c0 = 1.0925484305920792f, c2 = 0.31539156525252005f, c4 = 0.5462742152960396f;
        dx[4] = c0 * y;          dy[4] = c0 * x;
        dy[5] = -c0 * z;         dz[5] = -c0 * y;
        dx[6] = -2.0f * c2 * x;  dy[6] = -2.0f * c2 * y;  dz[6] = 4.0f * c2 * z;
        dx[7] = -c0 * z;         dz[7] = -c0 * x;
        dx[8] = 2.0f * c4 * x;   dy[8] = -2.0f * c4 * y;
    }
    if (deg < 3) return;
    {
        const float c0 = -0.5900435899266435f, c1 = 2.890611442640554f, c2 = -0.4570457994644658f,
                    c3 = 0.3731763325901154f, c5 = 1.445305721320277f;
        dx[9] = c0 * 6.0f * xy;                       dy[9] = c0 * (3.0f * xx - 3.0f * yy);
        dx[10] = c1 * yz;  dy[10] = c1 * xz;          dz[10] = c1 * xy;
        dx[11] = c2 * -2.0f * xy;  dy[11] = c2 * (4.0f * zz - xx - 3.0f * yy);  dz[11] = c2 * 8.0f * yz;
        dx[12] = c3 * -6.0f * xz;  dy[12] = c3 * -6.0f * yz;  dz[12] = c3 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        dx[13] = c2 * (4.0f * zz - 3.0f * xx - yy);  dy[13] = c2 * -2.0f * xy;  dz[13] = c2 * 8.0f * xz;
        dx[14] = c5 * 2.0f * xz;   dy[14] = c5 * -2.0f * yz;  dz[14] = c5 * (xx - yy);
        dx[15] = c0 * (3.0f * xx - 3.0f * yy);       dy[15] = c0 * -6.0f * xy;
    }
    if (deg < 4) return;
    {
        const float c0 = 2.5033429417967046f, c1 = -1.7701307697799304f, c2 = 0.9461746957575601f,
                    c3 = -0.6690465435572892f, c4 = 0.10578554691520431f, c6 = 0.47308734787878004f,
                    c8 = 0.6258357354491761f;
        dx[16] = c0 * (3.0f * xx * y - yy * y);       dy[16] = c0 * (xx * x - 3.0f * x * yy);
        dx[17] = c1 * 6.0f * xy * z;  dy[17] = c1 * z * (3.0f * xx - 3.0f * yy);  dz[17] = c1 * y * (3.0f * xx - yy);
        dx[18] = c2 * y * (7.0f * zz - 1.0f);  dy[18] = c2 * x * (7.0f * zz - 1.0f);  dz[18] = c2 * 14.0f * xy * z;
        dy[19] = c3 * z * (7.0f * zz - 3.0f);  dz[19] = c3 * y * (21.0f * zz - 3.0f);
        dz[20] = c4 * (140.0f * zz * z - 60.0f * z);
        dx[21] = c3 * z * (7.0f * zz - 3.0f);  dz[21] = c3 * x * (21.0f * zz - 3.0f);
        dx[22] = c6 * 2.0f * x * (7.0f * zz - 1.0f);  dy[22] = c6 * -2.0f * y * (7.0f * zz - 1.0f);  dz[22] = c6 * 14.0f * z * (xx - yy);
        dx[23] = c1 * z * (3.0f * xx - 3.0f * yy);  dy[23] = c1 * -6.0f * xy * z;  dz[23] = c1 * x * (xx - 3.0f * yy);
        dx[24] = c8 * (4.0f * xx * x - 12.0f * x * yy);  dy[24] = c8 * (4.0f * yy * y - 12.0f * xx * y);
    }
}

__device__ __forceinline__ int sh_index(int layout, int M, int k, int ch) {
    return layout == PS_SH_M3 ? k * 3 + ch : ch * M + k;
}

}  // namespace ps
