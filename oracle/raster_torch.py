"""Pure-PyTorch, differentiable CPU restatement of the tile rasterizer (one view).

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.c for the scope rules).  PARITY UNPINNED for
the same reason as the C oracle.  Two uses:
  * in float64, autograd through this file pins the hand-derived backward of the C oracle
    (tests/test_oracle_raster.py);
  * in float32 it is BASELINE.json's "pure-PyTorch CPU alpha-composite reference" (configs[0])
    and the `cpu_baseline` / `--impl reference` arm of bench.py.

Follows SURVEY.md Appendix A; argument conventions from the reference call site
src/model/decoder/cuda_splatting.py:99-124.  Upstream backward quirks are mirrored explicitly
(SURVEY A.6): straight-through min(0.99, .), hard skip masks, the terminating Gaussian
contributes nothing, the +-1.3 tanfov clamp kills d/dt.x (d/dt.y) only.
"""
from __future__ import annotations

import math

import torch

TILE = 16
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
         0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]
SH_C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
         0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
         0.6258357354491761]


SH_CONVENTION = 0   # 0 = 3DGS basis (default), 1 = e3nn basis (include/pixelsplat_b200.h PS_SH_BASIS_*)


def set_sh_basis(convention: int) -> None:
    global SH_CONVENTION
    SH_CONVENTION = int(convention)


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """d: [P,3] unit vectors -> [P,(deg+1)^2].  With SH_CONVENTION = 1 the e3nn basis:
    Y_e3nn,k(x, y, z) = (-1)^k Y_3dgs,k(z, x, y) (the basis the reference's rotate_sh rotates in,
    /root/reference/src/misc/sh_rotation.py:18-22)."""
    if SH_CONVENTION == 1:
        x, y, z = d.unbind(-1)
        sign = torch.tensor([(-1.0) ** k for k in range((deg + 1) ** 2)], dtype=d.dtype, device=d.device)
        return _sh_basis_3dgs(deg, torch.stack([z, x, y], dim=-1)) * sign
    return _sh_basis_3dgs(deg, d)


def _sh_basis_3dgs(deg: int, d: torch.Tensor) -> torch.Tensor:
    x, y, z = d.unbind(-1)
    b = [torch.full_like(x, SH_C0)]
    if deg >= 1:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz,
              SH_C2[4] * (xx - yy)]
    if deg >= 3:
        b += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
              SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    if deg >= 4:
        b += [SH_C4[0] * xy * (xx - yy), SH_C4[1] * yz * (3 * xx - yy), SH_C4[2] * xy * (7 * zz - 1),
              SH_C4[3] * yz * (7 * zz - 3), SH_C4[4] * (zz * (35 * zz - 30) + 3),
              SH_C4[5] * xz * (7 * zz - 3), SH_C4[6] * (xx - yy) * (7 * zz - 1),
              SH_C4[7] * xz * (xx - 3 * yy), SH_C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(b, dim=-1)


def _st(value: torch.Tensor, surrogate: torch.Tensor) -> torch.Tensor:
    """value in the forward pass, gradient of `surrogate` in the backward pass."""
    return surrogate + (value - surrogate).detach()


def preprocess(means, cov6, opac, sh, colors, vm, pm, campos, tanfovx, tanfovy, W, H, sh_degree):
    """All tensors torch (any float dtype). vm/pm: [16] column-major. Returns a dict."""
    dt = means.dtype
    P = means.shape[0]
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    px, py, pz = means.unbind(-1)
    tx = vm[0] * px + vm[4] * py + vm[8] * pz + vm[12]
    ty = vm[1] * px + vm[5] * py + vm[9] * pz + vm[13]
    tz = vm[2] * px + vm[6] * py + vm[10] * pz + vm[14]
    in_front = tz > 0.2
    tz_s = torch.where(in_front, tz, torch.ones_like(tz))  # keep culled lanes finite
    hx = pm[0] * px + pm[4] * py + pm[8] * pz + pm[12]
    hy = pm[1] * px + pm[5] * py + pm[9] * pz + pm[13]
    hw = pm[3] * px + pm[7] * py + pm[11] * pz + pm[15]
    p_w = 1.0 / (hw + 1e-7)
    projx, projy = hx * p_w, hy * p_w
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = tx / tz_s, ty / tz_s
    cl_x = (txtz < -limx) | (txtz > limx)
    cl_y = (tytz < -limy) | (tytz > limy)
    ctx_v = txtz.clamp(-limx, limx) * tz_s
    cty_v = tytz.clamp(-limy, limy) * tz_s
    ctx = _st(ctx_v, torch.where(cl_x, ctx_v.detach(), tx))
    cty = _st(cty_v, torch.where(cl_y, cty_v.detach(), ty))
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    j00, j02 = fx / tz_s, -(fx * ctx) / (tz_s * tz_s)
    j11, j12 = fy / tz_s, -(fy * cty) / (tz_s * tz_s)
    m0 = [j00 * vm[4 * j + 0] + j02 * vm[4 * j + 2] for j in range(3)]
    m1 = [j11 * vm[4 * j + 1] + j12 * vm[4 * j + 2] for j in range(3)]
    sxx, sxy, sxz, syy, syz, szz = cov6.unbind(-1)
    v0 = [sxx * m0[0] + sxy * m0[1] + sxz * m0[2], sxy * m0[0] + syy * m0[1] + syz * m0[2],
          sxz * m0[0] + syz * m0[1] + szz * m0[2]]
    v1 = [sxx * m1[0] + sxy * m1[1] + sxz * m1[2], sxy * m1[0] + syy * m1[1] + syz * m1[2],
          sxz * m1[0] + syz * m1[1] + szz * m1[2]]
    a = m0[0] * v0[0] + m0[1] * v0[1] + m0[2] * v0[2] + 0.3
    b = m0[0] * v1[0] + m0[1] * v1[1] + m0[2] * v1[2]
    c = m1[0] * v1[0] + m1[1] * v1[1] + m1[2] * v1[2] + 0.3
    det = a * c - b * b
    det_ok = det != 0
    det_s = torch.where(det_ok, det, torch.ones_like(det))
    conic = torch.stack([c / det_s, -b / det_s, a / det_s], dim=-1)
    mid = 0.5 * (a + c)
    sq = (mid * mid - det).clamp(min=0.1).sqrt()
    radius = torch.ceil(3.0 * torch.maximum(mid + sq, mid - sq).sqrt()).detach()
    pixx = ((projx + 1.0) * W - 1.0) * 0.5
    pixy = ((projy + 1.0) * H - 1.0) * 0.5
    r = radius.to(torch.int64).to(dt)

    def tile(v, g):  # C cast: truncation toward zero, then clamp to [0, g]
        return torch.trunc(v / TILE).clamp(0, g).to(torch.int64)

    with torch.no_grad():
        minx, miny = tile(pixx - r, gx), tile(pixy - r, gy)
        maxx, maxy = tile(pixx + r + (TILE - 1), gx), tile(pixy + r + (TILE - 1), gy)
        area = (maxx - minx) * (maxy - miny)
        visible = in_front & det_ok & (area > 0)
    if sh is not None:
        d = means - campos
        d = d / d.norm(dim=-1, keepdim=True)
        nb = (sh_degree + 1) ** 2
        basis = sh_basis(sh_degree, d)  # [P, nb]
        raw = (basis[:, :, None] * sh[:, :nb, :]).sum(dim=1) + 0.5
        clamped = raw < 0
        rgb = raw.clamp(min=0.0)
    else:
        rgb = colors
        clamped = torch.zeros_like(colors, dtype=torch.bool)
    return dict(depth=tz, radii=torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32),
                xy=torch.stack([pixx, pixy], -1), conic=conic, opacity=opac.reshape(-1), rgb=rgb,
                clamped=clamped, rect=torch.stack([minx, miny, maxx, maxy], -1), visible=visible,
                tiles_touched=torch.where(visible, area, torch.zeros_like(area)))


def bin_tiles(pre, W, H):
    """Returns (keys uint64-as-int64 [N], values int64 [N], ranges [tiles,2])."""
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    vis = pre["visible"].nonzero().squeeze(-1)
    rect = pre["rect"][vis]
    depth_bits = pre["depth"][vis].detach().to(torch.float32).view(torch.int32).to(torch.int64)
    ids, tiles = [], []
    w = rect[:, 2] - rect[:, 0]
    cnt = pre["tiles_touched"][vis]
    rep = torch.repeat_interleave(torch.arange(vis.numel()), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(rep.numel()) - start[rep]
    ty = rect[rep, 1] + local // w[rep]
    tx = rect[rep, 0] + local % w[rep]
    tile_id = ty * gx + tx
    keys = (tile_id << 32) | depth_bits[rep]
    order = torch.sort(keys, stable=True).indices
    keys, values = keys[order], vis[rep][order]
    tid = keys >> 32
    counts = torch.bincount(tid, minlength=gx * gy)
    ends = torch.cumsum(counts, 0)
    ranges = torch.stack([ends - counts, ends], -1)
    return keys, values, ranges


def composite(pre, values, ranges, bg, W, H):
    """Differentiable front-to-back composite. Returns color [3,H,W], final_T [H,W], n_contrib."""
    dt = pre["xy"].dtype
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    color = torch.zeros(3, H, W, dtype=dt) + bg.reshape(3, 1, 1) * torch.ones(1, H, W, dtype=dt)
    final_T = torch.ones(H, W, dtype=dt)
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    out_tiles = {}
    for t in range(gx * gy):
        s, e = int(ranges[t, 0]), int(ranges[t, 1])
        y0, x0 = (t // gx) * TILE, (t % gx) * TILE
        y1, x1 = min(y0 + TILE, H), min(x0 + TILE, W)
        if e == s:
            continue
        g = values[s:e]
        ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt),
                                indexing="ij")
        pxf, pyf = xs.reshape(-1, 1), ys.reshape(-1, 1)          # [pix,1]
        dx = pre["xy"][g, 0][None] - pxf                          # [pix,n]
        dy = pre["xy"][g, 1][None] - pyf
        A, B, C = pre["conic"][g].unbind(-1)
        power = -0.5 * (A[None] * dx * dx + C[None] * dy * dy) - B[None] * dx * dy
        G = torch.exp(power.clamp(max=0.0))
        a_raw = pre["opacity"][g][None] * G
        alpha = _st(a_raw.clamp(max=0.99), a_raw)
        keep = ((power <= 0) & (alpha >= 1.0 / 255.0)).detach()
        alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - alpha
        T_incl = torch.cumprod(one_m, dim=1)
        alive = (T_incl >= 0.0001).detach()   # monotone: false from the terminating Gaussian on
        alpha = torch.where(alive, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - alpha
        T_incl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        w = alpha * T_excl                                          # [pix,n]
        col = w @ pre["rgb"][g]                                     # [pix,3]
        Tf = T_incl[:, -1]
        contrib = (keep & alive)
        idx = torch.arange(1, e - s + 1)[None].expand_as(contrib)
        last = torch.where(contrib, idx, torch.zeros_like(idx)).max(dim=1).values
        out_tiles[t] = (col + Tf[:, None] * bg[None], Tf, last, (y0, y1, x0, x1))
    if out_tiles:
        color = color.clone()
        for t, (col, Tf, last, (y0, y1, x0, x1)) in out_tiles.items():
            hh, ww = y1 - y0, x1 - x0
            color[:, y0:y1, x0:x1] = col.T.reshape(3, hh, ww)
            final_T[y0:y1, x0:x1] = Tf.detach().reshape(hh, ww)
            n_contrib[y0:y1, x0:x1] = last.reshape(hh, ww)
    return color, final_T, n_contrib


def rasterize(means, cov6, opac, sh, colors, vm, pm, campos, tanfovx, tanfovy, bg, W, H,
              sh_degree):
    """Full single-view forward; differentiable w.r.t. means, cov6, opac, sh/colors."""
    pre = preprocess(means, cov6, opac, sh, colors, vm, pm, campos, tanfovx, tanfovy, W, H,
                     sh_degree)
    keys, values, ranges = bin_tiles(pre, W, H)
    color, final_T, n_contrib = composite(pre, values, ranges, bg, W, H)
    return color, dict(pre=pre, keys=keys, values=values, ranges=ranges, final_T=final_T,
                       n_contrib=n_contrib)


def get_projection_matrix(near, far, tanfovx, tanfovy):
    """Row-major projection as built by cuda_splatting.py:17-44 (z in [0,1], +z forward)."""
    m = torch.zeros(4, 4, dtype=torch.float64)
    top, right = tanfovy * near, tanfovx * near
    m[0, 0] = 2 * near / (2 * right)
    m[1, 1] = 2 * near / (2 * top)
    m[3, 2] = 1
    m[2, 2] = far / (far - near)
    m[2, 3] = -(far * near) / (far - near)
    return m


def camera_from_c2w(c2w: torch.Tensor, K: torch.Tensor, near: float, far: float, dtype):
    """Restates cuda_splatting.py:80-87 for one camera: returns (vm[16], pm[16], campos[3],
    tanfovx, tanfovy) with the column-major flattening the rasterizer consumes."""
    c2w = c2w.to(torch.float64)
    Kinv = torch.linalg.inv(K.to(torch.float64))

    def unit(v):
        v = Kinv @ torch.tensor(v, dtype=torch.float64)
        return v / v.norm()

    fov_x = torch.acos((unit([0, 0.5, 1]) * unit([1, 0.5, 1])).sum())
    fov_y = torch.acos((unit([0.5, 0, 1]) * unit([0.5, 1, 1])).sum())
    tx, ty = math.tan(0.5 * float(fov_x)), math.tan(0.5 * float(fov_y))
    proj = get_projection_matrix(near, far, tx, ty)
    view_t = torch.linalg.inv(c2w).T            # "b i j -> b j i" of extrinsics.inverse()
    full_t = view_t @ proj.T
    return (view_t.reshape(16).to(dtype), full_t.reshape(16).to(dtype), c2w[:3, 3].to(dtype), tx, ty)


def prepare_view(means, covariances, harmonics, opacities, extrinsics, intrinsics, near, far,
                 dtype=torch.float32, scale_invariant=True, use_sh=True):
    """Restates the host side of render_cuda (cuda_splatting.py:61-87, 115-123) for ONE view, in
    `dtype` on the CPU: scale normalisation, SH permutation to [P, M, 3], triu covariance packing,
    fov / view / projection matrices.  Returns the rasterizer's argument dict."""
    means, covariances = means.to(dtype), covariances.to(dtype)
    harmonics, opacities = harmonics.to(dtype), opacities.to(dtype)
    extrinsics, intrinsics = extrinsics.to(dtype).clone(), intrinsics.to(dtype)
    near, far = torch.as_tensor(near, dtype=dtype), torch.as_tensor(far, dtype=dtype)
    if scale_invariant:
        scale = 1 / near
        extrinsics[:3, 3] = extrinsics[:3, 3] * scale
        covariances = covariances * (scale ** 2)
        means = means * scale
        near, far = near * scale, far * scale
    d_sh = harmonics.shape[-1]
    degree = math.isqrt(d_sh) - 1
    shs = harmonics.permute(0, 2, 1).contiguous()  # "g xyz n -> g n xyz"
    kinv = intrinsics.inverse()

    def unit(v):
        v = kinv @ torch.tensor(v, dtype=dtype)
        return v / v.norm()

    fov_x = (unit([0, 0.5, 1]) * unit([1, 0.5, 1])).sum().acos()
    fov_y = (unit([0.5, 0, 1]) * unit([0.5, 1, 1])).sum().acos()
    tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    top, right = tan_y * near, tan_x * near
    proj = torch.zeros(4, 4, dtype=dtype)
    proj[0, 0] = 2 * near / (right - (-right))
    proj[1, 1] = 2 * near / (top - (-top))
    proj[3, 2] = 1
    proj[2, 2] = far / (far - near)
    proj[2, 3] = -(far * near) / (far - near)
    view_t = extrinsics.inverse().T
    full_t = view_t @ proj.T
    row, col = torch.triu_indices(3, 3)
    return dict(means=means, cov6=covariances[:, row, col].contiguous(),
                opac=opacities, sh=shs if use_sh else None,
                colors=None if use_sh else shs[:, 0, :].contiguous(),
                vm=view_t.reshape(16).contiguous(), pm=full_t.reshape(16).contiguous(),
                campos=extrinsics[:3, 3].contiguous(), tanfovx=float(tan_x), tanfovy=float(tan_y),
                sh_degree=degree)
