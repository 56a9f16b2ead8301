"""Synthetic scenes that stand in for re10k (no dataset or checkpoint is reachable offline).

Shapes, camera statistics and the Gaussian parameterisation mirror what the reference feeds its
rasterizer:
  * intrinsics: normalised K with fx = fy = 0.88 after the 256x256 crop (crop_shim.py:25-82);
  * two context cameras one baseline apart (dataset_re10k.py:143-152), near/far from the bounds
    shim formula (bounds_shim.py:9-37: disparity 3*256 px / 0.5 px -> near 0.293, far 450.6);
  * one Gaussian triple per context pixel: depth drawn in relative disparity
    (depth_predictor_monocular.py:50-68), opacity = bucket probability / gaussians_per_pixel
    (encoder_epipolar.py:97-110,170), scale = (0.5 + 14.5*sigmoid)*depth*0.1*(px_x + px_y),
    unit quaternion -> R S S^T R^T (gaussian_adapter.py:60-95, gaussians.py:33-44),
    SH = N(0,1) * sh_mask with sh_mask[l>=1] = 0.1*0.25^l (gaussian_adapter.py:41-46).
Everything is generated on the CPU with a fixed seed; callers move tensors where they need them.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor

RE10K_FOCAL = 0.88


@dataclass
class Scene:
    means: Tensor          # [P,3]
    covariances: Tensor    # [P,3,3]
    harmonics: Tensor      # [P,3,d_sh]
    opacities: Tensor      # [P]
    extrinsics: Tensor     # [T,4,4] camera-to-world (OpenCV), target views
    intrinsics: Tensor     # [T,3,3] normalised
    near: Tensor           # [T]
    far: Tensor            # [T]
    image_shape: tuple[int, int]
    background: Tensor     # [3]

    @property
    def num_gaussians(self) -> int:
        return self.means.shape[0]


def intrinsics_re10k(n: int = 1) -> Tensor:
    k = torch.eye(3, dtype=torch.float32)
    k[0, 0] = k[1, 1] = RE10K_FOCAL
    k[0, 2] = k[1, 2] = 0.5
    return k[None].repeat(n, 1, 1)


def bounds_from_baseline(baseline: float, h: int, w: int, near_disparity_px: float,
                         far_disparity_px: float) -> tuple[float, float]:
    """depth = baseline / (disparity * mean pixel size at depth 1)."""
    pixel = 0.5 * ((1.0 / w) / RE10K_FOCAL + (1.0 / h) / RE10K_FOCAL)
    return baseline / (near_disparity_px * pixel), baseline / (far_disparity_px * pixel)


def _quat_to_rot(q: Tensor) -> Tensor:
    i, j, k, r = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack([
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(*q.shape[:-1], 3, 3)


def sh_mask(sh_degree: int) -> Tensor:
    m = torch.ones((sh_degree + 1) ** 2)
    for l in range(1, sh_degree + 1):
        m[l * l:(l + 1) * (l + 1)] = 0.1 * 0.25 ** l
    return m


def _pixel_rays(h: int, w: int, c2w: Tensor, k: Tensor) -> tuple[Tensor, Tensor]:
    ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
    xy1 = torch.stack([xs, ys, torch.ones_like(xs)], -1).reshape(-1, 3)
    d = xy1 @ torch.linalg.inv(k).T
    d = d / d.norm(dim=-1, keepdim=True)
    d = d @ c2w[:3, :3].T
    return c2w[:3, 3].expand_as(d), d


def target_cameras(num_views: int, seed: int, baseline: float = 1.0) -> Tensor:
    """Target cameras between the two context cameras (x in (0, baseline)), small jitter."""
    g = torch.Generator().manual_seed(seed + 7919)
    ext = torch.eye(4)[None].repeat(num_views, 1, 1)
    for t in range(num_views):
        ext[t, 0, 3] = baseline * (t + 1) / (num_views + 1)
        if num_views > 1:
            ext[t, :3, 3] += 0.02 * baseline * torch.randn(3, generator=g)
    return ext


def scene_re10k_like(seed: int = 0, image_hw: tuple[int, int] = (256, 256), context_views: int = 2,
                     gaussians_per_pixel: int = 3, sh_degree: int = 4, target_views: int = 1,
                     num_buckets: int = 32, logit_scale: float = 2.0) -> Scene:
    """configs[1] (256x256, 2 context views, 3 gpp) and configs[4] (512x512, 3 views)."""
    h, w = image_hw
    g = torch.Generator().manual_seed(seed)
    near, far = bounds_from_baseline(1.0, h, w, 3.0 * min(h, w), 0.5)
    k = intrinsics_re10k(1)[0]
    d_sh = (sh_degree + 1) ** 2
    mask = sh_mask(sh_degree)
    means, covs, shs, opacs = [], [], [], []
    for v in range(context_views):
        c2w = torch.eye(4)
        c2w[0, 3] = float(v) / max(context_views - 1, 1)
        origins, dirs = _pixel_rays(h, w, c2w, k)
        r = origins.shape[0]
        # jitter the ray inside its pixel like the (sigmoid - 0.5) * pixel offset
        logits = logit_scale * torch.randn(r, num_buckets, generator=g)
        pdf = logits.softmax(-1)
        bucket = torch.multinomial(pdf, gaussians_per_pixel, replacement=True, generator=g)
        prob = pdf.gather(-1, bucket)
        u = (bucket + torch.rand(bucket.shape, generator=g)) / num_buckets  # relative disparity
        disp_near, disp_far = 1.0 / near, 1.0 / far
        depth = 1.0 / ((1.0 - u) * (disp_near - disp_far) + disp_far)
        mean = origins[:, None] + dirs[:, None] * depth[..., None]
        raw_scale = torch.randn(r, gaussians_per_pixel, 3, generator=g)
        mult = 0.1 * ((1.0 / w) / RE10K_FOCAL + (1.0 / h) / RE10K_FOCAL)
        scale = (0.5 + 14.5 * raw_scale.sigmoid()) * depth[..., None] * mult
        q = torch.randn(r, gaussians_per_pixel, 4, generator=g)
        q = q / q.norm(dim=-1, keepdim=True)
        rot = _quat_to_rot(q)
        s = torch.diag_embed(scale)
        cov = rot @ s @ s.transpose(-1, -2) @ rot.transpose(-1, -2)
        sh = torch.randn(r, gaussians_per_pixel, 3, d_sh, generator=g) * mask
        means.append(mean.reshape(-1, 3))
        covs.append(cov.reshape(-1, 3, 3))
        shs.append(sh.reshape(-1, 3, d_sh))
        opacs.append((prob / gaussians_per_pixel).reshape(-1))
    ext = target_cameras(target_views, seed)
    return Scene(torch.cat(means), torch.cat(covs), torch.cat(shs), torch.cat(opacs), ext,
                 intrinsics_re10k(target_views), torch.full((target_views,), near),
                 torch.full((target_views,), far), (h, w), torch.zeros(3))


def scene_random_frustum(seed: int = 0, image_hw: tuple[int, int] = (64, 64),
                         num_gaussians: int = 1000, sh_degree: int = 4,
                         background: tuple[float, float, float] = (0.0, 0.0, 0.0),
                         z_range: tuple[float, float] = (1.0, 20.0)) -> Scene:
    """configs[0]: Gaussians uniform in the frustum of one identity camera, random SPD
    covariances (screen sigma 0.5-3 px), opacity U(0.05, 0.9), SH ~ N(0,1)*sh_mask."""
    h, w = image_hw
    g = torch.Generator().manual_seed(seed)
    p = num_gaussians
    z = z_range[0] + (z_range[1] - z_range[0]) * torch.rand(p, generator=g)
    uv = -0.1 + 1.2 * torch.rand(p, 2, generator=g)   # slightly wider than the image
    xy = (uv - 0.5) / RE10K_FOCAL * z[:, None]
    means = torch.cat([xy, z[:, None]], -1)
    sigma_px = 0.5 + 2.5 * torch.rand(p, 3, generator=g)
    sigma = sigma_px * z[:, None] / (RE10K_FOCAL * w)
    q = torch.randn(p, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    rot = _quat_to_rot(q)
    s = torch.diag_embed(sigma)
    cov = rot @ s @ s.transpose(-1, -2) @ rot.transpose(-1, -2)
    d_sh = (sh_degree + 1) ** 2
    sh = torch.randn(p, 3, d_sh, generator=g) * sh_mask(sh_degree)
    opac = 0.05 + 0.85 * torch.rand(p, generator=g)
    ext = torch.eye(4)[None]
    return Scene(means, cov, sh, opac, ext, intrinsics_re10k(1), torch.tensor([0.5]),
                 torch.tensor([100.0]), (h, w), torch.tensor(background, dtype=torch.float32))
