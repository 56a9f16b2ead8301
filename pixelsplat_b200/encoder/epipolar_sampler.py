"""Drop-in for /root/reference/src/model/encoder/epipolar/epipolar_sampler.py:19-166.

`EpipolarSampler.forward` runs ONE CUDA kernel (ps_epipolar_geometry) instead of ~60 elementwise
torch kernels, 16 boolean-mask scatters and a grid_sample that materialises a
[b, v, ov, r, s, c] tensor (0.94 GB at configs[2]).  The returned `EpipolarSampling` exposes the
reference's fields; the expensive ones (`features`, and the per-sample coordinate tensors) are
produced lazily, because the training step never reads them -- only the visualisers do
(encoder_epipolar.py:186-187).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from .attention_fused import EpipolarGeometry, EpipolarKV, epipolar_geometry
from .heterogeneous_pairings import (generate_heterogeneous_index,
                                     generate_heterogeneous_index_transpose)


def sample_image_grid(shape: tuple[int, int], device) -> Tensor:
    """[h, w, 2] pixel-centre coordinates in (0, 1), xy order (projection.py:117-137)."""
    h, w = shape
    ys = (torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h
    xs = (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1)


def get_world_rays(xy: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> tuple[Tensor, Tensor]:
    """xy [r, 2], extrinsics [b, v, 4, 4], intrinsics [b, v, 3, 3] -> origins, directions
    [b, v, r, 3] (projection.py:91-114)."""
    xy1 = torch.cat([xy, torch.ones_like(xy[:, :1])], -1)
    d = torch.einsum("bvij,rj->bvri", intrinsics.inverse(), xy1)
    d = d / d.norm(dim=-1, keepdim=True)
    d = torch.einsum("bvij,bvrj->bvri", extrinsics[..., :3, :3], d)
    o = extrinsics[..., None, :3, 3].expand_as(d)
    return o, d


class EpipolarSampling:
    """Same attributes as the reference dataclass (epipolar_sampler.py:19-27)."""

    def __init__(self, geometry: EpipolarGeometry, kv_source: Optional[EpipolarKV], extrinsics: Tensor,
                 intrinsics: Tensor, num_samples: int):
        self._g, self._kv, self._e, self._k, self._s = geometry, kv_source, extrinsics, intrinsics, num_samples
        self._cache: dict = {}

    @property
    def valid(self) -> Tensor:                       # [b, v, ov, r] bool
        return self._g.valid.bool()

    def _xy(self, shift: float) -> Tensor:
        s = self._s
        u = (torch.arange(s, device=self._g.segments.device, dtype=torch.float32) + 0.5) / s + shift
        lo, hi = self._g.segments[..., None, :2], self._g.segments[..., None, 2:]
        return lo + u[:, None] * (hi - lo)

    @property
    def xy_sample(self) -> Tensor:                   # [b, v, ov, r, s, 2]
        return self._xy(0.0)

    @property
    def xy_sample_near(self) -> Tensor:
        return self._xy(-0.5 / self._s)

    @property
    def xy_sample_far(self) -> Tensor:
        return self._xy(0.5 / self._s)

    def _rays(self):
        if "rays" not in self._cache:
            h, w = self._g.grid
            xy = sample_image_grid((h, w), self._e.device).reshape(-1, 2)
            o, d = get_world_rays(xy, self._e.float(), self._k.float())
            self._cache["rays"] = (xy, o, d)
        return self._cache["rays"]

    @property
    def xy_ray(self) -> Tensor:                      # [b, v, r, 2]
        b, v = self._e.shape[:2]
        return self._rays()[0][None, None].expand(b, v, -1, -1)

    @property
    def origins(self) -> Tensor:
        return self._rays()[1]

    @property
    def directions(self) -> Tensor:
        return self._rays()[2]

    @property
    def features(self) -> Tensor:                    # [b, v, ov, r, s, c] -- materialised on demand
        if self._kv is None:
            raise RuntimeError("features are only available from EpipolarSampler.forward(images, ...)")
        if "features" not in self._cache:
            self._cache["features"] = self._kv.sample_features()
        return self._cache["features"]


class EpipolarSampler(nn.Module):
    def __init__(self, num_views: int, num_samples: int) -> None:
        super().__init__()
        self.num_samples = num_samples
        _, index_v = generate_heterogeneous_index(num_views)
        t_v, t_ov = generate_heterogeneous_index_transpose(num_views)
        self.register_buffer("index_v", index_v, persistent=False)
        self.register_buffer("transpose_v", t_v, persistent=False)
        self.register_buffer("transpose_ov", t_ov, persistent=False)

    def geometry(self, grid: tuple[int, int], extrinsics, intrinsics, near, far) -> EpipolarGeometry:
        return epipolar_geometry(extrinsics, intrinsics, near, far, grid, self.num_samples)

    def forward(self, images: Tensor, extrinsics: Tensor, intrinsics: Tensor, near: Tensor,
                far: Tensor) -> EpipolarSampling:
        b, v, _, h, w = images.shape
        g = self.geometry((h, w), extrinsics, intrinsics, near, far)
        kv = EpipolarKV(images, g, None, None, None)
        return EpipolarSampling(g, kv, extrinsics, intrinsics, self.num_samples)

    def generate_image_rays(self, images: Tensor, extrinsics: Tensor, intrinsics: Tensor):
        b, v, _, h, w = images.shape
        xy = sample_image_grid((h, w), images.device).reshape(-1, 2)
        o, d = get_world_rays(xy, extrinsics, intrinsics)
        return xy[None, None].expand(b, v, -1, -1), o, d

    def transpose(self, x: Tensor) -> Tensor:
        b = x.shape[0]
        t_b = torch.arange(b, device=x.device)[:, None, None]
        return x[t_b, self.transpose_v[None], self.transpose_ov[None]]

    def collect(self, target: Tensor) -> Tensor:
        """[b, v, ...] -> [b, v, v-1, ...]: for each view, all the other views."""
        b = target.shape[0]
        index_b = torch.arange(b, device=target.device)[:, None, None]
        return target[index_b, self.index_v[None]]
