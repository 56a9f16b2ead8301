"""The tail of `EncoderEpipolar.forward`: features after the epipolar
transformer -> the flattened `Gaussians` the decoder consumes.

Reference: /root/reference/src/model/encoder/encoder_epipolar.py:89-213 (`map_pdf_to_opacity`,
high-resolution skip, depth predictor, `to_gaussians`, sub-pixel offsets, `GaussianAdapter`, optional
per-pixel opacity, the final `rearrange`s).  Sub-module names equal the reference's
(`depth_predictor`, `to_gaussians`, `gaussian_adapter`, `high_resolution_skip`, `to_opacity`), so
those entries of an `EncoderEpipolar` state dict load into this module unchanged.  The backbone
and the transformer are separate modules (backbone: out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from ..decoder.decoder_splatting_cuda import Gaussians
from .depth_predictor_monocular import DepthPredictorMonocular
from .epipolar_sampler import sample_image_grid
from .gaussian_adapter import GaussianAdapter, GaussianAdapterCfg


@dataclass
class OpacityMappingCfg:
    initial: float
    final: float
    warm_up: int


@dataclass
class EncoderTailCfg:
    d_feature: int = 128
    num_monocular_samples: int = 32
    num_surfaces: int = 1
    predict_opacity: bool = False
    gaussians_per_pixel: int = 3
    use_transmittance: bool = False
    gaussian_adapter: GaussianAdapterCfg = None
    opacity_mapping: OpacityMappingCfg = None

    def __post_init__(self):
        if self.gaussian_adapter is None:
            self.gaussian_adapter = GaussianAdapterCfg(0.5, 15.0, 4)       # config/model/encoder/epipolar.yaml:18-21
        if self.opacity_mapping is None:
            self.opacity_mapping = OpacityMappingCfg(0.0, 0.0, 1)          # epipolar.yaml:6-9


def _wh(w: int, h: int, device) -> Tensor:
    """float32 [w, h] on `device`, built by fill kernels (torch.tensor(...) from a Python tuple is a pageable
    host-to-device copy, which a CUDA-graph capture rejects)."""
    return torch.cat([torch.full((1,), float(w), dtype=torch.float32, device=device),
                      torch.full((1,), float(h), dtype=torch.float32, device=device)])


class EncoderEpipolarTail(nn.Module):
    def __init__(self, cfg: EncoderTailCfg) -> None:
        super().__init__()
        self.cfg = cfg
        self.depth_predictor = DepthPredictorMonocular(cfg.d_feature, cfg.num_monocular_samples, cfg.num_surfaces,
                                                       cfg.use_transmittance)
        self.gaussian_adapter = GaussianAdapter(cfg.gaussian_adapter)
        if cfg.predict_opacity:
            self.to_opacity = nn.Sequential(nn.ReLU(), nn.Linear(cfg.d_feature, 1), nn.Sigmoid())
        self.to_gaussians = nn.Sequential(
            nn.ReLU(), nn.Linear(cfg.d_feature, cfg.num_surfaces * (2 + self.gaussian_adapter.d_in)))
        self.high_resolution_skip = nn.Sequential(nn.Conv2d(3, cfg.d_feature, 7, 1, 3), nn.ReLU())

    def map_pdf_to_opacity(self, pdf: Tensor, global_step: int) -> Tensor:
        cfg = self.cfg.opacity_mapping
        x = cfg.initial + min(global_step / cfg.warm_up, 1) * (cfg.final - cfg.initial)
        exponent = 2 ** x
        return 0.5 * (1 - (1 - pdf) ** exponent + pdf ** (1 / exponent))

    def forward(self, features: Tensor, context: dict, global_step: int = 0, deterministic: bool = False,
                visualization_dump: Optional[dict] = None) -> Gaussians:
        """features [b, v, c, h, w] (output of the epipolar transformer); context: image [b, v, 3, h, w],
        extrinsics [b, v, 4, 4], intrinsics [b, v, 3, 3], near / far [b, v]."""
        device = features.device
        b, v, _, h, w = context["image"].shape
        skip = self.high_resolution_skip(context["image"].reshape(b * v, 3, h, w))
        features = features + skip.reshape(b, v, -1, h, w)
        features = features.permute(0, 1, 3, 4, 2).reshape(b, v, h * w, -1)              # "b v (h w) c"
        gpp = self.cfg.gaussians_per_pixel
        depths, densities = self.depth_predictor(features, context["near"], context["far"], deterministic,
                                                 1 if deterministic else gpp)
        xy_ray = sample_image_grid((h, w), device)
        xy_ray = xy_ray.reshape(h * w, 1, 2)
        gaussians = self.to_gaussians(features)
        gaussians = gaussians.reshape(*gaussians.shape[:-1], self.cfg.num_surfaces, -1)   # "... (srf c) -> ... srf c"
        offset_xy = gaussians[..., :2].sigmoid()
        pixel_size = 1 / _wh(w, h, device)            # (no host->device copy: the step must be graph-capturable)
        xy_ray = xy_ray + (offset_xy - 0.5) * pixel_size
        g = self.gaussian_adapter(
            context["extrinsics"][:, :, None, None, None], context["intrinsics"][:, :, None, None, None],
            xy_ray[..., None, :], depths, self.map_pdf_to_opacity(densities, global_step) / gpp,
            gaussians[..., None, 2:], (h, w))
        if visualization_dump is not None:
            visualization_dump["depth"] = depths.reshape(b, v, h, w, *depths.shape[3:])
            visualization_dump["scales"] = g.scales.reshape(b, -1, 3)
            visualization_dump["rotations"] = g.rotations.reshape(b, -1, 4)
        opacity_multiplier = self.to_opacity(features)[..., None] if self.cfg.predict_opacity else 1
        return Gaussians(g.means.reshape(b, -1, 3), g.covariances.reshape(b, -1, 3, 3),
                         g.harmonics.reshape(b, -1, 3, self.gaussian_adapter.d_sh),
                         (opacity_multiplier * g.opacities).reshape(b, -1))
