"""`DepthPredictorMonocular` with the reference's parameter tree and sampling semantics
(/root/reference/src/model/encoder/epipolar/depth_predictor_monocular.py:10-81,
distribution_sampler.py:11-52, /root/reference/src/misc/discrete_probability_distribution.py:7-38,
conversions.py:5-27).  Part of SURVEY.md 8 row f-1: it feeds `GaussianAdapter` (depths, densities).

The tensors here are small ([b, v, r, srf, 32]); the projection is a library GEMM and the rest is a
handful of element-wise kernels, so this stays in torch.  `to_pdf` / `to_offset` remain modules so
that the reference's visualisation hooks keep working.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn


def relative_disparity_to_depth(relative_disparity: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    """0 = near, 1 = far (conversions.py:5-15)."""
    disp_near = 1 / (near + eps)
    disp_far = 1 / (far + eps)
    return 1 / ((1 - relative_disparity) * (disp_near - disp_far) + disp_far + eps)


def depth_to_relative_disparity(depth: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    disp_near = 1 / (near + eps)
    disp_far = 1 / (far + eps)
    disp = 1 / (depth + eps)
    return 1 - (disp - disp_far) / (disp_near - disp_far + eps)


def sample_discrete_distribution(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    *batch, bucket = pdf.shape
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    cdf = normalized_pdf.cumsum(dim=-1)
    samples = torch.rand((*batch, num_samples), device=pdf.device)
    index = torch.searchsorted(cdf, samples, right=True).clip(max=bucket - 1)
    return index, normalized_pdf.gather(dim=-1, index=index)


def gather_discrete_topk(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    index = pdf.topk(k=num_samples, dim=-1).indices
    return index, normalized_pdf.gather(dim=-1, index=index)


class DistributionSampler:
    def sample(self, pdf: Tensor, deterministic: bool, num_samples: int):
        if deterministic:
            return gather_discrete_topk(pdf, num_samples)
        return sample_discrete_distribution(pdf, num_samples)

    def gather(self, index: Tensor, target: Tensor) -> Tensor:
        bucket_dim = index.ndim - 1
        while len(index.shape) < len(target.shape):
            index = index[..., None]
        shape = list(target.shape)
        shape[bucket_dim] = index.shape[bucket_dim]
        index = index.broadcast_to(shape)
        if target.shape[bucket_dim] == 1:
            index = torch.zeros_like(index)
        return target.gather(dim=bucket_dim, index=index)


class DepthPredictorMonocular(nn.Module):
    def __init__(self, d_in: int, num_samples: int, num_surfaces: int, use_transmittance: bool) -> None:
        super().__init__()
        self.projection = nn.Sequential(nn.ReLU(), nn.Linear(d_in, 2 * num_samples * num_surfaces))
        self.sampler = DistributionSampler()
        self.num_samples = num_samples
        self.num_surfaces = num_surfaces
        self.use_transmittance = use_transmittance
        self.to_pdf = nn.Softmax(dim=-1)        # these exist for hooks to latch onto
        self.to_offset = nn.Sigmoid()

    def forward(self, features: Tensor, near: Tensor, far: Tensor, deterministic: bool,
                gaussians_per_pixel: int) -> tuple[Tensor, Tensor]:
        """features [b, v, r, c] -> (depth, opacity), each [b, v, r, srf, gaussians_per_pixel]."""
        s = self.num_samples
        features = self.projection(features)
        # "... (dpt srf c) -> c ... srf dpt"
        features = features.reshape(*features.shape[:-1], s, self.num_surfaces, 2)
        pdf_raw, offset_raw = features.movedim(-3, -2).unbind(-1)            # [..., srf, dpt]
        pdf = self.to_pdf(pdf_raw)
        offset = self.to_offset(offset_raw)
        index, pdf_i = self.sampler.sample(pdf, deterministic, gaussians_per_pixel)
        offset = self.sampler.gather(index, offset)
        relative_disparity = (index + offset) / s
        depth = relative_disparity_to_depth(relative_disparity, near[..., None, None, None], far[..., None, None, None])
        if self.use_transmittance:
            partial = pdf.cumsum(dim=-1)
            partial = torch.cat((torch.zeros_like(partial[..., :1]), partial[..., :-1]), dim=-1)
            opacity = pdf / (1 - partial + 1e-10)
            opacity = self.sampler.gather(index, opacity)
        else:
            opacity = pdf_i
        return depth, opacity
