"""Drop-in for /root/reference/src/model/encoder/epipolar/epipolar_transformer.py:19-183.

Same constructor arguments, parameter tree, forward signature and return
value `(features [b, v, c, H, W], EpipolarSampling)`; externally used attributes are kept
(`.epipolar_sampler`, `.transformer.layers[i][0].fn.attend`, `.cfg.downscale`).

What changed underneath: the sampler geometry + two-ray depth is one CUDA kernel,
and each layer's sampled cross-attention is one fused CUDA kernel that gathers from the feature map
-- the [b,v,ov,r,s,c] sample tensor, the two same-sized depth-encoding tensors and the 7.5 GB K/V
projection of the reference are never formed.  Convolutions, LayerNorms and the per-image
self-attention stay library code (cuDNN / cuBLAS through torch).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from functools import partial
from typing import Optional

import torch
from torch import Tensor, nn

from .attention_fused import EpipolarKV
from .epipolar_sampler import EpipolarSampler, EpipolarSampling
from .image_self_attention import ImageSelfAttention, ImageSelfAttentionCfg
from .positional_encoding import PositionalEncoding
from .transformer import Transformer


@dataclass
class EpipolarTransformerCfg:
    self_attention: ImageSelfAttentionCfg
    num_octaves: int
    num_layers: int
    num_heads: int
    num_samples: int
    d_dot: int
    d_mlp: int
    downscale: int


def _num_context_views_from_reference_cfg() -> int:
    """When dropped into the reference tree, read the same global config it reads
    (epipolar_transformer.py:46)."""
    try:
        from src.global_cfg import get_cfg  # type: ignore
    except Exception as exc:  # pragma: no cover - only inside the reference tree
        raise ValueError("pass num_context_views= (no reference global config is importable)") from exc
    return get_cfg().dataset.view_sampler.num_context_views


class ImageSelfAttentionWrapper(nn.Module):
    def __init__(self, self_attention_cfg: ImageSelfAttentionCfg, d_in: int, d_hidden: int, dropout: float):
        super().__init__()
        self.self_attention = ImageSelfAttention(self_attention_cfg, d_in, d_in)

    def forward(self, x: Tensor, b: int, v: int, h: int, w: int) -> Tensor:
        c = x.shape[-1]
        img = x.reshape(b * v, h, w, c).permute(0, 3, 1, 2)
        img = self.self_attention(img) + img
        return img.permute(0, 2, 3, 1).reshape(b * v * h * w, 1, c)


class EpipolarTransformer(nn.Module):
    def __init__(self, cfg: EpipolarTransformerCfg, d_in: int, num_context_views: Optional[int] = None) -> None:
        super().__init__()
        if num_context_views is None:
            num_context_views = _num_context_views_from_reference_cfg()
        self.cfg = cfg
        self.num_context_views = num_context_views
        self.conv_channels_last = os.environ.get("PIXELSPLAT_B200_CONV_NHWC", "0") == "1"
        self.epipolar_sampler = EpipolarSampler(num_context_views, cfg.num_samples)
        if cfg.num_octaves > 0:
            pe = PositionalEncoding(cfg.num_octaves)
            self.depth_encoding = nn.Sequential(pe, nn.Linear(pe.d_out(1), d_in))
        self.transformer = Transformer(d_in, cfg.num_layers, cfg.num_heads, cfg.d_dot, cfg.d_mlp, selfatt=False,
                                       kv_dim=d_in,
                                       feed_forward_layer=partial(ImageSelfAttentionWrapper, cfg.self_attention))
        self.downscaler = self.upscaler = self.upscale_refinement = None
        if cfg.downscale:
            self.downscaler = nn.Conv2d(d_in, d_in, cfg.downscale, cfg.downscale)
            self.upscaler = nn.ConvTranspose2d(d_in, d_in, cfg.downscale, cfg.downscale)
            self.upscale_refinement = nn.Sequential(nn.Conv2d(d_in, d_in * 2, 7, 1, 3), nn.GELU(),
                                                    nn.Conv2d(d_in * 2, d_in, 7, 1, 3))
        if num_context_views > 2:
            self.view_embeddings = nn.Embedding(num_context_views, d_in)

    def forward(self, features: Tensor, extrinsics: Tensor, intrinsics: Tensor, near: Tensor,
                far: Tensor) -> tuple[Tensor, EpipolarSampling]:
        b, v, c, H, W = features.shape
        if self.downscaler is not None:
            features = self.downscaler(features.flatten(0, 1)).unflatten(0, (b, v))
        h, w = features.shape[-2:]

        geometry = self.epipolar_sampler.geometry((h, w), extrinsics, intrinsics, near, far)
        emb = None
        if v > 2:
            # randomly permuted per-view embeddings (epipolar_transformer.py:126-131)
            shuffle = torch.randperm(v - 1, device=features.device)
            emb = self.view_embeddings(shuffle)
        depth_linear = self.depth_encoding[1] if self.cfg.num_octaves > 0 else None
        pe_module = self.depth_encoding[0] if self.cfg.num_octaves > 0 else None
        kv = EpipolarKV(features, geometry, depth_linear, pe_module, emb)
        sampling = EpipolarSampling(geometry, kv, extrinsics, intrinsics, self.cfg.num_samples)

        q = features.permute(0, 1, 3, 4, 2).reshape(b * v * h * w, 1, c)     # "(b v h w) () c"
        x = self.transformer(q, kv, b=b, v=v, h=h, w=w)
        features = x.reshape(b, v, h, w, c).permute(0, 1, 4, 2, 3)

        if self.upscaler is not None:
            x = features.flatten(0, 1)
            if self.conv_channels_last:
                # cuDNN's tensor-core kernels for the 7x7 refinement convolutions (the largest item of the step,
                # SURVEY.md 8 row f-2) are NHWC kernels; handing them NHWC activations removes the layout
                # transposes cuDNN otherwise wraps around every call.  Values are unchanged.
                x = x.contiguous(memory_format=torch.channels_last)
            f = self.upscaler(x)
            f = self.upscale_refinement(f) + f
            features = f.unflatten(0, (b, v))
        return features, sampling
