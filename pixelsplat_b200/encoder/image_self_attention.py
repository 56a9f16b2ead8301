"""Drop-in for /root/reference/src/model/encoder/epipolar/image_self_attention.py:12-79: a small
ViT over patch tokens (Conv patch-embed + ReLU, + Linear(PE(xy)), pre-LN MHA/FF blocks, ConvT
resampler).  Same parameter tree."""
from __future__ import annotations

from dataclasses import dataclass

from torch import Tensor, nn

from .epipolar_sampler import sample_image_grid
from .positional_encoding import PositionalEncoding
from .transformer import Transformer


@dataclass
class ImageSelfAttentionCfg:
    patch_size: int
    num_octaves: int
    num_layers: int
    num_heads: int
    d_token: int
    d_dot: int
    d_mlp: int


class ImageSelfAttention(nn.Module):
    def __init__(self, cfg: ImageSelfAttentionCfg, d_in: int, d_out: int):
        super().__init__()
        pe = PositionalEncoding(cfg.num_octaves)
        self.positional_encoding = nn.Sequential(pe, nn.Linear(pe.d_out(2), cfg.d_token))
        self.patch_embedder = nn.Sequential(nn.Conv2d(d_in, cfg.d_token, cfg.patch_size, cfg.patch_size), nn.ReLU())
        self.transformer = Transformer(cfg.d_token, cfg.num_layers, cfg.num_heads, cfg.d_dot, cfg.d_mlp)
        self.resampler = nn.ConvTranspose2d(cfg.d_token, d_out, cfg.patch_size, cfg.patch_size)

    def forward(self, image: Tensor) -> Tensor:
        tokens = self.patch_embedder(image)                       # [n, d_token, nh, nw]
        n, c, nh, nw = tokens.shape
        xy = self.positional_encoding(sample_image_grid((nh, nw), image.device))   # [nh, nw, d_token]
        tokens = tokens + xy.permute(2, 0, 1)
        tokens = self.transformer(tokens.flatten(2).transpose(1, 2))              # [n, nh*nw, d_token]
        tokens = tokens.transpose(1, 2).reshape(n, c, nh, nw)
        return self.resampler(tokens)
