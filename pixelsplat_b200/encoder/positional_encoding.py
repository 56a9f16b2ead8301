"""sin(x * 2 pi 2^k + {0, pi/2}) features, layout "(d f p)"
(/root/reference/src/model/encodings/positional_encoding.py:8-36).  Buffers are non-persistent
like the reference's, so state_dicts hold parameters only."""
import torch
from torch import Tensor, nn


class PositionalEncoding(nn.Module):
    def __init__(self, num_octaves: int):
        super().__init__()
        octaves = torch.arange(num_octaves).float()
        freq = (2 * torch.pi * 2 ** octaves)[:, None].repeat(1, 2)
        self.register_buffer("frequencies", freq, persistent=False)
        phases = torch.tensor([0, 0.5 * torch.pi], dtype=torch.float32)[None].repeat(num_octaves, 1)
        self.register_buffer("phases", phases, persistent=False)

    def forward(self, samples: Tensor) -> Tensor:
        s = samples[..., None, None] * self.frequencies
        return torch.sin(s + self.phases).flatten(-3)

    def d_out(self, dimensionality: int) -> int:
        return self.frequencies.numel() * dimensionality
