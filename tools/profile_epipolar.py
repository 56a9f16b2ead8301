#!/usr/bin/env python
"""Smallest run that launches every encoder-side kernel once at configs[2] per-scene shapes (2 views,
256x256 features -> 64x64 rays, 32 samples): meant to sit under `ncu -k regex:k_epi`.

    ncu --set full --clock-control none -k regex:"k_epi" -c 6 -o out python tools/profile_epipolar.py --batch 2
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    from pixelsplat_b200.encoder import EpipolarTransformer, EpipolarTransformerCfg, ImageSelfAttentionCfg
    from tests import golden_util as gu
    dev = torch.device("cuda:0")
    cfg = EpipolarTransformerCfg(ImageSelfAttentionCfg(4, 10, 2, 4, 128, 128, 256), 10, 2, 4, 32, 128, 256, 4)
    m = EpipolarTransformer(cfg, 128, num_context_views=2)
    gu.fill_parameters(m)
    m = m.to(dev)
    ext, K, near, far = [t.to(dev, torch.float32) for t in gu.camera_rig(args.batch, 2, "generic")]
    feats = torch.randn(args.batch, 2, 128, 256, 256, device=dev, requires_grad=True)
    out, _ = m(feats, ext, K, near, far)
    out.square().mean().backward()
    torch.cuda.synchronize()
    print("ok", float(out.abs().mean()))


if __name__ == "__main__":
    main()
