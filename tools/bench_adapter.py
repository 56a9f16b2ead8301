#!/usr/bin/env python
"""Times the fused GaussianAdapter kernels (csrc/gaussian_adapter.cu) against the explicit torch path
(the reference's op sequence) at the configs[2] shape: batch 7 x 2 views x 256x256 rays x 3 samples,
SH degree 4.  CUDA events, inputs >> L2 (the outputs alone are ~1 GB), prints one JSON line with the
achieved HBM GB/s of each direction against the algorithmic bytes.

    python tools/bench_adapter.py [--batch 7]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=7)
    ap.add_argument("--hw", type=int, default=256)
    args = ap.parse_args()
    from pixelsplat_b200 import synthetic
    from pixelsplat_b200.encoder.gaussian_adapter import GaussianAdapter, GaussianAdapterCfg
    dev = torch.device("cuda", 0)
    b, v, r, srf, spp, d_sh = args.batch, 2, args.hw * args.hw, 1, 3, 25
    ad = GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 4)).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    ext = torch.eye(4, device=dev).repeat(b, v, 1, 1)
    ext[:, 1, 0, 3] = 1.0
    ext[:, 1, :3, :3] = torch.tensor([[0.995, 0.0, 0.0998], [0.0, 1.0, 0.0], [-0.0998, 0.0, 0.995]], device=dev)
    K = synthetic.intrinsics_re10k(v)[None].repeat(b, 1, 1, 1).to(dev)
    coords = torch.rand(b, v, r, srf, 1, 2, device=dev, generator=g).requires_grad_(True)
    depths = (0.5 + 10 * torch.rand(b, v, r, srf, spp, device=dev, generator=g)).requires_grad_(True)
    opac = torch.rand(b, v, r, srf, spp, device=dev, generator=g)
    raw = torch.randn(b, v, r, srf, 1, 7 + 3 * d_sh, device=dev, generator=g).requires_grad_(True)
    E, Kb = ext[:, :, None, None, None], K[:, :, None, None, None]

    def fwd(fused):
        if fused:
            return ad(E, Kb, coords, depths, opac, raw, (args.hw, args.hw))
        return ad.forward_explicit(E, Kb, coords, depths, opac, raw, (args.hw, args.hw))

    def fwd_bwd(fused):
        gs = fwd(fused)
        (gs.means.sum() + gs.covariances.sum() + gs.harmonics.sum()).backward()
        coords.grad = depths.grad = raw.grad = None

    res = {}
    for name, fused in (("fused", True), ("explicit", False)):
        torch.cuda.reset_peak_memory_stats()
        with torch.no_grad():
            t_f = timed(lambda: fwd(fused))
        t_fb = timed(lambda: fwd_bwd(fused), iters=5, warmup=2)
        res[name] = {"forward_ms": t_f, "forward_backward_ms": t_fb, "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
    n_rays = b * v * r * srf
    raw_n = 7 + 3 * d_sh
    per_g = 3 + 9 + 3 * d_sh + 3                                   # means, covariances, harmonics, scales
    fwd_bytes = n_rays * 4 * (raw_n + 2 + spp + spp * per_g + 4)
    bwd_bytes = n_rays * 4 * (7 + 2 + spp + spp * (3 + 9 + 3 * d_sh) + raw_n + 2 + spp)
    f = res["fused"]
    bwd_ms = f["forward_backward_ms"] - f["forward_ms"]
    print(json.dumps({"what": f"GaussianAdapter, {b}x{v} views x {r} rays x {spp} samples, SH degree 4", **res,
                      "fused_forward_gbs": fwd_bytes / (f["forward_ms"] * 1e-3) / 1e9,
                      "fused_backward_gbs_incl_autograd_sums": bwd_bytes / (bwd_ms * 1e-3) / 1e9,
                      "forward_algorithmic_bytes": fwd_bytes, "backward_algorithmic_bytes": bwd_bytes}))


if __name__ == "__main__":
    main()
