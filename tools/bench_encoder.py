#!/usr/bin/env python
"""Encoder-half timing at BASELINE configs[2] shapes: EpipolarTransformer forward+backward on
features [B, 2, 128, 256, 256] (-> 64x64 rays, 32 samples), fused CUDA path vs the explicit path
(the reference's op sequence -- grid_sample, to_kv on every sample, soft-max -- restated with torch
ops inside the same module; it is NOT the reference itself, which cannot run on the GPU box).
Prints one JSON line.  Not the headline metric (bench.py is); documents the second half of the
hot path."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=7)
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--explicit", action="store_true", help="also time the explicit (materialised K/V) path")
    args = ap.parse_args()
    from pixelsplat_b200.encoder import EpipolarTransformer, EpipolarTransformerCfg, ImageSelfAttentionCfg
    from tests import golden_util as gu
    dev = torch.device("cuda:0")
    cfg = EpipolarTransformerCfg(ImageSelfAttentionCfg(4, 10, 2, 4, 128, 128, 256), 10, 2, 4, 32, 128, 256, 4)
    m = EpipolarTransformer(cfg, 128, num_context_views=args.views)
    gu.fill_parameters(m)
    m = m.to(dev)
    b, v = args.batch, args.views
    ext, K, near, far = [t.to(dev, torch.float32) for t in gu.camera_rig(b, v, "generic")]
    feats = torch.randn(b, v, 128, 256, 256, device=dev, requires_grad=True)

    def step():
        out, _ = m(feats, ext, K, near, far)
        out.square().mean().backward()
        m.zero_grad(set_to_none=True)
        feats.grad = None

    def timed(n_warm, n):
        for _ in range(n_warm):
            step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, torch.cuda.max_memory_allocated() / 2 ** 30

    res = {"config": f"configs[2] encoder half: features [{b},{v},128,256,256], 64x64 rays, 32 samples, 2 layers"}
    ms, mem = timed(args.warmup, args.steps)
    res["fused_ms_per_step"], res["fused_peak_gib"] = ms, mem
    # share of the fused attention kernels, from the profiler
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:12]
    res["top_kernels_ms"] = {r.key[:70]: round(r.device_time_total / 1e3, 3) for r in rows}
    if args.explicit:
        hooks = [layer[0].fn.attend.register_forward_hook(lambda *a: None) for layer in m.transformer.layers]
        ms, mem = timed(1, max(2, args.steps // 3))
        for h in hooks:
            h.remove()
        res["explicit_ms_per_step"], res["explicit_peak_gib"] = ms, mem
    print(json.dumps(res))


if __name__ == "__main__":
    main()
