#!/usr/bin/env python
"""A bare training-step harness for BASELINE configs[2] / configs[3]: the full hot path inside one
optimisation step, data-parallel over GPUs.

    features [B, 2, 128, 256, 256]  (synthetic stand-in for the out-of-scope DINO/ResNet backbone)
      -> EpipolarTransformer                       (hot path, rows a8-a15; trainable)
      -> EncoderEpipolarTail                       (row f-1: high-res skip, depth predictor, to_gaussians,
                                                    fused GaussianAdapter kernel; trainable)
      -> DecoderSplattingCUDA, 4 target views/scene (hot path, rows a1-a7; V cameras share Gaussians)
      -> MSE -> backward -> bucketed NCCL gradient all-reduce -> clip 0.5 -> Adam

Replaces, for measurement purposes only, the reference's Lightning loop
(/root/reference/src/main.py:89-134, model_wrapper.py:108-151).  One process per GPU:

    python tools/train_step.py --batch 7
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_step.py --batch 7

Prints one JSON line from rank 0 (scenes/s and views/s over all ranks, per-phase milliseconds).
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=7, help="scenes per GPU (README.md:87: batch is per GPU)")
    ap.add_argument("--target-views", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--hw", type=int, default=256)
    args = ap.parse_args()
    from pixelsplat_b200 import parallel, synthetic
    from pixelsplat_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from pixelsplat_b200.encoder import EpipolarTransformer, EpipolarTransformerCfg, ImageSelfAttentionCfg
    from pixelsplat_b200.encoder.encoder_tail import EncoderEpipolarTail, EncoderTailCfg
    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                     # identical initial parameters on every rank
    cfg = EpipolarTransformerCfg(ImageSelfAttentionCfg(4, 10, 2, 4, 128, 128, 256), 10, 2, 4, 32, 128, 256, 4)
    enc = EpipolarTransformer(cfg, 128, num_context_views=2).to(dev)
    head = EncoderEpipolarTail(EncoderTailCfg()).to(dev)
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda"),
                               type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).to(dev)
    params = list(enc.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=1.5e-4)
    B, T, HW = args.batch, args.target_views, args.hw
    g = torch.Generator().manual_seed(1234 + rank)          # per-rank data (main.py:106)
    feats = torch.randn(B, 2, 128, HW, HW, generator=g).to(dev)
    images = torch.rand(B, 2, 3, HW, HW, generator=g).to(dev)
    ctx_e = torch.eye(4).repeat(B, 2, 1, 1); ctx_e[:, 1, 0, 3] = 1.0
    ctx_k = synthetic.intrinsics_re10k(2)[None].repeat(B, 1, 1, 1)
    near_v, far_v = synthetic.bounds_from_baseline(1.0, HW, HW, 3.0 * HW, 0.5)
    tgt_e = torch.stack([synthetic.target_cameras(T, seed=rank * 100 + s) for s in range(B)])
    tgt_k = synthetic.intrinsics_re10k(T)[None].repeat(B, 1, 1, 1)
    target = torch.rand(B, T, 3, HW, HW, generator=g).to(dev)
    ctx_e, ctx_k, tgt_e, tgt_k = ctx_e.to(dev), ctx_k.to(dev), tgt_e.to(dev), tgt_k.to(dev)
    near_c, far_c = torch.full((B, 2), near_v, device=dev), torch.full((B, 2), far_v, device=dev)
    near_t, far_t = torch.full((B, T), near_v, device=dev), torch.full((B, T), far_v, device=dev)

    context = dict(image=images, extrinsics=ctx_e, intrinsics=ctx_k, near=near_c, far=far_c)
    phases = ["encoder", "head", "render", "backward", "allreduce", "optimizer"]
    acc = {p: 0.0 for p in phases}

    def step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
        ev[0].record()
        f, _ = enc(feats, ctx_e, ctx_k, near_c, far_c); ev[1].record()
        gs = head(f, context, global_step=0); ev[2].record()
        out = dec(gs, tgt_e, tgt_k, near_t, far_t, (HW, HW)); ev[3].record()
        loss = (out.color - target).square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward(); ev[4].record()
        parallel.allreduce_gradients(params); ev[5].record()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step(); ev[6].record()
        if timed:
            torch.cuda.synchronize()
            for i, p in enumerate(phases):
                acc[p] += ev[i].elapsed_time(ev[i + 1])
        return loss

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step(True)
    t1.record()
    torch.cuda.synchronize()
    sec = parallel.max_over_ranks(t0.elapsed_time(t1) * 1e-3, dev)
    if rank == 0:
        print(json.dumps({
            "config": f"configs[{2 if world == 1 else 3}]: 2-view {HW}x{HW}, batch {B}/GPU, {T} target views, "
                      f"EpipolarTransformer + splat render training step, world {world}",
            "scenes_per_s": world * B * args.steps / sec, "views_per_s": world * B * T * args.steps / sec,
            "ms_per_step": 1e3 * sec / args.steps, "phase_ms": {p: acc[p] / args.steps for p in phases},
            "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "loss": float(loss),
            "n_gpus": world, "data": "synthetic", "dtype": "f32"}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
