"""Generates tests/golden/epipolar_*.npz by running the REFERENCE's own modules from
/root/reference on the CPU (float64 and float32).  Run in the authoring container only:

    python oracle/make_epipolar_golden.py

TEST INFRASTRUCTURE.  Inputs and weights are regenerated from tests/golden_util.py on both sides;
the .npz files hold reference outputs only (small).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import epipolar_ref  # noqa: E402
from tests import golden_util as gu  # noqa: E402

OUT = ROOT / "tests" / "golden"


def geometry_case(ns, b, v, grid, S, case, dtype):
    torch.set_default_dtype(dtype)     # the reference builds its grids in the default dtype
    try:
        return _geometry_case(ns, b, v, grid, S, case, dtype)
    finally:
        torch.set_default_dtype(torch.float32)


def _geometry_case(ns, b, v, grid, S, case, dtype):
    ext, K, near, far = [t.to(dtype) for t in gu.camera_rig(b, v, case)]
    h, w = grid
    sampler = ns.EpipolarSampler(v, S)
    images = torch.zeros(b, v, 1, h, w, dtype=dtype)
    sampling = sampler.forward(images, ext, K, near, far)
    collect = sampler.collect
    depths = ns.get_depth(sampling.origins[:, :, None, :, None], sampling.directions[:, :, None, :, None],
                          sampling.xy_sample, collect(ext)[:, :, :, None, None], collect(K)[:, :, :, None, None])
    depths = depths.maximum(near[..., None, None, None]).minimum(far[..., None, None, None])
    rd = ns.depth_to_relative_disparity(depths, near[:, :, None, None, None], far[:, :, None, None, None])
    seg = torch.cat([sampling.xy_sample_near[..., 0, :] + 0, sampling.xy_sample_far[..., -1, :] + 0], -1)
    # xy_min = xy_sample_near[s=0] (sample_depth - half_span = 0), xy_max = xy_sample_far[s=S-1]
    return dict(segments=seg.numpy(), valid=sampling.valid.numpy(), rel_disparity=rd.numpy(),
                xy_sample=sampling.xy_sample.numpy(), origins=sampling.origins.numpy(),
                directions=sampling.directions.numpy())


def make_geometry():
    ns = epipolar_ref.load(2)
    out = {}
    for case, (b, v) in {"generic": (2, 2), "parallel": (1, 3), "diverging": (1, 2)}.items():
        ns = epipolar_ref.load(v)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = geometry_case(ns, b, v, (8, 8), 32, case, dt)
            for k, val in r.items():
                out[f"{case}_{tag}_{k}"] = val
    ns3 = epipolar_ref.load(3)
    r = geometry_case(ns3, 1, 3, (6, 10), 32, "generic", torch.float64)
    for k, val in r.items():
        out[f"generic3_f64_{k}"] = val
    np.savez_compressed(OUT / "epipolar_geometry.npz", **out)
    print("geometry:", {k: v.shape for k, v in out.items() if "f64" in k and "generic_" in k})


def transformer_case(v, dtype, HW=32, perm=None):
    ns = epipolar_ref.load(v)
    cfg = ns.EpipolarTransformerCfg(ns.ImageSelfAttentionCfg(4, 10, 2, 4, 128, 128, 256), 10, 2, 4, 32, 128, 256, 4)
    torch.set_default_dtype(dtype)
    try:
        return _transformer_case(ns, cfg, v, dtype, HW, perm)
    finally:
        torch.set_default_dtype(torch.float32)


def _transformer_case(ns, cfg, v, dtype, HW, perm):
    m = ns.EpipolarTransformer(cfg, 128).to(dtype)
    gu.fill_parameters(m)
    b = 1
    ext, K, near, far = [t.to(dtype) for t in gu.camera_rig(b, v, "generic")]
    feats = gu.seeded_like("features", (b, v, 128, HW, HW), 1.0, dtype).requires_grad_(True)
    wgt = gu.seeded_like("loss_weight", (b, v, 128, HW, HW), 1.0, dtype)
    captured = {}
    # the reference calls self.transformer.forward(...) directly (no hooks fire there), so the
    # transformer output is captured as the input of the upscaler: [(b v), c, h, w]
    hook = m.upscaler.register_forward_pre_hook(lambda mod, inp: captured.__setitem__("core", inp[0].detach()))
    if perm is not None:
        real = torch.randperm
        torch.randperm = lambda n, **kw: torch.tensor(perm)
    try:
        out, sampling = m(feats, ext, K, near, far)
    finally:
        if perm is not None:
            torch.randperm = real
        hook.remove()
    (out * wgt).sum().backward()
    res = dict(core=captured["core"].numpy(), out_sub=out.detach()[..., ::4, ::4].numpy(),
               out_mean=np.array(out.detach().mean().item()), out_abs=np.array(out.detach().abs().mean().item()),
               dfeat_sub=feats.grad[..., ::4, ::4].numpy())
    small = ["depth_encoding.1.weight", "depth_encoding.1.bias", "transformer.layers.0.0.norm.weight",
             "transformer.layers.1.0.fn.to_out.0.bias", "downscaler.bias"]
    big = ["transformer.layers.0.0.fn.to_q.weight", "transformer.layers.0.0.fn.to_kv.weight",
           "transformer.layers.1.0.fn.to_kv.weight", "transformer.layers.0.0.fn.to_out.0.weight",
           "downscaler.weight", "transformer.layers.0.1.fn.self_attention.patch_embedder.0.weight"]
    if v > 2:
        small.append("view_embeddings.weight")
    params = dict(m.named_parameters())
    for n in small:
        res["grad:" + n] = params[n].grad.numpy()
    for n in big:
        g = params[n].grad
        res["gradsub:" + n] = g.reshape(-1)[::97].numpy()
        res["gradnorm:" + n] = np.array(g.norm().item())
    return res


def make_transformer():
    for v, perm in ((2, None), (3, [1, 0])):
        out = {}
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = transformer_case(v, dt, perm=perm)
            for k, val in r.items():
                out[f"{tag}_{k}"] = val.astype(np.float64 if tag == "f64" else np.float32)
        if perm is not None:
            out["perm"] = np.array(perm)
        np.savez_compressed(OUT / f"epipolar_transformer_v{v}.npz", **out)
        e = np.abs(out["f64_core"] - out["f32_core"]).max() / np.abs(out["f64_core"]).max()
        print(f"transformer v={v}: core {out['f64_core'].shape}, reference fp32-vs-fp64 rel err {e:.2e}")


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    make_geometry()
    make_transformer()
