"""Generates tests/golden/adapter_*.npz by running the REFERENCE's GaussianAdapter and
DepthPredictorMonocular from /root/reference on the CPU (float64 and float32).  Authoring
container only:

    python oracle/make_adapter_golden.py

TEST INFRASTRUCTURE.  The reference's `rotate_sh` (src/misc/sh_rotation.py:10-30) is two e3nn calls
and e3nn is absent offline, so the reference module runs with its `rotate_sh` bound to
oracle/wigner_e3nn.rotate_sh_reference -- the same function body on a restatement of e3nn's published
`matrix_to_angles` / `wigner_D` (float64 generators; see that file for how it is pinned).  Everything
else is the reference's own code: means, covariances, scales, rotations, opacities, masked + broadcast +
ROTATED harmonics, and the gradients of tests/golden_util.adapter_loss.  Inputs and weights are
regenerated from tests/golden_util.py on both sides; only reference outputs are stored.
(adapter_v1.npz, round 1, had the rotation replaced by the identity.)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import epipolar_ref, wigner_e3nn  # noqa: E402
from tests import golden_util as gu  # noqa: E402

OUT = ROOT / "tests" / "golden"
IMAGE_SHAPE = (48, 64)     # (h, w): only enters through the pixel-size multiplier


def adapter(ns, dtype, tag, case, out):
    c = gu.adapter_case(case=case)
    t = {k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in c.items()}
    w = {k: v.to(dtype) for k, v in c["weights"].items()}
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("coordinates", "depths", "opacities", "raw")}
    ad = ns.GaussianAdapter(ns.GaussianAdapterCfg(0.5, 15.0, 4)).to(dtype)
    ad.sh_mask = ad.sh_mask.to(dtype)
    g = ad.forward(t["extrinsics"], t["intrinsics"], leaves["coordinates"], leaves["depths"], leaves["opacities"],
                   leaves["raw"], IMAGE_SHAPE)
    gu.adapter_loss(g, w).backward()
    pre = f"{case}_{tag}_"
    for k in ("means", "covariances", "scales", "opacities", "harmonics"):
        out[pre + k] = getattr(g, k).detach().numpy()
    out[pre + "rotations"] = g.rotations.detach()[..., :1, :].numpy()
    for k, leaf in leaves.items():
        out[pre + "d_" + k] = leaf.grad.numpy()


def depth_predictor(ns, dtype, tag, out):
    torch.manual_seed(0)
    m = ns.DepthPredictorMonocular(128, 32, 1, False).to(dtype)
    gu.fill_parameters(m)
    feats = gu.seeded_like("depth.features", (2, 2, 24, 128)).to(dtype).requires_grad_(True)
    _, _, near, far = gu.camera_rig(2, 2)
    near, far = near.to(dtype), far.to(dtype)
    for mode, det, gpp in (("topk", True, 1), ("sampled", False, 3)):
        torch.manual_seed(1234)
        depth, opacity = m.forward(feats, near, far, det, gpp)
        out[f"depth_{tag}_{mode}_depth"] = depth.detach().numpy()
        out[f"depth_{tag}_{mode}_opacity"] = opacity.detach().numpy()
    (depth.sum() * 0.01 + opacity.sum()).backward()
    out[f"depth_{tag}_d_features"] = feats.grad.numpy()
    m2 = ns.DepthPredictorMonocular(128, 32, 2, True).to(dtype)       # 2 surfaces + transmittance branch
    gu.fill_parameters(m2)
    torch.manual_seed(99)
    depth, opacity = m2.forward(feats.detach(), near, far, False, 3)
    out[f"depth_{tag}_srf2_depth"] = depth.detach().numpy()
    out[f"depth_{tag}_srf2_opacity"] = opacity.detach().numpy()


def main():
    ns = epipolar_ref.load_adapter()
    ns.module.rotate_sh = wigner_e3nn.rotate_sh_reference             # see the module docstring
    # the reference hard-codes a float32 pixel_size (gaussian_adapter.py:68); cast it so the module also
    # runs in float64 (value-preserving: 1/w and 1/h are first formed in float32 exactly as upstream)
    orig = ns.GaussianAdapter.get_scale_multiplier
    ns.GaussianAdapter.get_scale_multiplier = lambda self, K, ps, *a: orig(self, K, ps.to(K.dtype), *a)
    out = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(dtype)
        try:
            for case in ("generic", "diverging"):
                adapter(ns, dtype, tag, case, out)
            depth_predictor(ns, dtype, tag, out)
        finally:
            torch.set_default_dtype(torch.float32)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "adapter_v2.npz", **out)
    print("wrote", OUT / "adapter_v2.npz", {k: v.shape for k, v in list(out.items())[:8]}, len(out), "arrays")


if __name__ == "__main__":
    main()
