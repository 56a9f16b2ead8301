"""Shared helpers of the parity tests: build rasterizer arguments from a synthetic scene, run the
CPU oracle, compare."""
from __future__ import annotations

import numpy as np
import torch

from oracle import raster_oracle as ro
from oracle import raster_torch as rt
from pixelsplat_b200 import synthetic


def view_args(scene: synthetic.Scene, view: int = 0, use_sh: bool = True, scale_invariant=True):
    """Rasterizer-level arguments of one view (fp32, CPU), via the oracle's restatement of
    render_cuda's host code."""
    return rt.prepare_view(scene.means, scene.covariances, scene.harmonics, scene.opacities,
                           scene.extrinsics[view], scene.intrinsics[view], scene.near[view],
                           scene.far[view], dtype=torch.float32, scale_invariant=scale_invariant,
                           use_sh=use_sh)


def oracle_forward(a: dict, bg, W, H):
    n = lambda t: None if t is None else t.detach().cpu().numpy()
    return ro.forward(n(a["means"]), n(a["cov6"]), n(a["opac"]), n(a["sh"]), n(a["colors"]),
                      n(a["vm"]), n(a["pm"]), n(a["campos"]), a["tanfovx"], a["tanfovy"],
                      np.asarray(bg, np.float32), W, H, a["sh_degree"], dtype=np.float32)


def oracle_forward64(a: dict, bg, W, H):
    """The same view in float64 (inputs are the float32 values, widened): the yardstick of the gradient bars."""
    n = lambda t: None if t is None else t.detach().cpu().numpy().astype(np.float64)
    return ro.forward(n(a["means"]), n(a["cov6"]), n(a["opac"]), n(a["sh"]), n(a["colors"]),
                      n(a["vm"]), n(a["pm"]), n(a["campos"]), a["tanfovx"], a["tanfovy"],
                      np.asarray(bg, np.float64), W, H, a["sh_degree"], dtype=np.float64)


def oracle_backward(fwd, a: dict, d_img, bg, W, H):
    n = lambda t: None if t is None else t.detach().cpu().numpy()
    return ro.backward(fwd, np.asarray(d_img, np.float32), n(a["means"]), n(a["cov6"]), n(a["sh"]),
                       n(a["vm"]), n(a["pm"]), n(a["campos"]), a["tanfovx"], a["tanfovy"],
                       np.asarray(bg, np.float32), W, H, a["sh_degree"])


def upstream_keys_from_native(keys_i64: np.ndarray, tile_start: np.ndarray, tile_count: np.ndarray):
    """Native per-tile keys (depth_bits << 32 | gaussian) -> upstream's (tile << 32 | depth_bits,
    gaussian) pairs, for ONE view whose segments start at tile_start[0]."""
    k = keys_i64.astype(np.uint64)
    depth = k >> np.uint64(32)
    gauss = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    tile_of = np.repeat(np.arange(tile_count.size, dtype=np.uint64), tile_count.astype(np.int64))
    return (tile_of << np.uint64(32)) | depth, gauss


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """compute_psnr of the reference (src/evaluation/metrics.py:11-19) for one image."""
    a, b = np.clip(a, 0, 1), np.clip(b, 0, 1)
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else -10.0 * np.log10(mse)


# Gradient bars against the FLOAT64 oracle.  One global max per tensor lets an entry 100x
# smaller than the largest be 20 % wrong, so two more views of the same difference:
#   l2   = ||got - ref||_2 / ||ref||_2                                  (norm-wise relative error)
#   q999 = 99.9th percentile of |got - ref| / (ATOL_REL * max|ref| + RTOL * |ref|)   (element-wise, mixed)
# The percentile (not the max) because a pixel whose alpha sits on the 1/255 or T < 1e-4 decision boundary
# may legitimately take the other branch in float32 (ex2.approx on the GPU vs float64 exp in the oracle) and
# moves the handful of Gaussians that touch it.
GRAD_ATOL_REL, GRAD_RTOL = 1e-5, 1e-3


def grad_errors(got: np.ndarray, ref: np.ndarray) -> dict:
    got, ref = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    d = np.abs(got - ref)
    scale = max(float(np.abs(ref).max()), 1e-30)
    mixed = d / (GRAD_ATOL_REL * scale + GRAD_RTOL * np.abs(ref))
    return dict(max=float(d.max() / scale), l2=float(np.linalg.norm(d) / max(np.linalg.norm(ref), 1e-30)),
                q999=float(np.quantile(mixed, 0.999)) if mixed.size else 0.0)


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
