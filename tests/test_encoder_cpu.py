"""CPU tests of the encoder module surface: the drop-in's parameter tree
against the key list of the REFERENCE's own EpipolarTransformer (tests/golden/epipolar_state_dict_keys.json,
written by oracle/make_epipolar_golden.py from /root/reference).  Constructing the module needs no GPU; its
forward does."""
import json
from pathlib import Path

import pytest
import torch

from tests import golden_util as gu

KEYS = json.loads((Path(__file__).resolve().parent / "golden" / "epipolar_state_dict_keys.json").read_text())


def _module(v):
    from pixelsplat_b200.encoder import EpipolarTransformer, EpipolarTransformerCfg, ImageSelfAttentionCfg
    cfg = EpipolarTransformerCfg(ImageSelfAttentionCfg(4, 10, 2, 4, 128, 128, 256), 10, 2, 4, 32, 128, 256, 4)
    return EpipolarTransformer(cfg, 128, num_context_views=v)


@pytest.mark.parametrize("v", [2, 3])
def test_state_dict_is_the_references_key_for_key(v):
    ref = KEYS[f"v{v}"]
    m = _module(v)
    ours = [[k, list(t.shape)] for k, t in m.state_dict().items()]
    assert ours == ref["state_dict"]                                   # names, shapes AND order
    assert [k for k, _ in m.named_parameters()] == ref["parameters"]
    assert sum(p.numel() for p in m.parameters()) == ref["num_parameters"]
    assert ("view_embeddings.weight" in dict(ours)) == (v > 2)
    # buffers that the reference registers non-persistently must not leak into the state dict
    assert set(k for k, _ in m.named_buffers()) >= set(ref["buffers"])
    assert not (set(ref["buffers"]) & set(dict(ours)))


@pytest.mark.parametrize("v", [2, 3])
def test_a_reference_state_dict_loads_strictly(v):
    """A checkpoint with exactly the reference's keys / shapes (values seeded by name) loads with strict=True,
    nothing missing, nothing unexpected, every tensor taken."""
    ref = KEYS[f"v{v}"]
    ckpt = {k: gu.seeded_like("ckpt." + k, shape, 0.1, torch.float32) for k, shape in ref["state_dict"]}
    m = _module(v)
    res = m.load_state_dict(ckpt, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, t in m.state_dict().items():
        assert torch.equal(t, ckpt[k]), k
    with pytest.raises(RuntimeError):
        m.load_state_dict({**ckpt, "transformer.layers.0.0.fn.to_k.weight": torch.zeros(1)}, strict=True)
    # the attributes outside code reaches into (encoder_epipolar.py:232-235, encoder_visualizer_epipolar.py:53-56)
    assert hasattr(m, "epipolar_sampler") and hasattr(m.epipolar_sampler, "index_v") and callable(m.epipolar_sampler.collect)
    assert hasattr(m.transformer.layers[0][0].fn, "attend") and m.cfg.downscale == 4


def test_heterogeneous_index_tables_match_the_reference():
    """Row a15: generate_heterogeneous_index{,_transpose} (/root/reference/src/misc/heterogeneous_pairings.py:9-43)
    for 2..5 views against the reference's own tables (tests/golden/heterogeneous_pairings.json, written from
    /root/reference); the transpose is an involution."""
    from pixelsplat_b200.encoder.heterogeneous_pairings import (generate_heterogeneous_index,
                                                                generate_heterogeneous_index_transpose)
    gold = json.loads((Path(__file__).resolve().parent / "golden" / "heterogeneous_pairings.json").read_text())
    for v in (2, 3, 4, 5):
        a, b = generate_heterogeneous_index(v)
        c, d = generate_heterogeneous_index_transpose(v)
        g = gold[str(v)]
        assert a.tolist() == g["index_self"] and b.tolist() == g["index_other"]
        assert c.tolist() == g["t_v"] and d.tolist() == g["t_ov"]
        x = torch.arange(v * (v - 1)).reshape(v, v - 1)
        assert torch.equal(x[c, d][c, d], x)


def test_ray_generation_matches_the_reference_on_the_cpu():
    """Rows a9 / a11 pieces that are plain torch in the product (sample_image_grid, get_world_rays:
    /root/reference/src/geometry/projection.py:91-137) against the reference sampler's xy_ray / origins /
    directions (tests/golden/epipolar_geometry.npz, float64 run)."""
    import numpy as np
    from pixelsplat_b200.encoder.epipolar_sampler import get_world_rays, sample_image_grid
    gold = np.load(Path(__file__).resolve().parent / "golden" / "epipolar_geometry.npz")
    for case, b, v, grid, tag in (("generic", 2, 2, (8, 8), "generic"), ("generic", 1, 3, (6, 10), "generic3")):
        ext, K, _, _ = gu.camera_rig(b, v, case)
        xy = sample_image_grid(grid, "cpu").reshape(-1, 2).double()
        o, d = get_world_rays(xy, ext, K)
        assert np.abs(xy.numpy()[None, None] - gold[f"{tag}_f64_xy_ray"]).max() < 1e-7      # float32 grid, as upstream
        assert np.abs(o.numpy() - gold[f"{tag}_f64_origins"]).max() < 1e-12
        assert np.abs(d.numpy() - gold[f"{tag}_f64_directions"]).max() < 1e-6
