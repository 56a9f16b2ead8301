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
