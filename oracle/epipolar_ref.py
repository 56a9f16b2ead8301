"""Loader of the REFERENCE's own epipolar encoder modules from /root/reference (read-only).

TEST INFRASTRUCTURE ONLY, and usable ONLY in the authoring container: /root/reference does not
exist on the GPU box, so nothing at test/bench run time imports this module -- it is used by
oracle/make_epipolar_golden.py to generate the committed fixtures under tests/golden/, and by
CPU tests that skip when /root/reference is absent.

The encoder half of the hot path has a real oracle: the reference Python imports and
runs on CPU given (i) the omegaconf stub in oracle/_stubs and (ii) pre-registering
`src.model.encoder` / `src.dataset` as bare namespace packages so their __init__ files (which pull
in lightning, dacite, e3nn, lpips) never execute.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE = Path("/root/reference")
_STUBS = Path(__file__).resolve().parent / "_stubs"


def available() -> bool:
    return (REFERENCE / "src" / "model" / "encoder" / "epipolar" / "epipolar_transformer.py").exists()


def load(num_context_views: int = 2):
    """Returns a namespace with the reference classes/functions of the epipolar path."""
    if not available():
        raise RuntimeError("/root/reference is not present (expected on the GPU box)")
    for p in (str(_STUBS), str(REFERENCE)):
        if p not in sys.path:
            sys.path.insert(0, p)

    def bare(name, path):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [str(path)]
            sys.modules[name] = m

    bare("src.model.encoder", REFERENCE / "src/model/encoder")
    bare("src.dataset", REFERENCE / "src/dataset")
    from omegaconf import DictConfig
    from src.global_cfg import set_cfg
    set_cfg(DictConfig({"dataset": {"view_sampler": {"num_context_views": num_context_views}}}))
    from src.geometry import epipolar_lines, projection
    from src.model.encoder.epipolar import conversions, epipolar_sampler, epipolar_transformer
    from src.model.encoder.epipolar import image_self_attention
    from src.model.encodings import positional_encoding
    from src.model.transformer import attention, transformer
    ns = types.SimpleNamespace(
        EpipolarTransformer=epipolar_transformer.EpipolarTransformer,
        EpipolarTransformerCfg=epipolar_transformer.EpipolarTransformerCfg,
        ImageSelfAttentionCfg=image_self_attention.ImageSelfAttentionCfg,
        ImageSelfAttention=image_self_attention.ImageSelfAttention,
        EpipolarSampler=epipolar_sampler.EpipolarSampler,
        PositionalEncoding=positional_encoding.PositionalEncoding,
        Attention=attention.Attention, Transformer=transformer.Transformer,
        project_rays=epipolar_lines.project_rays, get_depth=epipolar_lines.get_depth,
        get_world_rays=projection.get_world_rays, sample_image_grid=projection.sample_image_grid,
        depth_to_relative_disparity=conversions.depth_to_relative_disparity,
    )
    return ns


def load_adapter():
    """The reference's GaussianAdapter / DepthPredictorMonocular.  The adapter
    module imports e3nn through src/misc/sh_rotation.py; oracle/_stubs/e3nn makes that import succeed
    and the returned namespace exposes the module object so the caller can replace `rotate_sh`
    (golden outputs are generated with the rotation factored out -- see make_adapter_golden.py)."""
    load(2)
    from src.model.encoder.common import gaussian_adapter
    from src.model.encoder.epipolar import depth_predictor_monocular
    return types.SimpleNamespace(
        module=gaussian_adapter, GaussianAdapter=gaussian_adapter.GaussianAdapter,
        GaussianAdapterCfg=gaussian_adapter.GaussianAdapterCfg,
        DepthPredictorMonocular=depth_predictor_monocular.DepthPredictorMonocular)
