"""ctypes binding of include/pixelsplat_b200.h.

There is no fallback: if the CUDA library is missing, importing this module raises.  Build it
with `python -c "import __graft_entry__ as g; g.build()"` or `make -C pixelsplat_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PIXELSPLAT_B200_LIB", _PKG / "_C" / "libpixelsplat_b200.so"))

PS_OK = 0
PS_SH_M3, PS_SH_3M = 0, 1
PS_COV_TRIU6, PS_COV_3X3 = 0, 1
PS_SH_BASIS_3DGS, PS_SH_BASIS_E3NN = 0, 1
TILE = 16

_ERR_NAMES = {1: "PS_ERR_INVALID_ARGUMENT", 2: "PS_ERR_CUDA", 3: "PS_ERR_UNSUPPORTED"}


class RasterDesc(ctypes.Structure):
    _fields_ = [
        ("n_scenes", ctypes.c_int32), ("views_per_scene", ctypes.c_int32),
        ("n_gaussians", ctypes.c_int32), ("sh_coeffs", ctypes.c_int32),
        ("sh_degree", ctypes.c_int32), ("sh_layout", ctypes.c_int32),
        ("cov_layout", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("sort_impl", ctypes.c_int32), ("sort_segment_hint", ctypes.c_int32),
        ("instance_capacity", ctypes.c_int64),
        ("sh_basis", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class RasterInputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "means", "cov", "opacities", "sh", "viewmatrix", "projmatrix", "campos", "tanfov",
        "background", "scene_scale")]


class RasterState(ctypes.Structure):
    _fields_ = [("geom", ctypes.c_void_p), ("geom_bytes", ctypes.c_size_t),
                ("binning", ctypes.c_void_p), ("binning_bytes", ctypes.c_size_t),
                ("image", ctypes.c_void_p), ("image_bytes", ctypes.c_size_t)]


class RasterSizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "geom_bytes", "binning_bytes", "image_bytes", "backward_bytes")]


class RasterLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "depth", "radii", "xy", "conic_opacity", "rgb", "rect", "clamped", "tile_count",
        "tile_start", "tile_cursor", "n_instances", "vis_pairs", "vis_any", "keys", "keys_alt", "final_T",
        "n_contrib", "cull", "color", "block_hits", "run_hits", "run_state")]


class EpipolarDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "batch", "views", "grid_h", "grid_w", "samples", "channels", "heads", "pe_dim")]


class EpipolarInputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "features", "segments", "valid", "rel_disparity", "q_feat", "q_pe", "bias")]


class AdapterDesc(ctypes.Structure):
    _fields_ = [("n_views", ctypes.c_int32), ("n_rays", ctypes.c_int32), ("n_samples", ctypes.c_int32),
                ("sh_coeffs", ctypes.c_int32), ("image_h", ctypes.c_int32), ("image_w", ctypes.c_int32),
                ("scale_min", ctypes.c_float), ("scale_max", ctypes.c_float), ("eps", ctypes.c_float),
                ("reserved", ctypes.c_int32)]


class AdapterInputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "extrinsics", "intrinsics", "sh_rotation", "sh_mask", "coordinates", "depths", "raw")]


class RasterLoss(ctypes.Structure):
    _fields_ = [("target", ctypes.c_void_p), ("sums", ctypes.c_void_p)]


LOSS_SLOTS = 64


class RasterGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "d_means", "d_cov", "d_opacities", "d_sh", "d_means2d")]


EXPORTS = ("ps_version", "ps_last_error", "ps_raster_sizes_query", "ps_raster_layout_query",
           "ps_raster_forward", "ps_raster_backward", "ps_camera_setup", "ps_launch_count",
           "ps_timing_enable", "ps_timing_read", "ps_epipolar_geometry",
           "ps_epipolar_attention_forward", "ps_epipolar_attention_backward",
           "ps_self_attention_forward", "ps_gaussian_adapter_forward", "ps_gaussian_adapter_backward",
           "ps_sh_rotation_matrices", "ps_set_option", "ps_raster_forward_loss", "ps_raster_backward_loss",
           "ps_self_attention_forward_stats", "ps_self_attention_backward")


class NativeLibraryMissing(ImportError):
    pass


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. Run "
            f"`make -C {_PKG / 'csrc'}` (or __graft_entry__.build()). There is no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.ps_version.restype = ctypes.c_int
    lib.ps_last_error.restype = ctypes.c_char_p
    P = ctypes.POINTER
    lib.ps_raster_sizes_query.argtypes = [P(RasterDesc), P(RasterSizes)]
    lib.ps_raster_layout_query.argtypes = [P(RasterDesc), P(RasterLayout)]
    lib.ps_raster_forward.argtypes = [P(RasterDesc), P(RasterInputs), P(RasterState), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ps_raster_backward.argtypes = [P(RasterDesc), P(RasterInputs), P(RasterState), ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_size_t, P(RasterGrads), ctypes.c_void_p]
    lib.ps_raster_forward_loss.argtypes = [P(RasterDesc), P(RasterInputs), P(RasterState), P(RasterLoss),
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ps_raster_backward_loss.argtypes = [P(RasterDesc), P(RasterInputs), P(RasterState), ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, P(RasterGrads),
                                            ctypes.c_void_p]
    lib.ps_raster_forward_loss.restype = ctypes.c_int
    lib.ps_raster_backward_loss.restype = ctypes.c_int
    lib.ps_camera_setup.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 4 + [ctypes.c_int32] + \
        [ctypes.c_void_p] * 6
    lib.ps_camera_setup.restype = ctypes.c_int
    lib.ps_launch_count.restype = ctypes.c_ulonglong
    lib.ps_timing_enable.argtypes = [ctypes.c_int]
    lib.ps_timing_enable.restype = None
    lib.ps_timing_read.argtypes = [ctypes.POINTER(ctypes.c_float)]
    lib.ps_timing_read.restype = ctypes.c_int
    lib.ps_epipolar_geometry.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 9
    lib.ps_epipolar_geometry.restype = ctypes.c_int
    lib.ps_epipolar_attention_forward.argtypes = [P(EpipolarDesc), P(EpipolarInputs)] + [ctypes.c_void_p] * 5
    lib.ps_epipolar_attention_forward.restype = ctypes.c_int
    lib.ps_epipolar_attention_backward.argtypes = [P(EpipolarDesc), P(EpipolarInputs)] + [ctypes.c_void_p] * 10
    lib.ps_epipolar_attention_backward.restype = ctypes.c_int
    lib.ps_self_attention_forward.argtypes = [ctypes.c_int32] * 4 + [ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p,
                                                                      ctypes.c_int32, ctypes.c_void_p]
    lib.ps_self_attention_forward.restype = ctypes.c_int
    lib.ps_self_attention_forward_stats.argtypes = [ctypes.c_int32] * 4 + [ctypes.c_void_p, ctypes.c_float,
                                                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ps_self_attention_forward_stats.restype = ctypes.c_int
    lib.ps_self_attention_backward.argtypes = [ctypes.c_int32] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_float,
                                                                                             ctypes.c_void_p, ctypes.c_void_p]
    lib.ps_self_attention_backward.restype = ctypes.c_int
    lib.ps_gaussian_adapter_forward.argtypes = [P(AdapterDesc), P(AdapterInputs)] + [ctypes.c_void_p] * 6
    lib.ps_gaussian_adapter_forward.restype = ctypes.c_int
    lib.ps_gaussian_adapter_backward.argtypes = [P(AdapterDesc), P(AdapterInputs)] + [ctypes.c_void_p] * 9
    lib.ps_gaussian_adapter_backward.restype = ctypes.c_int
    lib.ps_sh_rotation_matrices.argtypes = [ctypes.c_int32] * 4 + [ctypes.c_void_p] * 5
    lib.ps_sh_rotation_matrices.restype = ctypes.c_int
    lib.ps_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.ps_set_option.restype = ctypes.c_int
    for f in ("ps_raster_sizes_query", "ps_raster_layout_query", "ps_raster_forward", "ps_raster_backward"):
        getattr(lib, f).restype = ctypes.c_int
    return lib


lib = _load()


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != PS_OK:
        msg = lib.ps_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == 1 else NativeError
        raise exc(f"{what}: {_ERR_NAMES.get(rc, rc)}: {msg}")


def on_device(device, fn, *args) -> int:
    """Calls a library entry point with `device` as the CURRENT CUDA device: the library creates its side
    stream / events and sets kernel attributes on the current device, so a tensor living on another GPU than
    the caller's current one must switch first (torch.cuda.device is a no-op when it already is current)."""
    import torch
    with torch.cuda.device(device):
        return fn(*args)


def set_option(name: str, value: int) -> None:
    check(lib.ps_set_option(name.encode(), int(value)), "ps_set_option")


def sizes(desc: RasterDesc) -> RasterSizes:
    out = RasterSizes()
    check(lib.ps_raster_sizes_query(ctypes.byref(desc), ctypes.byref(out)), "ps_raster_sizes_query")
    return out


def layout(desc: RasterDesc) -> RasterLayout:
    out = RasterLayout()
    check(lib.ps_raster_layout_query(ctypes.byref(desc), ctypes.byref(out)), "ps_raster_layout_query")
    return out
