"""Host side of the rasterizer: autograd glue over the C ABI (include/pixelsplat_b200.h) and the
drop-in `GaussianRasterizationSettings` / `GaussianRasterizer` pair that the reference imports
at /root/reference/src/model/decoder/cuda_splatting.py:5-8 and calls at :99-124.

PyTorch is plumbing here (device memory, the current stream, autograd bookkeeping); every
arithmetic step of the hot path runs in the CUDA library.  There is no CPU path: tensors that
are not on a CUDA device are rejected.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from typing import NamedTuple, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import PS_COV_3X3, PS_COV_TRIU6, PS_SH_M3, TILE
from .sh import convention_id

# ------------------------------------------------------------------ instance-capacity policy
# The binning buffers are sized for `capacity` (tile, Gaussian) instances.  The exact count is
# only known on the device; the forward returns it asynchronously through pinned memory.
#   "sync"     (default): wait for the count right after enqueueing the forward (one event wait
#              per *batch* of views -- upstream syncs once per view) and transparently re-run with
#              a larger buffer if it overflowed.  Always correct.
#   "deferred": a forward that will be followed by a backward (grad enabled, an input requires grad)
#              does not block; its count is verified when its backward starts AND at the start of the
#              next forward of the same shape, whichever comes first, and a RuntimeError is raised
#              if it had overflowed.  Forwards with no backward coming (inference) are always checked
#              synchronously, so a truncated image is never returned silently.
_CHECK_MODE = os.environ.get("PIXELSPLAT_B200_CAPACITY_CHECK", "sync")
_HEADROOM = 1.25        # capacity kept for the next call of a shape = headroom x the instances it needed
_capacity_hint: dict[tuple, int] = {}
_segment_hint: dict[tuple, int] = {}
_pending: dict[tuple, "weakref.ref"] = {}       # last deferred (unverified) state per shape key

# Pinned host slots for the asynchronous instance count: every RasterOutputState OWNS its slot for as
# long as it lives (a captured CUDA graph keeps writing to it on every replay) and hands it back to the
# pool when it is garbage collected -- no ring, no aliasing.
_PINNED_BLOCK = 64
_pinned_blocks: list[Tensor] = []
_pinned_free: list[Tensor] = []
_pool_lock = threading.Lock()

# SH convention the rasterizer evaluates coefficients in (include/pixelsplat_b200.h PS_SH_BASIS_*).
_SH_BASIS = convention_id(os.environ.get("PIXELSPLAT_B200_SH_BASIS", "3dgs"))


def set_capacity_check(mode: str) -> None:
    global _CHECK_MODE
    if mode not in ("sync", "deferred"):
        raise ValueError("mode must be 'sync' or 'deferred'")
    _CHECK_MODE = mode


def set_capacity_headroom(factor: float) -> None:
    """Head-room factor (>= 1) applied to the instance count a forward needed when sizing the binning buffers of
    the NEXT forward of the same shape (default 1.25).  A step captured into a CUDA graph freezes its capacity,
    so a training loop whose Gaussians move between replays should capture with a generous factor (the count is
    still checked after every replay: `RasterOutputState.verify`)."""
    global _HEADROOM
    if not factor >= 1.0:
        raise ValueError("headroom factor must be >= 1")
    _HEADROOM = float(factor)


def set_sh_basis(convention) -> None:
    """Process-wide default SH convention of the rasterizer: "3dgs" (upstream 3DGS basis, the default) or
    "e3nn" (the basis the reference's rotate_sh rotates in; see pixelsplat_b200/sh.py).  The drop-in
    `GaussianRasterizationSettings` has no field for it (it mirrors the extension's NamedTuple), hence a
    module switch; `rasterize_gaussians(sh_basis=...)` overrides it per call."""
    global _SH_BASIS
    _SH_BASIS = convention_id(convention)


def get_sh_basis() -> int:
    return _SH_BASIS


def _pinned_slot() -> Tensor:
    with _pool_lock:
        if not _pinned_free:
            block = torch.zeros(2 * _PINNED_BLOCK, dtype=torch.int64).pin_memory()
            _pinned_blocks.append(block)
            _pinned_free.extend(block[2 * i:2 * i + 2] for i in range(_PINNED_BLOCK))
        return _pinned_free.pop()


def _release_slot(slot: Tensor) -> None:
    with _pool_lock:
        _pinned_free.append(slot)


def _ptr(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: Tensor, name: str, shape: tuple) -> Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (pixelsplat_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


class RasterOutputState:
    """Opaque forward->backward state (the geom / binning / image byte buffers)."""

    def __init__(self, desc, geom, binning, image, n_host, event, hint_key=None):
        self.desc, self.geom, self.binning, self.image = desc, geom, binning, image
        self.n_host, self.event, self.hint_key = n_host, event, hint_key
        self.verified = False
        # the slot returns to the pool when this state dies; until then nothing else writes to it
        weakref.finalize(self, _release_slot, n_host)

    def raw_state(self) -> _lib.RasterState:
        return _lib.RasterState(self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(),
                                self.binning.numel(), self.image.data_ptr(), self.image.numel())

    def num_instances(self) -> int:
        """Blocks until the forward's instance count has reached the host.  For a forward that was
        captured into a CUDA graph there is no event: the caller synchronises after a replay."""
        if self.event is not None:
            self.event.synchronize()
        if self.hint_key is not None:
            _segment_hint[self.hint_key] = max(int(self.n_host[1].item()), 1)
        return int(self.n_host[0].item())

    def verify(self) -> None:
        """Raises if the binning overflowed its buffer.  Inside a CUDA-graph capture nothing can be
        waited on, so the check is skipped there: call `verify()` again after replaying and
        synchronising (bench.py does)."""
        if self.verified:
            return
        if self.event is None and torch.cuda.is_current_stream_capturing():
            return
        n = self.num_instances()
        if n > self.desc.instance_capacity:
            if self.hint_key is not None:
                _capacity_hint[self.hint_key] = int(n * _HEADROOM) + 4096
            raise RuntimeError(
                f"rasterizer binning overflow: {n} instances needed, capacity was "
                f"{self.desc.instance_capacity}; the forward result of this call is invalid. "
                "Re-run (the capacity hint has been raised) or use the 'sync' capacity check.")
        self.verified = True

    # -- introspection for parity tests (bit-exact tile/bin indices) --
    def intermediates(self) -> dict:
        self.verify()
        d = self.desc
        lay = _lib.layout(d)
        vt = d.n_scenes * d.views_per_scene
        vp = vt * d.n_gaussians
        gx, gy = (d.width + TILE - 1) // TILE, (d.height + TILE - 1) // TILE
        tiles = gx * gy
        n = self.num_instances()

        def view(buf, off, dtype, count, shape):
            nbytes = count * torch.empty((), dtype=dtype).element_size()
            return buf[off:off + nbytes].view(dtype).reshape(shape)

        P = d.n_gaussians
        return dict(
            depth=view(self.geom, lay.depth, torch.float32, vp, (vt, P)),
            radii=view(self.geom, lay.radii, torch.int32, vp, (vt, P)),
            xy=view(self.geom, lay.xy, torch.float32, vp * 2, (vt, P, 2)),
            conic_opacity=view(self.geom, lay.conic_opacity, torch.float32, vp * 4, (vt, P, 4)),
            rgb=view(self.geom, lay.rgb, torch.float32, vp * 4, (vt, P, 4))[..., :3],
            rect=view(self.geom, lay.rect, torch.int16, vp * 4, (vt, P, 4)),
            clamped=view(self.geom, lay.clamped, torch.uint8, vp, (vt, P)),
            tile_count=view(self.geom, lay.tile_count, torch.int32, vt * tiles, (vt, tiles)),
            tile_start=view(self.geom, lay.tile_start, torch.int32, vt * tiles, (vt, tiles)),
            keys=view(self.binning, lay.keys, torch.int64, n, (n,)),
            final_T=view(self.image, lay.final_T, torch.float32, vt * d.height * d.width,
                         (vt, d.height, d.width)),
            n_contrib=view(self.image, lay.n_contrib, torch.int32, vt * d.height * d.width,
                           (vt, d.height, d.width)),
            num_instances=n,
        )


def _make_desc(S, V, P, M, deg, sh_layout, cov_layout, H, W, capacity, sort_impl, seg_hint=0,
               sh_basis=0) -> _lib.RasterDesc:
    return _lib.RasterDesc(S, V, P, M, deg, sh_layout, cov_layout, H, W, sort_impl, seg_hint, capacity,
                           sh_basis, 0)


def _forward_native(means, cov, opac, sh, cams, S, V, P, M, deg, sh_layout, cov_layout, H, W,
                    sort_impl, want_radii, sh_basis=0, backward_follows=False, loss_target=None, want_color=True):
    """Returns (color | None, radii | None, state) and, with `loss_target`, a 4th element: the
    [S*V, 2, LOSS_SLOTS] partial sums of the fused loss epilogue."""
    dev = means.device
    with torch.cuda.device(dev):            # the library works on the CURRENT device
        return _forward_on_device(means, cov, opac, sh, cams, S, V, P, M, deg, sh_layout, cov_layout, H, W,
                                  sort_impl, want_radii, sh_basis, backward_follows, loss_target, want_color)


def _forward_on_device(means, cov, opac, sh, cams, S, V, P, M, deg, sh_layout, cov_layout, H, W,
                       sort_impl, want_radii, sh_basis, backward_follows, loss_target=None, want_color=True):
    dev = means.device
    key = (dev.index, S, V, P, H, W)
    capturing = torch.cuda.is_current_stream_capturing()
    prev = _pending.pop(key, None)
    if prev is not None and not capturing:
        prev = prev()
        if prev is not None:
            prev.verify()                   # deferred check of the previous forward of this shape
    capacity = _capacity_hint.get(key)
    if capacity is None:
        capacity = max(4096, 3 * S * V * P)
    stream = torch.cuda.current_stream(dev)
    while True:
        desc = _make_desc(S, V, P, M, deg, sh_layout, cov_layout, H, W, capacity, sort_impl,
                          min(_segment_hint.get(key, 0), 1 << 30), sh_basis)
        sz = _lib.sizes(desc)
        geom = torch.empty(sz.geom_bytes, dtype=torch.uint8, device=dev)
        binning = torch.empty(sz.binning_bytes, dtype=torch.uint8, device=dev)
        image = torch.empty(sz.image_bytes, dtype=torch.uint8, device=dev)
        color = torch.empty((S * V, 3, H, W), dtype=torch.float32, device=dev) if want_color else None
        radii = torch.empty((S * V, P), dtype=torch.int32, device=dev) if want_radii else None
        sums = (torch.empty((S * V, 2, _lib.LOSS_SLOTS), dtype=torch.float32, device=dev)
                if loss_target is not None else None)
        n_host = _pinned_slot()
        inputs = _lib.RasterInputs(
            means.data_ptr(), cov.data_ptr(), opac.data_ptr(), sh.data_ptr(),
            cams["viewmatrix"].data_ptr(), cams["projmatrix"].data_ptr(), cams["campos"].data_ptr(),
            cams["tanfov"].data_ptr(), cams["background"].data_ptr(),
            cams["scene_scale"].data_ptr() if cams.get("scene_scale") is not None else None)
        state = _lib.RasterState(geom.data_ptr(), geom.numel(), binning.data_ptr(), binning.numel(),
                                 image.data_ptr(), image.numel())
        if loss_target is None:
            rc = _lib.lib.ps_raster_forward(ctypes.byref(desc), ctypes.byref(inputs), ctypes.byref(state),
                                            _ptr(color), _ptr(radii), ctypes.c_void_p(n_host.data_ptr()),
                                            ctypes.c_void_p(stream.cuda_stream))
            _lib.check(rc, "ps_raster_forward")
            ret = lambda st_: (color, radii, st_)
        else:
            loss = _lib.RasterLoss(loss_target.data_ptr(), sums.data_ptr())
            rc = _lib.lib.ps_raster_forward_loss(ctypes.byref(desc), ctypes.byref(inputs), ctypes.byref(state),
                                                 ctypes.byref(loss), _ptr(color), _ptr(radii),
                                                 ctypes.c_void_p(n_host.data_ptr()), ctypes.c_void_p(stream.cuda_stream))
            _lib.check(rc, "ps_raster_forward_loss")
            ret = lambda st_: (color, radii, st_, sums)
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: shapes and capacity are frozen into the graph; the count still
            # lands in pinned memory on every replay and is checked by the caller afterwards
            if key not in _capacity_hint:
                raise RuntimeError("run this shape once eagerly before capturing it in a CUDA graph "
                                   "(the binning capacity must be known)")
            return ret(RasterOutputState(desc, geom, binning, image, n_host, None, key))
        event = torch.cuda.Event()
        event.record(stream)
        st = RasterOutputState(desc, geom, binning, image, n_host, event, key)
        if _CHECK_MODE == "deferred" and key in _capacity_hint and backward_follows:
            _pending[key] = weakref.ref(st)
            return ret(st)
        n = st.num_instances()
        if n <= capacity:
            st.verified = True
            # keep ~25 % head-room for the next call of the same shape
            _capacity_hint[key] = max(_capacity_hint.get(key, 0), int(n * _HEADROOM) + 4096)
            return ret(st)
        capacity = int(n * _HEADROOM) + 4096
        _capacity_hint[key] = capacity


class _RasterizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, cov, opac, sh, means2d, cams, S, V, P, M, deg, sh_layout, cov_layout,
                H, W, sort_impl, state_out, sh_basis):
        backward_follows = any(ctx.needs_input_grad[:5])
        color, radii, st = _forward_native(means, cov, opac, sh, cams, S, V, P, M, deg, sh_layout,
                                           cov_layout, H, W, sort_impl, True, sh_basis, backward_follows)
        ctx.save_for_backward(means, cov, opac, sh)
        ctx.cams, ctx.st = cams, st
        ctx.want_m2d = means2d is not None and means2d.requires_grad
        if state_out is not None:
            state_out.append(st)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, d_color, _d_radii):
        means, cov, opac, sh = ctx.saved_tensors
        st: RasterOutputState = ctx.st
        st.verify()
        desc, cams = st.desc, ctx.cams
        dev = means.device
        d_color = d_color.contiguous()
        if d_color.dtype != torch.float32:
            d_color = d_color.float()
        sz = _lib.sizes(desc)
        scratch = torch.empty(sz.backward_bytes, dtype=torch.uint8, device=dev)
        d_means = torch.empty_like(means)
        d_cov = torch.empty_like(cov)
        d_opac = torch.empty_like(opac)
        d_sh = torch.empty_like(sh)
        VT = desc.n_scenes * desc.views_per_scene
        d_m2d = (torch.empty((VT, desc.n_gaussians, 3), dtype=torch.float32, device=dev)
                 if ctx.want_m2d else None)
        inputs = _lib.RasterInputs(
            means.data_ptr(), cov.data_ptr(), opac.data_ptr(), sh.data_ptr(),
            cams["viewmatrix"].data_ptr(), cams["projmatrix"].data_ptr(), cams["campos"].data_ptr(),
            cams["tanfov"].data_ptr(), cams["background"].data_ptr(),
            cams["scene_scale"].data_ptr() if cams.get("scene_scale") is not None else None)
        grads = _lib.RasterGrads(d_means.data_ptr(), d_cov.data_ptr(), d_opac.data_ptr(),
                                 d_sh.data_ptr(), d_m2d.data_ptr() if d_m2d is not None else None)
        state = st.raw_state()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            rc = _lib.lib.ps_raster_backward(ctypes.byref(desc), ctypes.byref(inputs), ctypes.byref(state),
                                             ctypes.c_void_p(d_color.data_ptr()),
                                             ctypes.c_void_p(scratch.data_ptr()), scratch.numel(),
                                             ctypes.byref(grads), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "ps_raster_backward")
        return (d_means, d_cov, d_opac, d_sh, d_m2d) + (None,) * 13


def rasterize_gaussians(
    means: Tensor,            # [S, P, 3]
    covariances: Tensor,      # [S, P, 6] (triu) or [S, P, 3, 3]
    opacities: Tensor,        # [S, P]
    colors: Tensor,           # SH [S, P, M, 3] / [S, P, 3, M], or precomputed RGB [S, P, 3]
    *,
    viewmatrix: Tensor,       # [S*V, 16] or [S*V, 4, 4], column-major flattening (see header)
    projmatrix: Tensor,       # same
    campos: Tensor,           # [S*V, 3]
    tanfov: Tensor,           # [S*V, 2]
    background: Tensor,       # [S*V, 3]
    image_shape: tuple[int, int],
    views_per_scene: int,
    sh_degree: int,
    use_sh: bool = True,
    sh_layout: int = PS_SH_M3,
    scene_scale: Optional[Tensor] = None,   # [S*V]
    sort_impl: int = 0,
    state_out: Optional[list] = None,
    means2d: Optional[Tensor] = None,       # [S*V, P, 3] gradient holder (upstream's means2D)
    sh_basis=None,                          # "3dgs" / "e3nn"; None = the module default (set_sh_basis)
) -> tuple[Tensor, Tensor]:
    """Batched differentiable rasterization: S scenes x V views in one set of launches.
    Returns (color [S*V, 3, H, W], radii [S*V, P] int32)."""
    means, covariances, opacities, colors, cams, S, V, P, M, cov_layout, H, W = _prepare(
        means, covariances, opacities, colors, viewmatrix, projmatrix, campos, tanfov, background, image_shape,
        views_per_scene, use_sh, sh_layout, scene_scale)
    VT = S * V
    if means2d is not None and tuple(means2d.shape) != (VT, P, 3):
        raise ValueError(f"means2d must be [S*V, P, 3], got {tuple(means2d.shape)}")
    return _RasterizeFn.apply(means, covariances, opacities, colors, means2d, cams, S, V, P, M,
                              int(sh_degree), sh_layout, cov_layout, int(H), int(W), int(sort_impl),
                              state_out, _SH_BASIS if sh_basis is None else convention_id(sh_basis))


class _RasterizeMseFn(torch.autograd.Function):
    """Rasterize + squared error against a target in one pass.  Differentiable output:
    sse [S*V] = sum over the view's pixels and channels of (C - target)^2; its gradient g [S*V] reaches the
    composite backward as the per-view scale 2 g of (C - target), formed in-kernel."""

    @staticmethod
    def forward(ctx, means, cov, opac, sh, target, cams, S, V, P, M, deg, sh_layout, cov_layout, H, W, sort_impl,
                state_out, sh_basis, want_color):
        backward_follows = any(ctx.needs_input_grad[:4])
        color, radii, st, sums = _forward_native(means, cov, opac, sh, cams, S, V, P, M, deg, sh_layout, cov_layout,
                                                 H, W, sort_impl, True, sh_basis, backward_follows,
                                                 loss_target=target, want_color=want_color)
        ctx.save_for_backward(means, cov, opac, sh, target)
        ctx.cams, ctx.st = cams, st
        if state_out is not None:
            state_out.append(st)
        totals = sums.sum(dim=-1)                                    # [S*V, 2]
        sse, sse_clipped = totals[:, 0].contiguous(), totals[:, 1].contiguous()
        if color is None:
            color = torch.empty(0, device=means.device)
        ctx.mark_non_differentiable(sse_clipped, color, radii)
        return sse, sse_clipped, color, radii

    @staticmethod
    def backward(ctx, d_sse, _d_clip, _d_color, _d_radii):
        means, cov, opac, sh, target = ctx.saved_tensors
        st: RasterOutputState = ctx.st
        st.verify()
        desc, cams = st.desc, ctx.cams
        dev = means.device
        scale = (2.0 * d_sse).to(torch.float32).contiguous()
        sz = _lib.sizes(desc)
        scratch = torch.empty(sz.backward_bytes, dtype=torch.uint8, device=dev)
        d_means, d_cov = torch.empty_like(means), torch.empty_like(cov)
        d_opac, d_sh = torch.empty_like(opac), torch.empty_like(sh)
        inputs = _lib.RasterInputs(
            means.data_ptr(), cov.data_ptr(), opac.data_ptr(), sh.data_ptr(),
            cams["viewmatrix"].data_ptr(), cams["projmatrix"].data_ptr(), cams["campos"].data_ptr(),
            cams["tanfov"].data_ptr(), cams["background"].data_ptr(),
            cams["scene_scale"].data_ptr() if cams.get("scene_scale") is not None else None)
        grads = _lib.RasterGrads(d_means.data_ptr(), d_cov.data_ptr(), d_opac.data_ptr(), d_sh.data_ptr(), None)
        state = st.raw_state()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            rc = _lib.lib.ps_raster_backward_loss(ctypes.byref(desc), ctypes.byref(inputs), ctypes.byref(state),
                                                  ctypes.c_void_p(target.data_ptr()), ctypes.c_void_p(scale.data_ptr()),
                                                  ctypes.c_void_p(scratch.data_ptr()), scratch.numel(),
                                                  ctypes.byref(grads), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "ps_raster_backward_loss")
        return (d_means, d_cov, d_opac, d_sh) + (None,) * 15


def _prepare(means, covariances, opacities, colors, viewmatrix, projmatrix, campos, tanfov, background, image_shape,
             views_per_scene, use_sh, sh_layout, scene_scale):
    """Argument checks shared by rasterize_gaussians and rasterize_gaussians_mse."""
    if means.dim() != 3 or means.shape[-1] != 3:
        raise ValueError(f"means must be [S, P, 3], got {tuple(means.shape)}")
    S, P, _ = means.shape
    V = int(views_per_scene)
    H, W = image_shape
    means = _req(means, "means", (S, P, 3))
    if covariances.dim() == 4:
        cov_layout = PS_COV_3X3
        covariances = _req(covariances, "covariances", (S, P, 3, 3))
    else:
        cov_layout = PS_COV_TRIU6
        covariances = _req(covariances, "covariances", (S, P, 6))
    opacities = _req(opacities.reshape(S, P), "opacities", (S, P))
    if use_sh:
        if colors.dim() != 4:
            raise ValueError("SH colours must be [S, P, M, 3] or [S, P, 3, M]")
        M = colors.shape[2] if sh_layout == PS_SH_M3 else colors.shape[3]
        shape = (S, P, M, 3) if sh_layout == PS_SH_M3 else (S, P, 3, M)
        colors = _req(colors, "sh", shape)
    else:
        M = 0
        colors = _req(colors, "colors_precomp", (S, P, 3))
    VT = S * V
    cams = dict(
        viewmatrix=_req(viewmatrix.reshape(VT, 16), "viewmatrix", (VT, 16)),
        projmatrix=_req(projmatrix.reshape(VT, 16), "projmatrix", (VT, 16)),
        campos=_req(campos, "campos", (VT, 3)),
        tanfov=_req(tanfov, "tanfov", (VT, 2)),
        background=_req(background, "background", (VT, 3)),
        scene_scale=None if scene_scale is None else _req(scene_scale, "scene_scale", (VT,)),
    )
    return means, covariances, opacities, colors, cams, S, V, P, M, cov_layout, int(H), int(W)


def rasterize_gaussians_mse(means: Tensor, covariances: Tensor, opacities: Tensor, colors: Tensor, target: Tensor, *,
                            viewmatrix: Tensor, projmatrix: Tensor, campos: Tensor, tanfov: Tensor, background: Tensor,
                            image_shape: tuple[int, int], views_per_scene: int, sh_degree: int, use_sh: bool = True,
                            sh_layout: int = PS_SH_M3, scene_scale: Optional[Tensor] = None, sort_impl: int = 0,
                            state_out: Optional[list] = None, sh_basis=None, want_color: bool = True):
    """rasterize_gaussians + the loss epilogue: returns (sse [S*V] differentiable, sse_clipped [S*V], color
    [S*V, 3, H, W] detached (empty when want_color=False), radii).  `target` is [S*V, 3, H, W]."""
    means, covariances, opacities, colors, cams, S, V, P, M, cov_layout, H, W = _prepare(
        means, covariances, opacities, colors, viewmatrix, projmatrix, campos, tanfov, background, image_shape,
        views_per_scene, use_sh, sh_layout, scene_scale)
    target = _req(target, "target", (S * V, 3, H, W))
    return _RasterizeMseFn.apply(means, covariances, opacities, colors, target, cams, S, V, P, M, int(sh_degree),
                                 sh_layout, cov_layout, H, W, int(sort_impl), state_out,
                                 _SH_BASIS if sh_basis is None else convention_id(sh_basis), bool(want_color))


# ------------------------------------------------------------------ drop-in extension surface
class GaussianRasterizationSettings(NamedTuple):
    """Field-for-field the NamedTuple the reference constructs by keyword at
    cuda_splatting.py:99-112."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


def _cov_from_scale_rotation(scales: Tensor, rotations: Tensor, mod: float) -> Tensor:
    """Upstream computeCov3D: quaternion (r, x, y, z) used un-normalised; Sigma = (S R)^T (S R)."""
    r, x, y, z = rotations.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * (mod * scales)[:, None, :]
    sigma = M @ M.transpose(1, 2)
    row, col = torch.triu_indices(3, 3, device=sigma.device)
    return sigma[:, row, col]


class GaussianRasterizer(torch.nn.Module):
    """Same constructor / forward signature and error behaviour as the extension class the
    reference instantiates at cuda_splatting.py:113."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if means3D.dim() != 2 or means3D.shape[-1] != 3:
            raise ValueError("means3D must have dimensions (num_points, 3)")
        P = means3D.shape[0]
        dev = means3D.device
        cov6 = cov3D_precomp if cov3D_precomp is not None else _cov_from_scale_rotation(
            scales, rotations, float(rs.scale_modifier))
        f32 = dict(dtype=torch.float32, device=dev)
        tanfov = torch.tensor([[float(rs.tanfovx), float(rs.tanfovy)]], **f32)
        color, radii = rasterize_gaussians(
            means3D[None], cov6[None], opacities.reshape(1, P),
            (shs if shs is not None else colors_precomp)[None],
            viewmatrix=rs.viewmatrix.reshape(1, 16).to(**f32),
            projmatrix=rs.projmatrix.reshape(1, 16).to(**f32),
            campos=rs.campos.reshape(1, 3).to(**f32), tanfov=tanfov,
            background=rs.bg.reshape(1, 3).to(**f32),
            image_shape=(int(rs.image_height), int(rs.image_width)), views_per_scene=1,
            sh_degree=int(rs.sh_degree), use_sh=shs is not None, sh_layout=PS_SH_M3,
            means2d=None if means2D is None else means2D[None])
        return color[0], radii[0]
