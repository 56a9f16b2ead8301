"""Real spherical harmonics, degree <= 4, in the basis the rasterizer evaluates
(csrc/raster_math.cuh `sh_for_each`), and rotation of coefficient vectors.

`rotate_sh` stands in for /root/reference/src/misc/sh_rotation.py:10-30, which builds Wigner-D
matrices with e3nn (absent offline, and in e3nn's own axis convention).  Here the rotation is DEFINED
by consistency with the rasterizer's basis: for a rotation R (camera-to-world), the rotated
coefficients c' satisfy  sum_i c'_i Y_i(d) = sum_i c_i Y_i(R^T d)  for every direction d -- a colour lobe
that pointed along p in the camera frame points along R p in the world frame.  The block-diagonal
matrix D(R) (blocks 1, 3, 5, 7, 9) is obtained by an exact least-squares fit over a fixed set of
directions in float64; there are only b*v distinct rotations per step, so this costs nothing.
Parity with the reference's e3nn call is UNPINNED; the properties that define a
correct rotation (identity, composition, orthogonality, function consistency) are tested.
"""
from __future__ import annotations

from functools import lru_cache

import torch
from torch import Tensor

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)
C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def sh_basis(dirs: Tensor, degree: int = 4) -> Tensor:
    """dirs [..., 3] (unit vectors) -> [..., (degree + 1)^2], index order l^2 + (m + l)."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, C0)]
    if degree >= 1:
        out += [-C1 * y, C1 * z, -C1 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if degree >= 3:
        out += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                C3[6] * x * (xx - 3 * yy)]
    if degree >= 4:
        out += [C4[0] * xy * (xx - yy), C4[1] * yz * (3 * xx - yy), C4[2] * xy * (7 * zz - 1),
                C4[3] * yz * (7 * zz - 3), C4[4] * (zz * (35 * zz - 30) + 3), C4[5] * xz * (7 * zz - 3),
                C4[6] * (xx - yy) * (7 * zz - 1), C4[7] * xz * (xx - 3 * yy),
                C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(out, dim=-1)


@lru_cache(maxsize=None)
def _fit_directions(n: int = 192) -> Tensor:
    """Fibonacci sphere, float64, CPU."""
    i = torch.arange(n, dtype=torch.float64) + 0.5
    z = 1.0 - 2.0 * i / n
    phi = i * (torch.pi * (3.0 - 5.0 ** 0.5))
    r = (1.0 - z * z).clamp_min(0).sqrt()
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1)


def sh_rotation_matrices(rotations: Tensor, degree: int = 4) -> Tensor:
    """rotations [..., 3, 3] -> D [..., n, n] (n = (degree + 1)^2, block diagonal) with c' = D c."""
    n = (degree + 1) ** 2
    lead = rotations.shape[:-2]
    R = rotations.reshape(-1, 3, 3).to(torch.float64)
    dirs = _fit_directions().to(R.device)                               # [m, 3]
    Y = sh_basis(dirs, degree)                                           # [m, n]
    Yr = sh_basis(torch.einsum("kji,mj->kmi", R, dirs), degree)          # Y(R^T d)  [k, m, n]
    D = torch.zeros((R.shape[0], n, n), dtype=torch.float64, device=R.device)
    for l in range(degree + 1):
        s = slice(l * l, (l + 1) ** 2)
        # per-degree blocks: Y_l D_l = Yr_l   (rotation never mixes degrees)
        D[:, s, s] = torch.linalg.lstsq(Y[:, s].expand(R.shape[0], -1, -1), Yr[:, :, s]).solution
    return D.reshape(*lead, n, n).to(rotations.dtype)


def rotate_sh(sh_coefficients: Tensor, rotations: Tensor) -> Tensor:
    """Same call shape as the reference's rotate_sh: coefficients [*#batch, n], rotations [*#batch, 3, 3]."""
    n = sh_coefficients.shape[-1]
    degree = int(round(n ** 0.5)) - 1
    D = sh_rotation_matrices(rotations, degree)
    return torch.einsum("...ij,...j->...i", D, sh_coefficients)


# ---- device-side construction (one tiny kernel per step: csrc/gaussian_adapter.cu `k_sh_rotation`) --------
_FIT_CACHE: dict = {}


def fit_operators(device, degree: int = 4) -> tuple[Tensor, Tensor]:
    """(fit directions [m, 3], per-degree pseudo-inverse of the basis there [n, m]), float32 on `device`,
    built once in float64: D_l(R) = pinv_l @ Y_l(R^T dirs)."""
    key = (str(device), degree)
    if key not in _FIT_CACHE:
        dirs = _fit_directions()
        Y = sh_basis(dirs, degree)                                       # [m, n] float64
        pinv = torch.zeros((Y.shape[1], Y.shape[0]), dtype=torch.float64)
        for l in range(degree + 1):
            s = slice(l * l, (l + 1) ** 2)
            pinv[s] = torch.linalg.pinv(Y[:, s])
        _FIT_CACHE[key] = (dirs.float().contiguous().to(device), pinv.float().contiguous().to(device))
    return _FIT_CACHE[key]


def camera_sh_rotations(extrinsics: Tensor, degree: int = 4) -> Tensor:
    """extrinsics [n, 4, 4] (CUDA, camera-to-world) -> D [n, (degree+1)^2, (degree+1)^2], float32, via the
    library kernel (no host synchronisation; equals sh_rotation_matrices(extrinsics[:, :3, :3]) to ~1e-6)."""
    import ctypes

    from . import _lib
    if not extrinsics.is_cuda:
        raise ValueError("pixelsplat_b200 has no CPU path: camera_sh_rotations needs a CUDA tensor")
    E = extrinsics.detach().contiguous().float()
    n = (degree + 1) ** 2
    dirs, pinv = fit_operators(E.device, degree)
    out = torch.empty((E.shape[0], n, n), dtype=torch.float32, device=E.device)
    stream = torch.cuda.current_stream(E.device)
    rc = _lib.lib.ps_sh_rotation_matrices(E.shape[0], n, dirs.shape[0], ctypes.c_void_p(E.data_ptr()),
                                          ctypes.c_void_p(dirs.data_ptr()), ctypes.c_void_p(pinv.data_ptr()),
                                          ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "ps_sh_rotation_matrices")
    return out
