"""Real spherical harmonics, degree <= 4, and rotation of coefficient vectors, in two conventions.

`rotate_sh` replaces /root/reference/src/misc/sh_rotation.py:10-30, which is two e3nn calls
(`wigner_D(l, *matrix_to_angles(R))`, e3nn absent offline).  e3nn's matrices are fully determined by
the property its documentation (and the reference's own demo, sh_rotation.py:33-80, which evaluates
`e3nn.o3.spherical_harmonics` on the rotated coefficients) states:

        Y_e3nn(R d) = D_e3nn(R) Y_e3nn(d)          for every direction d,

where Y_e3nn are the real spherical harmonics **with y as the polar axis and no Condon-Shortley sign**
(degree 1 is exactly (x, y, z), so D^1(R) = R).  The rasterizer evaluates the 3DGS basis (SURVEY.md
Appendix A.4: z polar, Condon-Shortley sign (-1)^m; degree 1 is (-y, z, -x)).  The two are related by an
axis permutation and a sign per coefficient:

        Y_e3nn,i(x, y, z) = s_i * Y_3dgs,i(z, x, y),        s_i = (-1)^m,  i = l^2 + l + m.

`convention` selects which of the two a function speaks:
  * "e3nn" (DEFAULT of `rotate_sh` / `camera_sh_rotations` / the fused GaussianAdapter): bit-for-bit
    what the reference computes -- D_e3nn(R) applied to coefficients l^2..(l+1)^2.  Checked in
    tests/test_adapter_cpu.py against an independent restatement of e3nn's published construction
    (YXY Euler angles -> matrix exponentials of the real so(3) generators, oracle/wigner_e3nn.py) and
    against the closed form D^1(R) = R.
  * "3dgs": the rotation that is physically consistent with the rasterizer's default basis (a lobe
    that points along p in the camera frame points along R p in the world frame).  Opt-in.
The rasterizer has the matching switch (`PS_SH_BASIS_E3NN`, pixelsplat_b200.rasterizer.set_sh_basis):
with basis "e3nn" + rotation "e3nn" the pair is physically consistent as well.

D(R) is obtained per degree by an exact least-squares fit of the defining identity over a fixed set of
directions in float64 (host) or by the library kernel `k_sh_rotation` (device, no host sync); there
are only b*v distinct rotations per step.
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

CONVENTIONS = {"3dgs": 0, "e3nn": 1}      # values = PS_SH_BASIS_* of include/pixelsplat_b200.h


def convention_id(convention) -> int:
    if convention in (0, 1):
        return int(convention)
    try:
        return CONVENTIONS[convention]
    except KeyError:
        raise ValueError(f"unknown SH convention {convention!r}; expected one of {sorted(CONVENTIONS)}") from None


def sh_signs(degree: int = 4, dtype=torch.float64, device=None) -> Tensor:
    """s_i = (-1)^m for i = l^2 + l + m: the Condon-Shortley signs the 3DGS basis carries and e3nn's does not."""
    s = [(-1.0) ** abs(i - l * l - l) for l in range(degree + 1) for i in range(l * l, (l + 1) ** 2)]
    return torch.tensor(s, dtype=dtype, device=device)


def _sh_basis_3dgs(dirs: Tensor, degree: int) -> Tensor:
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


def sh_basis(dirs: Tensor, degree: int = 4, convention="3dgs") -> Tensor:
    """dirs [..., 3] (unit vectors) -> [..., (degree + 1)^2], index order l^2 + (m + l).

    "3dgs": the basis the rasterizer evaluates by default (csrc/raster_math.cuh `sh_for_each`);
    "e3nn": orthonormal ("integral"-normalised) e3nn harmonics = s_i * Y_3dgs,i(z, x, y)."""
    if convention_id(convention) == 0:
        return _sh_basis_3dgs(dirs, degree)
    x, y, z = dirs.unbind(-1)
    return _sh_basis_3dgs(torch.stack([z, x, y], dim=-1), degree) * sh_signs(degree, dirs.dtype, dirs.device)


@lru_cache(maxsize=None)
def _fit_directions(n: int = 192) -> Tensor:
    """Fibonacci sphere, float64, CPU."""
    i = torch.arange(n, dtype=torch.float64) + 0.5
    z = 1.0 - 2.0 * i / n
    phi = i * (torch.pi * (3.0 - 5.0 ** 0.5))
    r = (1.0 - z * z).clamp_min(0).sqrt()
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1)


def sh_rotation_matrices(rotations: Tensor, degree: int = 4, convention="e3nn") -> Tensor:
    """rotations [..., 3, 3] -> D [..., n, n] (n = (degree + 1)^2, block diagonal) with c' = D c, where
    Y(R d) = D(R) Y(d) in the chosen convention's basis (equivalently: the function with coefficients
    c' evaluated at d equals the function with coefficients c evaluated at R^T d)."""
    n = (degree + 1) ** 2
    lead = rotations.shape[:-2]
    R = rotations.reshape(-1, 3, 3).to(torch.float64)
    dirs = _fit_directions().to(R.device)                               # [m, 3]
    Y = sh_basis(dirs, degree, convention)                               # [m, n]
    Yr = sh_basis(torch.einsum("kji,mj->kmi", R, dirs), degree, convention)   # Y(R^T d)  [k, m, n]
    D = torch.zeros((R.shape[0], n, n), dtype=torch.float64, device=R.device)
    for l in range(degree + 1):
        s = slice(l * l, (l + 1) ** 2)
        # per-degree blocks: Y_l(d)^T D_l = Y_l(R^T d)^T   (rotation never mixes degrees)
        D[:, s, s] = torch.linalg.lstsq(Y[:, s].expand(R.shape[0], -1, -1), Yr[:, :, s]).solution
    return D.reshape(*lead, n, n).to(rotations.dtype)


def rotate_sh(sh_coefficients: Tensor, rotations: Tensor, convention="e3nn") -> Tensor:
    """Same call shape as the reference's rotate_sh (sh_rotation.py:10-30): coefficients [*#batch, n],
    rotations [*#batch, 3, 3]; the default convention is the reference's (e3nn Wigner-D)."""
    n = sh_coefficients.shape[-1]
    degree = int(round(n ** 0.5)) - 1
    D = sh_rotation_matrices(rotations, degree, convention)
    return torch.einsum("...ij,...j->...i", D, sh_coefficients)


# ---- device-side construction (one tiny kernel per step: csrc/gaussian_adapter.cu `k_sh_rotation`) --------
_FIT_CACHE: dict = {}


def fit_operators(device, degree: int = 4) -> tuple[Tensor, Tensor]:
    """(fit directions [m, 3], per-degree pseudo-inverse of the 3DGS basis there [n, m]), float32 on
    `device`, built once in float64: D_l(R) = pinv_l @ Y_l(R^T dirs).  The kernel derives the e3nn
    convention from the same operators (axis permutation of R, sign per entry)."""
    key = (str(device), degree)
    if key not in _FIT_CACHE:
        dirs = _fit_directions()
        Y = sh_basis(dirs, degree, "3dgs")                               # [m, n] float64
        pinv = torch.zeros((Y.shape[1], Y.shape[0]), dtype=torch.float64)
        for l in range(degree + 1):
            s = slice(l * l, (l + 1) ** 2)
            pinv[s] = torch.linalg.pinv(Y[:, s])
        _FIT_CACHE[key] = (dirs.float().contiguous().to(device), pinv.float().contiguous().to(device))
    return _FIT_CACHE[key]


def camera_sh_rotations(extrinsics: Tensor, degree: int = 4, convention="e3nn") -> Tensor:
    """extrinsics [n, 4, 4] (CUDA, camera-to-world) -> D [n, (degree+1)^2, (degree+1)^2], float32, via the
    library kernel (no host synchronisation; equals sh_rotation_matrices(extrinsics[:, :3, :3], degree,
    convention) to ~1e-6)."""
    import ctypes

    from . import _lib
    if not extrinsics.is_cuda:
        raise ValueError("pixelsplat_b200 has no CPU path: camera_sh_rotations needs a CUDA tensor")
    E = extrinsics.detach().contiguous().float()
    n = (degree + 1) ** 2
    dirs, pinv = fit_operators(E.device, degree)
    out = torch.empty((E.shape[0], n, n), dtype=torch.float32, device=E.device)
    with torch.cuda.device(E.device):
        stream = torch.cuda.current_stream(E.device)
        rc = _lib.lib.ps_sh_rotation_matrices(E.shape[0], n, dirs.shape[0], convention_id(convention),
                                              ctypes.c_void_p(E.data_ptr()), ctypes.c_void_p(dirs.data_ptr()),
                                              ctypes.c_void_p(pinv.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "ps_sh_rotation_matrices")
    return out
