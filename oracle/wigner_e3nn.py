"""TEST INFRASTRUCTURE -- restatement of the two e3nn functions the reference's `rotate_sh` calls
(/root/reference/src/misc/sh_rotation.py:18-22: `matrix_to_angles`, `wigner_D`).

e3nn is a third-party dependency (requirements.txt, unpinned) that is absent from /root/reference and
from this image, so its published algorithm (e3nn.o3._rotation / e3nn.o3._wigner, 0.5.x) is restated
here in float64:
  * rotations are parametrised by YXY Euler angles:  R = Ry(alpha) Rx(beta) Ry(gamma);
  * D^l(alpha, beta, gamma) = exp(alpha X_y) exp(beta X_x) exp(gamma X_y) with X the so(3) generators of
    degree l in e3nn's REAL basis, obtained from the su(2) ladder operators by the real<->complex change
    of basis (with e3nn's extra factor (-i)^l).
Only tests/ import this file.  It is an independent route to the same matrices that
pixelsplat_b200/sh.py obtains from the function-space definition Y_e3nn(R d) = D(R) Y_e3nn(d); the two
agreeing (tests/test_adapter_cpu.py) is what pins the convention.  PARITY NOTE: no e3nn golden vector
exists offline, so this file is pinned by (a) D^1(R) = R, e3nn's documented property, and (b) the
function-space identity against e3nn's published degree-1/2 polynomials restated in the test.
"""
from __future__ import annotations

import math

import torch


def matrix_x(a: torch.Tensor) -> torch.Tensor:
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([o, z, z], -1), torch.stack([z, c, -s], -1), torch.stack([z, s, c], -1)], -2)


def matrix_y(a: torch.Tensor) -> torch.Tensor:
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def angles_to_matrix(alpha, beta, gamma):
    return matrix_y(alpha) @ matrix_x(beta) @ matrix_y(gamma)


def xyz_to_angles(xyz: torch.Tensor):
    xyz = torch.nn.functional.normalize(xyz, p=2, dim=-1).clamp(-1, 1)
    beta = torch.acos(xyz[..., 1])
    alpha = torch.atan2(xyz[..., 0], xyz[..., 2])
    return alpha, beta


def matrix_to_angles(R: torch.Tensor):
    """YXY Euler angles of a proper rotation (e3nn.o3.matrix_to_angles)."""
    x = R @ R.new_tensor([0.0, 1.0, 0.0])
    a, b = xyz_to_angles(x)
    R = angles_to_matrix(a, b, torch.zeros_like(a)).transpose(-1, -2) @ R
    c = torch.atan2(R[..., 0, 2], R[..., 0, 0])
    return a, b, c


def su2_generators(j: int) -> torch.Tensor:
    m = torch.arange(-j, j, dtype=torch.float64)
    raising = torch.diag(-torch.sqrt(j * (j + 1) - m * (m + 1)), diagonal=-1)
    m = torch.arange(-j + 1, j + 1, dtype=torch.float64)
    lowering = torch.diag(torch.sqrt(j * (j + 1) - m * (m - 1)), diagonal=1)
    m = torch.arange(-j, j + 1, dtype=torch.float64)
    return torch.stack([0.5 * (raising + lowering).to(torch.complex128),
                        torch.diag(1j * m.to(torch.complex128)),
                        -0.5j * (raising - lowering).to(torch.complex128)], dim=0)


def change_basis_real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2 ** 0.5
        q[l + m, l - abs(m)] = -1j / 2 ** 0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
    return (-1j) ** l * q


def so3_generators(l: int) -> torch.Tensor:
    X = su2_generators(l)
    Q = change_basis_real_to_complex(l)
    X = torch.conj(Q.T) @ X @ Q
    assert X.imag.abs().max() < 1e-12
    return X.real


def wigner_D(l: int, alpha: torch.Tensor, beta: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    alpha, beta, gamma = torch.broadcast_tensors(alpha, beta, gamma)
    alpha = alpha[..., None, None] % (2 * math.pi)
    beta = beta[..., None, None] % (2 * math.pi)
    gamma = gamma[..., None, None] % (2 * math.pi)
    X = so3_generators(l)
    return torch.matrix_exp(alpha * X[1]) @ torch.matrix_exp(beta * X[0]) @ torch.matrix_exp(gamma * X[1])


def rotate_sh_reference(sh_coefficients: torch.Tensor, rotations: torch.Tensor) -> torch.Tensor:
    """The body of the reference's rotate_sh (sh_rotation.py:10-30) on the restated e3nn functions."""
    n = sh_coefficients.shape[-1]
    alpha, beta, gamma = matrix_to_angles(rotations.double())
    out = []
    for degree in range(math.isqrt(n)):
        D = wigner_D(degree, alpha, beta, gamma).to(sh_coefficients.dtype)
        out.append(torch.einsum("...ij,...j->...i", D, sh_coefficients[..., degree ** 2:(degree + 1) ** 2]))
    return torch.cat(out, dim=-1)
