"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ORBIT math helpers on the hot path.

ORBIT (``omni.isaac.orbit.utils.math``, github.com/NVIDIA-Omniverse/orbit, unpinned, ~Feb-Mar 2024)
is a third-party dependency absent from ``/root/reference``; its published algorithm is restated here
from SURVEY.md Appendix A.1.  Call sites in the reference that anchor the semantics:
``terrain_importer.py:13`` (import), ``terrain_importer.py:99-101`` (``_update_command``),
``terrain_importer.py:103-106`` (``_update_metrics``).  Quaternions are ``(w, x, y, z)``; fp32 throughout.

PARITY UNPINNED for these helpers: the reference ships no test for them; they are
cross-checked by invariants in ``tests/test_oracle_math.py`` (rotation by yaw vs. explicit 2x2
rotation matrix in float64, norm preservation, wrap range).
"""
from __future__ import annotations

import math

import torch


def normalize(x: torch.Tensor, eps: float = 1e-9) -> torch.Tensor:
    """A.1 ``normalize``: ``x / ||x||.clamp(min=eps)``."""
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def yaw_quat(quat: torch.Tensor) -> torch.Tensor:
    """A.1 ``yaw_quat``: yaw-only quaternion of ``quat`` (w,x,y,z), re-normalised."""
    quat_yaw = quat.clone().view(-1, 4)
    qw = quat_yaw[:, 0]
    qx = quat_yaw[:, 1]
    qy = quat_yaw[:, 2]
    qz = quat_yaw[:, 3]
    yaw = torch.atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
    quat_yaw[:] = 0.0
    quat_yaw[:, 3] = torch.sin(yaw / 2)
    quat_yaw[:, 0] = torch.cos(yaw / 2)
    quat_yaw = normalize(quat_yaw)
    return quat_yaw


def quat_apply(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """A.1 ``quat_apply``: ``v + w*t + xyz x t`` with ``t = 2 * (xyz x v)``."""
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec + quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_apply_yaw(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """A.1 ``quat_apply_yaw = quat_apply(yaw_quat(q), v)``."""
    quat_yaw = yaw_quat(quat)
    return quat_apply(quat_yaw, vec)


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """A.1 ``quat_rotate_inverse``: ``a - b + c`` with ``a = v(2w^2-1)``, ``b = 2w (q_vec x v)``,
    ``c = 2 q_vec (q_vec . v)``."""
    shape = q.shape
    q_w = q[:, 0]
    q_vec = q[:, 1:]
    a = v * (2.0 * q_w**2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def wrap_to_pi(angles: torch.Tensor) -> torch.Tensor:
    """A.1 ``wrap_to_pi``: ``a %= 2pi; a -= 2pi * (a > pi)``."""
    angles = angles.clone()
    angles %= 2 * math.pi
    angles -= 2 * math.pi * (angles > math.pi)
    return angles


def heading_w(root_quat_w: torch.Tensor) -> torch.Tensor:
    """A.1 ``ArticulationData.heading_w``: ``atan2(f.y, f.x)``, ``f = quat_apply(q, (1,0,0))``."""
    fwd = torch.zeros(root_quat_w.shape[0], 3, dtype=root_quat_w.dtype, device=root_quat_w.device)
    fwd[:, 0] = 1.0
    f = quat_apply(root_quat_w, fwd)
    return torch.atan2(f[:, 1], f[:, 0])
