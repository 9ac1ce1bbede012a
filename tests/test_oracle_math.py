"""Invariants that pin oracle/orbit_math.py (restated ORBIT helpers; the reference ships no test for them): every helper
against float64 rotation-matrix algebra derived independently of the quaternion formulas."""
import math

import numpy as np
import torch

from oracle import orbit_math as M


def _rand_quat(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=1, keepdim=True)


def _rot_matrix64(q):
    """(w, x, y, z) -> 3x3 rotation matrices, float64, textbook formula."""
    w, x, y, z = q.double().unbind(1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], 1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], 1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1)], 1)


def _yaw64(q):
    r = _rot_matrix64(q)
    return torch.atan2(r[:, 1, 0], r[:, 0, 0])  # heading of the rotated x axis


def test_quat_apply_is_the_rotation_matrix():
    q, v = _rand_quat(500, 1), torch.randn(500, 3, generator=torch.Generator().manual_seed(2)) * 5
    want = torch.bmm(_rot_matrix64(q), v.double().unsqueeze(-1)).squeeze(-1)
    assert (M.quat_apply(q, v).double() - want).abs().max() < 5e-6
    assert (M.quat_apply(q, v).norm(dim=1) - v.norm(dim=1)).abs().max() < 5e-6  # rotations preserve norms


def test_yaw_quat_keeps_the_heading_and_drops_roll_pitch():
    q = _rand_quat(500, 3)
    qy = M.yaw_quat(q)
    assert (qy.norm(dim=1) - 1).abs().max() < 1e-6 and bool((qy[:, 1:3] == 0).all())
    yaw = _yaw64(q)
    err = (2 * torch.atan2(qy[:, 3].double(), qy[:, 0].double()) - yaw + math.pi) % (2 * math.pi) - math.pi
    assert err.abs().max() < 1e-5
    assert (M.heading_w(q).double() - yaw).abs().max() < 1e-5  # ArticulationData.heading_w is the same angle


def test_quat_apply_yaw_is_a_2x2_rotation_about_z():
    q, v = _rand_quat(500, 4), torch.randn(500, 3, generator=torch.Generator().manual_seed(5)) * 3
    yaw = _yaw64(q)
    c, s = torch.cos(yaw), torch.sin(yaw)
    vd = v.double()
    want = torch.stack([c * vd[:, 0] - s * vd[:, 1], s * vd[:, 0] + c * vd[:, 1], vd[:, 2]], 1)
    got = M.quat_apply_yaw(q, v)
    assert ((got.double() - want).abs().max(dim=1).values / vd.norm(dim=1)).max() < 5e-6  # fp32 atan2 / sincos chain
    assert torch.equal(got[:, 2], v[:, 2])  # z passes through untouched (what makes the height scan's rays vertical)


def test_quat_rotate_inverse_undoes_quat_apply():
    q, v = _rand_quat(500, 6), torch.randn(500, 3, generator=torch.Generator().manual_seed(7)) * 4
    want = torch.bmm(_rot_matrix64(q).transpose(1, 2), v.double().unsqueeze(-1)).squeeze(-1)
    assert (M.quat_rotate_inverse(q, v).double() - want).abs().max() < 5e-6
    assert (M.quat_rotate_inverse(q, M.quat_apply(q, v)) - v).abs().max() < 1e-5


def test_wrap_to_pi_range_and_congruence():
    a = torch.linspace(-25.0, 25.0, 4001)
    w = M.wrap_to_pi(a)
    assert float(w.max()) <= math.pi + 1e-6 and float(w.min()) > -math.pi - 1e-6
    k = ((a.double() - w.double()) / (2 * math.pi)).round()
    assert ((a.double() - w.double()) - k * 2 * math.pi).abs().max() < 1e-5  # differs by a whole number of turns
    assert torch.equal(M.wrap_to_pi(torch.tensor([0.0, 1.0, -1.0])), torch.tensor([0.0, 1.0, -1.0]))
    assert abs(float(M.wrap_to_pi(torch.tensor([math.pi + 0.1]))[0]) - (-math.pi + 0.1)) < 1e-6


def test_normalize_and_its_eps_guard():
    x = torch.tensor([[3.0, 4.0, 0.0], [0.0, 0.0, 0.0], [1e-12, 0.0, 0.0]])
    n = M.normalize(x)
    assert torch.allclose(n[0], torch.tensor([0.6, 0.8, 0.0])) and bool((n[1] == 0).all())
    assert np.isclose(float(n[2, 0]), 1e-3)  # ||x|| clamped to eps = 1e-9, as ORBIT does
