"""Seeded synthetic rover state that stands in for the PhysX articulation / contact views.

The closed PhysX solve is out of scope (BASELINE.json north_star); what the MDP terms read from it --
``robot.data.root_pos_w / root_quat_w`` and ``contact_sensor.data.force_matrix_w`` -- is generated here
with the distributions of SURVEY.md section 8(d).  Generated on the CPU with a seeded ``torch.Generator``
so that the oracle and the CUDA path can be fed byte-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class SyntheticStep:
    actions: torch.Tensor  # [N,2]  U(-1,1)
    root_pos_w: torch.Tensor  # [N,3]
    root_quat_w: torch.Tensor  # [N,4] (w,x,y,z): yaw U(-pi,pi) o pitch,roll N(0,0.1)
    force_matrix_w: torch.Tensor  # [N,B,1,3]: 0 w.p. 0.95 else N(0,2)
    spawn_perm: torch.Tensor  # [2N] int64  randperm(2N)            (randomizations.py:22)
    yaw_u: torch.Tensor  # [N]   U[0,1)                            (randomizations.py:30)
    heading_u: torch.Tensor  # [N]                                 (terrain_importer.py:94-95)
    theta_u: torch.Tensor  # [N,R]                                 (terrain_importer.py:169)

    def to(self, device, non_blocking: bool = False) -> "SyntheticStep":
        return SyntheticStep(**{k: v.to(device, non_blocking=non_blocking) for k, v in self.__dict__.items()})

    def pin(self) -> "SyntheticStep":
        return SyntheticStep(**{k: v.pin_memory() for k, v in self.__dict__.items()})


def quat_from_euler(roll, pitch, yaw):
    """(w,x,y,z) of yaw (Z) o pitch (Y) o roll (X)."""
    cr, sr = torch.cos(roll / 2), torch.sin(roll / 2)
    cp, sp = torch.cos(pitch / 2), torch.sin(pitch / 2)
    cy, sy = torch.cos(yaw / 2), torch.sin(yaw / 2)
    return torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy], dim=1)


def terrain_height_lookup(xy: torch.Tensor, vertices: torch.Tensor, size_m: float, grid_res: float) -> torch.Tensor:
    """Nearest-vertex height of the regular synthetic terrain (only used to place rovers near the ground)."""
    n = int(round(size_m / grid_res)) + 1
    i = torch.clamp((xy[:, 0] / grid_res).round().long(), 0, n - 1)
    j = torch.clamp((xy[:, 1] / grid_res).round().long(), 0, n - 1)
    return vertices[j * n + i, 2]


def make_poses(n: int, gen: torch.Generator, vertices: torch.Tensor, size_m: float, grid_res: float,
               margin: float = 20.0):
    """Root xy ~ U(margin, size-margin)^2, z = terrain + U(0.2, 0.4), yaw U(-pi,pi), roll/pitch N(0, 0.1)."""
    lo, hi = margin, size_m - margin
    xy = torch.rand(n, 2, generator=gen) * (hi - lo) + lo
    z = terrain_height_lookup(xy, vertices, size_m, grid_res) + torch.rand(n, generator=gen) * 0.2 + 0.2
    yaw = (torch.rand(n, generator=gen) * 2 - 1) * math.pi
    rp = torch.randn(n, 2, generator=gen) * 0.1
    return torch.cat([xy, z[:, None]], dim=1).contiguous(), quat_from_euler(rp[:, 0], rp[:, 1], yaw).contiguous()


def make_step(n: int, gen: torch.Generator, vertices: torch.Tensor, size_m: float, grid_res: float,
              num_bodies: int = 14, rounds: int = 16, margin: float = 20.0) -> SyntheticStep:
    pos, quat = make_poses(n, gen, vertices, size_m, grid_res, margin)
    actions = torch.rand(n, 2, generator=gen) * 2 - 1
    active = (torch.rand(n, 1, 1, 1, generator=gen) >= 0.95).float()
    force = active * torch.randn(n, num_bodies, 1, 3, generator=gen) * 2.0
    return SyntheticStep(actions, pos, quat, force.contiguous(), torch.randperm(2 * n, generator=gen),
                         torch.rand(n, generator=gen), torch.rand(n, generator=gen),
                         torch.rand(n, rounds, generator=gen))


def init_commands(n: int, gen: torch.Generator, root_pos_w: torch.Tensor):
    """Initial target = root + polar(r ~ U(0,12), theta ~ U) so that both d < 0.18 and d > 11 fire."""
    r = torch.rand(n, generator=gen) * 12.0
    r[: max(n // 50, 1)] *= 0.01  # a few inside the success radius
    th = torch.rand(n, generator=gen) * 2 * math.pi
    pos_cmd_w = root_pos_w + torch.stack([r * torch.cos(th), r * torch.sin(th), torch.zeros(n)], dim=1)
    heading_cmd_w = (torch.rand(n, generator=gen) * 2 - 1) * math.pi
    ep_len = torch.randint(0, 751, (n,), generator=gen)
    return pos_cmd_w.contiguous(), heading_cmd_w, ep_len
