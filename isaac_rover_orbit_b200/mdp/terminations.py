"""Termination terms -- rover_envs/envs/navigation/mdp/terminations.py:14-64 (+ ORBIT ``mdp.time_out``): bool ``[N]``
columns of the flags the fused pre-step kernel computed for the current step."""
from __future__ import annotations

import torch

__all__ = ["time_out", "is_success", "far_from_target", "collision_with_obstacles"]


def _flag(env, k: int) -> torch.Tensor:
    return env._buf.term_flags[:, k].bool()


def time_out(env) -> torch.Tensor:
    return _flag(env, 0)


def is_success(env, command_name: str, threshold: float) -> torch.Tensor:
    return _flag(env, 1)  # terminations.py:14-29


def far_from_target(env, command_name: str, threshold: float) -> torch.Tensor:
    return _flag(env, 2)  # terminations.py:32-47


def collision_with_obstacles(env, sensor_cfg, threshold: float) -> torch.Tensor:
    return _flag(env, 3)  # terminations.py:50-64
