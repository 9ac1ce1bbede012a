"""Termination terms -- rover_envs/envs/navigation/mdp/terminations.py:14-64 (+ ORBIT ``mdp.time_out``): bool ``[N]``
columns of the flags the fused pre-step kernel computed for the current step.  Arguments are checked against what the
kernel ran with (``_checks.py``); a differing threshold / unknown command / stale step raises."""
from __future__ import annotations

import torch

from ._checks import require_command, require_current, require_threshold

__all__ = ["time_out", "is_success", "far_from_target", "collision_with_obstacles"]


def _flag(env, k: int, what: str) -> torch.Tensor:
    require_current(env, what)
    return env._buf.term_flags.view(torch.bool)[:, k]  # 0 / 1 bytes: a view, not a conversion


def time_out(env) -> torch.Tensor:
    return _flag(env, 0, "time_out")


def is_success(env, command_name: str, threshold: float) -> torch.Tensor:
    require_command(env, command_name, "is_success")
    require_threshold(threshold, env.cfg.rewards.reached_threshold, "is_success", "reached_threshold")
    return _flag(env, 1, "is_success")  # terminations.py:14-29


def far_from_target(env, command_name: str, threshold: float) -> torch.Tensor:
    require_command(env, command_name, "far_from_target")
    require_threshold(threshold, env.cfg.rewards.far_threshold, "far_from_target", "far_threshold")
    return _flag(env, 2, "far_from_target")  # terminations.py:32-47


def collision_with_obstacles(env, sensor_cfg, threshold: float) -> torch.Tensor:
    env.scene.sensors[sensor_cfg.name]
    # ``threshold`` is ignored by the reference as well (terminations.py:50-64 hard-codes ``> 1``)
    return _flag(env, 3, "collision_with_obstacles")
