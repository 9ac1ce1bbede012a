"""Reward terms -- rover_envs/envs/navigation/mdp/rewards.py:14-137: same names and arguments, unweighted values
``[N]``.  Each returns a column of ``term_values`` computed by the fused pre-step kernel for the current step (the
RewardManager multiplies by ``weight * dt``; the kernel also keeps that product in ``term_rewards`` / ``reward``).
The arguments are honoured by checking them against what the kernel ran with (``_checks.py``): another threshold or
command, or a call before the terms of the current state exist, raises instead of returning a stale column."""
from __future__ import annotations

import torch

from ._checks import require_command, require_current, require_threshold

__all__ = ["distance_to_target_reward", "reached_target", "oscillation_penalty", "angle_to_target_penalty",
           "heading_soft_contraint", "collision_penalty", "far_from_target_reward"]


def _col(env, k: int, what: str) -> torch.Tensor:
    require_current(env, what)
    return env._buf.term_values[:, k]


def distance_to_target_reward(env, command_name: str) -> torch.Tensor:
    require_command(env, command_name, "distance_to_target_reward")
    return _col(env, 0, "distance_to_target_reward")  # rewards.py:14-32


def reached_target(env, command_name: str, threshold: float) -> torch.Tensor:
    require_command(env, command_name, "reached_target")
    require_threshold(threshold, env.cfg.rewards.reached_threshold, "reached_target", "reached_threshold")
    return _col(env, 1, "reached_target")  # rewards.py:35-53


def oscillation_penalty(env) -> torch.Tensor:
    return _col(env, 2, "oscillation_penalty")  # rewards.py:56-78


def angle_to_target_penalty(env, command_name: str) -> torch.Tensor:
    require_command(env, command_name, "angle_to_target_penalty")
    return _col(env, 3, "angle_to_target_penalty")  # rewards.py:81-96


def heading_soft_contraint(env, asset_cfg) -> torch.Tensor:
    env.scene[asset_cfg.name]  # KeyError for an unknown asset, like the reference's env.scene[asset_cfg.name]
    return _col(env, 4, "heading_soft_contraint")  # rewards.py:99-106 (sic: the reference's spelling)


def collision_penalty(env, sensor_cfg, threshold: float) -> torch.Tensor:
    env.scene.sensors[sensor_cfg.name]  # KeyError for an unknown sensor
    # ``threshold`` is accepted and ignored: the reference hard-codes ``> 1`` (rewards.py:109-124), and so does the kernel
    return _col(env, 5, "collision_penalty")


def far_from_target_reward(env, command_name: str, threshold: float) -> torch.Tensor:
    require_command(env, command_name, "far_from_target_reward")
    require_threshold(threshold, env.cfg.rewards.far_threshold, "far_from_target_reward", "far_threshold")
    return _col(env, 6, "far_from_target_reward")  # rewards.py:127-137
