"""Reward terms -- rover_envs/envs/navigation/mdp/rewards.py:14-137: same names and arguments, unweighted values
``[N]``.  Each returns a column of ``term_values`` computed by the fused pre-step kernel for the current step (the
RewardManager multiplies by ``weight * dt``; the kernel also keeps that product in ``term_rewards`` / ``reward``)."""
from __future__ import annotations

import torch

__all__ = ["distance_to_target_reward", "reached_target", "oscillation_penalty", "angle_to_target_penalty",
           "heading_soft_contraint", "collision_penalty", "far_from_target_reward"]


def _col(env, k: int) -> torch.Tensor:
    return env._buf.term_values[:, k]


def distance_to_target_reward(env, command_name: str) -> torch.Tensor:
    return _col(env, 0)  # rewards.py:14-32


def reached_target(env, command_name: str, threshold: float) -> torch.Tensor:
    return _col(env, 1)  # rewards.py:35-53 (threshold from cfg.rewards.reached_threshold)


def oscillation_penalty(env) -> torch.Tensor:
    return _col(env, 2)  # rewards.py:56-78


def angle_to_target_penalty(env, command_name: str) -> torch.Tensor:
    return _col(env, 3)  # rewards.py:81-96


def heading_soft_contraint(env, asset_cfg) -> torch.Tensor:
    return _col(env, 4)  # rewards.py:99-106 (sic: the reference's spelling)


def collision_penalty(env, sensor_cfg, threshold: float) -> torch.Tensor:
    return _col(env, 5)  # rewards.py:109-124 (``threshold`` is ignored by the reference too)


def far_from_target_reward(env, command_name: str, threshold: float) -> torch.Tensor:
    return _col(env, 6)  # rewards.py:127-137
