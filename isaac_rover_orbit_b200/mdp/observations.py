"""Observation terms -- rover_envs/envs/navigation/mdp/observations.py:15-45 (+ ORBIT ``mdp.last_action``).
Unscaled like the reference's functions (the ObservationManager applies ``scale``); the fused kernel writes the
scaled, concatenated observation ``[actions(2), distance*0.11, angle/pi, height_scan(961)]`` into ``env.obs_buf``."""
from __future__ import annotations

import torch

from ._checks import require_command

__all__ = ["last_action", "angle_to_target_observation", "distance_to_target_euclidean", "height_scan_rover"]


def last_action(env) -> torch.Tensor:
    return env.action_manager.action


def angle_to_target_observation(env, command_name: str) -> torch.Tensor:
    """observations.py:15-24 -> ``[N,1]``; column 3 of the fused observation divided by its scale."""
    require_command(env, command_name, "angle_to_target_observation")
    return env.obs_buf[:, 3:4] / env.cfg.obs_heading_scale


def distance_to_target_euclidean(env, command_name: str) -> torch.Tensor:
    """observations.py:27-32 -> ``[N,1]``."""
    require_command(env, command_name, "distance_to_target_euclidean")
    return env.obs_buf[:, 2:3] / env.cfg.obs_distance_scale


def height_scan_rover(env, sensor_cfg) -> torch.Tensor:
    """observations.py:35-45 -> ``[N,R]``: ``sensor.pos_w.z - ray_hits_w.z - 0.26878`` (-inf on a miss); a view of
    the scan columns the CUDA raycaster wrote for the current step (scale 1, rover_env_cfg.py:113-117)."""
    env.scene.sensors[sensor_cfg.name]  # KeyError for an unknown sensor, like the reference
    return env.obs_buf[:, 4:]
