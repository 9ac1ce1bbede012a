"""Reset randomisation -- rover_envs/envs/navigation/mdp/randomizations.py:12-39."""
from __future__ import annotations

import torch

from .. import _lib

__all__ = ["reset_root_state_rover"]


def reset_root_state_rover(env, env_ids: torch.Tensor, asset_cfg=None, z_offset: float = 0.5) -> None:
    """Spawn the envs in ``env_ids`` at ``spawn_table[randperm(2N)[:K]] + (0,0,z_offset)`` with a uniform yaw; updates
    ``env.scene.terrain.env_origins[env_ids]`` and the root pose (side effects only, like the reference)."""
    if abs(z_offset - env.cfg.spawn_z_offset) > 1e-12:
        env._params.spawn_z_offset = z_offset
    env._run_post(env_ids, _lib.PHASE_SPAWN)
    env._params.spawn_z_offset = env.cfg.spawn_z_offset
