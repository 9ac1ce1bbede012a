"""``TerrainBasedPositionCommand`` (rover_envs/envs/navigation/utils/terrains/terrain_importer.py:19-124) over the
phase-selectable post-step kernel: same buffers (views of the fused state), same private methods."""
from __future__ import annotations

import torch

from .. import _lib


class TerrainBasedPositionCommand:
    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        self.robot = env.scene["robot"]
        self.terrain = env.scene.terrain
        b = env._buf
        # (x, y, z, heading) command buffers -- terrain_importer.py:46-49
        self.pos_command_w = b.pos_cmd_w
        self.heading_command_w = b.heading_cmd_w
        self.pos_command_b = b.pos_cmd_b
        self.heading_command_b = b.heading_cmd_b
        self.time_left = b.time_left
        self.command_counter = b.command_counter
        self.metrics = {"error_pos": b.err_pos, "error_heading": b.err_heading}  # :51-52

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device

    @property
    def command(self) -> torch.Tensor:
        """The desired base position in the base frame, ``[N,3]`` (terrain_importer.py:65-68)."""
        return self.pos_command_b

    def _no_ids(self):
        return torch.zeros(0, dtype=torch.int64, device=self.device)

    def _resample_command(self, env_ids):
        """terrain_importer.py:74-95 (``simple_heading=False``).  Unlike ORBIT's ``_resample`` wrapper this touches
        neither ``time_left`` nor ``command_counter``."""
        saved = self.time_left[env_ids].clone(), self.command_counter[env_ids].clone()
        self._env._run_post(env_ids, _lib.PHASE_RESAMPLE)
        self.time_left[env_ids], self.command_counter[env_ids] = saved

    def _resample(self, env_ids):
        """ORBIT ``CommandTerm._resample`` (A.2): time_left, counter += 1, ``_resample_command``."""
        self._env._run_post(env_ids, _lib.PHASE_RESAMPLE)

    def _update_command(self):
        """terrain_importer.py:97-101."""
        self._env._run_post(self._no_ids(), _lib.PHASE_COMMAND)

    def _update_metrics(self):
        """terrain_importer.py:103-106."""
        self._env._run_post(self._no_ids(), _lib.PHASE_METRICS)

    def compute(self, dt: float):
        """ORBIT ``CommandTerm.compute`` (A.2): metrics, time_left -= dt, time-based resample, update command."""
        self._env._run_post(self._no_ids(), _lib.PHASE_METRICS | _lib.PHASE_TIME | _lib.PHASE_COMMAND)
