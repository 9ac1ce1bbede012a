"""``RoverEnv``: the reference's environment entry point over the fused CUDA kernels.

Mirrors ``rover_envs/envs/navigation/entrypoints/rover_env.py`` (``RoverEnv(RLTaskEnv)``: ``step`` :42-102,
``_reset_idx`` :27-39) and the slice of ORBIT's ``RLTaskEnv`` / manager objects that the reference's term functions
read (SURVEY.md 8b "What env must expose to terms"):

    env.num_envs, env.device, env.max_episode_length, env.episode_length_buf (int64),
    env.command_manager.get_command(name), env.action_manager.action / prev_action,
    env.scene.sensors[name].data.{pos_w, ray_hits_w, force_matrix_w},
    env.scene[asset].data.{root_pos_w, root_quat_w, heading_w, default_root_state}, env.scene.terrain

The step keeps the reference's order (process actions -> physics x decimation -> counters -> terminations ->
rewards -> reset -> command update -> observations) but runs it as three launches:
``rover_mdp_pre_step`` (ACTIONS | TERMS), ``rover_mdp_post_step`` and ``rover_height_scan``.
PhysX is out of scope: ``physics`` is any callable ``(env) -> None`` that updates ``robot.data.root_pos_w /
root_quat_w`` and ``contact_sensor.data.force_matrix_w`` in place (default: synthetic state, see synthetic.py).
"""
from __future__ import annotations

import types

import torch

from . import _lib, ops
from .config import REWARD_TERMS, TERMINATION_TERMS, RoverEnvCfg
from .dist import EpisodeStats, episode_log
from .terrain import TerrainTables


class _Data(types.SimpleNamespace):
    pass


class RobotArticulation:
    """The slice of ORBIT ``Articulation`` the terms touch (``scene["robot"]``)."""

    def __init__(self, n, device):
        self.data = _Data(
            root_pos_w=torch.zeros(n, 3, device=device), root_quat_w=torch.zeros(n, 4, device=device),
            default_root_state=torch.zeros(n, 13, device=device))
        self.data.root_quat_w[:, 0] = 1.0
        self.joint_pos_target = None
        self.joint_vel_target = None
        self._drive_ids, self._steer_ids = list(range(6)), list(range(6, 10))

    def find_joints(self, names):
        if any("Drive" in n for n in names):
            return self._drive_ids, [f"{p}_Drive_Continuous" for p in ("ML", "FL", "RL", "RR", "MR", "FR")]
        return self._steer_ids, [f"{p}_Steer_Revolute" for p in ("FL", "RL", "RR", "FR")]

    def set_joint_velocity_target(self, target, joint_ids=None):
        self.joint_vel_target = target

    def set_joint_position_target(self, target, joint_ids=None):
        self.joint_pos_target = target

    def write_root_pose_to_sim(self, pose, env_ids=None):
        self.data.root_pos_w[env_ids] = pose[:, :3]
        self.data.root_quat_w[env_ids] = pose[:, 3:]


class RayCaster:
    """ORBIT ``RayCaster`` facade: ``data.pos_w`` is the body pose, ``data.ray_hits_w`` is produced on demand by the
    CUDA raycaster (the fused path never materialises it: 12 B/ray of traffic the reference pays)."""

    def __init__(self, env, cfg):
        self.cfg = cfg
        self._env = env
        self.ray_pattern = ops.RayPattern.grid(env.device, cfg.pattern_cfg.resolution, cfg.pattern_cfg.size, cfg.offset_pos)
        self.ray_starts = self.ray_pattern.starts
        outer = self

        class _SensorData:
            @property
            def pos_w(self):
                return outer._env.scene["robot"].data.root_pos_w

            @property
            def quat_w(self):
                return outer._env.scene["robot"].data.root_quat_w

            @property
            def ray_hits_w(self):
                robot = outer._env.scene["robot"].data
                _, hits = ops.height_scan(robot.root_pos_w, robot.root_quat_w, outer.ray_pattern, outer._env.scan_grid,
                                          outer.cfg.max_distance, outer._env.cfg.height_scan_base_offset,
                                          return_hits=True)
                return hits

        self.data = _SensorData()


class ContactSensor:
    def __init__(self, n, num_bodies, device):
        self.data = _Data(force_matrix_w=torch.zeros(n, num_bodies, 1, 3, device=device))


class RoverTerrainImporter:
    """``RoverTerrainImporter`` (utils/terrains/terrain_importer.py:127-184) over the device tables."""

    def __init__(self, env, tables: TerrainTables):
        self._env = env
        self.tables = tables
        self.handle = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                              tables.resolution, env.device)
        self.target_distance = env.cfg.commands.target_distance

    @property
    def env_origins(self):
        return self._env._buf.env_origins

    def get_spawn_locations(self):
        """terrain_importer.py:177-184."""
        return self.handle.spawn_table

    def sample_new_targets(self, env_ids):
        """terrain_importer.py:134-157: rejection-sampled targets 9 m around the env origins -> ``[K,3]``."""
        cmd = self._env.command_manager.get_term("target_pose")
        saved = cmd.pos_command_w[env_ids].clone(), cmd.heading_command_w[env_ids].clone(), \
            cmd.time_left[env_ids].clone(), cmd.command_counter[env_ids].clone()
        self._env._run_post(env_ids, _lib.PHASE_RESAMPLE)
        out = cmd.pos_command_w[env_ids].clone()
        cmd.pos_command_w[env_ids], cmd.heading_command_w[env_ids] = saved[0], saved[1]
        cmd.time_left[env_ids], cmd.command_counter[env_ids] = saved[2], saved[3]
        return out


class Scene:
    def __init__(self, assets: dict, sensors: dict, terrain):
        self._assets = assets
        self.sensors = sensors
        self.terrain = terrain

    def __getitem__(self, name):
        return self._assets[name]


class ActionManager:
    """ORBIT ``ActionManager`` with one term (A.2): ``process_action`` shifts prev <- action <- new."""

    def __init__(self, env, term):
        self._env = env
        self._term = term

    @property
    def action(self):
        return self._env._buf.action

    @property
    def prev_action(self):
        return self._env._buf.prev_action

    def get_term(self, name="actions"):
        return self._term

    def process_action(self, action):
        self._term.process_actions(action)

    def apply_action(self):
        self._term.apply_actions()


class CommandManager:
    def __init__(self, env, terms: dict):
        self._env = env
        self._terms = terms

    def get_command(self, name):
        return self._terms[name].command

    def get_term(self, name):
        return self._terms[name]

    def compute(self, dt):
        for t in self._terms.values():
            t.compute(dt)


class RoverEnv:
    """Drop-in for ``RoverEnv.step`` on synthetic physics (see the module docstring)."""

    def __init__(self, cfg: RoverEnvCfg, terrain: TerrainTables, device="cuda:0", physics=None, seed: int = 0):
        from .mdp.actions import AckermannAction2
        from .mdp.commands import TerrainBasedPositionCommand

        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RoverEnv needs a CUDA device; there is no CPU fallback")
        self.num_envs = n = cfg.num_envs
        self.step_dt = cfg.step_dt
        self.max_episode_length = cfg.max_episode_length
        self.max_episode_length_s = cfg.episode_length_s
        self.common_step_counter = 0
        self._buf = ops.MdpBuffers.allocate(n, self.device)
        self._params = ops.mdp_params(cfg)
        self.episode_length_buf = self._buf.episode_length_buf
        self.scan_grid = ops.ScanGridHandle.from_mesh(terrain.vertices, terrain.faces, self.device)
        robot = RobotArticulation(n, self.device)
        self.scene = Scene({"robot": robot}, {}, None)
        self.scene.terrain = RoverTerrainImporter(self, terrain)
        self.scene.sensors["height_scanner"] = RayCaster(self, cfg.height_scanner)
        self.scene.sensors["contact_sensor"] = ContactSensor(n, cfg.num_contact_bodies, self.device)
        self.action_manager = ActionManager(self, AckermannAction2(cfg.actions, self))
        self.command_manager = CommandManager(self, {"target_pose": TerrainBasedPositionCommand(cfg.commands, self)})
        self.obs_buf = torch.zeros(n, 4 + self.scene.sensors["height_scanner"].ray_pattern.n_rays, device=self.device)
        self.reward_buf = self._buf.reward
        self.reset_buf = self._buf.reset_flags
        self.reset_terminated = self._buf.terminated
        self.reset_time_outs = self._buf.truncated
        self.extras = {}
        self._stats = EpisodeStats(self._buf)
        self._gen = torch.Generator(device=self.device).manual_seed(seed)
        self._physics = physics
        self._variates = None
        # rover_env.py:18-25: env origins are shifted by +100 m in x and y
        self._buf.env_origins[:, 0] += 100.0
        self._buf.env_origins[:, 1] += 100.0
        self._buf.time_left.fill_(cfg.commands.resampling_time_range[0])

    # ---------------------------------------------------------------------------------------------- helpers
    def _draw_variates(self):
        n, g = self.num_envs, self._gen
        return (torch.randperm(2 * n, device=self.device, generator=g),
                torch.rand(n, device=self.device, generator=g), torch.rand(n, device=self.device, generator=g),
                torch.rand(n, self.cfg.target_rounds, device=self.device, generator=g))

    def set_variates(self, spawn_perm, yaw_u, heading_u, theta_u):
        """Parity hook: feed the random variates of the next post-step explicitly (oracle and kernel share them)."""
        self._variates = (spawn_perm, yaw_u, heading_u, theta_u)

    def _run_post(self, env_ids, phases, obs=None):
        """Single-purpose use of the post-step kernel for the envs in ``env_ids`` (the reference's per-term calls)."""
        b = self._buf
        b.reset_flags.zero_()
        b.reset_flags[env_ids] = 1
        blocks = b.block_reset_counts.numel()
        padded = torch.zeros(blocks * _lib.MDP_BLOCK, dtype=torch.int32, device=self.device)
        padded[: self.num_envs] = b.reset_flags
        b.block_reset_counts.copy_(padded.view(blocks, _lib.MDP_BLOCK).sum(dim=1))
        v = self._variates or self._draw_variates()
        self._variates = None
        robot = self.scene["robot"].data
        ops.mdp_post_step(b, self._params, self.scene.terrain.handle, robot.root_pos_w, robot.root_quat_w, *v, obs=obs,
                          phases=phases)

    # ---------------------------------------------------------------------------------------------- API
    def reset(self):
        """Reset every env (spawn, targets), then compute the first observation."""
        ids = torch.arange(self.num_envs, device=self.device)
        self._reset_idx(ids)
        self._run_post(ids[:0], _lib.PHASE_METRICS | _lib.PHASE_COMMAND | _lib.PHASE_OBS, obs=self.obs_buf)
        self._scan()
        # a private copy: ``step`` refreshes ``obs_buf`` in place, and the reference's trainer keeps the tensor it got
        # from ``reset`` as its ``states`` buffer (``states.copy_(next_states)``, skrl_utils.py:148)
        return self.obs_buf.clone(), self.extras

    def _reset_idx(self, idx):
        """rover_env.py:27-39 / ORBIT ``RLTaskEnv._reset_idx`` (A.2) for explicit ids."""
        self._buf.stats.zero_()
        self._run_post(idx, _lib.PHASE_SPAWN | _lib.PHASE_MANAGERS | _lib.PHASE_RESAMPLE)
        self.extras["log"] = self._log_tensors()
        self.extras["episode"] = self.extras["log"]

    def _scan(self):
        robot = self.scene["robot"].data
        sensor = self.scene.sensors["height_scanner"]
        ops.height_scan(robot.root_pos_w, robot.root_quat_w, sensor.ray_pattern, self.scan_grid,
                        sensor.cfg.max_distance, self.cfg.height_scan_base_offset, out=self.obs_buf[:, 4:])

    def _log_tensors(self):
        """``extras["log"]`` as 0-dim device tensors (no host sync), ORBIT manager ``reset`` semantics (A.2)."""
        s = self._buf.stats
        k = torch.clamp(s[13], min=1.0)
        log = {f"Episode Reward/{name}": s[i] / k / self.max_episode_length_s for i, name in enumerate(REWARD_TERMS)}
        log.update({f"Episode Termination/{name}": s[7 + i] for i, name in enumerate(TERMINATION_TERMS)})
        log["Metrics/target_pose/error_pos"] = s[11] / k
        log["Metrics/target_pose/error_heading"] = s[12] / k
        return log

    def episode_log(self) -> dict:
        """Host-side copy of the last episode statistics (synchronises)."""
        return episode_log(self._buf.stats, self.max_episode_length_s)

    def step(self, action: torch.Tensor):
        """rover_env.py:42-102."""
        b, robot = self._buf, self.scene["robot"].data
        contact = self.scene.sensors["contact_sensor"].data
        # -- process actions; physics stepping (decimation x: identical joint targets, ackermann_actions.py:231-236)
        self.action_manager.process_action(action)
        self.action_manager.apply_action()
        if self._physics is not None:
            self._physics(self)
        # -- counters, terminations, rewards (one launch; reads the PREVIOUS command like the reference)
        ops.mdp_pre_step(b, self._params, None, contact.force_matrix_w, phases=_lib.PRE_TERMS)
        self.common_step_counter += 1
        # -- reset, command update, observation head (one launch)
        b.stats.zero_()
        v = self._variates or self._draw_variates()
        self._variates = None
        ops.mdp_post_step(b, self._params, self.scene.terrain.handle, robot.root_pos_w, robot.root_quat_w, *v,
                          obs=self.obs_buf)
        # -- height scan straight into the observation buffer (one launch)
        self._scan()
        self.extras["log"] = self._log_tensors()
        self.extras["episode"] = self.extras["log"]
        return self.obs_buf, self.reward_buf, self.reset_terminated.bool(), self.reset_time_outs.bool(), self.extras
