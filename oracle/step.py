"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one ``RoverEnv.step`` without physics.

Follows the reference's step ordering (``rover_envs/envs/navigation/entrypoints/rover_env.py:42-102``)
and the ORBIT manager combine rules it delegates to (third-party, absent from ``/root/reference``;
restated from SURVEY.md Appendix A.2).  PhysX is replaced by caller-supplied root-state tensors;
random draws are replaced by caller-supplied variates so that the CUDA path and this oracle consume
identical numbers:

* ``spawn_perm[j]``  -- spawn-table row of the j-th reset env (ascending env id, the order
  ``reset_buf.nonzero()`` yields; reference: ``randperm(len(spawn))[:K]``, randomizations.py:22);
* ``yaw_u[i]``       -- ``torch.rand`` variate of env ``i`` for the spawn yaw (randomizations.py:30);
* ``theta_u[i, r]``  -- variate of env ``i`` in rejection round ``r`` (terrain_importer.py:169);
* ``heading_u[i]``   -- variate for ``uniform_(-pi, pi)`` (terrain_importer.py:94-95).

The reference's unbounded ``while`` rejection loop (terrain_importer.py:143-151) is bounded to
``theta_u.shape[1]`` rounds; an env that exhausts them keeps its last candidate and is counted in
``stats["target_rounds_exhausted"]``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import orbit_math as om
from . import terms as T
from .terms import AAU_ROVER, RoverConstants


@dataclass
class TerrainTables:
    """Init-time lookup tables (terrain_utils.py:14-21, 117-127)."""

    heightmap: torch.Tensor  # [H, W] f32
    safe_mask: torch.Tensor  # [H, W] u8 (1 = rock / unsafe)
    offset_xy: torch.Tensor  # [2] f32 = (min_x, min_y)   (terrain_utils.py:21)
    spawn_table: torch.Tensor  # [2N, 3] f32            (terrain_utils.py:123-124)
    resolution: float = 0.05


@dataclass
class MdpState:
    """Everything the managers keep between steps (A.2)."""

    action: torch.Tensor
    prev_action: torch.Tensor
    pos_cmd_w: torch.Tensor
    heading_cmd_w: torch.Tensor
    pos_cmd_b: torch.Tensor
    heading_cmd_b: torch.Tensor
    time_left: torch.Tensor
    command_counter: torch.Tensor
    episode_length_buf: torch.Tensor  # int64
    episode_sums: torch.Tensor  # [N, 7] weighted * dt sums, order = terms.REWARD_TERMS
    env_origins: torch.Tensor
    err_pos: torch.Tensor
    err_heading: torch.Tensor

    @staticmethod
    def zeros(n: int) -> "MdpState":
        f = lambda *s: torch.zeros(*s, dtype=torch.float32)  # noqa: E731
        return MdpState(f(n, 2), f(n, 2), f(n, 3), f(n), f(n, 3), f(n), f(n),
                        torch.zeros(n, dtype=torch.int64), torch.zeros(n, dtype=torch.int64),
                        f(n, 7), f(n, 3), f(n), f(n))

    def clone(self) -> "MdpState":
        return MdpState(**{k: v.clone() for k, v in self.__dict__.items()})


@dataclass
class StepOutput:
    joint_pos: torch.Tensor
    joint_vel: torch.Tensor
    processed_actions: torch.Tensor
    reward: torch.Tensor
    term_rewards: torch.Tensor  # [N, 7] weight * value * dt
    terminated: torch.Tensor
    truncated: torch.Tensor
    term_flags: torch.Tensor  # [N, 4] bool, order = terms.TERMINATION_TERMS
    reset_ids: torch.Tensor
    spawn_index: torch.Tensor  # [K] int64 rows of the spawn table used
    root_pos_w: torch.Tensor  # after reset write-back
    root_quat_w: torch.Tensor
    obs_head: torch.Tensor  # [N, 4] = actions(2), distance*0.11, angle/pi
    stats: dict = field(default_factory=dict)


def sample_targets(ids, env_origins, theta_u, tables: TerrainTables, c: RoverConstants = AAU_ROVER):
    """terrain_importer.py:134-175 with bounded rounds.  Returns ``(target [K,3], exhausted [K] bool)``."""
    k = len(ids)
    target = torch.zeros(k, 3, dtype=torch.float32)
    pending = torch.ones(k, dtype=torch.bool)
    for r in range(theta_u.shape[1]):
        if not pending.any():
            break
        sel = pending.nonzero().squeeze(-1)
        theta = theta_u[ids[sel], r] * 2 * torch.pi  # :169
        target[sel, 0] = torch.cos(theta) * c.target_distance + env_origins[ids[sel], 0]  # :172
        target[sel, 1] = torch.sin(theta) * c.target_distance + env_origins[ids[sel], 1]  # :173
        bad = T.target_is_invalid(target[sel, 0:2], tables.safe_mask, tables.offset_xy, tables.resolution)
        pending[sel] = bad
    target[:, 2] = T.height_at(target[:, 0:2], tables.heightmap, tables.offset_xy, tables.resolution)  # :154
    return target, pending


def resample_command(state: MdpState, ids, theta_u, heading_u, tables, c: RoverConstants = AAU_ROVER):
    """A.2 ``CommandTerm._resample`` + terrain_importer.py:74-95 (``simple_heading=False``)."""
    state.time_left[ids] = c.resampling_time  # U(150,150)
    state.command_counter[ids] += 1
    tgt, exhausted = sample_targets(ids, state.env_origins, theta_u, tables, c)
    state.pos_cmd_w[ids] = tgt  # + default_root_state z (=0, aau_rover_simple.py init_state)
    lo, hi = c.heading_range
    state.heading_cmd_w[ids] = heading_u[ids] * (hi - lo) + lo  # Tensor.uniform_(lo, hi)
    return int(exhausted.sum())


def oracle_step(state: MdpState, actions, root_pos_w, root_quat_w, force_matrix_w, tables: TerrainTables,
                spawn_perm, yaw_u, theta_u, heading_u, c: RoverConstants = AAU_ROVER, spawn_by_env=None) -> StepOutput:
    """One env step; mutates ``state`` like the managers do.  Order = rover_env.py:61-102.
    ``spawn_by_env`` (instead of ``spawn_perm``): the spawn row of env ``i`` if it resets -- a random permutation of the
    table evaluated at the env ids; K reset envs draw K distinct uniformly random rows, the distribution of the
    reference's ``randperm(len)[:K]`` (randomizations.py:22) without the dependence on the reset order."""
    n = actions.shape[0]
    max_len = c.max_episode_length
    dt = c.step_dt
    root_pos_w = root_pos_w.clone()
    root_quat_w = root_quat_w.clone()

    # -- ActionManager.process_action (A.2) + AckermannAction2.process_actions/apply_actions
    state.prev_action[:] = state.action
    state.action[:] = actions
    processed = T.process_actions(state.action, c)
    joint_pos, joint_vel = T.ackermann2(processed[:, 0], processed[:, 1], c)  # decimation x identical

    # -- counters, terminations (rover_env.py:79-84); pos_cmd_b is the PREVIOUS step's command
    state.episode_length_buf += 1
    flags = torch.stack([
        T.term_time_out(state.episode_length_buf, max_len),
        T.term_is_success(state.pos_cmd_b, c.reached_threshold),
        T.term_far_from_target(state.pos_cmd_b, c.far_threshold),
        T.term_collision(force_matrix_w),
    ], dim=1)
    truncated = flags[:, 0]
    terminated = flags[:, 1] | flags[:, 2] | flags[:, 3]
    reset = truncated | terminated

    # -- RewardManager.compute(dt) (A.2): value * weight * dt, accumulated in declaration order
    vals = [
        (T.rew_distance_to_target(state.pos_cmd_b, max_len), c.w_distance),
        (T.rew_reached_target(state.pos_cmd_b, state.episode_length_buf, max_len, c.reached_threshold), c.w_reached),
        (T.rew_oscillation(state.action, state.prev_action, max_len), c.w_oscillation),
        (T.rew_angle_to_target(state.pos_cmd_b, max_len), c.w_angle),
        (T.rew_heading_soft_constraint(state.action, max_len), c.w_heading),
        (T.rew_collision(force_matrix_w), c.w_collision),
        (T.rew_far_from_target(state.pos_cmd_b, c.far_threshold), c.w_far),
    ]
    reward = torch.zeros(n, dtype=torch.float32)
    term_rewards = torch.zeros(n, 7, dtype=torch.float32)
    for i, (v, w) in enumerate(vals):
        contrib = v * w * dt
        reward += contrib
        state.episode_sums[:, i] += contrib
        term_rewards[:, i] = contrib

    # -- _reset_idx (rover_env.py:89-91, A.2 order)
    ids = reset.nonzero(as_tuple=False).squeeze(-1)
    k = len(ids)
    stats = {"num_resets": k, "target_rounds_exhausted": 0}
    spawn_index = spawn_by_env[ids].clone() if spawn_by_env is not None else spawn_perm[:k].clone()
    if k > 0:
        # randomization "reset": reset_root_state_rover (randomizations.py:12-39)
        pos = tables.spawn_table[spawn_index].clone()
        pos[:, 2] += c.spawn_z_offset
        angle = yaw_u[ids] * 2 * torch.pi
        quat = torch.zeros(k, 4, dtype=torch.float32)
        quat[:, 0] = torch.cos(angle / 2)
        quat[:, 3] = torch.sin(angle / 2)
        state.env_origins[ids] = pos
        root_pos_w[ids] = pos
        root_quat_w[ids] = quat
        # action manager reset
        state.action[ids] = 0.0
        state.prev_action[ids] = 0.0
        # reward manager reset: episodic sums of the reset envs (logged as mean / episode_length_s)
        stats["reward_sums"] = state.episode_sums[ids].sum(dim=0)
        state.episode_sums[ids] = 0.0
        # command term reset: log metrics, zero, counter = 0, resample
        stats["err_pos_sum"] = state.err_pos[ids].sum()
        stats["err_heading_sum"] = state.err_heading[ids].sum()
        state.err_pos[ids] = 0.0
        state.err_heading[ids] = 0.0
        state.command_counter[ids] = 0
        stats["target_rounds_exhausted"] += resample_command(state, ids, theta_u, heading_u, tables, c)
        # termination manager reset: per-term counts over the reset envs
        stats["term_counts"] = flags[ids].sum(dim=0)
        state.episode_length_buf[ids] = 0
    else:
        stats["reward_sums"] = torch.zeros(7)
        stats["err_pos_sum"] = torch.zeros(())
        stats["err_heading_sum"] = torch.zeros(())
        stats["term_counts"] = torch.zeros(4, dtype=torch.int64)

    # -- CommandManager.compute(dt) (A.2)
    state.err_pos, state.err_heading = T.update_metrics(state.pos_cmd_w, state.heading_cmd_w, root_pos_w, root_quat_w)
    state.time_left -= dt
    tids = (state.time_left <= 0.0).nonzero(as_tuple=False).squeeze(-1)
    if len(tids) > 0:
        stats["target_rounds_exhausted"] += resample_command(state, tids, theta_u, heading_u, tables, c)
    stats["num_time_resamples"] = len(tids)
    pos_b, head_b = T.update_command(state.pos_cmd_w, state.heading_cmd_w, root_pos_w, root_quat_w)
    state.pos_cmd_b[:] = pos_b
    state.heading_cmd_b[:] = head_b

    # -- ObservationManager.compute (A.2): func -> mul_(scale) -> cat; scan appended by the caller
    obs_head = torch.cat([
        state.action,
        T.obs_distance(state.pos_cmd_b) * c.obs_distance_scale,
        T.obs_angle(state.pos_cmd_b) * c.obs_heading_scale,
    ], dim=1)

    return StepOutput(joint_pos, joint_vel, processed, reward, term_rewards, terminated, truncated, flags, ids,
                      spawn_index, root_pos_w, root_quat_w, obs_head, stats)


def episode_log(stats: dict, c: RoverConstants = AAU_ROVER) -> dict:
    """``extras["log"]`` as the ORBIT managers' ``reset`` build it (A.2):
    reward sums -> mean / episode_length_s; terminations -> counts; metrics -> means."""
    k = max(int(stats["num_resets"]), 1)
    log = {}
    for i, name in enumerate(T.REWARD_TERMS):
        log[f"Episode Reward/{name}"] = float(stats["reward_sums"][i]) / k / c.episode_length_s
    for i, name in enumerate(T.TERMINATION_TERMS):
        log[f"Episode Termination/{name}"] = int(stats["term_counts"][i])
    log["Metrics/target_pose/error_pos"] = float(stats["err_pos_sum"]) / k
    log["Metrics/target_pose/error_heading"] = float(stats["err_heading_sum"]) / k
    return log


def grid_pattern(c: RoverConstants = AAU_ROVER) -> torch.Tensor:
    """ORBIT ``patterns.grid_pattern`` (A.3): ``arange(-s/2, s/2 + 1e-9, res)`` on both axes,
    ``meshgrid(indexing="xy")``, x fastest.  Returns local ray starts ``[R, 3]`` (z = 0)."""
    x = torch.arange(start=-c.scan_size[0] / 2, end=c.scan_size[0] / 2 + 1.0e-9, step=c.scan_resolution)
    y = torch.arange(start=-c.scan_size[1] / 2, end=c.scan_size[1] / 2 + 1.0e-9, step=c.scan_resolution)
    gx, gy = torch.meshgrid(x, y, indexing="xy")
    starts = torch.zeros(gx.numel(), 3, dtype=torch.float32)
    starts[:, 0] = gx.flatten()
    starts[:, 1] = gy.flatten()
    return starts


def ray_starts_world(pos_w, quat_w, c: RoverConstants = AAU_ROVER):
    """ORBIT ``RayCaster._initialize_rays_impl`` + ``_update_buffers_impl`` (A.3), ``attach_yaw_only``:
    local starts get the sensor offset (0,0,10), are rotated by the yaw-only quaternion and translated."""
    local = grid_pattern(c)
    local = local + torch.tensor([0.0, 0.0, c.scan_offset_z])
    n, r = pos_w.shape[0], local.shape[0]
    starts = local.unsqueeze(0).repeat(n, 1, 1)
    starts_w = om.quat_apply_yaw(quat_w.repeat(1, r), starts) + pos_w.unsqueeze(1)
    return starts_w


def height_scan(pos_w, quat_w, mesh, c: RoverConstants = AAU_ROVER):
    """observations.py:35-45 over ORBIT ``raycast_mesh`` (A.3).  ``mesh`` = oracle.raycast.Mesh.
    Returns ``(heights [N,R], ray_hits_w [N,R,3])``; a miss leaves the hit at +inf -> height -inf."""
    starts_w = ray_starts_world(pos_w, quat_w, c)
    n, r, _ = starts_w.shape
    dirs = torch.zeros(n * r, 3, dtype=torch.float32)
    dirs[:, 2] = -1.0
    hits = mesh.raycast(starts_w.reshape(-1, 3), dirs, c.scan_max_distance).view(n, r, 3)
    return T.obs_height_scan(pos_w, hits, c), hits
