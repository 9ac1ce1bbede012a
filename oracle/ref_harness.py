"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference functions on plain tensors.

Builds the minimal fake ``env`` the reference's manager-term functions read ("What env
must expose to terms") and calls the functions imported in place by :mod:`oracle.ref_loader`.
Works only where ``/root/reference`` exists (the build container); used to generate
``tests/golden/*.npz`` and to cross-check the restatements in :mod:`oracle.terms`.
"""
from __future__ import annotations

import types
from contextlib import contextmanager

import numpy as np
import torch

from . import orbit_math as om
from . import ref_loader
from .terms import AAU_ROVER, RoverConstants


class _Scene(dict):
    def __init__(self, sensors, terrain, **assets):
        super().__init__(**assets)
        self.sensors = sensors
        self.terrain = terrain


class _Asset:
    """Fake Articulation: records the joint targets / root pose the terms write."""

    def __init__(self, n, root_pos_w=None, root_quat_w=None):
        self.data = types.SimpleNamespace(
            root_pos_w=root_pos_w, root_quat_w=root_quat_w,
            default_root_state=torch.zeros(n, 13),
        )
        if root_quat_w is not None:
            self.data.heading_w = om.heading_w(root_quat_w)
        self.joint_vel_target = None
        self.joint_pos_target = None
        self.written_pose = None

    def find_joints(self, names):
        # drive: 6 joints, steer: 4 joints (env_cfg.py:29-30)
        k = 6 if "Drive" in names[0] else 4
        return list(range(k)), [f"j{i}" for i in range(k)]

    def set_joint_velocity_target(self, t, joint_ids=None):
        self.joint_vel_target = t

    def set_joint_position_target(self, t, joint_ids=None):
        self.joint_pos_target = t

    def write_root_pose_to_sim(self, pose, env_ids=None):
        self.written_pose = (pose.clone(), env_ids.clone())
        self.data.root_pos_w[env_ids] = pose[:, :3]
        self.data.root_quat_w[env_ids] = pose[:, 3:]


def make_env(n, *, pos_cmd_b=None, action=None, prev_action=None, episode_length_buf=None, force_matrix_w=None,
             sensor_pos_w=None, ray_hits_w=None, root_pos_w=None, root_quat_w=None, terrain=None,
             c: RoverConstants = AAU_ROVER):
    env = types.SimpleNamespace()
    env.num_envs = n
    env.device = "cpu"
    env.max_episode_length = c.max_episode_length
    env.episode_length_buf = episode_length_buf
    env.command_manager = types.SimpleNamespace(get_command=lambda name: pos_cmd_b)
    env.action_manager = types.SimpleNamespace(action=action, prev_action=prev_action)
    sensors = {
        "contact_sensor": types.SimpleNamespace(data=types.SimpleNamespace(force_matrix_w=force_matrix_w)),
        "height_scanner": types.SimpleNamespace(data=types.SimpleNamespace(pos_w=sensor_pos_w, ray_hits_w=ray_hits_w)),
    }
    env.scene = _Scene(sensors, terrain, robot=_Asset(n, root_pos_w, root_quat_w))
    return env


def _cfg(name):
    return types.SimpleNamespace(name=name)


def ref_ackermann2(actions: torch.Tensor, c: RoverConstants = AAU_ROVER):
    """``AckermannAction2`` driven through its ActionTerm API (ackermann_actions.py:186-236)."""
    mod = ref_loader.load("ackermann_actions")
    n = actions.shape[0]
    env = make_env(n)
    cfg = types.SimpleNamespace(
        asset_name="robot", scale=c.action_scale, offset=c.action_offset,
        wheelbase_length=c.wheelbase_length, middle_wheel_distance=c.middle_wheel_distance,
        rear_and_front_wheel_distance=c.rear_and_front_wheel_distance, wheel_radius=c.wheel_radius,
        min_steering_radius=c.min_steering_radius, steering_joint_names=[".*Steer_Revolute"],
        drive_joint_names=[".*Drive_Continuous"])
    term = mod.AckermannAction2(cfg, env)
    term.process_actions(actions)
    term.apply_actions()
    asset = env.scene["robot"]
    return term.processed_actions.clone(), asset.joint_pos_target.clone(), asset.joint_vel_target.clone()


def ref_ackermann_variants(actions: torch.Tensor, c: RoverConstants = AAU_ROVER):
    """``AckermannAction`` (ackermann_actions.py:19-158) through its ActionTerm API and the free function
    ``ackermann()`` of ``AckermannAction3`` (:423-505) on the processed actions."""
    mod = ref_loader.load("ackermann_actions")
    n = actions.shape[0]
    env = make_env(n)
    cfg = types.SimpleNamespace(
        asset_name="robot", scale=c.action_scale, offset=c.action_offset, wheelbase_length=c.wheelbase_length,
        middle_wheel_distance=c.middle_wheel_distance, rear_and_front_wheel_distance=c.rear_and_front_wheel_distance,
        wheel_radius=c.wheel_radius, min_steering_radius=c.min_steering_radius,
        steering_joint_names=[".*Steer_Revolute"], drive_joint_names=[".*Drive_Continuous"])
    t1 = mod.AckermannAction(cfg, env)
    t1.process_actions(actions)
    t1.apply_actions()
    asset = env.scene["robot"]
    v1 = asset.joint_pos_target.clone(), asset.joint_vel_target.clone()
    p = t1.processed_actions
    sa, wv = mod.ackermann(p[:, 0], p[:, 1], cfg, "cpu")
    return v1[0], v1[1], sa.clone(), wv.clone()


def ref_rewards_terminations(pos_cmd_b, action, prev_action, episode_length_buf, force_matrix_w,
                             c: RoverConstants = AAU_ROVER):
    """All 7 reward terms (unweighted) and the 3 first-party terminations, reference code."""
    rw = ref_loader.load("rewards")
    tm = ref_loader.load("terminations")
    n = pos_cmd_b.shape[0]
    env = make_env(n, pos_cmd_b=pos_cmd_b, action=action, prev_action=prev_action,
                   episode_length_buf=episode_length_buf, force_matrix_w=force_matrix_w)
    rewards = torch.stack([
        rw.distance_to_target_reward(env, "target_pose"),
        rw.reached_target(env, "target_pose", c.reached_threshold),
        rw.oscillation_penalty(env),
        rw.angle_to_target_penalty(env, "target_pose"),
        rw.heading_soft_contraint(env, _cfg("robot")),
        rw.collision_penalty(env, _cfg("contact_sensor"), 1.0),
        rw.far_from_target_reward(env, "target_pose", c.far_threshold),
    ], dim=1).to(torch.float32)
    terms = torch.stack([
        tm.is_success(env, "target_pose", c.reached_threshold),
        tm.far_from_target(env, "target_pose", c.far_threshold),
        tm.collision_with_obstacles(env, _cfg("contact_sensor"), 1.0),
    ], dim=1)
    return rewards, terms


def ref_observations(pos_cmd_b, sensor_pos_w, ray_hits_w):
    ob = ref_loader.load("observations")
    n = pos_cmd_b.shape[0]
    env = make_env(n, pos_cmd_b=pos_cmd_b, sensor_pos_w=sensor_pos_w, ray_hits_w=ray_hits_w)
    return (ob.distance_to_target_euclidean(env, "target_pose"), ob.angle_to_target_observation(env, "target_pose"),
            ob.height_scan_rover(env, _cfg("height_scanner")))


def ref_terrain_manager(vertices: np.ndarray, faces: np.ndarray, num_envs: int):
    """The reference ``TerrainManager`` with its USD loading bypassed (terrain_utils.py:92-127):
    the pure-Python ``mesh_to_heightmap``, ``find_rocks_in_heightmap`` and ``random_rover_spawns``
    run unmodified; the ``.cuda()``-only tensor attributes are set by hand on CPU."""
    tu = ref_loader.load("terrain_utils")
    hmm = object.__new__(tu.HeightmapManager)
    hmm.resolution_in_m = 0.05
    hmm.heightmap, hmm.min_x, hmm.min_y, hmm.max_x, hmm.max_y = hmm.mesh_to_heightmap(vertices, faces)
    hmm.heightmap_tensor = torch.from_numpy(hmm.heightmap)
    hmm.offset_tensor = torch.tensor([hmm.min_x, hmm.min_y])
    tmgr = object.__new__(tu.TerrainManager)
    tmgr.resolution_in_m = 0.05
    tmgr.gradient_threshold = 0.3
    tmgr._heightmap_manager = hmm
    tmgr.rock_mask, tmgr.safe_rock_mask = tmgr.find_rocks_in_heightmap(hmm.heightmap, tmgr.gradient_threshold)
    tmgr.spawn_locations = torch.from_numpy(tmgr.random_rover_spawns(
        rock_mask=tmgr.safe_rock_mask, heightmap=hmm.heightmap, n_spawns=num_envs * 2, seed=41))
    tmgr.rock_mask_tensor = torch.from_numpy(tmgr.safe_rock_mask).unsqueeze(-1)
    return tmgr


@contextmanager
def _patched(obj, name, fn):
    old = getattr(obj, name)
    setattr(obj, name, fn)
    try:
        yield
    finally:
        setattr(obj, name, old)


def ref_command_term(tmgr, num_envs, root_pos_w, root_quat_w, env_origins):
    """Reference ``TerrainBasedPositionCommand`` + ``RoverTerrainImporter`` wired on a fake env."""
    ti = ref_loader.load("terrain_importer")
    terrain = object.__new__(ti.RoverTerrainImporter)
    terrain.device = "cpu"
    terrain._cfg = types.SimpleNamespace(num_envs=num_envs)
    terrain._terrainManager = tmgr
    terrain.target_distance = 9.0
    terrain.env_origins = env_origins
    env = make_env(num_envs, root_pos_w=root_pos_w, root_quat_w=root_quat_w, terrain=terrain)
    cfg = types.SimpleNamespace(asset_name="robot", simple_heading=False, rel_standing_envs=0.0,
                                resampling_time_range=(150.0, 150.0),
                                ranges=types.SimpleNamespace(heading=AAU_ROVER.heading_range), debug_vis=False)
    term = ti.TerrainBasedPositionCommand(cfg, env)
    return term, terrain, env


def ref_resample_command(term, terrain, ids, theta_u, heading_u):
    """Runs the reference ``_resample_command`` (terrain_importer.py:74-95) with ``torch.rand`` /
    ``Tensor.uniform_`` replaced by table look-ups ``theta_u[env, round]`` / ``heading_u[env]`` so that
    the reference consumes exactly the variates the oracle and the CUDA kernel are fed.
    Returns the number of rejection rounds used."""
    rounds = {"r": 0, "ids": None}
    orig_generate = type(terrain).generate_random_targets

    def generate(self, env_ids, target_position):
        rounds["ids"] = env_ids
        out = orig_generate(self, env_ids, target_position)
        rounds["r"] += 1
        return out

    def fake_rand(k, device=None):
        assert k == len(rounds["ids"])
        return theta_u[rounds["ids"], rounds["r"]].clone()

    def fake_uniform_(self, lo, hi):
        self.copy_(heading_u[ids] * (hi - lo) + lo)
        return self

    with _patched(type(terrain), "generate_random_targets", generate), _patched(torch, "rand", fake_rand), \
            _patched(torch.Tensor, "uniform_", fake_uniform_):
        term._resample_command(ids)
    return rounds["r"]


def ref_reset_root_state(env, ids, spawn_perm, yaw_u):
    """Reference ``reset_root_state_rover`` (randomizations.py:12-39) with ``randperm``/``rand`` replaced
    by the supplied variates (``spawn_perm[:K]``, ``yaw_u[ids]``)."""
    rz = ref_loader.load("randomizations")

    def fake_randperm(n, device=None):
        return spawn_perm.clone()

    def fake_rand(k, device=None):
        return yaw_u[ids].clone()

    with _patched(torch, "randperm", fake_randperm), _patched(torch, "rand", fake_rand):
        rz.reset_root_state_rover(env, ids, _cfg("robot"))
    return env.scene["robot"].written_pose


def ref_value():
    """The reference ``DeterministicNeuralNetwork`` (models.py:105-162) with the ``value`` weights of ``best_agent.pt``."""
    md = ref_loader.load("models")
    space = types.SimpleNamespace(shape=(2,))
    net = md.DeterministicNeuralNetwork(observation_space=None, action_space=space, device="cpu", mlp_input_size=4,
                                        mlp_layers=[256, 160, 128], mlp_activation="leaky_relu",
                                        encoder_input_size=961, encoder_layers=[80, 60],
                                        encoder_activation="leaky_relu")
    sd = torch.load(ref_loader.POLICY_CHECKPOINT, map_location="cpu", weights_only=False)["value"]
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net


def ref_policy():
    """The reference ``GaussianNeuralNetwork`` (models.py:39-102) with ``best_agent.pt`` loaded strictly."""
    md = ref_loader.load("models")
    space = types.SimpleNamespace(shape=(2,))
    net = md.GaussianNeuralNetwork(observation_space=None, action_space=space, device="cpu", mlp_input_size=4,
                                   mlp_layers=[256, 160, 128], mlp_activation="leaky_relu", encoder_input_size=961,
                                   encoder_layers=[80, 60], encoder_activation="leaky_relu")
    sd = torch.load(ref_loader.POLICY_CHECKPOINT, map_location="cpu", weights_only=False)["policy"]
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net
