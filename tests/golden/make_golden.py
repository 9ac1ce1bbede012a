"""Generates the golden fixtures in this directory from the UNMODIFIED reference code.

Run in the build container (needs ``/root/reference``):

    python tests/golden/make_golden.py

Every array named ``ref_*`` is an output of the reference's own functions, imported in place by
``oracle/ref_loader.py`` (stubs only for the third-party packages that are not installed).
Arrays named ``in_*`` are the seeded inputs.  The fixtures let the oracle (and, through it, the CUDA
path) be checked on machines where the reference tree does not exist (the GPU box).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
from oracle import ref_loader  # noqa: E402

SMALL_TERRAIN = dict(size_m=48.0, grid_res=0.2, seed=3)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def special_actions():
    # both-zero, exactly-cancelled offset, point turns, reverse, saturations (SURVEY.md 8c spot values)
    return torch.tensor([[0.0, 0.0], [0.0135, 0.0135], [0.0135, 0.5], [0.5, 0.0135], [-0.5, 0.3], [0.3, -0.9],
                         [1.0, 1.0], [-1.0, -1.0], [1.0, 0.0], [0.02, 1.0], [0.0135, -1.0], [1.0, -0.0135]],
                        dtype=torch.float32)


def gen_terms(path):
    g = torch.Generator().manual_seed(1234)
    n = 2048
    act = torch.rand(n, 2, generator=g) * 2 - 1
    sp = special_actions()
    act[: len(sp)] = sp
    prev = torch.rand(n, 2, generator=g) * 2 - 1
    prev[:64] = act[:64] - 0.01  # small positive deltas around the 0.05/3 oscillation threshold
    r = torch.rand(n, generator=g) * 12.0
    r[:100] = torch.rand(100, generator=g) * 0.4  # many near the 0.18 success radius
    th = torch.rand(n, generator=g) * 2 * torch.pi
    pos_b = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.randn(n, generator=g)], dim=1)
    ep = torch.randint(0, 752, (n,), generator=g)
    force = torch.where(torch.rand(n, 1, 1, 1, generator=g) < 0.9, 0.0, 1.0) * torch.randn(n, 14, 1, 3, generator=g) * 2
    force[:32] *= 0.2  # sums close to the > 1 test
    processed, jpos, jvel = H.ref_ackermann2(act)
    v1_pos, v1_vel, v3_pos, v3_vel = H.ref_ackermann_variants(act)
    from oracle.terms import RoverConstants
    exomy = RoverConstants(wheelbase_length=0.29778, middle_wheel_distance=0.1548, rear_and_front_wheel_distance=0.1548,
                           wheel_radius=0.1, min_steering_radius=0.4, action_offset=0.0)
    _, ex_pos, ex_vel = H.ref_ackermann2(act, exomy)
    rewards, terms = H.ref_rewards_terminations(pos_b, act, prev, ep, force)
    sensor_pos = torch.randn(n, 3, generator=g)
    hits = torch.randn(n, 7, 3, generator=g)
    hits[::5, 3] = float("inf")
    d, a, h = H.ref_observations(pos_b, sensor_pos, hits)
    np.savez_compressed(
        path, in_actions=act.numpy(), in_prev_actions=prev.numpy(), in_pos_b=pos_b.numpy(), in_ep_len=ep.numpy(),
        in_force=force.numpy(), in_sensor_pos=sensor_pos.numpy(), in_hits=hits.numpy(),
        ref_processed=processed.numpy(), ref_joint_pos=jpos.numpy(), ref_joint_vel=jvel.numpy(),
        ref_v1_joint_pos=v1_pos.numpy(), ref_v1_joint_vel=v1_vel.numpy(), ref_v3_joint_pos=v3_pos.numpy(),
        ref_v3_joint_vel=v3_vel.numpy(), ref_exomy_joint_pos=ex_pos.numpy(), ref_exomy_joint_vel=ex_vel.numpy(),
        ref_rewards=rewards.numpy(), ref_terms=terms.numpy(), ref_obs_distance=d.numpy(), ref_obs_angle=a.numpy(),
        ref_obs_scan=h.numpy())


def gen_terrain_and_command(path):
    v, f = TR.make_synthetic_terrain(**SMALL_TERRAIN)
    n = 64
    tm = H.ref_terrain_manager(v, f, n)
    hm = tm._heightmap_manager.heightmap
    g = torch.Generator().manual_seed(4321)
    # -- command update / metrics (terrain_importer.py:97-106)
    root_pos = torch.stack([torch.rand(n, generator=g) * 8 + 20, torch.rand(n, generator=g) * 8 + 20,
                            torch.randn(n, generator=g) * 0.3], dim=1)
    yaw = (torch.rand(n, generator=g) * 2 - 1) * torch.pi
    rp = torch.randn(n, 2, generator=g) * 0.1
    # yaw * pitch * roll composed quaternion (w,x,y,z)
    cy, sy, cp, sp_, cr, sr = (torch.cos(yaw / 2), torch.sin(yaw / 2), torch.cos(rp[:, 0] / 2), torch.sin(rp[:, 0] / 2),
                               torch.cos(rp[:, 1] / 2), torch.sin(rp[:, 1] / 2))
    quat = torch.stack([cr * cp * cy + sr * sp_ * sy, sr * cp * cy - cr * sp_ * sy, cr * sp_ * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp_ * cy], dim=1)
    env_origins = root_pos.clone()
    term, terrain, env = H.ref_command_term(tm, n, root_pos.clone(), quat.clone(), env_origins.clone())
    term.pos_command_w[:] = root_pos + torch.randn(n, 3, generator=g) * 4
    term.heading_command_w[:] = (torch.rand(n, generator=g) * 2 - 1) * torch.pi
    in_pos_cmd_w = term.pos_command_w.clone()
    in_heading_cmd_w = term.heading_command_w.clone()
    term._update_command()
    term._update_metrics()
    ref_pos_b, ref_heading_b = term.pos_command_b.clone(), term.heading_command_b.clone()
    ref_err_pos, ref_err_heading = term.metrics["error_pos"].clone(), term.metrics["error_heading"].clone()
    # -- reset (randomizations.py:12-39) then command resample (terrain_importer.py:74-95, 134-175)
    ids = torch.tensor([1, 2, 5, 8, 13, 21, 34, 55, 60, 63])
    rounds_cap = 16
    spawn_perm = torch.randperm(2 * n, generator=g)
    yaw_u = torch.rand(n, generator=g)
    theta_u = torch.rand(n, rounds_cap, generator=g)
    heading_u = torch.rand(n, generator=g)
    pose, _ = H.ref_reset_root_state(env, ids, spawn_perm, yaw_u)
    ref_env_origins = terrain.env_origins.clone()
    rounds = H.ref_resample_command(term, terrain, ids, theta_u, heading_u)
    assert rounds <= rounds_cap, rounds
    # direct lookups (terrain_utils.py:62-84, 202-223) on random points, including out-of-range ones
    pts = torch.rand(512, 2, generator=g) * 60 - 6
    hts = tm._heightmap_manager.get_height_at(pts)
    bad_ids, _ = tm.check_if_target_is_valid(torch.arange(512), pts, device="cpu")
    invalid = torch.zeros(512, dtype=torch.bool)
    invalid[bad_ids] = True
    np.savez_compressed(
        path, terrain_size_m=SMALL_TERRAIN["size_m"], terrain_grid_res=SMALL_TERRAIN["grid_res"],
        terrain_seed=SMALL_TERRAIN["seed"], terrain_vertex_z=v[:, 2].copy(),
        ref_heightmap_sha=sha(hm), ref_rock_sha=sha(tm.rock_mask.astype(np.uint8)),
        ref_safe_sha=sha(tm.safe_rock_mask.astype(np.uint8)), ref_heightmap_dec=hm[::8, ::8].copy(),
        ref_safe_dec=tm.safe_rock_mask[::8, ::8].astype(np.uint8),
        ref_min_xy=np.array([tm._heightmap_manager.min_x, tm._heightmap_manager.min_y], dtype=np.float32),
        ref_spawn_table=tm.spawn_locations.numpy(),
        in_root_pos=root_pos.numpy(), in_root_quat=quat.numpy(), in_pos_cmd_w=in_pos_cmd_w.numpy(),
        in_heading_cmd_w=in_heading_cmd_w.numpy(), ref_pos_b=ref_pos_b.numpy(), ref_heading_b=ref_heading_b.numpy(),
        ref_err_pos=ref_err_pos.numpy(), ref_err_heading=ref_err_heading.numpy(),
        in_reset_ids=ids.numpy(), in_spawn_perm=spawn_perm.numpy(), in_yaw_u=yaw_u.numpy(), in_theta_u=theta_u.numpy(),
        in_heading_u=heading_u.numpy(), ref_reset_pose=pose.numpy(), ref_env_origins=ref_env_origins.numpy(),
        ref_resampled_pos_cmd_w=term.pos_command_w.numpy(), ref_resampled_heading_cmd_w=term.heading_command_w.numpy(),
        ref_rounds=rounds, in_points=pts.numpy(), ref_heights=hts.numpy(), ref_invalid=invalid.numpy())


def gen_policy(path):
    net = H.ref_policy()
    g = torch.Generator().manual_seed(99)
    n = 64
    obs = torch.cat([torch.rand(n, 2, generator=g) * 2 - 1, torch.rand(n, 1, generator=g) * 1.3,
                     torch.rand(n, 1, generator=g) * 2 - 1, torch.randn(n, 961, generator=g) * 0.15], dim=1)
    with torch.no_grad():
        mean, log_std, _ = net.compute({"states": obs}, role="policy")
    sd = {k.replace(".", "__"): v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(path, in_obs=obs.numpy(), ref_mean=mean.numpy(), ref_log_std=log_std.detach().numpy(), **sd)


def gen_value(path):
    net = H.ref_value()
    g = torch.Generator().manual_seed(101)
    n = 64
    obs = torch.cat([torch.rand(n, 2, generator=g) * 2 - 1, torch.rand(n, 1, generator=g) * 1.3,
                     torch.rand(n, 1, generator=g) * 2 - 1, torch.randn(n, 961, generator=g) * 0.15], dim=1)
    with torch.no_grad():
        value, _ = net.compute({"states": obs}, role="value")
    sd = {k.replace(".", "__"): v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(path, in_obs=obs.numpy(), ref_value=value.numpy(), **sd)


def main():
    assert ref_loader.available(), "run this where /root/reference exists"
    if "--value-only" in sys.argv:  # added after the other fixtures were committed: leaves them untouched
        gen_value(os.path.join(HERE, "value.npz"))
        return
    gen_terms(os.path.join(HERE, "terms.npz"))
    gen_terrain_and_command(os.path.join(HERE, "terrain_command.npz"))
    gen_policy(os.path.join(HERE, "policy.npz"))
    gen_value(os.path.join(HERE, "value.npz"))
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
