"""GPU: the reference-facing API (``RoverEnv.step`` + manager-term functions with the reference's names and
signatures) against the CPU oracle, closed loop over several steps.  Reads like the reference's own usage:
``mdp.distance_to_target_reward(env, "target_pose")``, ``AckermannAction2.process_actions`` ..."""
import types

import numpy as np
import pytest
import torch

from isaac_rover_orbit_b200 import mdp, synthetic
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.env import RoverEnv

pytestmark = pytest.mark.gpu

SIZE, RES, N = 48.0, 0.2, 192


def cfg_(name):
    return types.SimpleNamespace(name=name)


@pytest.fixture(scope="module")
def world(cuda_device):
    from oracle import raycast as oracle_raycast

    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    tables = TR.build_terrain_tables(v, f, N)
    return dict(v=v, f=f, tables=tables, mesh=oracle_raycast.Mesh(v, f), dev=cuda_device)


def _near(x, thr, tol):
    return (x - thr).abs() <= tol


def test_env_step_closed_loop_vs_oracle(world):
    from oracle import step as OS
    from oracle import terms as OT

    dev, tables = world["dev"], world["tables"]
    cfg = RoverEnvCfg(num_envs=N)
    cur = {}

    def physics(env):  # synthetic stand-in for PhysX: hands over root state and contact forces
        env.scene["robot"].data.root_pos_w.copy_(cur["st"].root_pos_w)
        env.scene["robot"].data.root_quat_w.copy_(cur["st"].root_quat_w)
        env.scene.sensors["contact_sensor"].data.force_matrix_w.copy_(cur["st"].force_matrix_w)

    env = RoverEnv(cfg, tables, dev, physics=physics, seed=1)
    assert env.max_episode_length == 750 and env.action_manager.get_term().action_dim == 2
    otab = OS.TerrainTables(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table)
    ost = OS.MdpState.zeros(N)
    ost.env_origins[:, :2] = 100.0  # rover_env.py:18-25
    ost.time_left[:] = 150.0
    gen = torch.Generator().manual_seed(77)
    vt = torch.from_numpy(world["v"])
    total_resets, compared = 0, 0
    for step in range(5):
        st = synthetic.make_step(N, gen, vt, SIZE, RES, margin=4.0)
        if step >= 2:  # keep the rover near its target so that not every env resets every step
            st.root_pos_w[:, :2] = ost.pos_cmd_w[:, :2] + torch.randn(N, 2, generator=gen) * 3.0
        cur["st"] = st.to(dev)
        env.set_variates(cur["st"].spawn_perm, cur["st"].yaw_u, cur["st"].heading_u, cur["st"].theta_u)
        pre_pos_b = ost.pos_cmd_b.clone()
        obs, rew, terminated, truncated, extras = env.step(cur["st"].actions)
        torch.cuda.synchronize()
        out = OS.oracle_step(ost, st.actions, st.root_pos_w, st.root_quat_w, st.force_matrix_w, otab, st.spawn_perm,
                             st.yaw_u, st.theta_u, st.heading_u)
        dist = pre_pos_b[:, :2].norm(dim=1)
        assert not (_near(dist, 0.18, 1e-6) | _near(dist, 11.0, 1e-5)).any(), "re-seed: sample on a threshold"
        total_resets += len(out.reset_ids)
        # ---- action term through its ActionTerm API
        term = env.action_manager.get_term()
        assert torch.equal(term.raw_actions.cpu(), st.actions)
        assert torch.equal(term.processed_actions.cpu(), out.processed_actions)
        robot = env.scene["robot"]
        torch.testing.assert_close(robot.joint_pos_target.cpu(), out.joint_pos, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(robot.joint_vel_target.cpu(), out.joint_vel, rtol=1e-5, atol=1e-6)
        # ---- step outputs
        assert torch.equal(terminated.cpu(), out.terminated) and torch.equal(truncated.cpu(), out.truncated)
        torch.testing.assert_close(rew.cpu(), out.reward, rtol=1e-5, atol=1e-8)
        # ---- reward / termination term functions, reference signatures (unweighted values)
        w = torch.tensor(cfg.rewards.weights) * cfg.step_dt
        vals = torch.stack([
            mdp.distance_to_target_reward(env, "target_pose"), mdp.reached_target(env, "target_pose", 0.18),
            mdp.oscillation_penalty(env), mdp.angle_to_target_penalty(env, "target_pose"),
            mdp.heading_soft_contraint(env, cfg_("robot")), mdp.collision_penalty(env, cfg_("contact_sensor"), 1.0),
            mdp.far_from_target_reward(env, "target_pose", 11.0)], dim=1).cpu()
        torch.testing.assert_close(vals * w, out.term_rewards, rtol=1e-5, atol=1e-9)
        flags = torch.stack([mdp.time_out(env), mdp.is_success(env, "target_pose", 0.18),
                             mdp.far_from_target(env, "target_pose", 11.0),
                             mdp.collision_with_obstacles(env, cfg_("contact_sensor"), 1.0)], dim=1).cpu()
        assert torch.equal(flags, out.term_flags)
        # ---- observations: manager output and the per-term functions
        assert obs.shape == (N, 965)
        torch.testing.assert_close(obs[:, :4].cpu(), out.obs_head, rtol=1e-5, atol=1e-5)
        assert torch.equal(mdp.last_action(env).cpu(), ost.action)
        torch.testing.assert_close(mdp.distance_to_target_euclidean(env, "target_pose").cpu(),
                                   OT.obs_distance(ost.pos_cmd_b), rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(mdp.angle_to_target_observation(env, "target_pose").cpu(),
                                   OT.obs_angle(ost.pos_cmd_b), rtol=1e-5, atol=1e-5)
        h_ref, _ = OS.height_scan(out.root_pos_w, out.root_quat_w, world["mesh"])
        h = mdp.height_scan_rover(env, cfg_("height_scanner")).cpu()
        assert torch.equal(torch.isinf(h), torch.isinf(h_ref))
        fin = ~torch.isinf(h_ref)
        assert (h[fin] - h_ref[fin]).abs().max() <= 1e-4  # 1e-5 x ray distance (~10 m)
        # ---- command term state and episode log
        cmd = env.command_manager.get_term("target_pose")
        torch.testing.assert_close(cmd.command.cpu(), ost.pos_cmd_b, rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(cmd.metrics["error_pos"].cpu(), ost.err_pos, rtol=1e-5, atol=2e-5)
        assert torch.equal(env.episode_length_buf.cpu(), ost.episode_length_buf)
        assert torch.equal(env.scene.terrain.env_origins.cpu(), ost.env_origins)
        log, ref_log = env.episode_log(), OS.episode_log(out.stats)
        for k, v in ref_log.items():
            assert abs(log[k] - v) <= 1e-5 * max(1.0, abs(v)), k
        assert set(extras["episode"]) == set(ref_log)
        compared += 1
    assert compared == 5 and total_resets >= N  # step 1 resets every env (zero command = "success")


def test_standalone_term_calls_vs_oracle(world):
    """reset_root_state_rover / _resample_command / sample_new_targets / _update_command called on their own,
    the way the ORBIT managers call the reference's functions."""
    from oracle import step as OS
    from oracle import terms as OT

    dev, tables = world["dev"], world["tables"]
    env = RoverEnv(RoverEnvCfg(num_envs=N), tables, dev, seed=2)
    otab = OS.TerrainTables(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table)
    gen = torch.Generator().manual_seed(5)
    st = synthetic.make_step(N, gen, torch.from_numpy(world["v"]), SIZE, RES, margin=4.0)
    d = st.to(dev)
    robot = env.scene["robot"].data
    robot.root_pos_w.copy_(d.root_pos_w)
    robot.root_quat_w.copy_(d.root_quat_w)
    env.scene.terrain.env_origins.copy_(d.root_pos_w)
    ids = torch.tensor([0, 3, 7, 64, 65, 100, 191])
    # -- reset_root_state_rover(env, env_ids, asset_cfg)
    env.set_variates(d.spawn_perm, d.yaw_u, d.heading_u, d.theta_u)
    mdp.reset_root_state_rover(env, ids.to(dev), cfg_("robot"))
    pos = tables.spawn_table[st.spawn_perm[: len(ids)]].clone()
    pos[:, 2] += 0.5
    assert torch.equal(robot.root_pos_w.cpu()[ids], pos) and torch.equal(env.scene.terrain.env_origins.cpu()[ids], pos)
    ang = st.yaw_u[ids] * 2 * torch.pi
    torch.testing.assert_close(robot.root_quat_w.cpu()[ids][:, [0, 3]],
                               torch.stack([torch.cos(ang / 2), torch.sin(ang / 2)], 1), rtol=1e-6, atol=1e-7)
    untouched = torch.ones(N, dtype=torch.bool)
    untouched[ids] = False
    assert torch.equal(robot.root_pos_w.cpu()[untouched], st.root_pos_w[untouched])
    # -- terrain.sample_new_targets(env_ids) and command._resample_command(env_ids)
    ost = OS.MdpState.zeros(N)
    ost.env_origins[:] = env.scene.terrain.env_origins.cpu()
    tgt_ref, exhausted = OS.sample_targets(ids, ost.env_origins, st.theta_u, otab)
    assert not exhausted.any()
    env.set_variates(d.spawn_perm, d.yaw_u, d.heading_u, d.theta_u)
    tgt = env.scene.terrain.sample_new_targets(ids.to(dev)).cpu()
    torch.testing.assert_close(tgt[:, :2], tgt_ref[:, :2], rtol=1e-6, atol=2e-5)
    cell = tgt_ref[:, :2] / 0.05 + tables.offset_xy
    ok = ~((cell - cell.round()).abs() < 1e-3).any(dim=1)
    assert torch.equal(tgt[ok, 2], tgt_ref[ok, 2])
    cmd = env.command_manager.get_term("target_pose")
    assert float(cmd.pos_command_w.abs().sum()) == 0.0, "sample_new_targets must not modify the command buffers"
    env.set_variates(d.spawn_perm, d.yaw_u, d.heading_u, d.theta_u)
    cmd._resample_command(ids.to(dev))
    torch.testing.assert_close(cmd.pos_command_w.cpu()[ids][:, :2], tgt_ref[:, :2], rtol=1e-6, atol=2e-5)
    torch.testing.assert_close(cmd.heading_command_w.cpu()[ids], st.heading_u[ids] * 2 * np.pi - np.pi, rtol=1e-6,
                               atol=1e-6)
    assert float(cmd.pos_command_w.cpu()[untouched].abs().sum()) == 0.0
    # -- _update_command / _update_metrics on the current state
    cmd._update_command()
    cmd._update_metrics()
    pb, hb = OT.update_command(cmd.pos_command_w.cpu(), cmd.heading_command_w.cpu(), robot.root_pos_w.cpu(),
                               robot.root_quat_w.cpu())
    torch.testing.assert_close(cmd.command.cpu(), pb, rtol=1e-5, atol=2e-5)
    ep, eh = OT.update_metrics(cmd.pos_command_w.cpu(), cmd.heading_command_w.cpu(), robot.root_pos_w.cpu(),
                               robot.root_quat_w.cpu())
    torch.testing.assert_close(cmd.metrics["error_pos"].cpu(), ep, rtol=1e-5, atol=2e-5)
    # -- RayCaster facade: sensor.data.ray_hits_w like the reference's consumer reads it (observations.py:42-45)
    sensor = env.scene.sensors["height_scanner"]
    hits = sensor.data.ray_hits_w
    h = sensor.data.pos_w[:, 2].unsqueeze(1) - hits[..., 2] - 0.26878
    h_ref, _ = OS.height_scan(robot.root_pos_w.cpu(), robot.root_quat_w.cpu(), world["mesh"])
    fin = ~torch.isinf(h_ref)
    assert torch.equal(torch.isinf(h.cpu()), ~fin) and (h.cpu()[fin] - h_ref[fin]).abs().max() <= 1e-4
