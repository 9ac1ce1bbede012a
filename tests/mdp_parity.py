"""Shared by tests/test_gpu_parity.py (small terrain) and tests/test_gpu_fullsize.py (cfg-3 size): the fused MDP kernels
against the CPU oracle, step by step.  Tolerances as stated in test_gpu_parity.py's header."""
import numpy as np
import torch

from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200.config import RoverEnvCfg


def _load_state(buf, ost):
    for k in ("action", "prev_action", "pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "time_left",
              "command_counter", "episode_length_buf", "episode_sums", "env_origins", "err_pos", "err_heading"):
        getattr(buf, k).copy_(getattr(ost, k))


def _near(x, thr, tol):
    return (x - thr).abs() <= tol


def check_mdp_steps_vs_oracle(dev, n, v, tables, size, res, n_steps=6, margin=4.0, max_ambiguous=2):
    """`n_steps` consecutive steps; before each step the CUDA buffers are loaded with the oracle's state so that every
    step is an independent single-step parity check over an evolving state distribution."""
    from oracle import step as OS
    from oracle import terms as OT

    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    otab = OS.TerrainTables(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table)
    # resampled targets are world coordinates: libm vs CUDA sinf/cosf differ in the last ulp, which after `+ origin` is one
    # ulp of a coordinate up to `size` metres; quantities derived from target - root inherit a few of those
    atol_w = max(2e-5, 4.0 * float(np.spacing(np.float32(size))))
    gen = torch.Generator().manual_seed(21)
    vt = torch.from_numpy(v)
    st0 = synthetic.make_step(n, gen, vt, size, res, margin=margin)
    pos_cmd_w, heading_cmd_w, ep_len = synthetic.init_commands(n, gen, st0.root_pos_w)
    ost = OS.MdpState.zeros(n)
    ost.pos_cmd_w[:] = pos_cmd_w
    ost.heading_cmd_w[:] = heading_cmd_w
    ost.episode_length_buf[:] = ep_len
    ost.env_origins[:] = st0.root_pos_w
    ost.time_left[:] = 150.0
    ost.time_left[5] = 0.1  # exercises the time-based resample branch (CommandTerm.compute)
    ost.pos_cmd_b[:], ost.heading_cmd_b[:] = OT.update_command(pos_cmd_w, heading_cmd_w, st0.root_pos_w,
                                                               st0.root_quat_w)
    buf = ops.MdpBuffers.allocate(n, dev)
    total_resets = 0
    for step in range(n_steps):
        st = synthetic.make_step(n, gen, vt, size, res, margin=margin)
        if step == 2:
            st.actions[:4] = torch.tensor([[0.0135, 0.0135], [0.0135, 0.7], [0.6, 0.0135], [-0.4, 0.2]])
        # Envs whose distance to the target sits within an ulp or two of a termination threshold would let libm and
        # CUDA disagree on the mask, and one flipped reset shifts the reset RANK (hence the spawn row) of every later
        # env.  Their command is moved off the threshold (same edit on both sides: the CUDA buffers are loaded from
        # the oracle state below), so that every check of the step runs on every env.
        dist0 = ost.pos_cmd_b[:, :2].norm(dim=1)
        amb0 = _near(dist0, 0.18, 1e-6) | _near(dist0, 11.0, 1e-5)
        assert amb0.sum() <= max_ambiguous
        ost.pos_cmd_b[amb0] *= 1.001
        _load_state(buf, ost)
        buf.stats.zero_()
        d = st.to(dev)
        pre_pos_b = ost.pos_cmd_b.clone()
        obs = torch.zeros(n, 965, device=dev)
        ops.mdp_pre_step(buf, params, d.actions, d.force_matrix_w)
        ops.mdp_post_step(buf, params, th, d.root_pos_w, d.root_quat_w, d.spawn_perm, d.yaw_u, d.heading_u,
                          d.theta_u, obs)
        torch.cuda.synchronize()
        out = OS.oracle_step(ost, st.actions, st.root_pos_w, st.root_quat_w, st.force_matrix_w, otab, st.spawn_perm,
                             st.yaw_u, st.theta_u, st.heading_u)
        # ---- masks: bit-exact away from the thresholds
        dist = pre_pos_b[:, :2].norm(dim=1)
        assert not (_near(dist, 0.18, 1e-6) | _near(dist, 11.0, 1e-5)).any()
        flags = buf.term_flags.cpu().bool()
        assert torch.equal(flags, out.term_flags)
        assert torch.equal(buf.terminated.cpu().bool(), out.terminated)
        assert torch.equal(buf.truncated.cpu().bool(), out.truncated)
        # ---- actions / rewards: 1e-5 relative
        torch.testing.assert_close(buf.processed_actions.cpu(), out.processed_actions, rtol=0, atol=0)
        torch.testing.assert_close(buf.joint_pos.cpu(), out.joint_pos, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(buf.joint_vel.cpu(), out.joint_vel, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(buf.term_rewards.cpu(), out.term_rewards, rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(buf.reward.cpu(), out.reward, rtol=1e-5, atol=1e-8)
        # ---- resets: indices bit-exact
        ids = out.reset_ids
        total_resets += len(ids)
        sp = buf.spawn_index.cpu()
        assert torch.equal((sp >= 0).nonzero().squeeze(-1), ids)
        assert torch.equal(sp[ids], out.spawn_index)
        torch.testing.assert_close(d.root_pos_w.cpu(), out.root_pos_w, rtol=0, atol=0)
        torch.testing.assert_close(d.root_quat_w.cpu(), out.root_quat_w, rtol=1e-6, atol=1e-7)
        # ---- state
        assert torch.equal(buf.episode_length_buf.cpu(), ost.episode_length_buf)
        assert torch.equal(buf.command_counter.cpu(), ost.command_counter)
        assert torch.equal(buf.action.cpu(), ost.action) and torch.equal(buf.prev_action.cpu(), ost.prev_action)
        assert torch.equal(buf.env_origins.cpu(), ost.env_origins)
        torch.testing.assert_close(buf.episode_sums.cpu(), ost.episode_sums, rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(buf.time_left.cpu(), ost.time_left, rtol=1e-6, atol=1e-6)
        # targets: sin/cos differ in the last ulp between libm and CUDA -> 9 m * 1e-6; the sampled cell (and so
        # z from the heightmap) is compared exactly unless the candidate sits within 1e-4 cells of a cell border
        torch.testing.assert_close(buf.pos_cmd_w.cpu()[:, :2], ost.pos_cmd_w[:, :2], rtol=1e-6, atol=atol_w)
        cell = ost.pos_cmd_w[:, :2] / 0.05 + tables.offset_xy
        border = ((cell - cell.round()).abs() < 1e-3).any(dim=1)
        assert torch.equal(buf.pos_cmd_w.cpu()[~border, 2], ost.pos_cmd_w[~border, 2])
        torch.testing.assert_close(buf.heading_cmd_w.cpu(), ost.heading_cmd_w, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(buf.pos_cmd_b.cpu(), ost.pos_cmd_b, rtol=1e-5, atol=atol_w)
        head_err = (buf.heading_cmd_b.cpu() - ost.heading_cmd_b).abs()
        assert (torch.minimum(head_err, (head_err - 2 * np.pi).abs()) < 1e-5).all()
        torch.testing.assert_close(buf.err_pos.cpu(), ost.err_pos, rtol=1e-5, atol=atol_w)
        torch.testing.assert_close(obs[:, :4].cpu(), out.obs_head, rtol=1e-5, atol=1e-5)
        # ---- episode statistics (deterministic reduction)
        s = buf.stats.cpu()
        torch.testing.assert_close(s[:7], out.stats["reward_sums"], rtol=1e-5, atol=1e-7)
        assert torch.equal(s[7:11].long(), out.stats["term_counts"].long())
        assert int(s[13]) == out.stats["num_resets"] and int(s[14]) == out.stats["target_rounds_exhausted"]
        assert int(s[15]) == out.stats["num_time_resamples"]
        torch.testing.assert_close(s[11], out.stats["err_pos_sum"], rtol=1e-5, atol=1e-5)
    assert total_resets > 10, "the fixture must exercise resets"
    return total_resets
