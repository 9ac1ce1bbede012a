"""GPU: the whole non-physics step in one launch (``rover_step_fused``: the MDP step run by one warp of every CTA of the
persistent height-scan kernel) against the two launches it replaces -- ``rover_mdp_step_v3`` with the in-kernel variates,
then ``rover_height_scan`` -- over several closed-loop steps: every per-env quantity bit-identical (state, outputs, root
poses, spawn rows, observation head + heights), statistics / episode log equal up to fp32 summation order."""
import pytest
import torch

from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.policy import alloc_obs

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
SIZE, RES = 48.0, 0.2
PER_ENV = ("action", "prev_action", "pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "time_left", "command_counter",
           "episode_length_buf", "episode_sums", "env_origins", "err_pos", "err_heading", "spawn_index", "reward", "reset_flags",
           "processed_actions", "joint_pos", "joint_vel", "term_rewards", "term_values", "terminated", "truncated", "term_flags")


@pytest.fixture(scope="module")
def world(cuda_device):
    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    return dict(v=v, f=f, grid=grid, dev=cuda_device, rays=ops.RayPattern.grid(cuda_device))


@pytest.mark.parametrize("n", [1, 31, 33, 148, 149, 148 * 32 + 5, 9000])
def test_step_fused_equals_mdp_step_then_scan(world, n):
    dev, grid, rays = world["dev"], world["grid"], world["rays"]
    tables = TR.build_terrain_tables(world["v"], world["f"], n)
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
    gen = torch.Generator().manual_seed(50 + n)
    vt = torch.from_numpy(world["v"])
    steps = [synthetic.make_step(n, gen, vt, SIZE, RES, margin=4.0).to(dev) for _ in range(4)]
    pc, hc, ep = synthetic.init_commands(n, gen, steps[0].root_pos_w.cpu())
    bufs, rngs, obss = [], [], []
    for _ in range(2):
        b = ops.MdpBuffers.allocate(n, dev)
        b.pos_cmd_w.copy_(pc)
        b.heading_cmd_w.copy_(hc)
        b.episode_length_buf.copy_(ep)
        b.env_origins.copy_(steps[0].root_pos_w)
        b.time_left.fill_(150.0)
        b.time_left[min(3, n - 1)] = 0.1
        bufs.append(b)
        rngs.append(ops.ResetRng(77, dev, step=5))
        obss.append(alloc_obs(n, dev))
    total_resets = 0.0
    for k, s in enumerate(steps):
        roots = [(s.root_pos_w.clone(), s.root_quat_w.clone()) for _ in range(2)]
        for b in bufs:
            b.stats.zero_()
        # the two launches
        ops.mdp_step(bufs[0], params, th, s.actions, s.force_matrix_w, *roots[0], obs=obss[0], rng=rngs[0])
        ops.height_scan(*roots[0], rays, grid, out=obss[0][:, 4:])
        # the one launch
        ops.step_fused(bufs[1], params, th, s.actions, s.force_matrix_w, *roots[1], rays, grid, obss[1], rngs[1])
        torch.cuda.synchronize()
        for name in PER_ENV:
            assert torch.equal(getattr(bufs[0], name), getattr(bufs[1], name)), (k, name)
        assert torch.equal(roots[0][0], roots[1][0]) and torch.equal(roots[0][1], roots[1][1]), k
        assert torch.equal(obss[0], obss[1]), k
        assert rngs[0].peek() == rngs[1].peek() == (77, 6 + k)
        torch.testing.assert_close(bufs[0].stats, bufs[1].stats, rtol=1e-5, atol=1e-5)
        assert torch.equal(bufs[0].stats[7:11], bufs[1].stats[7:11]) and torch.equal(bufs[0].stats[13:], bufs[1].stats[13:])
        torch.testing.assert_close(bufs[0].log, bufs[1].log, rtol=1e-5, atol=1e-6)
        total_resets += float(bufs[1].stats[13])
    if n >= 148:
        assert total_resets > 0
    # repeated launches from the same state are deterministic (statistics and log included)
    res = []
    for _ in range(2):
        a = ops.MdpBuffers.allocate(n, dev)
        a.pos_cmd_w.copy_(pc)
        a.heading_cmd_w.copy_(hc)
        a.episode_length_buf.copy_(ep)
        a.env_origins.copy_(steps[0].root_pos_w)
        a.time_left.fill_(150.0)
        o = alloc_obs(n, dev)
        r3 = ops.ResetRng(3, dev)
        for s in steps[:2]:
            ops.step_fused(a, params, th, s.actions, s.force_matrix_w, s.root_pos_w.clone(), s.root_quat_w.clone(), rays, grid, o, r3)
        torch.cuda.synchronize()
        res.append((a.stats.clone(), a.log.clone(), o.clone(), a.reward.clone(), a.pos_cmd_w.clone()))
    assert all(torch.equal(x, y) for x, y in zip(res[0], res[1]))


def test_step_fused_argument_errors(world):
    dev, grid, rays = world["dev"], world["grid"], world["rays"]
    n = 64
    tables = TR.build_terrain_tables(world["v"], world["f"], n)
    params = ops.mdp_params(RoverEnvCfg(num_envs=n))
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
    s = synthetic.make_step(n, torch.Generator().manual_seed(1), torch.from_numpy(world["v"]), SIZE, RES, margin=4.0).to(dev)
    b = ops.MdpBuffers.allocate(n, dev)
    b.time_left.fill_(150.0)
    with pytest.raises(RuntimeError, match="obs"):
        ops.step_fused(b, params, th, s.actions, s.force_matrix_w, s.root_pos_w, s.root_quat_w, rays, grid,
                       torch.zeros(n, 100, device=dev), ops.ResetRng(0, dev))
    small = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table[:10],
                                    tables.resolution, dev)
    with pytest.raises(RuntimeError, match="spawn table"):
        ops.step_fused(b, params, small, s.actions, s.force_matrix_w, s.root_pos_w, s.root_quat_w, rays, grid,
                       alloc_obs(n, dev), ops.ResetRng(0, dev))
