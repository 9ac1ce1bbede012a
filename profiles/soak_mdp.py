"""Soak: the single-launch MDP step (split CTA: env warps + kinematics warps, variates / spawn rows handed over through shared
memory, statistics published mid-way) against the two-launch pre-step + post-step from identical states and the same
{seed, step}, thousands of steps, every piece of state compared bit for bit.  Usage: python profiles/soak_mdp.py [steps] [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402

STATE = ("action", "prev_action", "pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "time_left", "command_counter",
         "episode_length_buf", "episode_sums", "env_origins", "err_pos", "err_heading", "spawn_index", "log", "reward", "reset_flags",
         "term_rewards", "joint_pos", "joint_vel", "processed_actions")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 + 37
dev = torch.device("cuda:0")
v, f = TR.make_synthetic_terrain(96.0, 0.2, seed=3)
tables = TR.build_terrain_tables(v, f, n, build_device=dev)
cfg = RoverEnvCfg(num_envs=n)
params = ops.mdp_params(cfg)
params.resampling_time = 0.35  # (> step_dt) the command timer runs out every few steps: time-based re-draws beside the resets
mask = torch.as_tensor(tables.safe_mask).clone()
g = torch.Generator().manual_seed(5)
mask[torch.rand(mask.shape, generator=g) < 0.6] = 1  # many rejected target candidates: several rounds, some exhausted
th = ops.TerrainTablesHandle(tables.heightmap, mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
vt = torch.from_numpy(v)
sets = [synthetic.make_step(n, g, vt, 96.0, 0.2, margin=4.0).to(dev) for _ in range(16)]
a, b = ops.MdpBuffers.allocate(n, dev), ops.MdpBuffers.allocate(n, dev)
for buf in (a, b):
    buf.env_origins.copy_(sets[0].root_pos_w)
    buf.time_left.copy_(torch.rand(n, generator=g).to(dev) * 3.0 + 0.05)  # timers run out all the time as well
    buf.pos_cmd_w.copy_(sets[0].root_pos_w + torch.tensor([5.0, 0.0, 0.0], device=dev))
ra, rb = ops.ResetRng(77, dev), ops.ResetRng(77, dev)
oa, ob = torch.zeros(n, 968, device=dev)[:, :965], torch.zeros(n, 968, device=dev)[:, :965]
resets = exhausted = timers = 0
for t in range(steps):
    s = sets[t % 16]
    drift = (torch.rand(n, 3, generator=g) * torch.tensor([4.0, 4.0, 0.0]) - torch.tensor([2.0, 2.0, 0.0])).to(dev)
    pa, qa = a.env_origins + drift, s.root_quat_w.clone()
    pb, qb = pa.clone(), qa.clone()
    a.stats.zero_(), b.stats.zero_()
    ops.mdp_step(a, params, th, s.actions, s.force_matrix_w, pa, qa, obs=oa, rng=ra, n_rounds=9)
    ops.mdp_pre_step(b, params, s.actions, s.force_matrix_w)
    ops.mdp_post_step(b, params, th, pb, qb, obs=ob, rng=rb, n_rounds=9)
    if t % 50 == 49 or t < 5:
        torch.cuda.synchronize()
        for name in STATE:
            if hasattr(a, name):
                assert torch.equal(getattr(a, name), getattr(b, name)), (t, name)
        assert torch.equal(pa, pb) and torch.equal(qa, qb) and torch.equal(oa[:, :4], ob[:, :4]), t
        assert torch.equal(a.stats[7:11], b.stats[7:11]) and torch.equal(a.stats[13:], b.stats[13:]), t
        torch.testing.assert_close(a.stats, b.stats, rtol=1e-5, atol=1e-5)
    resets += 0 if t % 50 != 49 else int(a.stats[13])
    exhausted += 0 if t % 50 != 49 else int(a.stats[14])
    timers += 0 if t % 50 != 49 else int(a.stats[15])
torch.cuda.synchronize()
print(f"SOAK_OK {steps} steps x {n} envs; sampled steps saw {resets} resets, {exhausted} exhausted target draws, {timers} timer re-draws")
