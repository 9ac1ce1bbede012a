"""Timing of the whole-step launch (rover_step_fused) against mdp_step + height_scan on the bench terrain (CUDA graphs,
CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402
from isaac_rover_orbit_b200.policy import alloc_obs  # noqa: E402
from isaac_rover_orbit_b200.trainer import capture_steps  # noqa: E402

dev = torch.device("cuda:0")
v, f, grid, tables = bench.build_world(16384, dev, dev)
vt = torch.from_numpy(v)
rays = ops.RayPattern.grid(dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
fl = lambda: flush_buf.fill_(1)  # noqa: E731
for n in [int(a) for a in sys.argv[1:]] or [256, 8192, 16384]:
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    gen = torch.Generator().manual_seed(n)
    sets = [synthetic.make_step(n, gen, vt, 200.0, 0.2, cfg.num_contact_bodies, cfg.target_rounds).to(dev) for _ in range(4)]
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table[: 2 * n], tables.resolution, dev)
    drift = [(torch.rand(n, 3, generator=gen) * torch.tensor([4.0, 4.0, 0.0]) - torch.tensor([2.0, 2.0, 0.0])).to(dev) for _ in range(4)]
    res = {}
    for mode in ("two launches", "one launch"):
        buf = ops.MdpBuffers.allocate(n, dev)
        buf.env_origins.copy_(sets[0].root_pos_w)
        buf.time_left.fill_(150.0)
        buf.pos_cmd_w.copy_(sets[0].root_pos_w + torch.tensor([9.0, 0.0, 0.0], device=dev))
        obs = alloc_obs(n, dev)
        rng = ops.ResetRng(1, dev)

        def step(i):
            s = sets[i % 4]
            torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)
            if mode == "one launch":
                ops.step_fused(buf, params, th, s.actions, s.force_matrix_w, s.root_pos_w, s.root_quat_w, rays, grid, obs, rng)
            else:
                ops.mdp_step(buf, params, th, s.actions, s.force_matrix_w, s.root_pos_w, s.root_quat_w, obs=obs, rng=rng)
                ops.height_scan(s.root_pos_w, s.root_quat_w, rays, grid, out=obs[:, 4:])

        ms = bench.time_steps(capture_steps(step, n_variants=4, warmup=1), 200, 5, fl, stream)
        res[mode] = ms.mean() * 1e3
    print(f"n={n}: " + ", ".join(f"{k} {t:.1f} us ({n / t:.1f} M env-steps/s)" for k, t in res.items()), flush=True)
