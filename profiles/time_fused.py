"""Timing of the fused scan + policy launch against the unfused pair on the bench terrain (CUDA events, L2 flushed)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16  # noqa: E402

dev = torch.device("cuda:0")
v, f = TR.make_synthetic_terrain(**bench.TERRAIN)
grid = ops.ScanGridHandle.from_mesh(v, f, dev)
rays = ops.RayPattern.grid(dev)
net = GaussianNeuralNetwork(device=dev)
g = torch.Generator().manual_seed(1)
net.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01) for k, t in net.state_dict().items()})
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
for n in [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]:
    gen = torch.Generator().manual_seed(n)
    sets = [tuple(t.to(dev) for t in synthetic.make_poses(n, gen, torch.from_numpy(v), 200.0, 0.2)) for _ in range(2)]
    obs = alloc_obs(n, dev)
    obs[:, :4] = torch.rand(n, 4, device=dev)
    obs_bf = alloc_obs_bf16(n, dev)
    t = {}
    fl = lambda: flush_buf.fill_(1)  # noqa: E731
    t["scan"] = bench.time_steps(lambda i: ops.height_scan(*sets[i % 2], rays, grid, out=obs[:, 4:]), 100, 3, fl, stream).mean()
    t["scan_obs(bf16 mirror)"] = bench.time_steps(lambda i: ops.height_scan_obs(*sets[i % 2], rays, grid, obs, obs_bf), 100, 3, fl, stream).mean()
    t["policy fp32 obs"] = bench.time_steps(lambda i: net.compute({"states": obs}), 100, 3, fl, stream).mean()
    t["policy bf16 obs"] = bench.time_steps(lambda i: net.compute({"states": obs_bf}), 100, 3, fl, stream).mean()
    t["fused write_obs"] = bench.time_steps(lambda i: ops.height_scan_policy(*sets[i % 2], rays, grid, obs, net, True), 100, 3, fl, stream).mean()
    t["fused no obs"] = bench.time_steps(lambda i: ops.height_scan_policy(*sets[i % 2], rays, grid, obs, net, False), 100, 3, fl, stream).mean()
    print(f"n={n}: " + ", ".join(f"{k} {v * 1e3:.1f} us" for k, v in t.items()), flush=True)
