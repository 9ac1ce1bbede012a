"""Soak: a closed loop (MDP step -> height scan -> policy + value -> Gaussian act -> next step) for many steps; prints a
checksum of the trajectory.  Run once with ROVER_PDL=1 and once with ROVER_PDL=0: programmatic dependent launch must not
change a bit.  Usage: python profiles/soak_pdl.py [steps] [envs]"""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402
from isaac_rover_orbit_b200.policy import DeterministicNeuralNetwork, GaussianNeuralNetwork, alloc_obs, policy_value_forward  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda:0")
v, f = TR.make_synthetic_terrain(96.0, 0.2, seed=3)
tables = TR.build_terrain_tables(v, f, n, build_device=dev)
grid = ops.ScanGridHandle.from_mesh(v, f, dev)
rays = ops.RayPattern.grid(dev)
cfg = RoverEnvCfg(num_envs=n)
params = ops.mdp_params(cfg)
th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
g = torch.Generator().manual_seed(9)
s = synthetic.make_step(n, g, torch.from_numpy(v), 96.0, 0.2, margin=4.0).to(dev)
nets = []
for cls in (GaussianNeuralNetwork, DeterministicNeuralNetwork):
    net = cls(device=dev)
    net.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01) for k, t in net.state_dict().items()})
    nets.append(net)
pol, val = nets
buf = ops.MdpBuffers.allocate(n, dev)
buf.env_origins.copy_(s.root_pos_w)
buf.time_left.fill_(150.0)
buf.pos_cmd_w.copy_(s.root_pos_w + torch.tensor([6.0, 0.0, 0.0], device=dev))
rng = ops.ResetRng(3, dev)
obs = alloc_obs(n, dev)
act = torch.zeros(n, 2, device=dev)
lp = torch.zeros(n, device=dev)
eps = [torch.randn(n, 2, generator=g).to(dev) for _ in range(8)]
drift = [(torch.rand(n, 3, generator=g) * torch.tensor([3.0, 3.0, 0.0])).to(dev) for _ in range(8)]
h = hashlib.sha256()
pos, quat = s.root_pos_w.clone(), s.root_quat_w.clone()
value_sum = torch.zeros((), device=dev, dtype=torch.float64)
for t in range(steps):  # no synchronisation inside the loop: kernels follow each other as closely as the launches allow
    torch.add(buf.env_origins, drift[t % 8], out=pos)
    ops.mdp_step(buf, params, th, act, s.force_matrix_w, pos, quat, obs=obs, rng=rng)
    ops.height_scan(pos, quat, rays, grid, out=obs[:, 4:])
    mean, value = policy_value_forward(pol, val, obs)
    torch.ops.rover_b200.gaussian_act_out(mean, pol.log_std_parameter, eps[t % 8], act, lp)
    value_sum += value.double().sum()
    if t % 50 == 49:
        torch.cuda.synchronize()
        for x in (obs, act, buf.reward, buf.episode_sums, buf.pos_cmd_w, pos):
            h.update(x.contiguous().cpu().numpy().tobytes())
torch.cuda.synchronize()
print(f"TRAJECTORY {h.hexdigest()[:32]} value_sum {float(value_sum):.6f} steps {steps} envs {n} pdl {os.environ.get('ROVER_PDL', '1')}")
