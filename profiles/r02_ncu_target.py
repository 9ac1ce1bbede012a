"""Round-2 ncu target: one warm-up launch and one profiled launch (L2 flushed before it) of every hot kernel.

    python profiles/r02_ncu_target.py && ncu --set full --clock-control none --import-source on \
        -k regex:'height_scan_paired|fused_scan_encoder|policy_mlp|mdp_fused_step|height_scan_step|policy_forward_ws' -s 10 -c 10 \
        -o gpurun_out/r02_full python profiles/r02_ncu_target.py

Launch order of the matching kernels (the same in both passes): scan @ 4096, 16384, 65536 envs; fused scan + encoder and
the MLP @ 16384; the single-launch MDP step and the whole-step launch @ 16384; policy forward on fp32 / bf16 observations and the policy + value pass @ 65536.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402
from isaac_rover_orbit_b200.policy import (DeterministicNeuralNetwork, GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16,  # noqa: E402
                                            policy_value_forward)

dev = torch.device("cuda:0")
v, f, grid, tables = bench.build_world(16384, dev, dev)
vt = torch.from_numpy(v)
rays = ops.RayPattern.grid(dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator().manual_seed(11)
poses = {n: tuple(t.to(dev) for t in synthetic.make_poses(n, gen, vt, 200.0, 0.2)) for n in (4096, 16384, 65536)}
outs = {n: torch.empty(n, 961, device=dev) for n in poses}
net = GaussianNeuralNetwork(device=dev)
g = torch.Generator().manual_seed(1)
net.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01) for k, t in net.state_dict().items()})
vnet = DeterministicNeuralNetwork(device=dev)
vnet.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01) for k, t in vnet.state_dict().items()})
obs16 = alloc_obs(16384, dev)
obs16[:, :4] = torch.rand(16384, 4, device=dev)
pol32 = alloc_obs(65536, dev)
pol32.copy_(torch.randn(65536, 965, device=dev) * 0.3)
pol16 = alloc_obs_bf16(65536, dev)
pol16.copy_(pol32)
cfg = RoverEnvCfg(num_envs=16384)
params = ops.mdp_params(cfg)
st = synthetic.make_step(16384, gen, vt, 200.0, 0.2, cfg.num_contact_bodies, cfg.target_rounds).to(dev)
buf = ops.MdpBuffers.allocate(16384, dev)
th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
buf.env_origins.copy_(st.root_pos_w)
buf.time_left.fill_(150.0)
buf.pos_cmd_w.copy_(st.root_pos_w + torch.tensor([9.0, 0.0, 0.0], device=dev))
buf.pos_cmd_b.copy_(torch.tensor([9.0, 0.0, 0.0], device=dev).expand(16384, 3))  # 9 m from the target: resets come from contacts (5 %)
rng = ops.ResetRng(5, dev)
mdp_obs = alloc_obs(16384, dev)
for rep in range(2):
    for n in (4096, 16384, 65536):
        flush_buf.fill_(1)
        ops.height_scan(*poses[n], rays, grid, out=outs[n])
    flush_buf.fill_(1)
    ops.height_scan_policy(*poses[16384], rays, grid, obs16, net, write_obs=True)
    flush_buf.fill_(1)
    ops.mdp_step(buf, params, th, st.actions, st.force_matrix_w, st.root_pos_w.clone(), st.root_quat_w.clone(), obs=mdp_obs,
                 rng=rng)
    flush_buf.fill_(1)
    ops.step_fused(buf, params, th, st.actions, st.force_matrix_w, st.root_pos_w.clone(), st.root_quat_w.clone(), rays, grid,
                   mdp_obs, rng)
    flush_buf.fill_(1)
    net.compute({"states": pol32})
    flush_buf.fill_(1)
    net.compute({"states": pol16})
    flush_buf.fill_(1)
    policy_value_forward(net, vnet, pol32)
    torch.cuda.synchronize()
print("ok")
