"""Bring-up check of the fused scan + policy kernel: sizes one by one, progress printed after each (run under `timeout`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs  # noqa: E402

dev = torch.device("cuda:0")
size = float(os.environ.get("SIZE", "48"))
v, f = TR.make_synthetic_terrain(size, 0.2, seed=3)
grid = ops.ScanGridHandle.from_mesh(v, f, dev)
rays = ops.RayPattern.grid(dev)
net = GaussianNeuralNetwork(device=dev)
g = torch.Generator().manual_seed(1)
net.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01) for k, t in net.state_dict().items()})
for n in [int(a) for a in sys.argv[1:]] or [1, 16, 17, 148, 149, 148 * 16 + 1, 5000]:
    gen = torch.Generator().manual_seed(n)
    p, q = synthetic.make_poses(n, gen, torch.from_numpy(v), size, 0.2, margin=4.0)
    p, q = p.to(dev), q.to(dev)
    head = (torch.rand(n, 4, generator=gen) * 2 - 1).to(dev)
    ref = alloc_obs(n, dev)
    ref[:, :4] = head
    ops.height_scan(p, q, rays, grid, out=ref[:, 4:])
    ref_mean = net.compute({"states": ref})[0]
    obs = alloc_obs(n, dev)
    obs[:, :4] = head
    torch.cuda.synchronize()
    print(f"n={n}: launching fused", flush=True)
    mean = ops.height_scan_policy(p, q, rays, grid, obs, net, write_obs=True)
    torch.cuda.synchronize()
    fin = torch.isfinite(ref_mean).all(dim=1)
    err = (mean[fin] - ref_mean[fin]).abs().max().item() if fin.any() else float("nan")
    print(f"n={n}: obs equal {torch.equal(obs, ref)}, finite {int(fin.sum())}/{n}, max |mean - ref| = {err:.3e}, "
          f"bit-equal {torch.equal(mean[fin], ref_mean[fin])}", flush=True)
    if err > 4e-3 or not torch.equal(obs, ref):
        bad = (mean - ref_mean).abs().max(dim=1).values
        print("  worst envs:", bad.topk(min(5, n)).indices.tolist(), bad.topk(min(5, n)).values.tolist(), flush=True)
        print("  mean[:4]", mean[:4].tolist(), "ref[:4]", ref_mean[:4].tolist(), flush=True)
