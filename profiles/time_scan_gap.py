"""Bring-up: how much of the event-timed scan is launch overhead?  Times the cfg-2 scan after an L2 flush (a) as bench.py
does and (b) with a 148-env warm launch of the same kernel between the flush and the first event (same shared-memory
carve-out, warm instruction cache)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
v, f, grid, _ = bench.build_world(None, dev, dev)
gen = torch.Generator().manual_seed(7)
p, q = synthetic.make_poses(4096, gen, torch.from_numpy(v), bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"])
p, q = p.to(dev), q.to(dev)
p2, q2 = p[:148].clone(), q[:148].clone()
rays = ops.RayPattern.grid(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty(4096, 961, device=dev)
out2 = torch.empty(148, 961, device=dev)
for mode in ("flush only", "flush + warm launch", "flush + torch add", "back-to-back x2 (second timed)"):
    ts = []
    for i in range(25):
        flush.fill_(1)
        if mode == "flush + warm launch":
            ops.height_scan(p2, q2, rays, grid, out=out2)
        elif mode == "flush + torch add":
            out2.add_(1.0)
        elif mode.startswith("back"):
            ops.height_scan(p, q, rays, grid, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.height_scan(p, q, rays, grid, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[5:])
    print(f"{mode}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
