"""Bring-up timing of the paired scan kernel built with ROVER_SCAN_DBG=1/2 (supply-only / compute-only)."""
import sys

import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = sys.argv[1]
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
v, f, grid, _ = bench.build_world(None, dev, dev)
gen = torch.Generator().manual_seed(7)
p, q = synthetic.make_poses(4096, gen, torch.from_numpy(v), bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"])
p, q = p.to(dev), q.to(dev)
rays = ops.RayPattern.grid(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty(4096, 961, device=dev)
for var in (4, 5):
    for cold in (True, False):
        ts = []
        for i in range(25):
            if cold:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.height_scan(p, q, rays, grid, out=out, variant=var)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[5:])
        print(f"{_lib.LIB_PATH.split('/')[-1]} variant {var} {'cold' if cold else 'warm'} L2: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
