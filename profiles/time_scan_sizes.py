"""Height scan (default variant) at several environment counts: launch time with a flushed L2 (CUDA events)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
v, f, grid, _ = bench.build_world(None, dev, dev)
rays = ops.RayPattern.grid(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (1024, 4096, 16384, 65536):
    gen = torch.Generator().manual_seed(7)
    p, q = synthetic.make_poses(n, gen, torch.from_numpy(v), bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"])
    p, q = p.to(dev), q.to(dev)
    out = torch.empty(n, 961, device=dev)
    ts = []
    for i in range(30):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.height_scan(p, q, rays, grid, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[5:])
    med = ts[len(ts) // 2]
    alg = 4.0 * n * 961 + 28.0 * n + 36.0e6
    print(f"envs {n:6d}: median {med:7.1f} us  {n * 961 / med / 1e3:7.1f} Grays/s  algorithmic {alg / 1e6:6.1f} MB -> "
          f"{alg / med / 1e3:6.0f} GB/s = {100 * alg / med / 1e3 / 6554.6:4.1f} % of the measured HBM peak")
