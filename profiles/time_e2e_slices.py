"""e2e (host buffers) through rover_height_scan_host for several slice counts, against torch copies around one launch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
n = 4096
v, f, grid, tables = bench.build_world(n, dev, dev)
rays = ops.RayPattern.grid(dev)
gen = torch.Generator().manual_seed(2)
pin = [tuple(t.pin_memory() for t in synthetic.make_poses(n, gen, torch.from_numpy(v), 200.0, 0.2)) for _ in range(8)]
host_out = torch.empty(n, 961).pin_memory()
work = ops.HostScanWork(n, 961, dev)
p_d, q_d, out = torch.empty(n, 3, device=dev), torch.empty(n, 4, device=dev), torch.empty(n, 961, device=dev)


def torch_copies(i):
    p, q = pin[i % 8]
    p_d.copy_(p, non_blocking=True)
    q_d.copy_(q, non_blocking=True)
    ops.height_scan(p_d, q_d, rays, grid, out=out)
    host_out.copy_(out, non_blocking=True)


def run(fn, steps=200):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e6


for rep in range(2):
    print(f"torch copies + one launch: {run(torch_copies):.1f} us/step")
    for k in (1, 2, 4, 8):
        print(f"rover_height_scan_host, {k} slice(s): {run(lambda i: ops.height_scan_host(*pin[i % 8], rays, grid, host_out, work, n_slices=k)):.1f} us/step")
