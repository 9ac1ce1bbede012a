"""GPU check: every plane-cell scan variant returns bit-identical heights on the bench workload and on ragged
pattern sizes / strided outputs; prints a quick timing per variant (CUDA events, L2 flushed between launches).

    python profiles/cmp_scan_variants.py [path/to/librover_b200.so]
"""
import sys

import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:  # an alternative build of the library (bring-up experiments)
    _lib.LIB_PATH = sys.argv[1]
import bench  # noqa: E402
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
v, f, grid, _ = bench.build_world(None, dev, dev)
vt = torch.from_numpy(v)
gen = torch.Generator().manual_seed(7)
p, q = synthetic.make_poses(4096, gen, vt, bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"])
p, q = p.to(dev), q.to(dev)
rays = ops.RayPattern.grid(dev)
ref = ops.height_scan(p, q, rays, grid, variant=2)
ok = True
for var in (3, 4, 5):
    h = ops.height_scan(p, q, rays, grid, variant=var)
    same = torch.equal(h, ref)
    ok &= same
    print(f"variant {var} == variant 2 on 4096 x 961: {same}", "" if same else (h != ref).sum().item())
# poses near / outside the terrain border (misses, closed far border, deferred rays)
pb = p.clone()
pb[:, 0] = torch.linspace(-3.0, bench.TERRAIN["size_m"] + 3.0, pb.shape[0], device=dev)
refb = ops.height_scan(pb, q, rays, grid, variant=2)
for var in (4, 5):
    h = ops.height_scan(pb, q, rays, grid, variant=var)
    same = torch.equal(h, refb)
    ok &= same
    print(f"variant {var} border sweep: {same}; misses {torch.isinf(refb).sum().item()}")
# ragged pattern sizes and a strided output row (the observation buffer layout)
full = rays.starts.clone()
for n in (1, 2, 63, 64, 65, 255, 256, 257, 511, 960, 961):
    sub = ops.RayPattern(full[:n].contiguous(), dev)
    a = ops.height_scan(p[:300], q[:300], sub, grid, variant=2)
    obs = torch.full((300, n + 7), 123.0, device=dev)
    ops.height_scan(p[:300], q[:300], sub, grid, out=obs[:, 4:4 + n], variant=5)
    same = torch.equal(obs[:, 4:4 + n], a) and bool((obs[:, :4] == 123.0).all()) and bool((obs[:, 4 + n:] == 123.0).all())
    ok &= same
    print(f"n_rays {n}: variant 5 == variant 2 (strided out, padding untouched): {same}")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty(4096, 961, device=dev)
for var in (2, 4, 5):
    ts = []
    for i in range(25):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.height_scan(p, q, rays, grid, out=out, variant=var)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[5:])
    print(f"variant {var}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
