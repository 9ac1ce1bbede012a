"""Bring-up: timeline (SM cycles) of CTA 0 / CTA 100 of the paired scan kernel built with -DROVER_SCAN_DBG=3."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200 import _lib  # noqa: E402

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
for i in range(5):
    flush.fill_(1)
    ops.height_scan(p, q, rays, grid, out=out, variant=5)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = np.zeros((2, 512), dtype=np.uint64)
rc = lib.rover_debug_scan_timeline(buf.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
for b in range(2):
    t = buf[b].astype(np.int64)
    print(f"--- CTA {0 if b == 0 else 100}: frames ready {t[1]}, consumers filled {t[3]}, prologue sync {t[0]}, producer done {t[2]} (cycles)")
    print("producer issue times:", [int(x) for x in t[16:16 + 28]])
    for name, base in (("warp0", 64), ("warp6", 192), ("warp13", 128)):
        rows = [(int(t[base + 4 * j]), int(t[base + 4 * j + 1]), int(t[base + 4 * j + 2])) for j in range(9) if t[base + 4 * j + 2] > 0]
        print(name, "(begin, got data, done):", rows)
        waits = sum(r[1] - r[0] for r in rows)
        work = sum(r[2] - r[1] for r in rows)
        print(f"   total wait {waits} cycles, total work {work} cycles, chunks {len(rows)}")
cta = np.zeros((2, 256), dtype=np.uint64)
assert lib.rover_debug_scan_ctas(cta.ctypes.data_as(C.c_void_p)) == 0
st, en = cta[0, :148].astype(np.int64), cta[1, :148].astype(np.int64)
t0 = st.min()
print(f"CTA starts (ns after first): max {int((st - t0).max())}; CTA ends: min {int((en - t0).min())} median {int(np.median(en - t0))} max {int((en - t0).max())}")
dur = en - st
order = np.argsort(dur)
print("CTA durations ns: min", int(dur.min()), "median", int(np.median(dur)), "max", int(dur.max()), "slowest CTAs", order[-6:].tolist(), "fastest", order[:4].tolist())
