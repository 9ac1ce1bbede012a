import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from isaac_rover_orbit_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import bench
from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs
dev = torch.device("cuda:0")
n = 16384
v, f, grid, tables = bench.build_world(n, dev, dev)
rays = ops.RayPattern.grid(dev)
gen = torch.Generator().manual_seed(2)
p, q = (t.to(dev) for t in synthetic.make_poses(n, gen, torch.from_numpy(v), 200.0, 0.2))
net = GaussianNeuralNetwork(device=dev)
obs = alloc_obs(n, dev); obs[:, :4] = 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = C.CDLL(_lib.LIB_PATH)
for mode in ("cold obs (flushed)", "obs just written by the scan"):
    for _ in range(3):
        flush.fill_(1)
        ops.height_scan(p, q, rays, grid, out=obs[:, 4:])
        if mode.startswith("cold"): flush.fill_(1)
        net.compute({"states": obs})
    torch.cuda.synchronize()
    t = np.zeros(2048, dtype=np.int64)
    assert lib.rover_debug_policy_timeline(t.ctypes.data_as(C.c_void_p)) == 0
    names = ["begin", "D0 ready", "epi0+MMA1 issued", "MMA1 done", "epi1+MMA2 issued", "MMA2 done", "epi2+MMA3a issued", "MMA3a done",
             "W3hi+MMA3b done", "epi3+MMA4 issued", "MMA4 done", "epi4+MMA5 issued", "MMA5 done", "epi5 done"]
    row = t[1024:1024 + 14]
    print(mode, ": producer issue", t[0:31:6].tolist(), "MMA issued", t[512:543:6].tolist())
    c = np.arange(8, 30)
    print("   steady state per chunk (cycles): TMA issue -> data landed", int((t[128 + c] - t[c]).mean()), "| landed -> A slot free", int((t[256 + c] - t[128 + c]).mean()),
          "| convert", int((t[384 + c] - t[256 + c]).mean()), "| converted -> MMA issued", int((t[512 + c] - t[384 + c]).mean()),
          "| chunk period", int(np.diff(t[512 + c]).mean()))
    print("   layer group tile 0:", {nm: int(v) for nm, v in zip(names, row)})
    cta = np.zeros((2, 256), dtype=np.uint64)
    assert lib.rover_debug_policy_ctas(cta.ctypes.data_as(C.c_void_p)) == 0
    st, en = cta[0, :148].astype(np.int64), cta[1, :148].astype(np.int64)
    print("   kernel span", int(en.max() - st.min()), "ns; CTA durations median", int(np.median(en - st)))
