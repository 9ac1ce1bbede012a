"""Bring-up: per-block globaltimer stamps of the single-launch MDP step built with -DROVER_MDP_DBG=1.
    nvcc ... -DROVER_MDP_DBG=1 -o profiles/build/librover_mdpdbg.so ...; ROVER_B200_LIB=profiles/build/librover_mdpdbg.so python profiles/mdp_timeline.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from isaac_rover_orbit_b200 import _lib, ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402
from isaac_rover_orbit_b200.policy import alloc_obs  # noqa: E402

dev = torch.device("cuda:0")
n = 16384
v, f, grid, tables = bench.build_world(n, dev, dev)
vt = torch.from_numpy(v)
cfg = RoverEnvCfg(num_envs=n)
params = ops.mdp_params(cfg)
gen = torch.Generator().manual_seed(3)
st = synthetic.make_step(n, gen, vt, 200.0, 0.2, cfg.num_contact_bodies, cfg.target_rounds).to(dev)
th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
buf = ops.MdpBuffers.allocate(n, dev)
buf.env_origins.copy_(st.root_pos_w)
buf.time_left.fill_(150.0)
buf.pos_cmd_w.copy_(st.root_pos_w + torch.tensor([9.0, 0.0, 0.0], device=dev))
buf.pos_cmd_b.copy_(torch.tensor([9.0, 0.0, 0.0], device=dev).expand(n, 3))
rng = ops.ResetRng(5, dev)
obs = alloc_obs(n, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = C.CDLL(os.environ["ROVER_B200_LIB"])
for rep in range(4):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.mdp_step(buf, params, th, st.actions, st.force_matrix_w, st.root_pos_w.clone(), st.root_quat_w.clone(), obs=obs, rng=rng)
    b.record()
    torch.cuda.synchronize()
    t = np.zeros((8, 1024), dtype=np.uint64)
    assert lib.rover_debug_mdp_timeline(t.ctypes.data_as(C.c_void_p)) == 0
    nb = (n + 63) // 64
    t = t[:, :nb].astype(np.int64)
    t0 = t[0].min()
    names = ["block start", "prefetches issued", "pre-step done", "post work done", "partials stored+fenced"]
    print(f"rep {rep}: event time {a.elapsed_time(b) * 1e3:.1f} us, resets {int(buf.stats[13])}")
    for k, nm in enumerate(names):
        d = t[k] - t0
        print(f"  {nm:26s} min {d.min():6d}  median {int(np.median(d)):6d}  p90 {int(np.percentile(d, 90)):6d}  max {d.max():6d} ns")
    last = t[6].max() - t0
    print(f"  launch-wide reduction done {last} ns;   env warps' lifetime (start -> post work done) median {int(np.median(t[3] - t[0]))} max {int((t[3] - t[0]).max())} ns")
    buf.stats.zero_()
