"""Per-kernel CUDA-event timing of the MDP kernels (L2 flushed before every launch). Usage: python profiles/time_mdp.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from isaac_rover_orbit_b200 import ops, synthetic, _lib
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg

dev = torch.device("cuda:0")
N = int(os.environ.get("N", 16384))
size = 200.0
v, f = TR.make_synthetic_terrain(size, 0.2, seed=0)
tables = TR.build_terrain_tables(v, f, N, build_device=dev)
cfg = RoverEnvCfg(num_envs=N)
buf = ops.MdpBuffers.allocate(N, dev)
params = ops.mdp_params(cfg)
th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
gen = torch.Generator().manual_seed(1)
vt = torch.from_numpy(v)
s = synthetic.make_step(N, gen, vt, size, 0.2).to(dev)
buf.env_origins.copy_(s.root_pos_w); buf.time_left.fill_(150.0)
buf.pos_cmd_w.copy_(s.root_pos_w + torch.tensor([9.0, 0, 0], device=dev))
obs = torch.zeros(N, 968, device=dev)[:, :965]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()

def timeit(fn, reps=50, flush=True):
    ts = []
    for i in range(reps + 5):
        if flush: flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

pre = lambda: ops.mdp_pre_step(buf, params, s.actions, s.force_matrix_w)
post = lambda: ops.mdp_post_step(buf, params, th, s.root_pos_w, s.root_quat_w, s.spawn_perm, s.yaw_u, s.heading_u, s.theta_u, obs)
def post_noreset():
    buf.reset_flags.zero_(); buf.block_reset_counts.zero_(); torch.cuda.synchronize()
for flush in (True, False):
    print("flush", flush, "empty-ish kernel (torch add)", timeit(lambda: torch.add(buf.err_pos, 1.0, out=buf.err_pos), flush=flush))
    print("flush", flush, "pre_step us", timeit(pre, flush=flush))
    pre(); torch.cuda.synchronize(); print("  resets", int(buf.reset_flags.sum()))
    print("flush", flush, "post_step us", timeit(lambda: (pre(), post())[1], flush=flush), "(pre+post)")
    for ph, name in ((_lib.PHASE_METRICS | _lib.PHASE_COMMAND | _lib.PHASE_OBS, "metrics+command+obs"), (_lib.PHASE_OBS, "obs only"), (_lib.PHASE_ALL, "all, with resets")):
        pre(); torch.cuda.synchronize()
        print("flush", flush, "post phases", name, timeit(lambda: ops.mdp_post_step(buf, params, th, s.root_pos_w, s.root_quat_w, s.spawn_perm, s.yaw_u, s.heading_u, s.theta_u, obs, phases=ph), flush=flush))

# ---- how much of the cold launch is instruction fetch?  (ncu: 44 % of the post-step's stall samples are `no_inst`.)
# After the L2 flush, a 64-env launch of the same kernels on separate buffers brings their code back into L2 / the SM
# instruction caches; the timed 16384-env launch then still sees cold DATA.
small = 64
tb2 = TR.build_terrain_tables(v, f, small, build_device=dev)
buf2 = ops.MdpBuffers.allocate(small, dev)
th2 = ops.TerrainTablesHandle(tb2.heightmap, tb2.safe_mask, tb2.offset_xy, tb2.spawn_table, tb2.resolution, dev)
s2 = synthetic.make_step(small, gen, vt, size, 0.2).to(dev)
buf2.time_left.fill_(150.0)


def timeit_warm_code(fn, reps=50):
    ts = []
    for i in range(reps + 5):
        flush_buf.fill_(1)
        ops.mdp_pre_step(buf2, params, s2.actions, s2.force_matrix_w)
        ops.mdp_post_step(buf2, params, th2, s2.root_pos_w, s2.root_quat_w, s2.spawn_perm, s2.yaw_u, s2.heading_u, s2.theta_u)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


print("code warm, data cold: pre_step us", timeit_warm_code(pre))
pre(); torch.cuda.synchronize()
print("code warm, data cold: post_step (all phases, with resets) us", timeit_warm_code(post))
