"""torchrun worker (one rank per GPU): every rank steps its own env shard with the statistics published through the
P2P mailbox; the global totals read from the local mailbox must equal an NCCL all-reduce of the ranks' totals."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_rover_orbit_b200 import ops, synthetic  # noqa: E402
from isaac_rover_orbit_b200 import terrain as TR  # noqa: E402
from isaac_rover_orbit_b200.config import RoverEnvCfg  # noqa: E402
from isaac_rover_orbit_b200.dist import P2PStats  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # ROVER_P2P_SAME_DEVICE=1: every rank is a separate process on cuda:0 -- the mailboxes are still mapped through CUDA
    # IPC and written by "peer" stores, so the sequence-lock protocol runs on a one-GPU box (NCCL refuses two ranks on
    # one device: the handles and the reference sum travel over gloo there)
    same_device = os.environ.get("ROVER_P2P_SAME_DEVICE", "0") == "1"
    local = 0 if same_device else local
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    n = 1024
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=3)
    tables = TR.build_terrain_tables(v, f, n)
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    buf = ops.MdpBuffers.allocate(n, dev)
    buf.time_left.fill_(150.0)
    p2p = P2PStats(dev)
    gen = torch.Generator().manual_seed(100 + rank)
    vt = torch.from_numpy(v)
    mine = torch.zeros(16, dtype=torch.float64, device=dev)
    steps = 20 + (3 + int(os.environ.get("ROVER_P2P_EXTRA_STEPS", "0"))) * rank  # ranks publish different numbers of steps
    for _ in range(steps):
        st = synthetic.make_step(n, gen, vt, 48.0, 0.2, margin=4.0).to(dev)
        buf.stats.zero_()
        ops.mdp_pre_step(buf, params, st.actions, st.force_matrix_w)
        ops.mdp_post_step(buf, params, th, st.root_pos_w, st.root_quat_w, st.spawn_perm, st.yaw_u, st.heading_u,
                          st.theta_u, None, xchg=p2p)
        mine += buf.stats.double()
        p2p.read()  # concurrent readers while peers write: must never hang or fault
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    want = mine.cpu() if same_device else mine.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    got = p2p.read().clone()
    torch.cuda.synchronize()
    ok = bool(torch.equal(got.cpu(), want.cpu())) and float(got[13]) > 0
    # ---- second phase: the single-launch step with in-kernel variates publishes at the START of a launch the totals of
    #      the launches before it (CTA 0's kinematics warps), so the mailboxes run one launch behind until a flush;
    #      mixed here with an end-of-launch publisher in between
    rng = ops.ResetRng(7 + rank, dev)
    steps2 = 9 + 2 * rank
    for k in range(steps2):
        st = synthetic.make_step(n, gen, vt, 48.0, 0.2, margin=4.0).to(dev)
        buf.stats.zero_()
        if k == 4:
            ops.mdp_pre_step(buf, params, st.actions, st.force_matrix_w)
            ops.mdp_post_step(buf, params, th, st.root_pos_w, st.root_quat_w, rng=rng, xchg=p2p)
        else:
            ops.mdp_step(buf, params, th, st.actions, st.force_matrix_w, st.root_pos_w, st.root_quat_w, rng=rng, xchg=p2p)
        mine += buf.stats.double()
        p2p.read()
    torch.cuda.synchronize()
    assert torch.equal(p2p.local_totals(), mine), "the rank's own running totals"
    want2 = mine.cpu() if same_device else mine.clone()
    dist.all_reduce(want2, op=dist.ReduceOp.SUM)
    got2 = p2p.totals().clone()  # flush + barrier + read
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(got2.cpu(), want2.cpu())) and float(got2[13]) > float(got[13])
    got, want = got2, want2
    flag = torch.tensor([1.0 if ok else 0.0], device="cpu" if same_device else dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("P2P_STATS_OK" if flag.item() == 1.0 else f"P2P_STATS_MISMATCH got {got.tolist()} want {want.tolist()}", flush=True)
    dist.barrier()
    p2p.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
