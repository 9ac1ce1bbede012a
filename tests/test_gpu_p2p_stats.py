"""GPU: episode statistics through the P2P mailbox (dist.P2PStats).  World size 1 here (the mailbox of the rank itself);
``tests/p2p_stats_worker.py`` is the world-size-2 check, launched by this test when two GPUs are visible and by
``gpurun --gpus 2`` during development."""
import os
import subprocess
import sys

import pytest
import torch

from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.dist import P2PStats, episode_log

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_p2p_stats_world1_equals_accumulator(cuda_device):
    n = 512
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=3)
    tables = TR.build_terrain_tables(v, f, n)
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, cuda_device)
    buf = ops.MdpBuffers.allocate(n, cuda_device)
    buf.time_left.fill_(150.0)
    p2p = P2PStats(cuda_device, rank=0, world=1)
    gen = torch.Generator().manual_seed(3)
    vt = torch.from_numpy(v)
    total = torch.zeros(16, dtype=torch.float64)
    for step in range(5):
        st = synthetic.make_step(n, gen, vt, 48.0, 0.2, margin=4.0).to(cuda_device)
        buf.stats.zero_()
        ops.mdp_pre_step(buf, params, st.actions, st.force_matrix_w)
        ops.mdp_post_step(buf, params, th, st.root_pos_w, st.root_quat_w, st.spawn_perm, st.yaw_u, st.heading_u,
                          st.theta_u, None, xchg=p2p)
        total += buf.stats.double().cpu()
        if step == 2:
            first = p2p.interval().cpu()
            assert torch.equal(first, total)  # fp64 running totals of fp32 per-launch statistics: exact
    torch.cuda.synchronize()
    assert torch.equal(p2p.read().cpu(), total) and total[13] > 0
    log = episode_log(p2p.interval())
    assert log["num_resets"] == int(total[13] - first[13])
    # the single-launch step with in-kernel variates publishes one launch behind (at the START of the next launch)
    rng = ops.ResetRng(5, cuda_device)
    for step in range(3):
        st = synthetic.make_step(n, gen, vt, 48.0, 0.2, margin=4.0).to(cuda_device)
        buf.stats.zero_()
        before = total.clone()
        ops.mdp_step(buf, params, th, st.actions, st.force_matrix_w, st.root_pos_w, st.root_quat_w, rng=rng, xchg=p2p)
        total += buf.stats.double().cpu()
        torch.cuda.synchronize()
        seen = p2p.read().cpu()  # (ROVER_MDP_SPLIT=0, the A/B switch of the split CTA, publishes at the end of the launch)
        assert torch.equal(seen, before) or torch.equal(seen, total), "mailbox = totals of the launches before this one"
        assert torch.equal(p2p.local_totals().cpu(), total)
    assert torch.equal(p2p.totals().cpu(), total)  # flush + read
    p2p.close()


def _run_world2(env_extra: dict, port: int):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "p2p_stats_worker.py")], capture_output=True, text=True, timeout=600,
                         env=env)
    assert out.returncode == 0 and "P2P_STATS_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_p2p_stats_world2_same_device(cuda_device):
    """World size 2 on ONE GPU: two processes share cuda:0, map each other's mailbox through CUDA IPC and publish with
    the same stores / fences / sequence numbers as over NVLink -- the protocol test the one-GPU driver box can run."""
    _run_world2({"ROVER_P2P_SAME_DEVICE": "1"}, 29534)


def test_p2p_stats_world2(cuda_device):
    """World size 2 over NVLink peer mappings when the box has two GPUs; on a one-GPU box the same-device run above is
    the world-2 coverage, so this case passes through it instead of skipping."""
    if torch.cuda.device_count() < 2:
        _run_world2({"ROVER_P2P_SAME_DEVICE": "1", "ROVER_P2P_EXTRA_STEPS": "7"}, 29535)
        return
    _run_world2({}, 29533)
