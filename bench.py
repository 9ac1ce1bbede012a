#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native AAURoverEnv-v0 non-physics MDP hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), host cores

Headline metric (BASELINE.json): height-scan rays/s on cfg-2 -- 4096 envs x 961 rays per GPU on the synthetic
200 m x 200 m / 2,000,000-triangle terrain.  A "step" is one height-scan pass over one batch of synthetic poses.
The same run also measures the fused non-physics step (cfg-3 at N=1, cfg-5 = 8192 envs/GPU + the NCCL
episode-statistics all-reduce under torchrun) and reports it under "extra".

Timing: CUDA events on the launching stream around every timed step, L2 flushed (256 MiB write) between steps and
excluded from the timed region, max over ranks.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TERRAIN = dict(size_m=200.0, grid_res=0.2, seed=0)
SCAN_ENVS_PER_GPU = 4096  # cfg-2
STEP_ENVS_1GPU = 16384  # cfg-3
STEP_ENVS_PER_GPU_MULTI = 8192  # cfg-5
N_RAYS = 961
TERRAIN_BYTES = 36.0e6  # compulsory mesh footprint (12.0 MB vertices + 24.0 MB indices)
POSE_SETS = 8


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel: str):
    """DRAM bytes per launch from the committed ncu capture (profiles/ncu_summary.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def recorded_metric(kernel: str, key: str):
    """One number of the committed ncu capture of `kernel` (profiles/ncu_summary.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get(kernel, {}).get(key)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs (pynvml, ~20 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("clock sampling unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def build_world(need_tables_for: int | None, build_device, device):
    """Synthetic terrain + scan grid (+ resample tables for `need_tables_for` envs)."""
    from isaac_rover_orbit_b200 import ops
    from isaac_rover_orbit_b200 import terrain as TR

    t0 = time.time()
    v, f = TR.make_synthetic_terrain(**TERRAIN)
    grid = ops.ScanGridHandle.from_mesh(v, f, device)
    log(f"terrain {v.shape[0]} verts / {f.shape[0]} tris, home grid {grid.grid.nbytes() / 1e6:.1f} MB, plane cells "
        f"{grid.cells.nbytes() / 1e6:.1f} MB ({grid.cells.n_general} general cells) ({time.time() - t0:.1f}s)")
    tables = None
    if need_tables_for:
        t0 = time.time()
        tables = TR.build_terrain_tables(v, f, need_tables_for, build_device=build_device)
        log(f"terrain tables {tuple(tables.heightmap.shape)} ({time.time() - t0:.1f}s)")
    return v, f, grid, tables


def time_steps(fn, steps, warmup, flush, stream):
    """fn(i) enqueues one step; returns per-step milliseconds (CUDA events, flush excluded)."""
    for i in range(warmup):
        flush()
        fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush()
        ev[i][0].record(stream)
        fn(warmup + i)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product path has no CPU fallback); "
                           "use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from isaac_rover_orbit_b200 import ops, synthetic
    from isaac_rover_orbit_b200.config import RoverEnvCfg
    from isaac_rover_orbit_b200.dist import EpisodeStats

    n_scan = SCAN_ENVS_PER_GPU
    n_step = STEP_ENVS_1GPU if world == 1 else STEP_ENVS_PER_GPU_MULTI
    v, f, grid, tables = build_world(n_step, "cpu" if args.init_on_cpu else dev, dev)
    vt = torch.from_numpy(v)
    cfg = RoverEnvCfg(num_envs=n_step)
    rays = ops.RayPattern.grid(dev)
    stream = torch.cuda.current_stream(dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    # ------------------------------------------------------------------ headline: height scan (cfg-2 per GPU)
    gen = torch.Generator().manual_seed(1234 + rank)
    poses = [synthetic.make_poses(n_scan, gen, vt, TERRAIN["size_m"], TERRAIN["grid_res"]) for _ in range(POSE_SETS)]
    poses_d = [(p.to(dev), q.to(dev)) for p, q in poses]
    out = torch.empty(n_scan, N_RAYS, device=dev)

    def scan_step(i):
        p, q = poses_d[i % POSE_SETS]
        ops.height_scan(p, q, rays, grid, out=out, variant=args.variant)

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        ms = time_steps(scan_step, args.steps, max(args.warmup, 3), flush, stream)  # never fewer than 3 untimed steps
        # keep the GPU busy long enough for >= a few clock samples
        t_end = time.time() + 0.3
        while time.time() < t_end:
            scan_step(0)
        torch.cuda.synchronize()
    total_ms = torch.tensor([ms.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    rays_total = n_scan * N_RAYS * args.steps * world
    value = rays_total / (total_ms * 1e-3)
    t_launch = float(ms.mean()) * 1e-3
    alg_bytes = 4.0 * n_scan * N_RAYS + 28.0 * n_scan + TERRAIN_BYTES
    peak, peak_src = measured_peak()
    achieved = alg_bytes / t_launch / 1e9
    kname = {0: "height_scan_direct_kernel", 1: "height_scan_staged_kernel", 2: "height_scan_cells_kernel",
             3: "height_scan_cells_tma_kernel", 4: "height_scan_pipelined_kernel",
             5: "height_scan_paired_kernel"}[args.variant]

    # ------------------------------------------------------------------ e2e: host buffers through the public API
    pin = [(p.pin_memory(), q.pin_memory()) for p, q in poses]
    host_out = torch.empty(n_scan, N_RAYS).pin_memory()
    p_d, q_d = torch.empty(n_scan, 3, device=dev), torch.empty(n_scan, 4, device=dev)

    def e2e_step(i):
        p, q = pin[i % POSE_SETS]
        p_d.copy_(p, non_blocking=True)
        q_d.copy_(q, non_blocking=True)
        ops.height_scan(p_d, q_d, rays, grid, out=out, variant=args.variant)
        host_out.copy_(out, non_blocking=True)

    e2e_steps = max(min(args.steps, 200), 3)
    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
        torch.cuda.synchronize()  # the caller reads the heights of step i before issuing step i+1
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_scan * N_RAYS * e2e_steps * world / float(e2e_t.item())

    # the same calls double-buffered on two streams: step i+1's H2D and kernel overlap step i's D2H; the host waits for
    # step i-2 before it reuses that step's buffers (reported next to the synchronous figure, which stays the headline)
    e2e_pipe = None
    try:
        streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        dbuf = [(torch.empty(n_scan, 3, device=dev), torch.empty(n_scan, 4, device=dev),
                 torch.empty(n_scan, N_RAYS, device=dev), torch.empty(n_scan, N_RAYS).pin_memory()) for _ in range(2)]
        done = [None, None]

        def piped(i):
            k = i & 1
            if done[k] is not None:
                done[k].synchronize()  # the heights of step i-2 are on the host: the caller may consume them now
            pp, qq, oo, hh = dbuf[k]
            with torch.cuda.stream(streams[k]):
                p, q = pin[i % POSE_SETS]
                pp.copy_(p, non_blocking=True)
                qq.copy_(q, non_blocking=True)
                ops.height_scan(pp, qq, rays, grid, out=oo, variant=args.variant)
                hh.copy_(oo, non_blocking=True)
                done[k] = torch.cuda.Event()
                done[k].record(streams[k])

        for i in range(4):
            piped(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            piped(i)
        torch.cuda.synchronize()
        tp_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp_, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(dbuf[(e2e_steps - 1) & 1][3], dbuf[(e2e_steps - 1) & 1][2].cpu()))
        e2e_pipe = {"value": n_scan * N_RAYS * e2e_steps * world / float(tp_.item()), "unit": "rays/s",
                    "how": "two streams, double-buffered pinned host buffers; every step still moves its poses in and "
                           "its heights out", "host_copy_equals_device": ok}
    except Exception as e:  # the synchronous figure above is the contract; this one is informative
        e2e_pipe = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ extra: fused non-physics step
    extra = {}
    try:
        buf = ops.MdpBuffers.allocate(n_step, dev)
        params = ops.mdp_params(cfg)
        th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                     tables.resolution, dev)
        gen2 = torch.Generator().manual_seed(99 + rank)
        sets = [synthetic.make_step(n_step, gen2, vt, TERRAIN["size_m"], TERRAIN["grid_res"],
                                    cfg.num_contact_bodies, cfg.target_rounds).to(dev) for _ in range(4)]
        pc, hc, ep = synthetic.init_commands(n_step, gen2, sets[0].root_pos_w.cpu())
        buf.pos_cmd_w.copy_(pc)
        buf.heading_cmd_w.copy_(hc)
        buf.episode_length_buf.copy_(ep)
        buf.env_origins.copy_(sets[0].root_pos_w)
        buf.time_left.fill_(150.0)
        obs = torch.zeros(n_step, 4 + N_RAYS, device=dev)
        stats = EpisodeStats(buf, world)
        # several GPUs: the episode statistics travel as peer stores issued by the post-step kernel itself
        # (dist.P2PStats); ROVER_STATS=nccl keeps the all-reduce per step (EpisodeStats)
        p2p = None
        if world > 1 and os.environ.get("ROVER_STATS", "p2p") != "nccl":
            from isaac_rover_orbit_b200.dist import P2PStats

            p2p = P2PStats(dev, rank, world)

        # synthetic "physics": every step the rover sits at its env origin + a bounded offset, so that the
        # far/success terminations stay rare and resets come from contacts (5 %) and time-outs
        drift = [(torch.rand(n_step, 3, generator=gen2) * torch.tensor([4.0, 4.0, 0.0]) - torch.tensor([2.0, 2.0, 0.0])
                  ).to(dev) for _ in range(4)]
        buf.pos_cmd_w.copy_(sets[0].root_pos_w + torch.tensor([9.0, 0.0, 0.0], device=dev))

        def mdp(buf, params, th, actions, s, obs_, xchg=None):
            """pre-step + post-step: two launches (the faster arrangement, profiles/r01_mdp_v1.md); --single-launch-mdp
            runs both in one launch (rover_mdp_step, reset rank by decoupled look-back)"""
            if not args.single_launch_mdp:
                ops.mdp_pre_step(buf, params, actions, s.force_matrix_w)
                ops.mdp_post_step(buf, params, th, s.root_pos_w, s.root_quat_w, s.spawn_perm, s.yaw_u, s.heading_u,
                                  s.theta_u, obs_, xchg=xchg)
            else:
                ops.mdp_step(buf, params, th, actions, s.force_matrix_w, s.root_pos_w, s.root_quat_w, s.spawn_perm,
                             s.yaw_u, s.heading_u, s.theta_u, obs_, xchg=xchg)

        def full_step(i):
            s = sets[i % 4]
            torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)  # stand-in for PhysX, not one of our kernels
            mdp(buf, params, th, s.actions, s, obs)
            ops.height_scan(s.root_pos_w, s.root_quat_w, rays, grid, out=obs[:, 4:], variant=args.variant)
            if world > 1:
                stats.all_reduce_async()

        def mdp_only(i):
            s = sets[i % 4]
            torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)
            mdp(buf, params, th, s.actions, s, obs)

        ksteps = max(min(args.steps, 200), 3)
        # the step is launch-bound on the host side (3 ctypes launches + 1 torch op ~ 40 us of Python per step):
        # capture each input set's step once in a CUDA graph and replay it (one cudaGraphLaunch per step)
        def graphed(fn):
            from isaac_rover_orbit_b200.trainer import capture_steps

            return fn if args.no_graph else capture_steps(fn, n_variants=4, warmup=1)

        step_body = full_step
        if world > 1:
            def step_body(i):  # noqa: F811  (kernels only; an NCCL all-reduce, if used, is issued after the replay)
                s = sets[i % 4]
                torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)
                mdp(buf, params, th, s.actions, s, obs, xchg=p2p)
                ops.height_scan(s.root_pos_w, s.root_quat_w, rays, grid, out=obs[:, 4:], variant=args.variant)
        g_full = graphed(step_body)
        g_mdp = graphed(mdp_only)

        def run_full(i):
            g_full(i)
            if world > 1 and p2p is None:
                stats.all_reduce_async()

        ms_full = time_steps(run_full, ksteps, 3, flush, stream)
        ms_mdp = time_steps(g_mdp, ksteps, 3, flush, stream)
        buf.stats.zero_()
        mdp_only(0)
        torch.cuda.synchronize()
        resets_per_step = float(buf.stats[13].item())
        tf = torch.tensor([ms_full.sum(), ms_mdp.sum()], dtype=torch.float64, device=dev)
        stats_check = None
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            if p2p is not None:
                # one-off check outside the timed region: the mailbox totals equal an NCCL all-reduce of the ranks' totals
                dist.barrier()
                torch.cuda.synchronize()
                mine = p2p._cumulative.clone()
                dist.all_reduce(mine, op=dist.ReduceOp.SUM)
                got = p2p.read().clone()
                torch.cuda.synchronize()
                stats_check = {"exchange": "p2p mailbox stores from mdp_post_step (no collective launch on the step path)",
                               "equals_nccl_all_reduce": bool(torch.equal(got, mine)),
                               "global_resets": float(got[13].item())}
            else:
                stats_check = {"exchange": "nccl all_reduce per step"}
        full_bytes = n_step * (414.0 + 4.0 * N_RAYS) + TERRAIN_BYTES
        extra = {
            "fused_step": {
                "workload": f"cfg-{'3' if world == 1 else '5'}: {n_step} envs/GPU, pre_step + post_step + height scan"
                            + ("" if world == 1 else " + episode statistics through P2P mailboxes (peer stores from the "
                               "post-step kernel)" if p2p is not None else " + NCCL episode-stat all-reduce"),
                "env_steps_per_s": n_step * world * ksteps / (float(tf[0]) * 1e-3),
                "ms_per_step": float(tf[0]) / ksteps,
                "gpu_launches_per_step": 2 if args.single_launch_mdp else 3, "cuda_graph": not args.no_graph,
                "roofline_frac_hbm": full_bytes / (float(tf[0]) / ksteps * 1e-3) / 1e9 / peak,
                "resets_in_one_step": resets_per_step,
                "episode_stats": stats_check,
            },
            "mdp_only": {
                "env_steps_per_s": n_step * world * ksteps / (float(tf[1]) * 1e-3),
                "ms_per_step": float(tf[1]) / ksteps,
                "roofline_frac_hbm": n_step * 414.0 / (float(tf[1]) / ksteps * 1e-3) / 1e9 / peak,
                "note": "414 B/env: launch-latency bound at this N (SURVEY.md 8d); pre + post launch (one launch with --single-launch-mdp, measured 1.4 us slower)",
            },
        }
    except Exception as e:  # the headline number must survive a failure of the extra measurements
        extra = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ extra: policy forward (cfg-4)
    try:
        from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16

        n_pol = 65536
        net = GaussianNeuralNetwork(device=dev)
        gen3 = torch.Generator().manual_seed(7 + rank)
        net.load_state_dict({k: (torch.randn(v.shape, generator=gen3) * (0.05 if v.dim() == 2 else 0.01))
                             for k, v in net.state_dict().items()})
        pol_obs = alloc_obs(n_pol, dev)
        pol_obs.copy_(torch.randn(n_pol, 965, device=dev) * 0.3)
        ksteps = max(min(args.steps, 100), 3)
        ms_pol = time_steps(lambda i: net.compute({"states": pol_obs}), ksteps, 3, flush, stream)
        tp = torch.tensor([ms_pol.sum()], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        t_pol = float(tp[0]) / ksteps * 1e-3
        # the same forward on the bf16 observation mirror (what the closed loop uses: written by the height scan)
        pol_obs_bf = alloc_obs_bf16(n_pol, dev)
        pol_obs_bf.copy_(pol_obs)
        ms_bf = time_steps(lambda i: net.compute({"states": pol_obs_bf}), ksteps, 3, flush, stream)
        tb = torch.tensor([ms_bf.sum()], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        t_bf = float(tb[0]) / ksteps * 1e-3
        same = bool(torch.equal(net.compute({"states": pol_obs})[0], net.compute({"states": pol_obs_bf})[0]))
        bf16_pol = {"env_forwards_per_s": n_pol * world / t_bf, "us_per_launch": t_bf * 1e6,
                    "bit_identical_to_fp32_path": same,
                    "tensor_pipe_active_pct_ncu": recorded_metric("policy_forward_ws_kernel<bf16 observation>",
                                                                  "tensor_pipe_active_pct"),
                    "roofline_frac_hbm": n_pol * (964 * 2 + 8) / t_bf / 1e9 / measured_peak()[0]}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = float(peaks.get("bf16_tflops", 1590.0))
        extra["policy_forward"] = {
            "workload": f"cfg-4: skrl Gaussian policy (961->80->60 (+4) ->256->160->128->2), {n_pol} envs/GPU, bf16 operands "
                        "/ fp32 accumulate on tcgen05, random-init weights",
            "env_forwards_per_s": n_pol * world / t_pol, "us_per_launch": t_pol * 1e6,
            "tflops": n_pol * 319520 / t_pol / 1e12,
            "roofline_frac_tensor": n_pol * 319520 / t_pol / 1e12 / tf_peak, "tensor_peak_tflops": tf_peak,
            "roofline_frac_hbm": n_pol * (965 * 4 + 8) / t_pol / 1e9 / peak,
            "tensor_pipe_active_pct_ncu": recorded_metric("policy_forward_ws_kernel", "tensor_pipe_active_pct"),
            "kernel": "policy_forward_ws_kernel" if os.environ.get("ROVER_POLICY_KERNEL", "ws")[:2] != "v1"
                      else "policy_forward_kernel",
            "bf16_observation": bf16_pol,
            "note": "standalone forward reads 3860 B/env of fp32 observations: 83 FLOP/B < ridge, HBM-bound (SURVEY 8d)",
        }
    except Exception as e:
        extra["policy_forward"] = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ extra: closed non-physics loop with the policy
    # obs -> policy -> action -> (stand-in physics) -> pre_step -> post_step -> height scan -> obs, one CUDA graph per step
    try:
        loop_obs = alloc_obs(n_step, dev)
        act_buf = torch.zeros(n_step, 2, device=dev)
        eps_sets = [torch.randn(n_step, 2, generator=gen3).to(dev) for _ in range(4)]

        def closed_step(i):
            s = sets[i % 4]
            torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)  # stand-in for PhysX
            mdp(buf, params, th, act_buf, s, loop_obs)
            ops.height_scan(s.root_pos_w, s.root_quat_w, rays, grid, out=loop_obs[:, 4:], variant=args.variant)
            actions, _, _ = net.act({"states": loop_obs}, eps=eps_sets[i % 4])
            act_buf.copy_(actions)

        loop_obs_bf = alloc_obs_bf16(n_step, dev)

        def closed_step_bf16(i):  # the scan also writes the bf16 observation; the policy reads only that
            s = sets[i % 4]
            torch.add(buf.env_origins, drift[i % 4], out=s.root_pos_w)
            mdp(buf, params, th, act_buf, s, loop_obs)
            ops.height_scan_obs(s.root_pos_w, s.root_quat_w, rays, grid, loop_obs, loop_obs_bf)
            actions, _, _ = net.act({"states": loop_obs_bf}, eps=eps_sets[i % 4])
            act_buf.copy_(actions)

        ksteps = max(min(args.steps, 200), 3)
        ms_loop = time_steps(graphed(closed_step), ksteps, 3, flush, stream)
        act_buf.zero_()
        ms_loop_bf = time_steps(graphed(closed_step_bf16), ksteps, 3, flush, stream)
        tl = torch.tensor([ms_loop.sum(), ms_loop_bf.sum()], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        extra["closed_loop_step"] = {
            "workload": f"{n_step} envs/GPU: pre_step + post_step + height scan + policy forward (tcgen05) + Gaussian act, "
                        "actions fed back to the next step; physics replaced by a synthetic pose update",
            "env_steps_per_s": n_step * world * ksteps / (float(tl[0]) * 1e-3),
            "ms_per_step": float(tl[0]) / ksteps, "gpu_launches_per_step": 4 if args.single_launch_mdp else 5,
            "cuda_graph": not args.no_graph,
            "finite_actions": bool(torch.isfinite(act_buf).all().item()),
            "bf16_observation": {"env_steps_per_s": n_step * world * ksteps / (float(tl[1]) * 1e-3),
                                 "ms_per_step": float(tl[1]) / ksteps,
                                 "how": "rover_height_scan_obs writes the bf16 mirror, rover_policy_forward_bf16 reads it"},
        }
    except Exception as e:
        extra["closed_loop_step"] = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(v, f, steps=None, warmup=1, n_envs=SCAN_ENVS_PER_GPU, budget_s=12.0)

    line = None
    if rank == 0:
        line = {
            "metric": "height_scan_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg-2 height-scan raycast only: 4096 envs x 961 rays per GPU, synthetic 200 m x "
                                   "200 m Mars-like terrain, 2,000,000 triangles",
                       "envs_per_gpu": n_scan, "rays_per_env": N_RAYS, "kernel_variant": args.variant,
                       "l2": "flushed between timed steps (256 MiB write, outside the timed region)",
                       "pose_sets": POSE_SETS, "untimed_warmup_steps": max(args.warmup, 3)},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_scan * 28,
                    "d2h_bytes_per_step": n_scan * N_RAYS * 4, "steps": e2e_steps,
                    "how": "pinned host poses -> H2D -> rover_height_scan -> D2H heights, synchronised per step",
                    "pipelined": e2e_pipe},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": recorded_traffic(kname),
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_us": t_launch * 1e6},
            "cpu_baseline": cpu,
            "extra": extra,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def cpu_reference(v, f, steps, warmup, n_envs, budget_s=None):
    """The reference's CPU path for the height scan, restated (oracle port): torch pose->ray transform
    (ORBIT RayCaster) + BVH/watertight closest hit (warp mesh_query_ray) + height_scan_rover, all host threads.
    ``budget_s``: size the sample from a probe so that the timed part takes about that long -- ``steps`` is then
    chosen here (cpu_baseline leg) or, with ``steps`` fixed by the caller, the envs per step are (reference arm)."""
    from isaac_rover_orbit_b200 import synthetic
    from oracle import raycast as oracle_raycast
    from oracle import step as OS

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.time()
    mesh = oracle_raycast.Mesh(v, f)
    log(f"oracle BVH over {len(f)} triangles ({time.time() - t0:.1f}s)")
    gen = torch.Generator().manual_seed(4321)
    vt = torch.from_numpy(v)
    if budget_s is not None:
        probe = synthetic.make_poses(256, gen, vt, TERRAIN["size_m"], TERRAIN["grid_res"])
        OS.height_scan(*probe, mesh)
        t0 = time.perf_counter()
        OS.height_scan(*probe, mesh)
        t_env = (time.perf_counter() - t0) / 256  # seconds per env (961 rays)
        if steps is None:   # cpu_baseline leg: whole cfg-2 batches, as many as fit the budget
            steps = int(min(max(round(budget_s / (t_env * n_envs)), 3), 200))
        else:               # reference arm: K steps are given, bound the envs per step
            n_envs = int(min(n_envs, max(64, budget_s / (t_env * (steps + warmup)))))
    sets = [synthetic.make_poses(n_envs, gen, vt, TERRAIN["size_m"], TERRAIN["grid_res"]) for _ in range(4)]
    for i in range(warmup):
        OS.height_scan(*sets[i % 4], mesh)
    t0 = time.perf_counter()
    for i in range(steps):
        OS.height_scan(*sets[i % 4], mesh)
    dt = time.perf_counter() - t0
    return {"value": n_envs * N_RAYS * steps / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {n_envs} envs x {N_RAYS} rays on the same 2,000,000-triangle terrain, "
                      f"{dt:.1f} s of CPU work (oracle/: torch ray transform + C BVH raycast with OpenMP)",
            "ms_per_step": dt / steps * 1e3, "envs_per_step": n_envs}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from isaac_rover_orbit_b200 import terrain as TR

    v, f = TR.make_synthetic_terrain(**TERRAIN)
    # every step = one cfg-2 batch (4096 envs) when K + W such steps fit ~2 minutes of host time, else a bounded sample
    cpu = cpu_reference(v, f, steps=max(args.steps, 1), warmup=max(args.warmup, 1), n_envs=SCAN_ENVS_PER_GPU,
                        budget_s=120.0)
    n_envs = cpu["envs_per_step"]
    line = {
        "impl": "reference", "metric": "height_scan_rays_per_s", "value": cpu["value"], "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg-2 height-scan raycast only: 4096 envs x 961 rays per GPU, synthetic 200 m x "
                               "200 m Mars-like terrain, 2,000,000 triangles",
                   "sample": f"each step = {n_envs} envs x {N_RAYS} rays of that workload (bounded CPU sample)"},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class StdoutToStderr:
    """Everything but the final JSON line goes to stderr (NCCL prints its version banner on fd 1)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--single-launch-mdp", action="store_true",
                    help="fused step through the single rover_mdp_step launch instead of rover_mdp_pre_step + rover_mdp_post_step")
    ap.add_argument("--variant", type=int, default=int(os.environ.get("ROVER_SCAN_VARIANT", "5")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the fused step kernel by kernel instead of replaying CUDA graphs")
    ap.add_argument("--init-on-cpu", action="store_true", help="build the init-time tables on the host (profiling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        with StdoutToStderr():
            line = run_ours(args)
        if line is not None:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
