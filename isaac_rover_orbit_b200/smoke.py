"""``__graft_entry__.smoke()``: one small invocation of the hot path on cuda:0, checked against the CPU oracle.

The oracle (``oracle/``) is used here only as the checker; everything measured or shipped runs through
``librover_b200.so``.
"""
from __future__ import annotations

import torch

from . import ops, synthetic
from . import terrain as TR
from .config import RoverEnvCfg


def run(n: int = 64, size_m: float = 48.0, seed: int = 7) -> dict:
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs cuda:0; there is no CPU fallback")
    from oracle import raycast as oracle_raycast
    from oracle import step as OS

    dev = torch.device("cuda:0")
    cfg = RoverEnvCfg(num_envs=n)
    v, f = TR.make_synthetic_terrain(size_m, 0.2, seed=3)
    tables = TR.build_terrain_tables(v, f, n)
    vt = torch.from_numpy(v)
    gen = torch.Generator().manual_seed(seed)
    st = synthetic.make_step(n, gen, vt, size_m, 0.2, cfg.num_contact_bodies, cfg.target_rounds, margin=4.0)

    # ---- height scan
    grid = ops.ScanGridHandle.from_mesh(v, f, dev)
    rays = ops.RayPattern.grid(dev, cfg.height_scanner.pattern_cfg.resolution, cfg.height_scanner.pattern_cfg.size,
                               cfg.height_scanner.offset_pos)
    h_direct = ops.height_scan(st.root_pos_w.to(dev), st.root_quat_w.to(dev), rays, grid,
                               cfg.height_scanner.max_distance, cfg.height_scan_base_offset, variant=0)
    h_gpu = ops.height_scan(st.root_pos_w.to(dev), st.root_quat_w.to(dev), rays, grid,
                            cfg.height_scanner.max_distance, cfg.height_scan_base_offset)  # default: plane cells
    if not torch.equal(torch.isinf(h_direct), torch.isinf(h_gpu)) or \
            (h_direct - h_gpu).nan_to_num(0.0, 0.0, 0.0).abs().max().item() > 1e-5:
        raise AssertionError("smoke: plane-cell and direct height-scan kernels disagree")
    torch.cuda.synchronize()
    h_ref, _ = OS.height_scan(st.root_pos_w, st.root_quat_w, oracle_raycast.Mesh(v, f))
    h_gpu = h_gpu.cpu()
    if not torch.equal(torch.isinf(h_gpu), torch.isinf(h_ref)):
        raise AssertionError("smoke: hit mask differs from the oracle")
    fin = ~torch.isinf(h_ref)
    # tolerance: 1e-5 relative on the ray distance t (~10 m)  => 1e-4 m absolute on the height
    err = (h_gpu[fin] - h_ref[fin]).abs().max().item()
    if err > 1e-4:
        raise AssertionError(f"smoke: height scan differs from the oracle by {err}")

    # ---- fused MDP step
    pos_cmd_w, heading_cmd_w, ep_len = synthetic.init_commands(n, gen, st.root_pos_w)
    ost = OS.MdpState.zeros(n)
    ost.pos_cmd_w[:] = pos_cmd_w
    ost.heading_cmd_w[:] = heading_cmd_w
    ost.episode_length_buf[:] = ep_len
    ost.env_origins[:] = st.root_pos_w
    ost.time_left[:] = 150.0
    from oracle import terms as OT

    pb, hb = OT.update_command(ost.pos_cmd_w, ost.heading_cmd_w, st.root_pos_w, st.root_quat_w)
    ost.pos_cmd_b[:] = pb
    ost.heading_cmd_b[:] = hb
    buf = ops.MdpBuffers.allocate(n, dev)
    for k in ("pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "episode_length_buf", "env_origins",
              "time_left"):
        getattr(buf, k).copy_(getattr(ost, k))
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    d = st.to(dev)
    obs = torch.zeros(n, 965, device=dev)
    ops.mdp_pre_step(buf, params, d.actions, d.force_matrix_w)
    ops.mdp_post_step(buf, params, th, d.root_pos_w, d.root_quat_w, d.spawn_perm, d.yaw_u, d.heading_u, d.theta_u, obs)
    ops.height_scan(d.root_pos_w, d.root_quat_w, rays, grid, out=obs[:, 4:])
    torch.cuda.synchronize()
    otab = OS.TerrainTables(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table)
    out = OS.oracle_step(ost, st.actions, st.root_pos_w, st.root_quat_w, st.force_matrix_w, otab, st.spawn_perm,
                         st.yaw_u, st.theta_u, st.heading_u)
    if not torch.equal(buf.reset_flags.cpu().bool(), out.terminated | out.truncated):
        raise AssertionError("smoke: reset mask differs from the oracle")
    torch.testing.assert_close(buf.reward.cpu(), out.reward, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(buf.joint_vel.cpu(), out.joint_vel, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(obs[:, :4].cpu(), out.obs_head, rtol=1e-5, atol=1e-5)
    # ---- policy forward on the tensor cores (tcgen05), on the observation just produced, vs the fp32 oracle network
    from oracle import policy as OP

    from .policy import GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16

    net = GaussianNeuralNetwork(device=dev)
    g2 = torch.Generator().manual_seed(seed + 1)
    sd = {k: torch.randn(t.shape, generator=g2) * (0.05 if t.dim() == 2 else 0.01) for k, t in net.state_dict().items()}
    net.load_state_dict(sd)
    pobs = alloc_obs(n, dev)
    pobs.copy_(obs.nan_to_num(0.0, 0.0, 0.0))  # rays that miss are -inf: keep the comparison finite
    mean = net.compute({"states": pobs})[0]
    pobs_bf = alloc_obs_bf16(n, dev)
    ops.height_scan_obs(d.root_pos_w, d.root_quat_w, rays, grid, pobs.clone(), pobs_bf)  # exercises the bf16 mirror
    pobs_bf.copy_(pobs)
    mean_bf = net.compute({"states": pobs_bf})[0]
    torch.cuda.synchronize()
    ref_mean = OP.policy_mean(pobs.cpu(), sd)
    pol_err = (mean.cpu() - ref_mean).abs().max().item()
    if pol_err > 3e-2:  # bf16 operands vs the fp32 reference network (tests/test_gpu_policy.py uses the same bound)
        raise AssertionError(f"smoke: policy forward differs from the fp32 oracle by {pol_err}")
    if not torch.equal(mean, mean_bf):
        raise AssertionError("smoke: bf16-observation policy path differs from the fp32-observation path")
    return {"n_envs": n, "rays": int(h_ref.numel()), "scan_max_abs_err": err,
            "resets": int(out.stats["num_resets"]), "policy_max_abs_err": pol_err}


if __name__ == "__main__":
    print(run())
