"""GPU, BASELINE.json full sizes: properties that do not need the (slow) oracle at that size.

cfg-2: 4096 envs x 961 rays on the 2,000,000-triangle terrain; cfg-3: 16384 envs through the fused MDP step; cfg-4:
65536 envs through the policy.  Checked: every kernel variant returns the same bits, repeated launches are
deterministic, results are equivariant under a permutation of the environments and invariant under q -> -q, the
episode statistics equal plain torch reductions of the per-env outputs, the reset ranks reproduce
``reset_buf.nonzero()`` order (index-exact), a 48-env sample of the full-size scan agrees with the CPU oracle, and
EVERY ray of the cfg-2 launch agrees with a float64 interpolation of the heightfield derived independently of both.
"""
import numpy as np
import pytest
import torch

import bench
from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.policy import WEIGHT_KEYS, GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full(cuda_device):
    v, f, grid, tables = bench.build_world(bench.STEP_ENVS_CFG3, cuda_device, cuda_device)
    gen = torch.Generator().manual_seed(2024)
    vt = torch.from_numpy(v)
    pos, quat = synthetic.make_poses(bench.SCAN_ENVS_PER_GPU, gen, vt, bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"])
    return dict(v=v, f=f, vt=vt, grid=grid, tables=tables, pos=pos.to(cuda_device), quat=quat.to(cuda_device),
                rays=ops.RayPattern.grid(cuda_device), dev=cuda_device, gen=gen)


def test_scan_cfg2_variants_determinism_and_symmetries(full):
    pos, quat, rays, grid = full["pos"], full["quat"], full["rays"], full["grid"]
    assert pos.shape[0] == 4096 and rays.n_rays == 961
    ref = ops.height_scan(pos, quat, rays, grid, variant=2)
    assert float(torch.isfinite(ref).float().mean()) > 0.99  # the bench poses keep (nearly) every ray over the terrain
    for variant in (4, 5):
        assert torch.equal(ops.height_scan(pos, quat, rays, grid, variant=variant), ref), f"variant {variant}"
    h1 = ops.height_scan(pos, quat, rays, grid)
    h2 = ops.height_scan(pos, quat, rays, grid)
    assert torch.equal(h1, h2) and torch.equal(h1, ref)  # deterministic, default variant included
    perm = torch.randperm(pos.shape[0], generator=torch.Generator().manual_seed(1)).to(pos.device)
    assert torch.equal(ops.height_scan(pos[perm].contiguous(), quat[perm].contiguous(), rays, grid), ref[perm])
    assert torch.equal(ops.height_scan(pos, (-quat).contiguous(), rays, grid), ref)  # q and -q are the same rotation
    # lifting the sensor by dz lifts every height by dz up to fp32 rounding of (z + dz) - hit - offset
    up = pos.clone()
    up[:, 2] += 0.5
    lifted = ops.height_scan(up, quat, rays, grid)
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(lifted), fin)
    torch.testing.assert_close(lifted[fin], ref[fin] + 0.5, rtol=0, atol=2e-6)
    # checksum of checksums: per-env sums (fp64) of two variants agree exactly because the heights do
    assert torch.equal(h1.double().sum(1), ref.double().sum(1))


def test_scan_cfg2_sample_against_oracle(full):
    from oracle import raycast as oracle_raycast
    from oracle import step as OS

    idx = torch.arange(0, 4096, 86)[:48]
    pos, quat = full["pos"][idx.to(full["dev"])].cpu(), full["quat"][idx.to(full["dev"])].cpu()
    h = ops.height_scan(pos.to(full["dev"]), quat.to(full["dev"]), full["rays"], full["grid"]).cpu()
    h_ref, _ = OS.height_scan(pos, quat, oracle_raycast.Mesh(full["v"], full["f"]))
    assert torch.equal(torch.isinf(h), torch.isinf(h_ref))
    fin = ~torch.isinf(h_ref)
    assert (h[fin] - h_ref[fin]).abs().max().item() <= 1e-4  # 1e-5 relative on the ray distance t ~ 10 m


def test_scan_cfg2_every_ray_against_float64_heightfield_interpolation(full):
    """All 3,936,256 rays of the cfg-2 launch against geometry derived independently of the kernels AND of the oracle, in
    float64 on the same fp32 inputs: ORBIT's yaw-only ray transform, then the barycentric interpolation of the lattice
    triangle under the ray (terrain.make_synthetic_terrain splits every quad along (1,0)-(0,1)).
    Tolerance: the north-star 1e-5 relative on the ray distance (t ~ 10 m -> 1e-4 m); a ray within rounding of a cell or
    diagonal border may be evaluated on the neighbouring triangle by either side, which changes z by slope * 1e-6 m."""
    dev = full["dev"]
    pos, quat, rays = full["pos"].double(), full["quat"].double(), full["rays"]
    res = bench.TERRAIN["grid_res"]
    n = int(round(bench.TERRAIN["size_m"] / res)) + 1
    z = torch.from_numpy(full["v"][:, 2].astype(np.float64)).reshape(n, n).to(dev)  # [y index, x index]
    xs = torch.from_numpy(full["v"][:n, 0].astype(np.float64)).to(dev)             # the fp32 lattice lines, exactly
    w_, x_, y_, z_ = quat.unbind(1)
    yaw = torch.atan2(2.0 * (w_ * z_ + x_ * y_), 1.0 - 2.0 * (y_ * y_ + z_ * z_))
    c, s = torch.cos(yaw)[:, None], torch.sin(yaw)[:, None]
    loc = rays.starts.double()
    X = pos[:, 0:1] + c * loc[None, :, 0] - s * loc[None, :, 1]
    Y = pos[:, 1:2] + s * loc[None, :, 0] + c * loc[None, :, 1]
    ci = (torch.searchsorted(xs, X.contiguous(), right=True) - 1).clamp_(0, n - 2)
    cj = (torch.searchsorted(xs, Y.contiguous(), right=True) - 1).clamp_(0, n - 2)
    u = (X - xs[ci]) / (xs[ci + 1] - xs[ci])
    w = (Y - xs[cj]) / (xs[cj + 1] - xs[cj])
    z00, z10, z01, z11 = z[cj, ci], z[cj, ci + 1], z[cj + 1, ci], z[cj + 1, ci + 1]
    lower = (u + w) <= 1.0
    zhit = torch.where(lower, z00 + u * (z10 - z00) + w * (z01 - z00),
                       z11 + (1.0 - u) * (z01 - z11) + (1.0 - w) * (z10 - z11))
    want = pos[:, 2:3] - zhit - 0.26878
    got = ops.height_scan(full["pos"], full["quat"], rays, full["grid"])
    inside = (X > xs[0]) & (X < xs[-1]) & (Y > xs[0]) & (Y < xs[-1])
    assert float(inside.double().mean()) > 0.99 and bool(torch.isfinite(got[inside]).all())
    err = (got.double() - want)[inside].abs()
    assert float(err.max()) <= 1e-4, float(err.max())
    assert float(err.mean()) < 3e-6  # typical error is fp32 rounding of a ~10 m ray parameter


def test_mdp_cfg3_steps_against_oracle(full):
    """cfg-3 size (16384 envs, 200 m terrain tables): three consecutive fused MDP steps against the vectorised CPU oracle
    with the same checks as the small-terrain parity test -- masks and reset indices bit-exact, floats 1e-5 relative."""
    from mdp_parity import check_mdp_steps_vs_oracle

    n = bench.STEP_ENVS_CFG3
    assert n == 16384
    resets = check_mdp_steps_vs_oracle(full["dev"], n, full["v"], full["tables"].to("cpu"), bench.TERRAIN["size_m"],
                                       bench.TERRAIN["grid_res"], n_steps=3, margin=20.0, max_ambiguous=4)
    assert resets > 500


def test_mdp_cfg3_statistics_ranks_and_determinism(full):
    dev, tables, n = full["dev"], full["tables"], bench.STEP_ENVS_CFG3
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    gen = torch.Generator().manual_seed(5)
    st = synthetic.make_step(n, gen, full["vt"], bench.TERRAIN["size_m"], bench.TERRAIN["grid_res"],
                             cfg.num_contact_bodies, cfg.target_rounds).to(dev)
    pc, hc, ep = synthetic.init_commands(n, gen, st.root_pos_w.cpu())

    def run():
        buf = ops.MdpBuffers.allocate(n, dev)
        buf.pos_cmd_w.copy_(pc)
        buf.heading_cmd_w.copy_(hc)
        buf.episode_length_buf.copy_(ep)
        buf.env_origins.copy_(st.root_pos_w)
        buf.time_left.fill_(150.0)
        gb = torch.Generator().manual_seed(8)  # previous body-frame command: 1..7 m away, so few envs terminate
        ang = torch.rand(n, generator=gb) * 6.2831853
        rad = torch.rand(n, generator=gb) * 6.0 + 1.0
        buf.pos_cmd_b.copy_(torch.stack([rad * torch.cos(ang), rad * torch.sin(ang), torch.zeros(n)], 1))
        buf.heading_cmd_b.copy_(torch.rand(n, generator=gb) * 2 - 1)
        buf.episode_sums.copy_(torch.rand(n, 7, generator=torch.Generator().manual_seed(6)))
        sums_before = buf.episode_sums.clone()
        pos, quat = st.root_pos_w.clone(), st.root_quat_w.clone()
        obs = torch.zeros(n, 965, device=dev)
        ops.mdp_pre_step(buf, params, st.actions, st.force_matrix_w)
        sums_mid = buf.episode_sums.clone()
        ops.mdp_post_step(buf, params, th, pos, quat, st.spawn_perm, st.yaw_u, st.heading_u, st.theta_u, obs)
        torch.cuda.synchronize()
        return buf, pos, quat, obs, sums_before, sums_mid

    buf, pos, quat, obs, sums_before, sums_mid = run()
    reset = buf.reset_flags.bool()
    k = int(reset.sum())
    assert 0 < k < n
    # index-exact: the j-th reset env (ascending env id, = reset_buf.nonzero()) takes spawn row spawn_perm[j]
    ids = reset.nonzero().squeeze(1)
    assert torch.equal(buf.spawn_index[ids], st.spawn_perm[:k])
    assert bool((buf.spawn_index[~reset] == -1).all())
    assert torch.equal(pos[ids, :2], th.spawn_table[st.spawn_perm[:k], :2])
    assert torch.equal(pos[~reset], st.root_pos_w[~reset])
    # reward bookkeeping: episodic sums advance by the weighted term rewards, the step reward is their sum in order
    torch.testing.assert_close(sums_mid, sums_before + buf.term_rewards, rtol=0, atol=1e-6)
    total = torch.zeros(n, device=dev)
    for j in range(7):
        total = total + buf.term_rewards[:, j]
    assert torch.equal(total, buf.reward)
    # statistics vector = torch reductions over the reset envs (sums in another order: fp32 tolerance)
    stats = buf.stats.cpu().double()
    torch.testing.assert_close(stats[:7], sums_mid[reset].double().sum(0).cpu(), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(stats[7:11], buf.term_flags[reset].double().sum(0).cpu(), rtol=0, atol=0)
    assert stats[13] == k
    assert bool((buf.episode_sums[reset] == 0).all()) and bool((buf.episode_length_buf[reset] == 0).all())
    # the observation head is finite and repeated runs are bit-identical (deterministic statistics included)
    assert torch.isfinite(obs[:, :4]).all()
    buf2, pos2, quat2, obs2, _, _ = run()
    assert torch.equal(buf.stats, buf2.stats) and torch.equal(obs, obs2) and torch.equal(pos, pos2)
    assert torch.equal(buf.pos_cmd_w, buf2.pos_cmd_w) and torch.equal(buf.reward, buf2.reward)


def test_policy_cfg4_kernels_agree_and_rows_are_independent(full, monkeypatch):
    dev, n = full["dev"], 65536
    net = GaussianNeuralNetwork(device=dev)
    g = torch.Generator().manual_seed(9)
    net.load_state_dict({k: torch.randn(t.shape, generator=g) * (0.05 if t.dim() == 2 else 0.01)
                         for k, t in net.state_dict().items()})
    obs = alloc_obs(n, dev)
    obs.copy_(torch.randn(n, 965, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) * 0.3)
    mean = net.compute({"states": obs})[0]
    assert torch.isfinite(mean).all() and float(mean.abs().max()) <= 1.0
    monkeypatch.setenv("ROVER_POLICY_KERNEL", "v1")
    assert torch.equal(net.compute({"states": obs})[0], mean)
    monkeypatch.setenv("ROVER_POLICY_KERNEL", "ws")
    ob = alloc_obs_bf16(n, dev)
    ob.copy_(obs)
    assert torch.equal(net.compute({"states": ob})[0], mean)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2)).to(dev)
    obs_p = alloc_obs(n, dev)
    obs_p.copy_(obs[perm])
    assert torch.equal(net.compute({"states": obs_p})[0], mean[perm])  # tiles / rounds do not leak between rows
    assert torch.equal(net.compute({"states": obs[:1000]})[0], mean[:1000])  # another tile height, same rows
    # every one of the 65536 rows against a torch emulation of the same numerics on the GPU (bf16 operands, fp32
    # accumulation; true fp32 matmuls, TF32 off) and against the plain fp32 network
    from test_gpu_policy import emulate_bf16

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = {k: t.to(dev) for k, t in net.state_dict().items()}
        emu = emulate_bf16(obs[:, :965], sd)
        assert float((mean - emu).abs().max()) <= 4e-3
        lin = lambda k, x: x @ sd[k + ".weight"].T + sd[k + ".bias"]  # noqa: E731
        lr = lambda x: torch.nn.functional.leaky_relu(x, 0.01)  # noqa: E731
        keys = WEIGHT_KEYS
        h = torch.cat([obs[:, 0:4], lr(lin(keys[1], lr(lin(keys[0], obs[:, 3:964]))))], dim=1)
        for k in keys[2:5]:
            h = lr(lin(k, h))
        ref = torch.tanh(lin(keys[5], h))
        # bf16 operands cannot meet 1e-5: the kernel may be no further from the fp32 network than
        # the emulation's own operand-rounding distance plus the accumulation-order allowance above
        assert float((mean - ref).abs().max()) <= float((emu - ref).abs().max()) + 4e-3
        assert float((mean - ref).abs().max()) < 3e-2
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
