"""GPU: the in-kernel counter-based variates (csrc/rng.cuh).  A post-step launched with ``rng_state = {seed, step}`` must
produce, bit for bit, what the same kernel produces when it is fed the explicit arrays that ``rover_rng_variates``
(the same functions compiled for the host, pinned by tests/test_rng_cpu.py) returns for that (seed, step) -- and hence
what the oracle produces from those arrays.  Also: the step word advances by one per launch (CUDA-graph safe)."""
import pytest
import torch

from isaac_rover_orbit_b200 import _lib, ops, synthetic
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg

pytestmark = pytest.mark.gpu
SIZE, RES = 48.0, 0.2
STATE = ("action", "prev_action", "pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "time_left",
         "command_counter", "episode_length_buf", "episode_sums", "env_origins", "err_pos", "err_heading", "spawn_index",
         "stats", "log", "reward", "reset_flags")


@pytest.fixture(scope="module")
def world(cuda_device):
    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    return dict(v=v, f=f, dev=cuda_device)


def _setup(world, n, seed):
    dev = world["dev"]
    tables = TR.build_terrain_tables(world["v"], world["f"], n)
    cfg = RoverEnvCfg(num_envs=n)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    gen = torch.Generator().manual_seed(seed)
    vt = torch.from_numpy(world["v"])
    steps = [synthetic.make_step(n, gen, vt, SIZE, RES, margin=4.0).to(dev) for _ in range(4)]
    pc, hc, ep = synthetic.init_commands(n, gen, steps[0].root_pos_w.cpu())
    return cfg, ops.mdp_params(cfg), th, steps, (pc, hc, ep), tables


def _fresh(n, dev, init, root):
    buf = ops.MdpBuffers.allocate(n, dev)
    buf.pos_cmd_w.copy_(init[0])
    buf.heading_cmd_w.copy_(init[1])
    buf.episode_length_buf.copy_(init[2])
    buf.env_origins.copy_(root)
    buf.time_left.fill_(150.0)
    buf.time_left[min(3, n - 1)] = 0.1  # the time-based resample draws from the env's variates as well
    return buf


@pytest.mark.parametrize("n,seed,rounds", [(1, 5, 16), (63, 6, 3), (64, 7, 16), (1000, 8, 9), (4096 + 17, 2 ** 41 + 11, 16)])
def test_kernel_rng_equals_explicit_variates_from_the_host_export(world, n, seed, rounds):
    dev = world["dev"]
    cfg, params, th, steps, init, tables = _setup(world, n, seed)
    a, b = _fresh(n, dev, init, steps[0].root_pos_w), _fresh(n, dev, init, steps[0].root_pos_w)
    rng = ops.ResetRng(seed, dev, step=2 ** 32 - 2)  # crosses the 32-bit boundary of the step counter
    obs_a, obs_b = torch.zeros(n, 965, device=dev), torch.zeros(n, 965, device=dev)
    for k, s in enumerate(steps):
        pa, qa = s.root_pos_w.clone(), s.root_quat_w.clone()
        pb, qb = s.root_pos_w.clone(), s.root_quat_w.clone()
        assert rng.peek() == (seed, 2 ** 32 - 2 + k)
        sp_env, yaw, head, theta = (t.to(dev) for t in rng.variates(n, rounds, th.n_spawns))
        for buf in (a, b):
            ops.mdp_pre_step(buf, params, s.actions, s.force_matrix_w)
        sp_full = ops.ResetRng.by_rank(sp_env, b.reset_flags)  # explicit interface: rows by reset rank
        ops.mdp_post_step(a, params, th, pa, qa, obs=obs_a, rng=rng, n_rounds=rounds)
        ops.mdp_post_step(b, params, th, pb, qb, sp_full, yaw, head, theta, obs=obs_b)
        torch.cuda.synchronize()
        for name in STATE:
            assert torch.equal(getattr(a, name), getattr(b, name)), (k, name)
        assert torch.equal(pa, pb) and torch.equal(qa, qb) and torch.equal(obs_a, obs_b)
        if n >= 64:
            assert float(a.stats[13]) > 0, "the fixture must reset some envs"
        used = a.spawn_index[a.spawn_index >= 0]
        assert used.unique().numel() == used.numel(), "spawn rows are drawn without replacement"
    assert rng.peek()[1] == 2 ** 32 - 2 + len(steps)


def test_kernel_rng_against_the_oracle_and_single_launch(world):
    """The oracle step fed with the exported variates == the RNG-mode kernels (two launches and the single launch)."""
    from oracle import step as OS
    from oracle import terms as OT

    dev, n, seed = world["dev"], 512, 99
    cfg, params, th, steps, init, tables = _setup(world, n, seed)
    s = steps[1]
    otab = OS.TerrainTables(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table)
    ost = OS.MdpState.zeros(n)
    ost.pos_cmd_w[:], ost.heading_cmd_w[:], ost.episode_length_buf[:] = init
    ost.env_origins[:] = s.root_pos_w.cpu()
    ost.time_left[:] = 150.0
    ost.pos_cmd_b[:], ost.heading_cmd_b[:] = OT.update_command(init[0], init[1], s.root_pos_w.cpu(), s.root_quat_w.cpu())
    dist0 = ost.pos_cmd_b[:, :2].norm(dim=1)
    ost.pos_cmd_b[((dist0 - 0.18).abs() < 1e-6) | ((dist0 - 11.0).abs() < 1e-5)] *= 1.001
    bufs = [ops.MdpBuffers.allocate(n, dev) for _ in range(2)]
    for buf in bufs:
        for k in ("pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "episode_length_buf", "env_origins", "time_left"):
            getattr(buf, k).copy_(getattr(ost, k))
    rngs = [ops.ResetRng(seed, dev, step=41), ops.ResetRng(seed, dev, step=41)]
    sp, yaw, head, theta = rngs[0].variates(n, 16, th.n_spawns)
    roots = [(s.root_pos_w.clone(), s.root_quat_w.clone()) for _ in range(2)]
    ops.mdp_pre_step(bufs[0], params, s.actions, s.force_matrix_w)
    ops.mdp_post_step(bufs[0], params, th, *roots[0], rng=rngs[0])
    ops.mdp_step(bufs[1], params, th, s.actions, s.force_matrix_w, *roots[1], rng=rngs[1])
    torch.cuda.synchronize()
    for name in STATE:
        assert torch.equal(getattr(bufs[0], name), getattr(bufs[1], name)), name
    out = OS.oracle_step(ost, s.actions.cpu(), s.root_pos_w.cpu(), s.root_quat_w.cpu(), s.force_matrix_w.cpu(), otab, None,
                         yaw, theta, head, spawn_by_env=sp)
    ids = out.reset_ids
    assert len(ids) > 10
    got = bufs[0].spawn_index.cpu()
    assert torch.equal((got >= 0).nonzero().squeeze(-1), ids) and torch.equal(got[ids], out.spawn_index)
    assert torch.equal(roots[0][0].cpu(), out.root_pos_w)
    torch.testing.assert_close(bufs[0].pos_cmd_w.cpu()[:, :2], ost.pos_cmd_w[:, :2], rtol=1e-6, atol=2e-5)
    torch.testing.assert_close(bufs[0].heading_cmd_w.cpu(), ost.heading_cmd_w, rtol=1e-6, atol=1e-6)
    assert rngs[0].peek()[1] == 42 and rngs[1].peek()[1] == 42


def test_log_vector_is_written_only_by_a_launch_that_resets(world):
    """extras["log"] semantics of the ORBIT managers (A.2; rover_env.py:89-91 calls _reset_idx only when something
    reset): the kernel refreshes ``buf.log`` when >= 1 env reset and leaves the previous values otherwise."""
    dev, n = world["dev"], 256
    cfg, params, th, steps, init, tables = _setup(world, n, 3)
    buf = _fresh(n, dev, init, steps[0].root_pos_w)
    rng = ops.ResetRng(1, dev)
    s = steps[0]
    ops.mdp_pre_step(buf, params, s.actions, s.force_matrix_w)
    buf.stats.zero_()
    ops.mdp_post_step(buf, params, th, s.root_pos_w.clone(), s.root_quat_w.clone(), rng=rng)
    torch.cuda.synchronize()
    st, log = buf.stats.clone(), buf.log.clone()
    k = float(st[13])
    assert k > 0
    torch.testing.assert_close(log[:7], st[:7] / k / 150.0, rtol=1e-6, atol=0)
    assert torch.equal(log[7:11], st[7:11]) and float(log[13]) == k
    torch.testing.assert_close(log[11:13], st[11:13] / k, rtol=1e-6, atol=0)
    # a launch without resets: flags cleared by hand (quiet step), the log must not move
    buf.reset_flags.zero_()
    buf.block_reset_counts.zero_()
    buf.stats.zero_()
    ops.mdp_post_step(buf, params, th, s.root_pos_w.clone(), s.root_quat_w.clone(), rng=rng)
    torch.cuda.synchronize()
    assert float(buf.stats[13]) == 0 and torch.equal(buf.log, log)


def test_post_step_argument_errors(world):
    dev, n = world["dev"], 128
    cfg, params, th, steps, init, tables = _setup(world, n, 4)
    buf = _fresh(n, dev, init, steps[0].root_pos_w)
    s = steps[0]
    small = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table[: n // 2],
                                    tables.resolution, dev)
    with pytest.raises(RuntimeError, match="spawn table"):
        ops.mdp_post_step(buf, params, small, s.root_pos_w, s.root_quat_w, rng=ops.ResetRng(0, dev))
    with pytest.raises(RuntimeError):
        ops.mdp_post_step(buf, params, th, s.root_pos_w, s.root_quat_w)  # neither variates nor rng
    with pytest.raises(RuntimeError):
        ops.ResetRng(0, "cpu")
    bad = _lib.MdpParams.from_buffer_copy(params)
    bad.resampling_time = 0.1
    with pytest.raises(RuntimeError, match="resampling_time"):
        ops.mdp_post_step(buf, bad, th, s.root_pos_w, s.root_quat_w, rng=ops.ResetRng(0, dev))


@pytest.mark.parametrize("n,rounds,unsafe", [(96, 16, 0.85), (333, 9, 0.7), (64, 3, 0.5), (1024, 16, 0.97)])
def test_cooperative_target_draw_under_dense_rejections_against_the_oracle(world, n, rounds, unsafe):
    """The target rejection loop is warp-cooperative (four resampling envs per pass, eight rounds side by side).  Worst
    case for that scheme: EVERY env resets (32 owners per warp = 8 passes), most candidates are rejected (several batches
    of 8 rounds, exhausted envs keep the last candidate tried) and ``n_rounds`` is not a multiple of 8.  Explicit variates
    and in-kernel variates, two launches and one launch, all against the oracle's sequential loop."""
    from oracle import step as OS
    from oracle import terms as OT

    dev, seed = world["dev"], 1234 + n
    cfg, params, th0, steps, init, tables = _setup(world, n, seed)
    g = torch.Generator().manual_seed(seed)
    mask = torch.as_tensor(tables.safe_mask).clone()
    mask[torch.rand(mask.shape, generator=g) < unsafe] = 1          # 1 = not a valid target cell (terrain_utils.py:220)
    th = ops.TerrainTablesHandle(tables.heightmap, mask, tables.offset_xy, tables.spawn_table, tables.resolution, dev)
    otab = OS.TerrainTables(tables.heightmap, mask, tables.offset_xy, tables.spawn_table)
    s = steps[0]

    def oracle_state():
        ost = OS.MdpState.zeros(n)
        ost.pos_cmd_w[:], ost.heading_cmd_w[:], ost.episode_length_buf[:] = init
        ost.env_origins[:] = s.root_pos_w.cpu()
        ost.time_left[:] = 150.0
        ost.pos_cmd_b[:] = torch.tensor([20.0, 0.0, 0.0])           # far_from_target: every env resets
        return ost

    ost = oracle_state()
    rngs = [ops.ResetRng(seed, dev, step=7) for _ in range(2)]
    sp, yaw, head, theta = rngs[0].variates(n, rounds, th.n_spawns)
    out = OS.oracle_step(ost, s.actions.cpu(), s.root_pos_w.cpu(), s.root_quat_w.cpu(), s.force_matrix_w.cpu(), otab, None,
                         yaw, theta, head, spawn_by_env=sp)
    assert len(out.reset_ids) == n
    if unsafe > 0.9:
        assert out.stats["target_rounds_exhausted"] > 0, "the fixture must exhaust some envs"

    def fresh():
        buf = ops.MdpBuffers.allocate(n, dev)
        src = oracle_state()
        for k in ("pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "episode_length_buf", "env_origins", "time_left"):
            getattr(buf, k).copy_(getattr(src, k))
        return buf

    results = []
    # (a) in-kernel variates, two launches; (b) in-kernel variates, one launch; (c) explicit arrays, two launches
    for mode in ("rng2", "rng1", "explicit"):
        buf, p, q = fresh(), s.root_pos_w.clone(), s.root_quat_w.clone()
        if mode == "rng1":
            ops.mdp_step(buf, params, th, s.actions, s.force_matrix_w, p, q, rng=rngs[1], n_rounds=rounds)
        else:
            ops.mdp_pre_step(buf, params, s.actions, s.force_matrix_w)
            if mode == "rng2":
                ops.mdp_post_step(buf, params, th, p, q, rng=rngs[0], n_rounds=rounds)
            else:
                ops.mdp_post_step(buf, params, th, p, q, ops.ResetRng.by_rank(sp.to(dev), buf.reset_flags), yaw.to(dev),
                                  head.to(dev), theta.to(dev))
        torch.cuda.synchronize()
        results.append((buf, p))
        assert int(buf.stats[13]) == n and int(buf.stats[14]) == out.stats["target_rounds_exhausted"], mode
        torch.testing.assert_close(buf.pos_cmd_w.cpu(), ost.pos_cmd_w, rtol=1e-6, atol=2e-5)
        torch.testing.assert_close(buf.heading_cmd_w.cpu(), ost.heading_cmd_w, rtol=1e-6, atol=1e-6)
        assert torch.equal(p.cpu(), out.root_pos_w)
    for buf, p in results[1:]:
        for name in ("pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "time_left", "command_counter", "env_origins"):
            assert torch.equal(getattr(buf, name), getattr(results[0][0], name)), name
