"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the CPU oracle.

Tolerances (BASELINE.json north_star): bit-exact hit / termination masks and resampled indices; 1e-5 relative
for float rewards, actions and ray distances.  A ray distance is t ~ 10 m, so heights are compared with
``atol = 1e-5 * t`` (1e-4 m).  Threshold-exact masks are compared on the envs whose compared quantity is not
within a few ulp of the threshold (CPU libm and CUDA atan2f/sinf/cosf differ in the last ulp).
"""
import os

import numpy as np
import pytest
import torch

from isaac_rover_orbit_b200 import ops, synthetic
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from mdp_parity import _load_state, _near  # noqa: F401  (shared with tests/test_gpu_fullsize.py)

pytestmark = pytest.mark.gpu

SIZE, RES = 48.0, 0.2


@pytest.fixture(scope="module")
def world(cuda_device):
    from oracle import raycast as oracle_raycast

    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    n = 256
    tables = TR.build_terrain_tables(v, f, n)
    return dict(v=v, f=f, n=n, tables=tables, mesh=oracle_raycast.Mesh(v, f),
                grid=ops.ScanGridHandle.from_mesh(v, f, cuda_device), dev=cuda_device,
                rays=ops.RayPattern.grid(cuda_device))


def _scan_compare(h_gpu, h_ref, t_ref=10.0):
    h_gpu = h_gpu.cpu()
    assert torch.equal(torch.isinf(h_gpu), torch.isinf(h_ref)), "hit mask must be bit-exact"
    fin = ~torch.isinf(h_ref)
    assert (h_gpu[~fin] == -float("inf")).all()
    if fin.any():
        err = (h_gpu[fin] - h_ref[fin]).abs().max().item()
        assert err <= 1e-5 * max(t_ref, 1.0), f"height error {err}"


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
def test_height_scan_vs_oracle_small_terrain(world, variant):
    from oracle import step as OS

    gen = torch.Generator().manual_seed(11)
    n = world["n"]
    pos, quat = synthetic.make_poses(n, gen, torch.from_numpy(world["v"]), SIZE, RES, margin=0.5)
    # rays leaving the mesh (misses), exact axis-aligned yaw on grid lines (rays on edges / vertices)
    pos[:4, :2] = torch.tensor([[0.3, 0.3], [47.9, 20.0], [24.0, -0.7], [10.0, 10.2]])
    quat[4:8] = torch.tensor([[1.0, 0, 0, 0], [0.70710678, 0, 0, 0.70710678], [0.0, 0, 0, 1.0], [1.0, 0, 0, 0]])
    pos[4:8, :2] = torch.tensor([[10.0, 10.0], [20.2, 30.0], [5.0, 5.0], [40.1, 40.1]])
    dev = world["dev"]
    h, hits = ops.height_scan(pos.to(dev), quat.to(dev), world["rays"], world["grid"], return_hits=True,
                              variant=variant)
    h_ref, hits_ref = OS.height_scan(pos, quat, world["mesh"])
    assert torch.isinf(h_ref).any() and (~torch.isinf(h_ref)).any()
    _scan_compare(h, h_ref)
    hits = hits.cpu()
    miss = torch.isinf(hits_ref[..., 2])
    assert torch.equal(torch.isinf(hits).all(-1), miss)
    torch.testing.assert_close(hits[~miss], hits_ref[~miss], rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
def test_height_scan_analytic_plane(cuda_device, variant):
    """Independent of the oracle: on the plane z = 0.3x - 0.2y + 1 the scan equals body_z - z(x,y) - 0.26878 at the
    961 yaw-rotated grid points (SURVEY.md 8c invariant)."""
    v = np.array([[-50, -50, 0], [50, -50, 0], [50, 50, 0], [-50, 50, 0]], dtype=np.float32)
    v[:, 2] = 0.3 * v[:, 0] - 0.2 * v[:, 1] + 1.0
    f = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    # a 1 x 1 lattice with a closed-form cell: the home grid is deferred until variant 0 asks for it
    assert grid.has_home_grid == (grid.cells.n_general != 0)
    grid.ensure_home_grid()
    assert len(grid.grid.levels) == 1 and grid.grid.levels[0].cell > 50  # two giant triangles -> a coarse level
    gen = torch.Generator().manual_seed(5)
    n = 33
    pos = torch.cat([torch.rand(n, 2, generator=gen) * 60 - 30, torch.rand(n, 1, generator=gen) * 3 + 20], 1)
    yaw = (torch.rand(n, generator=gen) * 2 - 1) * np.pi
    quat = synthetic.quat_from_euler(torch.randn(n, generator=gen) * 0.2, torch.randn(n, generator=gen) * 0.2, yaw)
    rays = ops.grid_pattern()
    h = ops.height_scan(pos.to(cuda_device), quat.to(cuda_device), ops.RayPattern(rays, cuda_device), grid,
                        variant=variant).cpu().double()
    # float64 expectation with the true yaw of the full quaternion
    w, x, y, z = quat.double().unbind(1)
    yaw_true = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    c, s = torch.cos(yaw_true)[:, None], torch.sin(yaw_true)[:, None]
    rx, ry = rays[:, 0].double()[None], rays[:, 1].double()[None]
    X = pos[:, 0:1].double() + c * rx - s * ry
    Y = pos[:, 1:2].double() + s * rx + c * ry
    expect = pos[:, 2:3].double() - (0.3 * X - 0.2 * Y + 1.0) - 0.26878
    torch.testing.assert_close(h, expect, rtol=0, atol=2e-4)


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
def test_height_scan_mixed_mesh_levels_and_degenerates(cuda_device, variant):
    """Ground of two huge triangles + small rock pyramids + vertical / zero-area / flipped triangles."""
    from oracle import raycast as oracle_raycast
    from oracle import step as OS

    rng = np.random.default_rng(2)
    verts = [[-20, -20, 0.1], [20, -20, -0.1], [20, 20, 0.2], [-20, 20, 0.0]]
    faces = [[0, 1, 2], [0, 3, 2]]  # second one is clockwise: double-sided must still hit
    for _ in range(300):
        cx, cy = rng.uniform(-15, 15, 2)
        r, hgt = rng.uniform(0.1, 0.8), rng.uniform(0.2, 1.0)
        b = len(verts)
        verts += [[cx - r, cy - r, 0.05], [cx + r, cy - r, 0.05], [cx + r, cy + r, 0.05], [cx - r, cy + r, 0.05],
                  [cx, cy, hgt]]
        faces += [[b, b + 1, b + 4], [b + 1, b + 2, b + 4], [b + 2, b + 3, b + 4], [b + 3, b, b + 4]]
    b = len(verts)
    verts += [[1, 1, 0], [1, 1, 3], [2, 1, 0], [3, 3, 1], [3, 3, 1], [4, 4, 1]]  # vertical wall, zero-area
    faces += [[b, b + 1, b + 2], [b + 3, b + 4, b + 5]]
    v = np.array(verts, dtype=np.float32)
    f = np.array(faces, dtype=np.int32)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device, home_grid=True)
    assert len(grid.grid.levels) >= 2 and grid.grid.n_dropped == 2
    gen = torch.Generator().manual_seed(9)
    n = 128
    pos = torch.cat([torch.rand(n, 2, generator=gen) * 44 - 22, torch.rand(n, 1, generator=gen) + 0.5], 1)
    quat = synthetic.quat_from_euler(torch.zeros(n), torch.zeros(n), (torch.rand(n, generator=gen) * 2 - 1) * np.pi)
    h = ops.height_scan(pos.to(cuda_device), quat.to(cuda_device), ops.RayPattern.grid(cuda_device), grid,
                        variant=variant)
    h_ref, _ = OS.height_scan(pos, quat, oracle_raycast.Mesh(v, f))
    # a 40 m ground triangle lives on a coarse level whose fp32 edge functions resolve ~3e-5 m (as does the
    # reference's own fp32 test at that size): rays that close to the OUTER border of the ground are excluded
    starts = OS.ray_starts_world(pos, quat)
    border = ((starts[..., :2].abs() - 20.0).abs() < 1e-3).any(dim=-1)
    assert border.sum() < 200
    h = torch.where(border.to(h.device), torch.zeros_like(h), h)
    h_ref = torch.where(border, torch.zeros_like(h_ref), h_ref)
    _scan_compare(h, h_ref)


def _heightfield_mesh(nx, ny, dx, dy, seed):
    """Regular heightfield, two triangles per quad (same diagonal), vertices x-fastest."""
    rng = np.random.default_rng(seed)
    xs, ys = np.arange(nx + 1) * dx, np.arange(ny + 1) * dy
    X, Y = np.meshgrid(xs, ys)
    Z = 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y) + 0.05 * rng.standard_normal(X.shape)
    v = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny))
    a = (j * (nx + 1) + i).ravel()
    f = np.concatenate([np.stack([a, a + 1, a + nx + 2], 1), np.stack([a, a + nx + 2, a + nx + 1], 1)]).astype(np.int32)
    return v, f


@pytest.mark.parametrize("shape", ["more lines than the shared tables hold", "cells so small that windows do not fit",
                                   "narrow strip"])
def test_height_scan_pipelined_variants_fallback_paths(cuda_device, shape):
    """Variants 4 / 5 stage table windows in shared memory; tables whose grid lines do not fit the shared line tables,
    whose windows exceed the staged 26 x 26 cells, or that are narrower than a window take the global-memory paths and
    must return exactly what variant 2 returns."""
    if shape.startswith("more lines"):
        v, f = _heightfield_mesh(1100, 30, 0.2, 0.2, 1)     # 1100 > 1024 lines in x
        lo, hi = (2.0, 2.0), (218.0, 4.0)
    elif shape.startswith("cells so small"):
        v, f = _heightfield_mesh(400, 400, 0.05, 0.05, 2)   # a 3 m pattern spans ~85 cells > 26
        lo, hi = (1.0, 1.0), (19.0, 19.0)
    else:
        v, f = _heightfield_mesh(300, 9, 0.2, 0.2, 3)        # 9 rows: every window hangs over both borders in y
        lo, hi = (-1.0, -1.0), (61.0, 3.0)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    gen = torch.Generator().manual_seed(11)
    n = 700
    u = torch.rand(n, 2, generator=gen)
    pos = torch.cat([u * (torch.tensor(hi) - torch.tensor(lo)) + torch.tensor(lo), torch.rand(n, 1, generator=gen) + 0.5], 1)
    quat = synthetic.quat_from_euler(torch.zeros(n), torch.zeros(n), (torch.rand(n, generator=gen) * 2 - 1) * np.pi)
    rays = ops.RayPattern.grid(cuda_device)
    pos, quat = pos.to(cuda_device), quat.to(cuda_device)
    ref = ops.height_scan(pos, quat, rays, grid, variant=2)
    assert torch.isfinite(ref).any() and torch.isinf(ref).any()  # hits and misses (rays beyond the border) both occur
    for variant in (4, 5):
        h = ops.height_scan(pos, quat, rays, grid, variant=variant)
        assert torch.equal(h, ref), f"variant {variant} differs from variant 2 ({shape})"


def test_height_scan_variant5_without_planar_table_runs_as_variant4(cuda_device):
    """RoverPlaneCells.entries_planar is optional (ABI v2): a NULL pointer must not break variant 5."""
    v, f = _heightfield_mesh(120, 120, 0.2, 0.2, 4)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    gen = torch.Generator().manual_seed(12)
    n = 200
    pos = torch.cat([torch.rand(n, 2, generator=gen) * 20 + 2, torch.rand(n, 1, generator=gen) + 0.5], 1).to(cuda_device)
    quat = synthetic.quat_from_euler(torch.zeros(n), torch.zeros(n), (torch.rand(n, generator=gen) * 2 - 1) * np.pi).to(cuda_device)
    rays = ops.RayPattern.grid(cuda_device)
    ref = ops.height_scan(pos, quat, rays, grid, variant=5)
    saved = grid.cells_struct.entries_planar
    grid.cells_struct.entries_planar = None
    try:
        h = ops.height_scan(pos, quat, rays, grid, variant=5)
    finally:
        grid.cells_struct.entries_planar = saved
    assert torch.equal(h, ref)


@pytest.mark.parametrize("case", ["fused (variant 5 with the extra stores)", "fallback (no planar table)",
                                  "fallback (cells too small for a window)"])
def test_height_scan_obs_writes_fp32_heights_and_bf16_mirror(cuda_device, case):
    """rover_height_scan_obs: the fp32 heights are exactly rover_height_scan's, the bf16 buffer is the round-to-nearest
    mirror of obs[:, :4 + R] (head columns included), nothing else is touched -- on the fused path, on the generic
    fallback, and with border envs whose rays miss (-inf) or take the deferred path."""
    if case.endswith("window)"):
        v, f = _heightfield_mesh(300, 300, 0.05, 0.05, 6)
        span = 15.0
    else:
        v, f = _heightfield_mesh(150, 150, 0.2, 0.2, 5)
        span = 30.0
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    gen = torch.Generator().manual_seed(13)
    n = 333
    pos = torch.cat([torch.rand(n, 2, generator=gen) * (span + 2.0) - 1.0, torch.rand(n, 1, generator=gen) + 0.5], 1).to(cuda_device)
    quat = synthetic.quat_from_euler(torch.zeros(n), torch.zeros(n), (torch.rand(n, generator=gen) * 2 - 1) * np.pi).to(cuda_device)
    rays = ops.RayPattern.grid(cuda_device)
    ref = ops.height_scan(pos, quat, rays, grid, variant=2)
    assert torch.isinf(ref).any() and torch.isfinite(ref).any()
    obs = torch.full((n, 968), 7.0, device=cuda_device)
    obs[:, :4] = torch.randn(n, 4, generator=gen).to(cuda_device)
    head = obs[:, :4].clone()
    obs_bf = torch.full((n, 968), 9.0, dtype=torch.bfloat16, device=cuda_device)
    saved = grid.cells_struct.entries_planar
    if case.startswith("fallback (no planar"):
        grid.cells_struct.entries_planar = None
    try:
        ops.height_scan_obs(pos, quat, rays, grid, obs[:, :965], obs_bf[:, :965])
    finally:
        grid.cells_struct.entries_planar = saved
    torch.cuda.synchronize()
    assert torch.equal(obs[:, 4:965], ref) and torch.equal(obs[:, :4], head) and bool((obs[:, 965:] == 7.0).all())
    assert torch.equal(obs_bf[:, :965], obs[:, :965].to(torch.bfloat16))
    assert bool((obs_bf[:, 965:] == 9.0).all())


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
def test_height_scan_max_distance_and_empty(cuda_device, variant):
    v = np.array([[-5, -5, -95.0], [5, -5, -95.0], [0, 5, -95.0]], dtype=np.float32)
    f = np.array([[0, 1, 2]], dtype=np.int32)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    rays = ops.RayPattern.grid(cuda_device)
    pos = torch.tensor([[0.0, 0.0, -5.5], [0.0, 0.0, -4.5], [0.0, 0.0, -200.0]], device=cuda_device)
    quat = torch.tensor([[1.0, 0, 0, 0]] * 3, device=cuda_device)
    h = ops.height_scan(pos, quat, rays, grid, variant=variant).cpu()
    assert torch.isfinite(h[0, 480])  # t = 99.5 < 100
    assert torch.isinf(h[1]).all()  # t = 100.5: beyond max_distance
    assert torch.isinf(h[2]).all()  # triangle above the ray start (t < 0)
    # zero envs / empty mesh
    assert ops.height_scan(pos[:0], quat[:0], rays, grid, variant=variant).shape == (0, 961)
    empty = ops.ScanGridHandle.from_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), cuda_device)
    assert torch.isinf(ops.height_scan(pos, quat, rays, empty, variant=variant)).all()


def test_mdp_step_vs_oracle(world):
    """Six consecutive steps of the fused MDP kernels against the oracle (tests/mdp_parity.py)."""
    from mdp_parity import check_mdp_steps_vs_oracle

    check_mdp_steps_vs_oracle(world["dev"], world["n"], world["v"], world["tables"], SIZE, RES)


@pytest.mark.parametrize("n", [64, 5000, 16384 + 17])
def test_mdp_fused_single_launch_equals_two_launches(cuda_device, n):
    """rover_mdp_step (pre + post in one launch, reset rank by decoupled look-back) against the two-launch sequence over
    consecutive steps with carried state: every buffer bit-identical (ranks, spawn rows, resampled commands, rewards,
    statistics); then the fused launch replayed from a CUDA graph (the kernel maintains its own epoch / tickets)."""
    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    tables = TR.build_terrain_tables(v, f, n)
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, cuda_device)
    gen = torch.Generator().manual_seed(31)
    vt = torch.from_numpy(v)
    bufs = [ops.MdpBuffers.allocate(n, cuda_device) for _ in range(2)]
    for b in bufs:
        b.time_left.fill_(150.0)
        b.time_left[: n // 50] = 0.05  # time-based resamples as well
    obs = [torch.zeros(n, 965, device=cuda_device) for _ in range(2)]
    names = ("action", "prev_action", "pos_cmd_w", "heading_cmd_w", "pos_cmd_b", "heading_cmd_b", "time_left",
             "command_counter", "episode_length_buf", "episode_sums", "env_origins", "err_pos", "err_heading",
             "processed_actions", "joint_pos", "joint_vel", "reward", "term_rewards", "term_values", "terminated",
             "truncated", "term_flags", "reset_flags", "spawn_index", "stats")
    steps = [synthetic.make_step(n, gen, vt, SIZE, RES, margin=4.0).to(cuda_device) for _ in range(4)]
    total_resets = 0
    for k, st in enumerate(steps):
        pos = [st.root_pos_w.clone(), st.root_pos_w.clone()]
        quat = [st.root_quat_w.clone(), st.root_quat_w.clone()]
        ops.mdp_pre_step(bufs[0], params, st.actions, st.force_matrix_w)
        ops.mdp_post_step(bufs[0], params, th, pos[0], quat[0], st.spawn_perm, st.yaw_u, st.heading_u, st.theta_u, obs[0])
        ops.mdp_step(bufs[1], params, th, st.actions, st.force_matrix_w, pos[1], quat[1], st.spawn_perm, st.yaw_u,
                     st.heading_u, st.theta_u, obs[1])
        torch.cuda.synchronize()
        for name in names:
            assert torch.equal(getattr(bufs[0], name), getattr(bufs[1], name)), f"step {k}: {name}"
        assert torch.equal(pos[0], pos[1]) and torch.equal(quat[0], quat[1]) and torch.equal(obs[0], obs[1])
        total_resets += int(bufs[0].reset_flags.sum())
    assert total_resets > 0 and float(bufs[1].stats[13]) == total_resets
    # CUDA-graph replay of the fused launch: epoch and tickets live on the device
    st = steps[0]
    pos_g, quat_g = st.root_pos_w.clone(), st.root_quat_w.clone()
    g = torch.cuda.CUDAGraph()
    ops.mdp_step(bufs[1], params, th, st.actions, st.force_matrix_w, pos_g, quat_g, st.spawn_perm, st.yaw_u, st.heading_u,
                 st.theta_u, obs[1])  # warm-up outside capture (state advances on both sides below)
    ops.mdp_pre_step(bufs[0], params, st.actions, st.force_matrix_w)
    pos_e, quat_e = st.root_pos_w.clone(), st.root_quat_w.clone()
    ops.mdp_post_step(bufs[0], params, th, pos_e, quat_e, st.spawn_perm, st.yaw_u, st.heading_u, st.theta_u, obs[0])
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        ops.mdp_step(bufs[1], params, th, st.actions, st.force_matrix_w, pos_g, quat_g, st.spawn_perm, st.yaw_u,
                     st.heading_u, st.theta_u, obs[1])
    for _ in range(3):
        pos_g.copy_(st.root_pos_w)
        quat_g.copy_(st.root_quat_w)
        g.replay()
        pos_e.copy_(st.root_pos_w)
        quat_e.copy_(st.root_quat_w)
        ops.mdp_pre_step(bufs[0], params, st.actions, st.force_matrix_w)
        ops.mdp_post_step(bufs[0], params, th, pos_e, quat_e, st.spawn_perm, st.yaw_u, st.heading_u, st.theta_u, obs[0])
        torch.cuda.synchronize()
        for name in names:
            assert torch.equal(getattr(bufs[0], name), getattr(bufs[1], name)), f"graph replay: {name}"
        assert torch.equal(pos_e, pos_g) and torch.equal(obs[0], obs[1])


def test_mdp_operators_reject_wrong_dtype_device_and_shape(cuda_device):
    """The kernels read raw fp32 / int64: anything else must raise instead of being reinterpreted (no CPU path either)."""
    n = 64
    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    tables = TR.build_terrain_tables(v, f, n)
    params = ops.mdp_params(RoverEnvCfg(num_envs=n))
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, cuda_device)
    buf = ops.MdpBuffers.allocate(n, cuda_device)
    st = synthetic.make_step(n, torch.Generator().manual_seed(1), torch.from_numpy(v), SIZE, RES, margin=4.0).to(cuda_device)
    good = dict(root_pos_w=st.root_pos_w, root_quat_w=st.root_quat_w, spawn_perm=st.spawn_perm, yaw_u=st.yaw_u,
                heading_u=st.heading_u, theta_u=st.theta_u)
    ops.mdp_pre_step(buf, params, st.actions, st.force_matrix_w)
    ops.mdp_post_step(buf, params, th, **good)  # the unmodified arguments are accepted
    for key, bad in (("theta_u", st.theta_u.double()), ("yaw_u", st.yaw_u.half()), ("root_pos_w", st.root_pos_w.double()),
                     ("spawn_perm", st.spawn_perm.int()), ("root_quat_w", st.root_quat_w[:, :3].contiguous()),
                     ("heading_u", st.heading_u.cpu()), ("theta_u", st.theta_u[: n // 2])):
        with pytest.raises(RuntimeError):
            ops.mdp_post_step(buf, params, th, **{**good, key: bad})
        with pytest.raises(RuntimeError):
            ops.mdp_step(buf, params, th, st.actions, st.force_matrix_w, **{**good, key: bad})
    with pytest.raises(RuntimeError):
        ops.mdp_step(buf, params, th, st.actions.double(), st.force_matrix_w, **good)
    with pytest.raises(RuntimeError):
        ops.mdp_pre_step(buf, params, st.actions.double(), st.force_matrix_w)
    with pytest.raises(RuntimeError):
        ops.mdp_pre_step(buf, params, st.actions.cpu(), st.force_matrix_w)


def test_mdp_terms_against_reference_golden(cuda_device, golden_dir):
    """The fused kernel against outputs of the UNMODIFIED reference functions (tests/golden/terms.npz)."""
    z = np.load(os.path.join(golden_dir, "terms.npz"))
    n = len(z["in_actions"])
    cfg = RoverEnvCfg(num_envs=n)
    params = ops.mdp_params(cfg)
    buf = ops.MdpBuffers.allocate(n, cuda_device)
    t = lambda k: torch.from_numpy(z[k]).to(cuda_device)  # noqa: E731
    buf.action.copy_(t("in_prev_actions"))  # becomes prev_action inside the kernel
    buf.pos_cmd_b.copy_(t("in_pos_b"))
    buf.episode_length_buf.copy_(t("in_ep_len") - 1)  # the kernel increments before the terms read it
    ops.mdp_pre_step(buf, params, t("in_actions").contiguous(), t("in_force").contiguous())
    torch.cuda.synchronize()
    assert torch.equal(buf.processed_actions.cpu(), torch.from_numpy(z["ref_processed"]))
    torch.testing.assert_close(buf.joint_pos.cpu(), torch.from_numpy(z["ref_joint_pos"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(buf.joint_vel.cpu(), torch.from_numpy(z["ref_joint_vel"]), rtol=1e-5, atol=1e-6)
    w = torch.tensor(cfg.rewards.weights) * cfg.step_dt
    dist = torch.from_numpy(z["in_pos_b"])[:, :2].norm(dim=1)
    ok = ~(_near(dist, 0.18, 1e-6) | _near(dist, 11.0, 1e-5))
    ref_r = torch.from_numpy(z["ref_rewards"]) * w
    torch.testing.assert_close(buf.term_rewards.cpu()[ok], ref_r[ok], rtol=1e-5, atol=1e-9)
    flags = buf.term_flags.cpu().bool()
    assert torch.equal(flags[ok][:, 1:], torch.from_numpy(z["ref_terms"])[ok])
    assert torch.equal(flags[:, 0], torch.from_numpy(z["in_ep_len"]) >= 750)


def test_ackermann_variants_against_reference_golden(cuda_device, golden_dir):
    """the three action-term variants and the Exomy constants against outputs of the unmodified
    reference classes (tests/golden/terms.npz)."""
    from isaac_rover_orbit_b200 import mdp
    from isaac_rover_orbit_b200.config import exomy_action_cfg
    from isaac_rover_orbit_b200.env import RobotArticulation

    z = np.load(os.path.join(golden_dir, "terms.npz"))
    a = torch.from_numpy(z["in_actions"]).to(cuda_device)
    params = ops.mdp_params(RoverEnvCfg(num_envs=len(a)))
    for variant, kp, kv in ((2, "ref_joint_pos", "ref_joint_vel"), (1, "ref_v1_joint_pos", "ref_v1_joint_vel"),
                            (3, "ref_v3_joint_pos", "ref_v3_joint_vel")):
        processed, jp, jv = ops.ackermann(a, params, variant)
        assert torch.equal(processed.cpu(), torch.from_numpy(z["ref_processed"]))
        torch.testing.assert_close(jp.cpu(), torch.from_numpy(z[kp]), rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(jv.cpu(), torch.from_numpy(z[kv]), rtol=1e-5, atol=2e-6)
    ex = ops.mdp_params(RoverEnvCfg(num_envs=len(a), actions=exomy_action_cfg()))
    _, jp, jv = ops.ackermann(a, ex, 2)
    torch.testing.assert_close(jp.cpu(), torch.from_numpy(z["ref_exomy_joint_pos"]), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(jv.cpu(), torch.from_numpy(z["ref_exomy_joint_vel"]), rtol=1e-5, atol=2e-6)
    # AckermannAction3 through its own (cfg, robot, num_envs, device) API
    robot = RobotArticulation(len(a), cuda_device)
    ctl = mdp.AckermannAction3(RoverEnvCfg().actions, robot, len(a), cuda_device)
    ctl.process_actions(a)
    ctl.apply_actions()
    torch.testing.assert_close(robot.joint_pos_target.cpu(), torch.from_numpy(z["ref_v3_joint_pos"]), rtol=1e-5, atol=2e-6)
    assert ctl.action_dim == 2 and torch.equal(ctl.processed_actions.cpu(), torch.from_numpy(z["ref_processed"]))


@pytest.mark.parametrize("n,slices", [(1, 8), (7, 3), (257, 1), (1000, 8), (1000, 64)])
def test_height_scan_host_buffers_equal_the_device_call(world, n, slices):
    """rover_height_scan_host (poses and heights in page-locked HOST memory, the scan in slices whose heights travel back
    while the next slice runs) returns exactly what rover_height_scan returns for the same poses, whatever the slicing;
    a work area that is too small, unpinned host tensors and CUDA tensors in the host slots are refused."""
    dev = world["dev"]
    rays = ops.RayPattern.grid(dev)
    grid = world["grid"]
    gen = torch.Generator().manual_seed(900 + n)
    p, q = synthetic.make_poses(n, gen, torch.from_numpy(world["v"]), SIZE, RES, margin=4.0)
    ref = ops.height_scan(p.to(dev), q.to(dev), rays, grid)
    work = ops.HostScanWork(n, rays.starts.shape[0], dev)
    out = torch.full((n, rays.starts.shape[0]), float("nan")).pin_memory()
    for rep in range(2):  # the second call reuses the work area and the internal stream / events
        ops.height_scan_host(p.pin_memory(), q.pin_memory(), rays, grid, out, work, n_slices=slices)
        torch.cuda.synchronize()
        assert torch.equal(out, ref.cpu())
        out.fill_(float("nan"))
    if n == 7:
        with pytest.raises(RuntimeError, match="page-locked"):
            ops.height_scan_host(p, q, rays, grid, out, work)
        with pytest.raises(RuntimeError):
            ops.height_scan_host(p.to(dev), q.to(dev), rays, grid, out, work)
        with pytest.raises(RuntimeError, match="work area"):
            ops.height_scan_host(p.pin_memory(), q.pin_memory(), rays, grid, out, ops.HostScanWork(3, rays.starts.shape[0], dev))
        with pytest.raises(RuntimeError, match="n_slices"):
            ops.height_scan_host(p.pin_memory(), q.pin_memory(), rays, grid, out, work, n_slices=0)
