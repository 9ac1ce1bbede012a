"""GPU: the torch custom-operator boundary (``torch.ops.rover_b200.*``): schemas hold (``opcheck``:
declared mutations are the only mutations, fake implementations agree with the real ones on shapes / dtypes), the ops
run under ``torch.cuda.graph``, and the step through ``RoverEnv`` is graph-replayable with identical results."""
import types

import pytest
import torch

from isaac_rover_orbit_b200 import mdp, ops, synthetic, torch_ops
from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.env import RoverEnv

pytestmark = pytest.mark.gpu
SIZE, RES, N = 48.0, 0.2, 320


@pytest.fixture(scope="module")
def world(cuda_device):
    v, f = TR.make_synthetic_terrain(SIZE, RES, seed=3)
    tables = TR.build_terrain_tables(v, f, N)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    return dict(v=v, f=f, tables=tables, grid=grid, dev=cuda_device)


def test_every_abi_entry_point_is_a_registered_operator():
    for name in torch_ops.OPS:
        op = getattr(torch.ops.rover_b200, name)
        assert op.default._schema.name == f"rover_b200::{name}"
    mutated = {n for n in torch_ops.OPS if getattr(torch.ops.rover_b200, n).default._schema.is_mutable}
    assert {"height_scan_out", "height_scan_obs", "height_scan_host", "gaussian_act_out", "mdp_pre_step", "mdp_post_step",
            "mdp_step", "policy_pack"} <= mutated


def test_opcheck_height_scan_and_policy(world):
    dev, grid = world["dev"], world["grid"]
    rays = ops.RayPattern.grid(dev)
    gen = torch.Generator().manual_seed(1)
    p, q = synthetic.make_poses(33, gen, torch.from_numpy(world["v"]), SIZE, RES, margin=4.0)
    args = (p.to(dev), q.to(dev), rays.starts, rays.box_t, grid.desc, grid.cells_desc, 100.0, 0.26878, 5)
    for tests in ("test_schema", "test_faketensor"):
        torch.library.opcheck(torch.ops.rover_b200.height_scan.default, args, test_utils=tests)
        torch.library.opcheck(torch.ops.rover_b200.height_scan_hits.default, args[:-1] + (4,), test_utils=tests)
    out = torch.zeros(33, 965, device=dev)
    torch.library.opcheck(torch.ops.rover_b200.height_scan_out.default, args + (out[:, 4:],), test_utils="test_schema")
    from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs

    net = GaussianNeuralNetwork(device=dev)
    net._pack()
    obs = alloc_obs(40, dev)
    obs.copy_(torch.randn(40, 965, device=dev) * 0.1)
    for tests in ("test_schema", "test_faketensor"):
        torch.library.opcheck(torch.ops.rover_b200.policy_forward.default, (obs, net._packed, False), test_utils=tests)
        torch.library.opcheck(torch.ops.rover_b200.policy_value_forward.default, (obs, net._packed, net._packed), test_utils=tests)
        torch.library.opcheck(torch.ops.rover_b200.gaussian_act.default,
                              (torch.zeros(40, 2, device=dev), torch.zeros(2, device=dev), torch.randn(40, 2, device=dev)),
                              test_utils=tests)


def test_opcheck_mdp_ops_declare_their_mutations(world):
    dev, tables = world["dev"], world["tables"]
    cfg = RoverEnvCfg(num_envs=N)
    params = ops.mdp_params(cfg)
    th = ops.TerrainTablesHandle(tables.heightmap, tables.safe_mask, tables.offset_xy, tables.spawn_table,
                                 tables.resolution, dev)
    buf = ops.MdpBuffers.allocate(N, dev)
    buf.time_left.fill_(150.0)
    s = synthetic.make_step(N, torch.Generator().manual_seed(2), torch.from_numpy(world["v"]), SIZE, RES, margin=4.0).to(dev)
    torch.library.opcheck(torch.ops.rover_b200.mdp_pre_step.default,
                          (s.actions, s.force_matrix_w, ops.params_desc(params), buf.state_list(), buf.out_list(), 3),
                          test_utils="test_schema")
    rng = ops.ResetRng(5, dev)
    obs = torch.zeros(N, 968, device=dev)[:, :965]
    torch.library.opcheck(torch.ops.rover_b200.mdp_post_step.default,
                          (s.root_pos_w, s.root_quat_w, ops.params_desc(params), buf.state_list(), buf.out_list(), th.desc,
                           [], rng.state, 16, buf.spawn_index, buf.stats, buf.scratch, buf.log, obs, 127, None),
                          test_utils="test_schema")
    torch.library.opcheck(torch.ops.rover_b200.ackermann.default, (s.actions, ops.params_desc(params)),
                          test_utils=("test_schema", "test_faketensor"))


def test_cpu_tensors_have_no_implementation(world):
    rays, grid = ops.RayPattern.grid(world["dev"]), world["grid"]
    with pytest.raises(RuntimeError):  # NotImplementedError: no CPU kernel is registered -- there is no fallback
        torch.ops.rover_b200.height_scan(torch.zeros(2, 3), torch.zeros(2, 4), rays.starts.cpu(), rays.box_t, grid.desc,
                                         grid.cells_desc, 100.0, 0.26878, 5)
    with pytest.raises(RuntimeError, match="descriptor|bytes"):
        torch.ops.rover_b200.height_scan(torch.zeros(2, 3, device=world["dev"]), torch.zeros(2, 4, device=world["dev"]),
                                         rays.starts, rays.box_t, torch.zeros(8, dtype=torch.uint8), None, 100.0, 0.2, 0)


def _make_env(world, graph, seed=11):
    dev, tables = world["dev"], world["tables"]
    gen = torch.Generator().manual_seed(8)
    drift = [(torch.rand(N, 3, generator=gen) * torch.tensor([3.0, 3.0, 0.0])).to(dev) for _ in range(4)]
    forces = [synthetic.make_step(N, gen, torch.from_numpy(world["v"]), SIZE, RES).force_matrix_w.to(dev) for _ in range(4)]
    tick = torch.zeros(1, dtype=torch.int64, device=dev)

    def physics(env):  # graph-safe stand-in for PhysX: device-side step counter selects the drift / contact set
        k = (tick % 4).expand(N)
        d = torch.stack(drift)[k, torch.arange(N, device=dev)]
        pos = env.scene["robot"].data.root_pos_w
        pos.copy_(env._buf.env_origins + d)
        pos[:, 2] = 0.3
        env.scene.sensors["contact_sensor"].data.force_matrix_w.copy_(torch.stack(forces)[k, torch.arange(N, device=dev)])
        tick.add_(1)

    env = RoverEnv(RoverEnvCfg(num_envs=N), tables, dev, physics=physics, seed=seed, physics_needs_targets=graph != "eager4")
    env.reset()
    if graph == "graph":
        env.enable_cuda_graph(warmup=2)
    else:
        a0 = torch.zeros(N, 2, device=dev)
        for _ in range(2):  # the same two warm-up steps the capture takes
            env.step(a0)
    return env


def test_env_step_graph_replay_equals_eager(world):
    dev = world["dev"]
    eager, graph, eager4 = _make_env(world, "eager"), _make_env(world, "graph"), _make_env(world, "eager4")
    gen = torch.Generator().manual_seed(4)
    total = 0.0
    for k in range(12):
        a = (torch.rand(N, 2, generator=gen) * 2 - 1).to(dev)
        outs = [e.step(a) for e in (eager, graph, eager4)]
        torch.cuda.synchronize()
        for o in outs[1:]:
            assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]), k
            assert torch.equal(outs[0][2], o[2]) and torch.equal(outs[0][3], o[3])
        assert outs[0][2].dtype == torch.bool and outs[0][3].dtype == torch.bool
        for name in ("pos_cmd_w", "episode_length_buf", "episode_sums", "env_origins", "log", "action", "prev_action"):
            assert torch.equal(getattr(eager._buf, name), getattr(graph._buf, name)), (k, name)
        total += float(eager._buf.log[13])
    assert total > 0 and eager.common_step_counter == graph.common_step_counter


def test_env_step_launch_count(world):
    """``RoverEnv.step`` = 2 of our launches (the whole MDP step in one, then the height scan) when nothing has to run
    between the action term and the reward terms, 4 when the physics callable must see the joint targets -- plus ONE
    copy of the action into the term's ``raw_actions``; no generator, reduction or logging kernels."""
    from torch.profiler import ProfilerActivity, profile

    dev, tables = world["dev"], world["tables"]
    for physics, want in ((None, 2), (lambda env: None, 4)):
        env = RoverEnv(RoverEnvCfg(num_envs=N), tables, dev, physics=physics, seed=1)
        env.reset()
        a = torch.zeros(N, 2, device=dev)
        env.step(a)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            env.step(a)
            torch.cuda.synchronize()
        kernels = [e.name for e in prof.events()
                   if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
                   and "memset" not in e.name.lower()]
        ours = [k for k in kernels if "rover::" in k]
        assert len(ours) == want, kernels
        assert len(kernels) <= want + 1, kernels


def test_term_functions_check_their_arguments_and_staleness(world):
    env = _make_env(world, "eager")
    a = torch.zeros(N, 2, device=world["dev"])
    env.step(a)
    cfg_ = lambda name: types.SimpleNamespace(name=name)  # noqa: E731
    assert mdp.reached_target(env, "target_pose", 0.18).shape == (N,)
    assert mdp.is_success(env, "target_pose", 0.18).dtype == torch.bool
    with pytest.raises(ValueError, match="threshold"):
        mdp.reached_target(env, "target_pose", 0.5)
    with pytest.raises(ValueError, match="threshold"):
        mdp.far_from_target(env, "target_pose", 9.0)
    with pytest.raises(KeyError):
        mdp.distance_to_target_reward(env, "no_such_command")
    with pytest.raises(KeyError):
        mdp.collision_penalty(env, cfg_("no_such_sensor"), 1.0)
    mdp.collision_penalty(env, cfg_("contact_sensor"), 123.0)  # ignored by the reference as well (hard-coded > 1)
    env.action_manager.process_action(a)  # a new action: the reward columns now describe the previous one
    with pytest.raises(RuntimeError, match="current state"):
        mdp.oscillation_penalty(env)
    with pytest.raises(RuntimeError, match="current state"):
        mdp.time_out(env)


def test_extras_log_persists_between_resets(world):
    """on steps without a reset the reference keeps the last reset's extras; so does the kernel-written log."""
    dev = world["dev"]
    env = RoverEnv(RoverEnvCfg(num_envs=8), world["tables"], dev, seed=3)
    env.reset()
    robot = env.scene["robot"].data
    cmd = env.command_manager.get_term("target_pose")
    a = torch.full((8, 2), 0.3, device=dev)
    robot.root_pos_w.copy_(cmd.pos_command_w)  # on the target: every env succeeds and resets
    env.step(a)
    env.step(a)
    torch.cuda.synchronize()
    first = {k: float(v) for k, v in env.extras["log"].items()}
    assert env.episode_log()["num_resets"] == 8 and first["Episode Termination/is_success"] == 8
    robot.root_pos_w.copy_(cmd.pos_command_w - torch.tensor([4.0, 0.0, 0.0], device=dev))  # 4 m away: nothing resets
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize()
    assert not bool(env.reset_buf.any())
    assert {k: float(v) for k, v in env.extras["log"].items()} == first
    assert env.extras["episode"] is env.extras["log"]


def test_env_rejects_a_spawn_table_smaller_than_num_envs(world):
    with pytest.raises(ValueError, match="spawn table"):
        RoverEnv(RoverEnvCfg(num_envs=4 * N), world["tables"], world["dev"])
