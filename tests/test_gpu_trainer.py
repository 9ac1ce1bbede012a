"""GPU, SURVEY.md 8 (f-2): act -> step -> record through the trainer mirror on the real env and the tcgen05 policy."""
import pytest
import torch

from isaac_rover_orbit_b200 import terrain as TR
from isaac_rover_orbit_b200.config import RoverEnvCfg
from isaac_rover_orbit_b200.env import RoverEnv
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork
from isaac_rover_orbit_b200.trainer import RolloutAgent, RolloutMemory, SkrlSequentialLogTrainer, capture_steps

pytestmark = pytest.mark.gpu


def test_train_loop_records_the_rollout(cuda_device):
    n, steps = 160, 6
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=5)
    tables = TR.build_terrain_tables(v, f, n)
    gen = torch.Generator().manual_seed(3)
    drift = [(torch.rand(n, 3, generator=gen) * torch.tensor([3.0, 3.0, 0.0])).to(cuda_device) for _ in range(steps)]
    seen = {"obs": [], "act": [], "rew": []}

    def physics(env):  # stand-in for PhysX: the rover sits near its spawn point
        k = env.common_step_counter % steps
        pos = env.scene["robot"].data.root_pos_w
        pos.copy_(env._buf.env_origins + drift[k])
        pos[:, 2] = 0.3

    env = RoverEnv(RoverEnvCfg(num_envs=n), tables, cuda_device, physics=physics, seed=2)
    step0 = env.step

    def spy_step(actions):
        out = step0(actions)
        seen["act"].append(actions.clone())
        seen["obs"].append(out[0].clone())
        seen["rew"].append(out[1].clone())
        return out

    env.step = spy_step
    net = GaussianNeuralNetwork(device=cuda_device)
    g2 = torch.Generator().manual_seed(4)
    net.load_state_dict({k: torch.randn(t.shape, generator=g2) * (0.05 if t.dim() == 2 else 0.01)
                         for k, t in net.state_dict().items()})
    mem = RolloutMemory(memory_size=steps, num_envs=n, device=cuda_device)
    agent = RolloutAgent(net, mem)
    SkrlSequentialLogTrainer(env=env, agents=agent, cfg={"timesteps": steps, "disable_progressbar": True}).train()
    torch.cuda.synchronize()
    assert mem.filled and mem.memory_index == 0
    st, ac = mem.get_tensor_by_name("states"), mem.get_tensor_by_name("actions")
    for t in range(steps):
        assert torch.equal(ac[t], seen["act"][t]), "the recorded action is the one the env stepped with"
        assert torch.equal(mem.get_tensor_by_name("rewards")[t, :, 0], seen["rew"][t])
        if t + 1 < steps:
            assert torch.equal(st[t + 1], seen["obs"][t]), "states of step t+1 = observation returned by step t"
    assert torch.isfinite(st[:, :, :4]).all() and (ac.abs() <= 1.0).all()
    assert torch.isfinite(mem.get_tensor_by_name("log_prob")).all()
    assert any(k.startswith("EpisodeInfo / ") for k in agent.tracking_data)


def test_rollout_agent_with_value_network_records_the_values(cuda_device):
    """skrl PPO.record_transition evaluates the value network on the states ``act`` saw and stores ``values``; the agent
    does both networks in one pass over the observation.  The recorded values equal ``value.compute(states)`` bit for
    bit, and the actions equal those of the policy-only agent under the same torch seed."""
    from isaac_rover_orbit_b200.policy import DeterministicNeuralNetwork

    n, steps = 200, 4
    g2 = torch.Generator().manual_seed(4)
    net, vnet = GaussianNeuralNetwork(device=cuda_device), DeterministicNeuralNetwork(device=cuda_device)
    for m in (net, vnet):
        m.load_state_dict({k: torch.randn(t.shape, generator=g2) * (0.05 if t.dim() == 2 else 0.01)
                           for k, t in m.state_dict().items()})
    mem = RolloutMemory(memory_size=steps, num_envs=n, device=cuda_device)
    action_in = torch.zeros(n, 2, device=cuda_device)  # stands for env.action_input
    agent, plain = RolloutAgent(net, mem, value=vnet, action_out=action_in), RolloutAgent(net)
    assert "values" in mem.get_tensor_names()
    for t in range(steps):
        states = torch.randn(n, 965, generator=g2).to(cuda_device) * 0.3
        torch.manual_seed(100 + t)
        a1, lp1, out1 = agent.act(states, t, steps)
        torch.manual_seed(100 + t)
        a0, lp0, out0 = plain.act(states, t, steps)
        assert torch.equal(a1, a0) and torch.equal(lp1, lp0) and torch.equal(out1["mean_actions"], out0["mean_actions"])
        assert a1 is action_in  # sampled in place: nothing to copy on the way to the env
        agent.record_transition(states, a1, torch.zeros(n, device=cuda_device), states, torch.zeros(n, dtype=torch.bool, device=cuda_device),
                                torch.zeros(n, dtype=torch.bool, device=cuda_device), {}, t, steps)
        assert torch.equal(mem.get_tensor_by_name("values")[t], vnet.compute({"states": states})[0])


def test_eval_loop_runs_on_the_real_env_and_agent(cuda_device):
    """``SkrlSequentialLogTrainer.eval()`` (the reference's eval.py path, skrl_utils.py:150-206) with the
    shipped ``RoverEnv`` + ``RolloutAgent``: records through the base agent (tracking only, no rollout write) and calls
    ``env.render()`` unless headless."""
    n, steps = 96, 5
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=5)
    tables = TR.build_terrain_tables(v, f, n)
    env = RoverEnv(RoverEnvCfg(num_envs=n), tables, cuda_device, seed=2)
    net = GaussianNeuralNetwork(device=cuda_device)
    g2 = torch.Generator().manual_seed(4)
    net.load_state_dict({k: torch.randn(t.shape, generator=g2) * (0.05 if t.dim() == 2 else 0.01)
                         for k, t in net.state_dict().items()})
    mem = RolloutMemory(memory_size=steps, num_envs=n, device=cuda_device)
    agent = RolloutAgent(net, mem)
    for headless in (False, True):
        SkrlSequentialLogTrainer(env=env, agents=agent, cfg={"timesteps": steps, "headless": headless}).eval()
    torch.cuda.synchronize()
    assert mem.memory_index == 0 and not mem.filled, "evaluation must not write the rollout memory"
    assert not agent.training
    stats = agent.finished_episode_stats()
    assert stats["episodes"] >= 0 and env.common_step_counter == 2 * steps


def test_capture_steps_replays_the_enqueued_work(cuda_device):
    acc = torch.zeros(4, device=cuda_device)
    inc = [torch.full((4,), float(i + 1), device=cuda_device) for i in range(3)]
    replay = capture_steps(lambda i: acc.add_(inc[i]), n_variants=3, warmup=1)
    base = acc.clone()  # warm-up already added 1 + 2 + 3
    for i in range(6):
        replay(i)
    torch.cuda.synchronize()
    assert torch.equal(acc, base + 12.0)


def test_recorder_collects_from_the_real_env_on_the_gpu(cuda_device):
    """f-4: recorder/orbit.py:24-36 over RoverEnv with CUDA tensors -- staging ring on the device, page-locked transfer
    every `chunk_steps` steps; the rows equal what a per-step `.cpu()` copy (the reference's way) collects."""
    import types

    import numpy as np
    from isaac_rover_orbit_b200.recorder import HDF5DataRecorder, SequentialCollector
    from oracle.recorder import FakeH5, reference_recorder_run

    n, steps = 96, 11
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=5)
    tables = TR.build_terrain_tables(v, f, n)
    gen = torch.Generator().manual_seed(8)
    drift = [(torch.rand(n, 3, generator=gen) * torch.tensor([9.0, 9.0, 0.0])).to(cuda_device) for _ in range(steps)]

    def physics(env):
        pos = env.scene["robot"].data.root_pos_w
        pos.copy_(env._buf.env_origins + drift[env.common_step_counter % steps])
        pos[:, 2] = 0.3

    env = RoverEnv(RoverEnvCfg(num_envs=n), tables, cuda_device, physics=physics, seed=4)
    net = GaussianNeuralNetwork(device=cuda_device)
    seen = []
    step0 = env.step

    def spy_step(action):  # what the reference's recorder would have been handed, copied to the host per step
        obs_before = env.obs_buf.clone()
        out = step0(action)
        seen.append((obs_before.cpu().numpy(), action.cpu().numpy(), out[1].cpu().numpy(), out[2].cpu().numpy(), {}))
        return out

    env.step = spy_step
    spaces = types.SimpleNamespace(observation_space=types.SimpleNamespace(shape=(965,), dtype=np.float32),
                                   action_space=types.SimpleNamespace(shape=(2,), dtype=np.float32))
    h5, ref = FakeH5(), FakeH5()
    rec = HDF5DataRecorder("roll", n, spaces, max_rows=400, chunk_steps=4, backend=h5)
    SequentialCollector(env, net, rec, predict_fn=lambda m, o: m.compute({"states": o})[0], num_episodes=steps).collect()
    rec.close()
    reference_recorder_run(ref, "roll", n, 965, 2, np.float32, np.float32, {}, 400, seen)
    assert list(h5.files) == list(ref.files) and len(h5.files) >= 2
    for name in ref.files:
        assert h5.files[name]["__attrs__"] == ref.files[name]["__attrs__"]
        for k in ("observations", "actions", "rewards", "terminated"):
            assert np.array_equal(h5.files[name][k].data, ref.files[name][k].data), (name, k)
