"""GPU: tcgen05 policy forward against (i) a torch emulation of the kernel's numerics (bf16 operands, fp32
accumulation, bf16 activations between layers) and (ii) the fp32 oracle / the reference network's own outputs
(tests/golden/policy.npz, generated from best_agent.pt by the unmodified reference model)."""
import os

import numpy as np
import pytest
import torch

from isaac_rover_orbit_b200.policy import (DeterministicNeuralNetwork, GaussianNeuralNetwork, WEIGHT_KEYS, alloc_obs,
                                            alloc_obs_bf16)

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def emulate_bf16(obs, sd):
    """The kernel's arithmetic: bf16(operands) x bf16(weights) accumulated in fp32, fp32 bias + activation."""
    act = lambda x: torch.nn.functional.leaky_relu(x, 0.01)  # noqa: E731
    lin = lambda k, x: _bf(x) @ _bf(sd[k + ".weight"]).T + sd[k + ".bias"]  # noqa: E731
    e = act(lin(WEIGHT_KEYS[1], act(lin(WEIGHT_KEYS[0], obs[:, 3:964]))))
    h = torch.cat([obs[:, 0:4], e], dim=1)
    for k in WEIGHT_KEYS[2:5]:
        h = act(lin(k, h))
    return torch.tanh(lin(WEIGHT_KEYS[5], h))


@pytest.fixture(scope="module")
def golden(golden_dir):
    from oracle import policy as OP

    z = np.load(os.path.join(golden_dir, "policy.npz"))
    return z, OP.load_golden_weights(z)


def test_policy_forward_reference_checkpoint(cuda_device, golden):
    from oracle import policy as OP

    z, sd = golden
    net = GaussianNeuralNetwork(device=cuda_device)
    net.load_state_dict(sd, strict=True)
    assert set(net.state_dict()) == set(sd)
    obs = torch.from_numpy(z["in_obs"])
    buf = alloc_obs(obs.shape[0], cuda_device)
    buf.copy_(obs)
    mean, log_std, _ = net.compute({"states": buf}, role="policy")
    torch.cuda.synchronize()
    mean = mean.cpu()
    assert torch.equal(log_std.cpu(), sd["log_std_parameter"])
    emu = emulate_bf16(obs, sd)
    torch.testing.assert_close(mean, emu, rtol=0, atol=4e-3)          # same numerics: bf16 rounding noise only
    ref = torch.from_numpy(z["ref_mean"])                             # the reference network in fp32
    # bf16 operands cannot meet 1e-5.  On these inputs the bf16 emulation itself sits 7.1e-3 from the
    # fp32 network (max; mean 1.4e-3) -- pure operand rounding, measured on the CPU -- so the kernel is held to 1e-2
    # and to the emulation's own distance (+ the 4e-3 accumulation-order allowance above), not to a looser bound.
    emu_err = float((emu - ref).abs().max())
    assert emu_err < 8e-3, emu_err
    assert (mean - ref).abs().max() < 1e-2, (mean - ref).abs().max()
    assert (mean - ref).abs().max() <= emu_err + 4e-3
    torch.testing.assert_close(OP.policy_mean(obs, sd), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 4096 + 37])
def test_policy_forward_tiles_and_tails(cuda_device, golden, n):
    _, sd = golden
    net = GaussianNeuralNetwork(device=cuda_device)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(n)
    obs = torch.cat([torch.rand(n, 2, generator=g) * 2 - 1, torch.rand(n, 1, generator=g) * 1.3,
                     torch.rand(n, 1, generator=g) * 2 - 1, torch.randn(n, 961, generator=g) * 0.3], dim=1)
    # an unaligned [N,965] tensor exercises the re-homing path of compute(); the last ray must not matter
    states = obs.to(cuda_device)
    mean, _, _ = net.compute({"states": states})
    states2 = states.clone()
    states2[:, 964] = 123.0
    mean2, _, _ = net.compute({"states": states2})
    torch.cuda.synchronize()
    assert torch.equal(mean, mean2), "obs[:, -1] is not an input of the policy (models.py:95)"
    torch.testing.assert_close(mean.cpu(), emulate_bf16(obs, sd), rtol=0, atol=4e-3)


@pytest.mark.parametrize("n", [1, 129, 3000, 18944 + 5, 40000])
def test_policy_kernels_agree_bitwise(cuda_device, golden, n, monkeypatch):
    """The warp-specialised kernel (default) and the tile-serial v1 (ROVER_POLICY_KERNEL=v1) run the same arithmetic in
    the same order: identical means for every tiling (tile height varies with N: whole rounds of the grid)."""
    _, sd = golden
    net = GaussianNeuralNetwork(device=cuda_device)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(100 + n)
    obs = alloc_obs(n, cuda_device)
    obs.copy_((torch.randn(n, 965, generator=g) * 0.4).to(cuda_device))
    monkeypatch.setenv("ROVER_POLICY_KERNEL", "v1")
    m1 = net.compute({"states": obs})[0].clone()
    monkeypatch.setenv("ROVER_POLICY_KERNEL", "ws")
    m2 = net.compute({"states": obs})[0].clone()
    torch.cuda.synchronize()
    assert torch.equal(m1, m2)
    assert torch.isfinite(m2).all()


def test_gaussian_act(cuda_device, golden):
    from oracle import policy as OP

    z, sd = golden
    net = GaussianNeuralNetwork(device=cuda_device)
    net.load_state_dict(sd)
    obs = torch.from_numpy(z["in_obs"]).to(cuda_device)
    g = torch.Generator().manual_seed(3)
    eps = torch.randn(obs.shape[0], 2, generator=g) * 3
    actions, log_prob, out = net.act({"states": obs}, eps=eps.to(cuda_device))
    a_ref, lp_ref = OP.gaussian_act(out["mean_actions"].cpu(), sd["log_std_parameter"], eps)
    torch.testing.assert_close(actions.cpu(), a_ref, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(log_prob.cpu(), lp_ref, rtol=1e-4, atol=1e-4)
    assert actions.abs().max() <= 1.0 and (actions.abs() == 1.0).any()
    # the same sample written straight into a caller's buffer (the env's action input): no new tensor, same numbers
    buf = torch.full_like(actions, 7.0)
    a2, lp2, _ = net.act({"states": obs}, eps=eps.to(cuda_device), out_actions=buf)
    assert a2 is buf and torch.equal(buf, actions) and torch.equal(lp2, log_prob)


def test_policy_rejects_other_shapes(cuda_device):
    with pytest.raises(ValueError):
        GaussianNeuralNetwork(device=cuda_device, mlp_layers=(64, 64))
    with pytest.raises(ValueError):
        GaussianNeuralNetwork(device=cuda_device, mlp_activation="relu")
    net = GaussianNeuralNetwork(device=cuda_device)
    with pytest.raises(KeyError):
        net.load_state_dict({"mlp.0.weight": torch.zeros(256, 64)}, strict=True)


def emulate_value_bf16(obs, sd):
    act = lambda x: torch.nn.functional.leaky_relu(x, 0.01)  # noqa: E731
    lin = lambda k, x: _bf(x) @ _bf(sd[k + ".weight"]).T + sd[k + ".bias"]  # noqa: E731
    e = act(lin(WEIGHT_KEYS[1], act(lin(WEIGHT_KEYS[0], obs[:, 3:964]))))
    h = torch.cat([obs[:, 0:4], e], dim=1)
    for k in WEIGHT_KEYS[2:5]:
        h = act(lin(k, h))
    return lin(WEIGHT_KEYS[5], h)


@pytest.mark.parametrize("kernel", ["ws", "v1"])
def test_value_forward_reference_checkpoint(cuda_device, golden_dir, kernel, monkeypatch):
    """DeterministicNeuralNetwork (models.py:105-162) with the ``value`` weights of best_agent.pt:
    bf16-emulation tolerance 2e-2 abs on values of magnitude ~10 (unbounded linear output), 1e-1 vs the fp32 reference."""
    from oracle import policy as OP

    monkeypatch.setenv("ROVER_POLICY_KERNEL", kernel)
    z = np.load(os.path.join(golden_dir, "value.npz"))
    sd = OP.load_golden_weights(z)
    net = DeterministicNeuralNetwork(device=cuda_device)
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd)
    obs = torch.from_numpy(z["in_obs"])
    value, outputs = net.compute({"states": obs.to(cuda_device)}, role="value")
    assert value.shape == (obs.shape[0], 1) and outputs == {}
    ref = torch.from_numpy(z["ref_value"])
    scale = float(ref.abs().max())
    torch.testing.assert_close(value.cpu(), emulate_value_bf16(obs, sd), rtol=0, atol=2e-3 * max(scale, 1.0))
    torch.testing.assert_close(value.cpu(), ref, rtol=0, atol=2e-2 * max(scale, 1.0))
    v2, none, _ = net.act({"states": obs.to(cuda_device)})
    assert none is None and torch.equal(v2, value)
    n = 5000  # several tiles, ragged tail
    g = torch.Generator().manual_seed(5)
    big = torch.cat([torch.rand(n, 4, generator=g) * 2 - 1, torch.randn(n, 961, generator=g) * 0.3], dim=1)
    vb = net.compute({"states": big.to(cuda_device)})[0]
    emu = emulate_value_bf16(big, sd)
    # same arithmetic, different fp32 summation order inside the MMAs; bf16 re-rounding of the activations amplifies
    # it: 4e-3 of the output range (the policy test uses 4e-3 on tanh outputs in [-1, 1])
    torch.testing.assert_close(vb.cpu(), emu, rtol=0, atol=4e-3 * max(float(emu.abs().max()), 1.0))


@pytest.mark.parametrize("n", [1, 129, 3000, 18944 + 5])
def test_bf16_observation_path_is_bit_identical(cuda_device, golden, golden_dir, n):
    """rover_policy_forward_bf16 / rover_value_forward_bf16: the SWIZZLE_128B TMA tile of the bf16 observation is the
    MMA operand; the fp32 entry points round the same observation to the same bf16 values, so the outputs are equal
    bit for bit -- also when the column the reference drops (964) holds -inf (a missed ray)."""
    from oracle import policy as OP

    _, sd = golden
    pol = GaussianNeuralNetwork(device=cuda_device)
    pol.load_state_dict(sd)
    val = DeterministicNeuralNetwork(device=cuda_device)
    val.load_state_dict(OP.load_golden_weights(np.load(os.path.join(golden_dir, "value.npz"))))
    g = torch.Generator().manual_seed(200 + n)
    obs = alloc_obs(n, cuda_device)
    obs.copy_((torch.randn(n, 965, generator=g) * 0.4).to(cuda_device))
    ob = alloc_obs_bf16(n, cuda_device)
    ob.copy_(obs)
    ob[:, 964] = float("-inf")
    for net in (pol, val):
        a = net.compute({"states": obs})[0]
        b = net.compute_bf16({"states": ob})[0]
        torch.cuda.synchronize()
        assert torch.equal(a, b) and torch.isfinite(b).all(), type(net).__name__
    with pytest.raises(RuntimeError):
        pol.compute_bf16({"states": obs})  # fp32 tensor on the bf16 entry point


@pytest.mark.parametrize("n", [1, 129, 3000, 18944 + 5, 40000])
def test_policy_and_value_in_one_pass_are_bit_identical(cuda_device, golden, golden_dir, n):
    """rover_policy_value_forward (one pass for policy + value, models.py:89-102 + :151-162): the
    observation tile meets both heightmap encoders, the layer group runs once per network (layer 2 as two N halves
    through a 192-column accumulator).  Same arithmetic per network, so mean and value equal the two single passes bit
    for bit -- with the weights of best_agent.pt, one tile, ragged tiles, several tiles per SM, a non-aligned tensor."""
    from oracle import policy as OP
    from isaac_rover_orbit_b200.policy import policy_value_forward

    _, sd = golden
    pol = GaussianNeuralNetwork(device=cuda_device)
    pol.load_state_dict(sd)
    val = DeterministicNeuralNetwork(device=cuda_device)
    val.load_state_dict(OP.load_golden_weights(np.load(os.path.join(golden_dir, "value.npz"))))
    g = torch.Generator().manual_seed(300 + n)
    obs = alloc_obs(n, cuda_device)
    obs.copy_((torch.randn(n, 965, generator=g) * 0.4).to(cuda_device))
    obs[:, 964] = float("-inf")  # the column the reference's slicing drops (a missed ray) must not matter
    mean, value = policy_value_forward(pol, val, obs)
    torch.cuda.synchronize()
    assert mean.shape == (n, 2) and value.shape == (n, 1)
    assert torch.equal(mean, pol.compute({"states": obs})[0]) and torch.equal(value, val.compute({"states": obs})[0])
    assert torch.isfinite(mean).all() and torch.isfinite(value).all()
    if n == 129:
        packed_dense = obs.contiguous()  # [N, 965] rows are not 16-byte aligned: re-homed once, like compute()
        m2, v2 = policy_value_forward(pol, val, packed_dense[:, :965].clone())
        assert torch.equal(m2, mean) and torch.equal(v2, value)
        with pytest.raises(RuntimeError):
            policy_value_forward(pol, val, obs.cpu())
        with pytest.raises(RuntimeError):
            torch.ops.rover_b200.policy_value_forward(obs.double(), pol._packed, val._packed)


def test_bf16_only_observation_scan_changes_nothing_downstream(cuda_device, golden):
    """rover_height_scan_obs_bf16: the scan stores only the bf16 observation.  The mirror equals the one the fp32 + bf16
    scan writes, the fp32 height columns are left alone, and the policy's means equal those on the fp32 observation
    bit for bit (it rounds fp32 observations to these very bf16 values)."""
    from isaac_rover_orbit_b200 import ops, synthetic
    from isaac_rover_orbit_b200 import terrain as TR

    _, sd = golden
    net = GaussianNeuralNetwork(device=cuda_device)
    net.load_state_dict(sd)
    v, f = TR.make_synthetic_terrain(48.0, 0.2, seed=3)
    grid = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    rays = ops.RayPattern.grid(cuda_device)
    n = 777
    g = torch.Generator().manual_seed(12)
    p, q = (t.to(cuda_device) for t in synthetic.make_poses(n, g, torch.from_numpy(v), 48.0, 0.2, margin=4.0))
    head = (torch.rand(n, 4, generator=g) * 2 - 1).to(cuda_device)
    obs_a, obs_b = alloc_obs(n, cuda_device), alloc_obs(n, cuda_device)
    bf_a, bf_b = alloc_obs_bf16(n, cuda_device), alloc_obs_bf16(n, cuda_device)
    for o in (obs_a, obs_b):
        o.fill_(123.0)
        o[:, :4] = head
    ops.height_scan_obs(p, q, rays, grid, obs_a, bf_a)
    ops.height_scan_obs(p, q, rays, grid, obs_b, bf_b, bf16_only=True)
    torch.cuda.synchronize()
    assert torch.equal(bf_a[:, :965].view(torch.int16), bf_b[:, :965].view(torch.int16))
    assert bool((obs_b[:, 4:] == 123.0).all()) and torch.equal(obs_b[:, :4], head)  # fp32 heights never written
    m_fp32 = net.compute({"states": obs_a})[0]
    m_bf16 = net.compute_bf16({"states": bf_b})[0]
    assert torch.equal(m_fp32, m_bf16)
