import sys, torch
sys.path.insert(0, ".")
from isaac_rover_orbit_b200.policy import (DeterministicNeuralNetwork, GaussianNeuralNetwork, alloc_obs, policy_value_forward)
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(7)
nets = []
for cls in (GaussianNeuralNetwork, DeterministicNeuralNetwork):
    net = cls(device=dev)
    net.load_state_dict({k: (torch.randn(v.shape, generator=gen) * (0.05 if v.dim() == 2 else 0.01)) for k, v in net.state_dict().items()})
    nets.append(net)
pol, val = nets
ok = True
for n in (1, 129, 1000, 5000, 18949, 65536):
    obs = alloc_obs(n, dev); obs.copy_(torch.randn(n, 965, device=dev) * 0.3)
    m0 = pol.compute({"states": obs})[0]; v0 = val.compute({"states": obs})[0]
    m1, v1 = policy_value_forward(pol, val, obs)
    torch.cuda.synchronize()
    same = torch.equal(m0, m1) and torch.equal(v0, v1)
    ok &= same
    print(f"n={n}: dual == separate: {same}  max|dm| {float((m0-m1).abs().max()):.2e} max|dv| {float((v0-v1).abs().max()):.2e}", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (4096, 16384, 65536):
    obs = alloc_obs(n, dev); obs.copy_(torch.randn(n, 965, device=dev) * 0.3)
    for name, fn in (("policy then value", lambda: (pol.compute({"states": obs}), val.compute({"states": obs}))), ("one pass", lambda: policy_value_forward(pol, val, obs))):
        ts = []
        for i in range(30):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[5:])
        print(f"n={n} {name}: median {ts[len(ts)//2]:.1f} us", flush=True)
print("IDENTICAL" if ok else "DIFFERENT")
