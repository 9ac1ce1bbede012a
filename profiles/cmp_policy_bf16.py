"""GPU check + timing: policy / value forward on the bf16 observation mirror vs the fp32 path (bitwise)."""
import sys

import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200.policy import (DeterministicNeuralNetwork, GaussianNeuralNetwork, alloc_obs,  # noqa: E402
                                            alloc_obs_bf16)

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(7)
nets = []
for cls in (GaussianNeuralNetwork, DeterministicNeuralNetwork):
    net = cls(device=dev)
    net.load_state_dict({k: (torch.randn(v.shape, generator=gen) * (0.05 if v.dim() == 2 else 0.01))
                         for k, v in net.state_dict().items()})
    nets.append(net)
ok = True
for n in (1, 129, 1000, 18949, 65536):
    obs = alloc_obs(n, dev)
    obs.copy_(torch.randn(n, 965, device=dev) * 0.3)
    ob = alloc_obs_bf16(n, dev)
    ob.copy_(obs)  # round-to-nearest-even, as the kernels' converter does
    ob[:, 964] = float("-inf")  # the ray the reference drops must not matter (here: a miss)
    for net in nets:
        a = net.compute({"states": obs})[0]
        b = net.compute_bf16({"states": ob})[0]
        torch.cuda.synchronize()
        same = torch.equal(a, b)
        ok &= same
        print(f"n={n} {type(net).__name__}: bf16 path == fp32 path bitwise: {same}  max|diff| {float((a - b).abs().max()):.3e}")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
obs = alloc_obs(65536, dev)
obs.copy_(torch.randn(65536, 965, device=dev) * 0.3)
ob = alloc_obs_bf16(65536, dev)
ob.copy_(obs)
for name, fn in (("fp32 obs", lambda: nets[0].compute({"states": obs})), ("bf16 obs", lambda: nets[0].compute_bf16({"states": ob}))):
    ts = []
    for i in range(25):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[5:])
    print(f"{name}: cfg-4 (65536 envs) median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
print("IDENTICAL" if ok else "DIFFERENT")
