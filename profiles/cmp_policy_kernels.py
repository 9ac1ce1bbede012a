"""GPU check + timing: warp-specialised policy kernel (default) vs the tile-serial v1 (ROVER_POLICY_KERNEL=v1)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs  # noqa: E402

dev = torch.device("cuda:0")
net = GaussianNeuralNetwork(device=dev)
gen = torch.Generator().manual_seed(7)
net.load_state_dict({k: (torch.randn(v.shape, generator=gen) * (0.05 if v.dim() == 2 else 0.01))
                     for k, v in net.state_dict().items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ok = True
for n in (1, 127, 128, 129, 1000, 18944 + 5, 65536):
    obs = alloc_obs(n, dev)
    obs.copy_(torch.randn(n, 965, device=dev) * 0.3)
    os.environ["ROVER_POLICY_KERNEL"] = "v1"
    m1 = net.compute({"states": obs})[0].clone()
    os.environ["ROVER_POLICY_KERNEL"] = "ws"
    m2 = net.compute({"states": obs})[0].clone()
    same = torch.equal(m1, m2)
    ok &= same
    print(f"n={n}: ws == v1 bitwise: {same}  max|diff| {float((m1 - m2).abs().max()):.3e}")
obs = alloc_obs(65536, dev)
obs.copy_(torch.randn(65536, 965, device=dev) * 0.3)
for which in ("v1", "ws"):
    os.environ["ROVER_POLICY_KERNEL"] = which
    ts = []
    for i in range(25):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        net.compute({"states": obs})
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[5:])
    print(f"{which}: cfg-4 (65536 envs) median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
print("IDENTICAL" if ok else "DIFFERENT")
