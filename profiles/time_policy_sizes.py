import sys, torch
sys.path.insert(0, ".")
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16
dev = torch.device("cuda:0")
net = GaussianNeuralNetwork(device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (4096, 16384, 65536):
    obs = alloc_obs(n, dev); obs.copy_(torch.randn(n, 965, device=dev) * 0.3)
    ob = alloc_obs_bf16(n, dev); ob.copy_(obs)
    for name, fn in (("fp32", lambda: net.compute({"states": obs})), ("bf16", lambda: net.compute_bf16({"states": ob}))):
        for warm in (False, True):
            ts = []
            for i in range(30):
                if not warm: flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts = sorted(ts[5:])
            print(f"n={n} {name} {'warm L2' if warm else 'flushed'}: median {ts[len(ts)//2]:.1f} us", flush=True)
