import sys, os, torch
sys.path.insert(0, os.getcwd())
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs
dev = torch.device("cuda:0")
net = GaussianNeuralNetwork(device=dev)
torch.manual_seed(0)
sd = {k: torch.randn_like(v) * 0.05 for k, v in net.state_dict().items()}
net.load_state_dict(sd)
obs = alloc_obs(256, dev); obs.copy_(torch.randn(256, 965, device=dev))
try:
    m, _, _ = net.compute({"states": obs}); torch.cuda.synchronize(); print("stage", os.environ.get("ROVER_POLICY_DEBUG_STOP"), "OK", m[:2].tolist())
except Exception as e:
    print("stage", os.environ.get("ROVER_POLICY_DEBUG_STOP"), "FAIL", str(e)[:100])
