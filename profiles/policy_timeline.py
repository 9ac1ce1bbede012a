"""Bring-up: timeline (SM cycles) of CTA 0 of the warp-specialised policy kernel built with -DROVER_POLICY_DBG=1."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from isaac_rover_orbit_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]
from isaac_rover_orbit_b200.policy import GaussianNeuralNetwork, alloc_obs, alloc_obs_bf16  # noqa: E402

dev = torch.device("cuda:0")
net = GaussianNeuralNetwork(device=dev)
obs = alloc_obs(65536, dev)
obs.copy_(torch.randn(65536, 965, device=dev) * 0.3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if len(sys.argv) > 2 and sys.argv[2] == "bf16":  # the bf16-observation mode (16 chunks of 64 columns per tile)
    ob = alloc_obs_bf16(65536, dev)
    ob.copy_(obs)
    obs = ob
for _ in range(3):
    flush.fill_(1)
    net.compute({"states": obs})
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
t = np.zeros(2048, dtype=np.int64)
assert lib.rover_debug_policy_timeline(t.ctypes.data_as(C.c_void_p)) == 0
n = 64 if obs.dtype == torch.bfloat16 else 124
print("producer issue      :", t[0:n:4].tolist())
print("converter got data  :", t[128:128 + n:4].tolist())
print("converter got A slot:", t[256:256 + n:4].tolist())
print("converter done      :", t[384:384 + n:4].tolist())
print("MMA issued          :", t[512:512 + n:4].tolist())
d = np.diff(t[512:512 + n])
print("MMA issue interval per chunk: median", int(np.median(d)), "mean", int(d.mean()))
print("wait for data (conv)", int((t[128:128 + n] - np.maximum(t[384 - 1:384 + n - 1], 0)).clip(0).mean()), "cycles/chunk avg; convert", int((t[384:384 + n] - t[256:256 + n]).mean()), "cycles/chunk")
names = ["begin", "D0 ready", "epi0+MMA1 issued", "MMA1 done", "epi1+MMA2 issued", "MMA2 done", "epi2+MMA3a issued", "MMA3a done",
         "W3hi+MMA3b done", "epi3+MMA4 issued", "MMA4 done", "epi4+MMA5 issued", "MMA5 done", "epi5 done"]
for i in range(4):
    row = t[1024 + 16 * i:1024 + 16 * i + 14]
    print(f"layer group tile {i}:", {nm: int(v) for nm, v in zip(names, row)})
cta = np.zeros((2, 256), dtype=np.uint64)
assert lib.rover_debug_policy_ctas(cta.ctypes.data_as(C.c_void_p)) == 0
st, en = cta[0, :148].astype(np.int64), cta[1, :148].astype(np.int64)
t0 = st.min()
dur = en - st
print(f"CTA starts spread {int((st - t0).max())} ns; kernel span (first start -> last end) {int(en.max() - t0)} ns; CTA durations min/median/max "
      f"{int(dur.min())}/{int(np.median(dur))}/{int(dur.max())} ns; CTA 0: {int(dur[0])} ns for {int(t[1024 + 16 * 3 + 13])} cycles")
