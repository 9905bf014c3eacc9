#!/usr/bin/env python
"""Short driver for ncu: a few env.step launches at N envs (default 8192)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import prng
from rsr_mjx_b200.envs import AirbotPlayBase

kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(prng.split(prng.PRNGKey(0), N))
a = torch.rand(steps, N, 5, device="cuda") * 2 - 1
for t in range(steps):
    env.step(st, a[t])
torch.cuda.synchronize()
print("ok", int(st._buf["status"].max()))
