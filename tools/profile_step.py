#!/usr/bin/env python
"""Driver for ncu: env.step launches at N envs (default 8192), optionally after advancing the envs to the stationary
episode-phase distribution the bench times (bench.stagger_episode_phases + 1200 steps), so that the captured launch is a
representative one.  usage: profile_step.py kind N steps [stationary]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import prng
from rsr_mjx_b200.envs import AirbotPlayBase

kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
stationary = len(sys.argv) > 4 and sys.argv[4] == "stationary"
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(prng.split(prng.PRNGKey(0), N))
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand(64, N, 5, device="cuda", generator=gen) * 2 - 1
if stationary:
    import bench
    bench.stagger_episode_phases(env, st, gen)
    for t in range(1200):
        env.step(st, a[t % 64])
for t in range(steps):
    env.step(st, a[t % 64])
torch.cuda.synchronize()
print("ok", int(st._buf["status"].max()))
