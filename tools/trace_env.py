#!/usr/bin/env python
"""Record the whole trajectory (all state buffers + action per step) of chosen envs of the bench workload."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
from rsr_mjx_b200 import sharding
from rsr_mjx_b200.envs import AirbotPlayBase
kind, N, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
envs = [int(x) for x in sys.argv[4:]]
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(sharding.shard_keys(0, N, 0, 1))
gen = torch.Generator(device="cuda").manual_seed(1)
actions = torch.rand(64, N, env.action_size, device="cuda", generator=gen) * 2 - 1
names = ("data", "first_data", "obs", "first_obs", "reward", "done", "info", "metrics", "status")
idx = torch.tensor(envs, device="cuda")
rec = {k: [] for k in names}
rec["action"] = []
for t in range(T):
    for k in names:
        rec[k].append(st._buf[k][idx].cpu().numpy())
    rec["action"].append(actions[t % 64][idx].cpu().numpy())
    env.step(st, actions[t % 64])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"trace_{kind}.npz"), envs=np.array(envs), **{k: np.array(v) for k, v in rec.items()})
print("ok")
