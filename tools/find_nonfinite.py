#!/usr/bin/env python
"""Roll the bench workload and save the pre-state of env-steps that end non-finite (or hit a given status bit)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from rsr_mjx_b200 import prng, sharding, _lib
from rsr_mjx_b200.envs import AirbotPlayBase
kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3700
bit = int(sys.argv[4]) if len(sys.argv) > 4 else 1
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(sharding.shard_keys(0, N, 0, 1))
gen = torch.Generator(device="cuda").manual_seed(1)
actions = torch.rand(64, N, env.action_size, device="cuda", generator=gen) * 2 - 1
names = ("data", "first_data", "obs", "first_obs", "reward", "done", "info", "metrics")
cases = []
for t in range(T):
    before = {k: st._buf[k].clone() for k in names}
    st._buf["status"].zero_()
    env.step(st, actions[t % 64])
    idx = torch.nonzero(st._buf["status"] & bit).flatten()
    for i in idx.tolist():
        cases.append(dict(t=t, e=i, pre={k: before[k][i].cpu().numpy() for k in names}, post={k: st._buf[k][i].cpu().numpy() for k in names},
                          action=actions[t % 64][i].cpu().numpy(), status=int(st._buf["status"][i])))
        print("step", t, "env", i, "status", int(st._buf["status"][i]), "steps", float(before["info"][i, _lib.INFO["STEPS"]]))
    if len(cases) > 40:
        break
np.save(os.path.join(ROOT, "gpurun_out", f"nonfinite_{kind}.npy"), np.array(cases, dtype=object), allow_pickle=True)
print(len(cases), "cases")
