#!/usr/bin/env python
"""How often does the active-contact cap bind?  Long random-action rollout at N envs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import prng, _lib
from rsr_mjx_b200.envs import AirbotPlayBase
kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(prng.split(prng.PRNGKey(0), N))
gen = torch.Generator("cuda").manual_seed(0)
hist = np.zeros(64, np.int64)
for t in range(T):
    env.step(st, torch.rand(N, 5, device="cuda", generator=gen) * 2 - 1)
    if t % 10 == 9:
        d = env.physics_step_debug(st._buf["data"].clone())
        nc = d[:, 480].long().clamp(max=63)
        hist += torch.bincount(nc, minlength=64).cpu().numpy()
status = st._buf["status"].cpu().numpy()
print("cap", _lib.lib().rsrx_max_contacts(), "envs flagged overflow:", int(((status & 2) != 0).sum()), "of", N, "| nonfinite:", int(((status & 1) != 0).sum()), "| solver cap:", int(((status & 4) != 0).sum()))
tot = hist.sum()
cum = np.cumsum(hist) / tot
print("ncon histogram (sampled substeps):", {i: int(h) for i, h in enumerate(hist) if h})
for q in (0.5, 0.9, 0.99, 0.999, 0.9999):
    print(f"  p{q}: {int(np.searchsorted(cum, q))}")
