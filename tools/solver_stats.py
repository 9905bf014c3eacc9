#!/usr/bin/env python
"""Distribution of contacts / Newton iterations / line-search iterations per physics substep,
GPU vs the f32 oracle on identical states."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import parity_utils as P
from oracle import oracle as O
from rsr_mjx_b200 import prng
from rsr_mjx_b200.envs import AirbotPlayBase
from rsr_mjx_b200.model import pack_model

kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
st = env.reset(prng.split(prng.PRNGKey(0), N))
gen = torch.Generator("cuda").manual_seed(0)
blob = pack_model(env.model)
S = {k: [] for k in ("ncon", "nefc", "niter", "ls")}
So = {k: [] for k in ("niter", "ls")}
for t in range(60):
    a = torch.rand(N, 5, device="cuda", generator=gen) * 2 - 1
    env.step(st, a)
    if t % 6 == 5:
        d = env.physics_step_debug(st._buf["data"].clone()).cpu().numpy()
        S["ncon"] += list(d[:, 480]); S["nefc"] += list(d[:, 481]); S["niter"] += list(d[:, 482]); S["ls"] += list(d[:, -1])
        b = P.buffers_to_numpy(st)
        for e in range(24):
            so = P.gpu_to_oracle_states(env, {k: v[e:e + 1] for k, v in b.items()})[0]
            O.step(blob, so.d, 1, precision="f32")
            So["niter"].append(so.d.solver_niter); So["ls"].append(so.d.ls_total)
            if t == 59 and e < 6:
                print(f"  env {e}: gpu niter {int(d[e,482])} ls {int(d[e,-1])} | oracle-f32 niter {so.d.solver_niter} ls {so.d.ls_total}")
for k, v in S.items():
    v = np.array(v)
    print(f"GPU {k:6s}: mean {v.mean():7.2f} median {np.median(v):6.1f} p90 {np.percentile(v,90):6.1f} p99 {np.percentile(v,99):6.1f} max {v.max():6.1f}")
for k, v in So.items():
    v = np.array(v)
    print(f"ORC {k:6s}: mean {v.mean():7.2f} median {np.median(v):6.1f} p90 {np.percentile(v,90):6.1f} max {v.max():6.1f}")
print("status bits:", np.unique(st._buf["status"].cpu().numpy(), return_counts=True))
