#!/usr/bin/env python
"""Teacher-forced GPU-vs-oracle run that saves the env-steps on which the two disagree (pre-state, action, both
post-states, and the kernel's per-substep debug dumps) for offline analysis: tools/find_parity_outliers.py kind N T thr"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import parity_utils as P
from oracle import oracle as O
from rsr_mjx_b200 import airbot_spec as A, prng
from rsr_mjx_b200.envs import AirbotPlayBase
from rsr_mjx_b200.model import pack_model

kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1202
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-5
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 7
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
keys = prng.split(prng.PRNGKey(seed), N)
ic = A.sample_reset(env.model, kind, keys)
st = env.reset_from(*ic)
blob, L, m = pack_model(env.model), env.layout, env.model
actions = np.random.default_rng(seed + 1).uniform(-1, 1, (T, N, m.nu)).astype(np.float32)
arr, view = P.oracle_state_array(N)
cases = []
for t in range(T):
    b0 = P.buffers_to_numpy(st)
    env.step(st, torch.from_numpy(actions[t]).cuda())
    torch.cuda.synchronize()
    b1 = P.buffers_to_numpy(st)
    P.fill_oracle_states(env, view, b0)
    P.oracle_step_batch(blob, env.cfg, arr, actions[t])
    ref = P.oracle_states_to_buffers(env, view)
    eq = P.elem_err_rows(b1["data"][:, L.qpos:L.qpos + m.nq], ref["data"][:, L.qpos:L.qpos + m.nq])
    for e in np.nonzero(eq > thr)[0]:
        if b1["done"][e]:
            continue
        # the kernel's substeps, one at a time, from the same pre-state with the shaped ctrl
        row = torch.from_numpy(b0["data"][e:e + 1].copy()).cuda()
        row[:, L.ctrl:L.ctrl + m.nu] = torch.from_numpy(b1["data"][e:e + 1, L.ctrl:L.ctrl + m.nu]).cuda()
        one = AirbotPlayBase(kind, num_envs=1, episode_length=1200) if "one" not in globals() else one
        dumps, rows = [], []
        for f in range(env.cfg.n_frames):
            rows.append(row.cpu().numpy()[0].copy())
            d = one.physics_step_debug(row)
            torch.cuda.synchronize()
            dumps.append(d.cpu().numpy()[0])
        rows.append(row.cpu().numpy()[0].copy())
        cases.append(dict(t=t, e=int(e), err=float(eq[e]), pre={k: b0[k][e] for k in b0}, post_gpu={k: b1[k][e] for k in b1},
                          post_ref={k: ref[k][e] for k in ref}, action=actions[t, e], dumps=np.array(dumps), rows=np.array(rows)))
print(f"{kind}: {len(cases)} env-steps of {N * T} with qpos err > {thr}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"outliers_{kind}.npy"), np.array(cases, dtype=object), allow_pickle=True)
for c in sorted(cases, key=lambda c: -c["err"])[:20]:
    print(c["t"], c["e"], f"{c['err']:.2e}", "ncon/substep", [int(d[480]) for d in c["dumps"]], "niter", [int(d[482]) for d in c["dumps"]])
