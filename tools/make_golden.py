#!/usr/bin/env python
"""Generate tests/golden/*.npz from the float64 CPU oracle.

PARITY UNPINNED: the reference ships no golden vectors and mujoco-mjx cannot be
installed here, so these fixtures pin the *oracle* (regression)
and give the CUDA path a fixed target that travels to the GPU box.
Inputs are reproducible: reset keys = jax-style split(PRNGKey(seed), N), actions
= numpy default_rng(seed+1).uniform(-1, 1).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from oracle import oracle as O
from rsr_mjx_b200 import airbot_spec as A, domain_randomize as DR, prng
from rsr_mjx_b200.model import pack_model

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def rollout(kind, N, T, seed, episode_length, dr=False):
    m = A.load_model(kind)
    cfg = A.make_env_cfg(m, kind, episode_length=episode_length)
    keys = prng.split(prng.PRNGKey(seed), N)
    qpos, qvel, ctrl = A.sample_reset(m, kind, keys)
    per_env = None
    if dr:
        per_env = DR.domain_randomize_arrays(m, prng.split(prng.PRNGKey(seed + 7), N))
        blobs = [pack_model(m.replace_arrays(**{k: v[i] for k, v in per_env.items()})) for i in range(N)]
    else:
        blobs = [pack_model(m)] * N
    actions = np.random.default_rng(seed + 1).uniform(-1, 1, (T, N, m.nu)).astype(np.float32)
    states = [O.env_reset(blobs[i], cfg, qpos[i], qvel[i], ctrl[i]) for i in range(N)]
    rec = dict(qpos0=qpos, qvel0=qvel, ctrl0=ctrl, actions=actions, keys=keys,
               reset_obs=np.array([np.array(s.obs) for s in states]),
               reset_qpos=np.array([np.array(s.d.qpos)[:m.nq] for s in states]),
               reset_warm=np.array([np.array(s.d.qacc_warmstart)[:m.nv] for s in states]))
    names = ("qpos", "qvel", "ctrl", "obs", "reward", "done", "steps", "truncation", "ncon")
    out = {k: [] for k in names}
    for t in range(T):
        row = {k: [] for k in names}
        for i in range(N):
            s = O.env_step(blobs[i], cfg, states[i], actions[t, i])
            row["qpos"].append(np.array(s.d.qpos)[:m.nq]); row["qvel"].append(np.array(s.d.qvel)[:m.nv])
            row["ctrl"].append(np.array(s.d.ctrl)[:m.nu]); row["obs"].append(np.array(s.obs))
            row["reward"].append(s.reward); row["done"].append(s.done); row["steps"].append(s.steps)
            row["truncation"].append(s.truncation); row["ncon"].append(s.d.ncon_active)
        for k in names:
            out[k].append(np.array(row[k]))
    rec.update({k: np.array(v) for k, v in out.items()})
    if per_env is not None:
        rec.update({"dr_" + k: v for k, v in per_env.items()})
    rec["episode_length"] = np.array(episode_length)
    return rec


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    O.build()
    jobs = [("sf", 4, 40, 42, 1200, False), ("cube", 4, 40, 43, 1200, False), ("T", 4, 40, 44, 1200, False),
            ("cube_dr", 4, 30, 45, 1200, True), ("sf_short", 4, 14, 46, 5, False)]
    for name, N, T, seed, ep, dr in jobs:
        kind = name.split("_")[0]
        rec = rollout(kind, N, T, seed, ep, dr)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
        print(name, "done: mean reward", rec["reward"].mean(), "dones", rec["done"].sum(), "max ncon", rec["ncon"].max())
