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


def host_layout(m):
    """the data-row layout rsrx_model_create computes (csrc/rsrx_api.cu::build_dmodel), restated for the generator"""
    from rsr_mjx_b200 import _lib
    L, o = _lib.Layout(), 0
    for name, n in (("qpos", m.nq), ("qvel", m.nv), ("ctrl", m.nu), ("qacc_warmstart", m.nv), ("time", 1),
                    ("xpos", m.nbody * 3), ("xquat", m.nbody * 4), ("site_xpos", m.nsite * 3), ("geom_xpos", m.ngeom * 3)):
        setattr(L, name, o)
        o += n
    L.data_stride = (o + 3) // 4 * 4
    L.obs_stride, L.info_stride, L.metrics_stride = _lib.OBS_STRIDE, _lib.INFO_STRIDE, _lib.METRICS_STRIDE
    L.obs_size = 16 if m.nq == 15 else 23  # T-shape env : cube envs
    L.nq, L.nv, L.nu, L.nbody, L.nsite, L.ngeom = m.nq, m.nv, m.nu, m.nbody, m.nsite, m.ngeom
    return L


def rollout_teacher_forced(kind, N, T, seed, episode_length, crafted=False):
    """Long TEACHER-FORCED golden: every pre-step state is stored in the device layout (float32), so a float32
    implementation can be checked step by step over hundreds of contact-rich steps without the chaotic free-running
    drift (f32 vs f64 free-running differ by O(1) after ~50 steps).  tf_*[t] is the state before step t (t = 0: after
    reset), tf_*[t + 1] what the oracle made of it.  Stepped by the FLOAT32 oracle: contact dynamics has discrete
    events (a contact entering its margin, stick/slip of a friction row, a manifold tie), and at those a float64
    evaluation of the same step lands elsewhere (measured: 1 of 440 stored steps differs by more than 1e-4 in qpos,
    2.5 % by more than 2e-3 in qvel; tests/test_golden_oracle.py keeps that statistic), so a float32 kernel can only be
    held to 1e-4 on EVERY step against float32 arithmetic in the oracle's operation order."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import parity_utils as P

    class Env:
        pass
    m = A.load_model(kind)
    env = Env()
    env.model, env.layout = m, host_layout(m)
    cfg = A.make_env_cfg(m, kind, episode_length=episode_length)
    blob = pack_model(m)
    keys = prng.split(prng.PRNGKey(seed), N)
    qpos, qvel, ctrl = A.sample_reset(m, kind, keys)
    if crafted:  # env-triggered termination: see tests/test_gpu_parity.py::_crafted_done_ic
        ids = A.env_ids(m, kind)
        b = ids["_box_qposadr"]
        dadr = int(m.jnt_dofadr[m.body_jntadr[ids["cube_id"]]])
        if kind == "sf":
            qpos[:, b:b + 3] = qpos[:, ids["_site_qposadr"]:ids["_site_qposadr"] + 3]
        else:
            qpos[:, b:b + 3] = np.array([0.3, 3.0, 0.615 if kind == "cube" else 0.6025], np.float32)
            qvel[:, dadr + 2] = -1.0
    actions = np.random.default_rng(seed + 1).uniform(-1, 1, (T, N, m.nu)).astype(np.float32)
    arr, view = P.oracle_state_array(N)
    import ctypes as C
    for i in range(N):
        s = O.env_reset(blob, cfg, qpos[i], qvel[i], ctrl[i], precision="f32")
        C.memmove(C.byref(arr[i]), C.byref(s), C.sizeof(s))
    names = ("data", "obs", "reward", "done", "info", "metrics")
    f32 = lambda b: {k: b[k].astype(np.float32) for k in names}
    cur = f32(P.oracle_states_to_buffers(env, view))
    first = dict(first_data=cur["data"].copy(), first_obs=cur["obs"].copy())
    seq = {k: [cur[k]] for k in names}
    ncon = []
    for t in range(T):
        P.fill_oracle_states(env, view, {**cur, **first})  # the oracle starts from the float32-rounded state
        O.rollout(blob, cfg, arr, actions[t][None].astype(np.float64), precision="f32")
        cur = f32(P.oracle_states_to_buffers(env, view))
        ncon.append(view["d"]["ncon_active"].copy())
        for k in names:
            seq[k].append(cur[k])
    L = env.layout
    rec = {"tf_" + k: np.array(v) for k, v in seq.items()}
    rec.update(first)
    rec.update(qpos0=qpos, qvel0=qvel, ctrl0=ctrl, actions=actions, keys=keys, ncon=np.array(ncon),
               episode_length=np.array(episode_length), crafted=np.array(int(crafted)), oracle_precision=np.array("f32"),
               layout=np.array([getattr(L, f) for f, _ in L._fields_], np.int32))
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
    # round 2: long teacher-forced goldens (>= 200 steps) and env-triggered termination
    for name, N, T, seed, ep, crafted in [("sf_tf", 2, 220, 52, 1200, False), ("T_tf", 2, 220, 54, 1200, False),
                                          ("cube_done_tf", 2, 24, 55, 1200, True), ("sf_done_tf", 2, 6, 56, 1200, True)]:
        rec = rollout_teacher_forced(name.split("_")[0], N, T, seed, ep, crafted)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
        print(name, "steps", T, "dones", rec["tf_done"].sum(), "max ncon", rec["ncon"].max())
