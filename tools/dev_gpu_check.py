#!/usr/bin/env python
"""Developer smoke check on a real GPU: CUDA env vs CPU oracle (teacher-forced),
plus a quick timing.  Usage: python tools/dev_gpu_check.py [kind] [N] [T]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch

from oracle import oracle as O
from rsr_mjx_b200 import _lib, airbot_spec as A, prng
from rsr_mjx_b200.envs import AirbotPlayBase
from rsr_mjx_b200.model import pack_model
import parity_utils as P

np.set_printoptions(precision=5, suppress=True, linewidth=180)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
    m, L = env.model, env.layout
    blob, cfg = pack_model(m), env.cfg
    keys = prng.split(prng.PRNGKey(0), N)
    qpos, qvel, ctrl = A.sample_reset(m, kind, keys)
    st = env.reset_from(qpos, qvel, ctrl)
    torch.cuda.synchronize()
    b = P.buffers_to_numpy(st)
    print("status after reset:", np.unique(b["status"]))
    # ---- reset parity
    worst = {}
    for e in range(min(N, 16)):
        so = O.env_reset(blob, cfg, qpos[e], qvel[e], ctrl[e], precision="f32")
        row = P.oracle_row(env, so.d)
        for name, off, n in (("qpos", L.qpos, m.nq), ("warm", L.qacc_warmstart, m.nv), ("xpos", L.xpos, 3 * m.nbody),
                             ("xquat", L.xquat, 4 * m.nbody), ("site", L.site_xpos, 3 * m.nsite), ("geom", L.geom_xpos, 3 * m.ngeom)):
            worst[name] = max(worst.get(name, 0), P.rel_err(b["data"][e, off:off + n], row[off:off + n]))
        worst["obs"] = max(worst.get("obs", 0), P.rel_err(b["obs"][e], np.array(so.obs)))
        worst["info"] = max(worst.get("info", 0), P.rel_err(b["info"][e], P.oracle_info(so)))
    print("reset parity vs oracle-f32 (max rel err):", {k: f"{v:.2e}" for k, v in worst.items()})
    # ---- internals of one forward
    data = st._buf["data"].clone()
    dump = env.physics_step_debug(data).cpu().numpy()
    torch.cuda.synchronize()
    nv = m.nv
    D = _lib.lib().rsrx_debug_stride()
    from rsr_mjx_b200 import _lib as LL
    worst = {}
    MAXC = 32
    offM, offB, offQS, offQ, offFC = 0, 400, 420, 440, 460
    for e in range(min(N, 16)):
        so = P.gpu_to_oracle_states(env, {k: v[e:e + 1] for k, v in b.items()})[0]
        ins = O.inspect(blob, so.d, precision="f32")
        Mg = dump[e, offM:offM + nv * nv].reshape(nv, nv)
        worst["M"] = max(worst.get("M", 0), P.rel_err(Mg, ins["M"]))
        worst["bias"] = max(worst.get("bias", 0), P.rel_err(dump[e, offB:offB + nv], np.array(so.d.qfrc_bias)[:nv]))
        worst["qacc_smooth"] = max(worst.get("qacc_smooth", 0), P.rel_err(dump[e, offQS:offQS + nv], np.array(so.d.qacc_smooth)[:nv]))
        worst["qacc"] = max(worst.get("qacc", 0), P.rel_err(dump[e, offQ:offQ + nv], np.array(so.d.qacc)[:nv]))
        worst["qfrc_c"] = max(worst.get("qfrc_c", 0), P.rel_err(dump[e, offFC:offFC + nv], np.array(so.d.qfrc_constraint)[:nv]))
        ncon_g, nefc_g, niter_g = dump[e, 480:483]
        if int(ncon_g) != so.d.ncon or e == 0:
            print(f" env {e}: ncon gpu {int(ncon_g)} oracle {so.d.ncon} | nefc {int(nefc_g)} vs {so.d.nefc} | niter {int(niter_g)} vs {so.d.solver_niter}")
    print("forward internals vs oracle-f32 (max rel err):", {k: f"{v:.2e}" for k, v in worst.items()})
    # ---- teacher-forced env steps
    rng = np.random.default_rng(1)
    stats = []
    done_mismatch = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, m.nu)).astype(np.float32)
        b0 = P.buffers_to_numpy(st)
        st = env.step(st, torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        b1 = P.buffers_to_numpy(st)
        for e in range(min(N, 8)):
            so = P.gpu_to_oracle_states(env, {k: v[e:e + 1] for k, v in b0.items()})[0]
            O.env_step(blob, cfg, so, a[e], precision="f32")
            row = P.oracle_row(env, so.d)
            eq = P.rel_err(b1["data"][e, L.qpos:L.qpos + m.nq], row[L.qpos:L.qpos + m.nq])
            ev = P.rel_err(b1["data"][e, L.qvel:L.qvel + nv], row[L.qvel:L.qvel + nv])
            eo = P.rel_err(b1["obs"][e], np.array(so.obs))
            er = abs(b1["reward"][e] - so.reward) / max(1, abs(so.reward))
            done_mismatch += int(b1["done"][e] != so.done)
            stats.append((eq, ev, eo, er))
    s = np.array(stats)
    for i, name in enumerate(("qpos", "qvel", "obs", "reward")):
        print(f" step parity {name}: median {np.median(s[:, i]):.2e} p90 {np.percentile(s[:, i], 90):.2e} p99 {np.percentile(s[:, i], 99):.2e} max {s[:, i].max():.2e}")
    print(" done mismatches:", done_mismatch, "status:", np.unique(P.buffers_to_numpy(st)["status"]))
    # ---- timing
    for NN in (1024, 8192):
        env2 = AirbotPlayBase(kind, num_envs=NN, episode_length=1200)
        keys = prng.split(prng.PRNGKey(1), NN)
        s2 = env2.reset(keys)
        act = torch.rand(NN, m.nu, device="cuda") * 2 - 1
        for _ in range(5):
            env2.step(s2, act)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 20
        ev0.record()
        for _ in range(K):
            env2.step(s2, act)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        print(f" N={NN}: {ms:.3f} ms/step -> {NN / ms * 1e3:.3e} env-steps/s; status {np.unique(s2._buf['status'].cpu().numpy())}")


if __name__ == "__main__":
    main()
