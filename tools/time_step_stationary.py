#!/usr/bin/env python
"""ms per env.step launch at N envs on the stationary episode-phase distribution (as bench.py times it)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import prng
from rsr_mjx_b200.envs import AirbotPlayBase
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
for N in [int(x) for x in (sys.argv[2:] or ["8192"])]:
    env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
    st = env.reset(prng.split(prng.PRNGKey(0), N))
    g = torch.Generator("cuda").manual_seed(1)
    a = torch.rand(64, N, 5, device="cuda", generator=g) * 2 - 1
    bench.stagger_episode_phases(env, st, g)
    for t in range(1200):
        env.step(st, a[t % 64])
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    ms = []
    for t in range(60):
        flush.fill_(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(st, a[t % 64]); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = np.array(ms[10:])
    import hashlib
    digest = hashlib.sha256(st._buf["data"].cpu().numpy().tobytes() + st.obs.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f" N={N}: {ms.mean():.3f} ms/step (min {ms.min():.3f}) -> {N / ms.mean() * 1e3:.3e} env-steps/s   state sha {digest}")
