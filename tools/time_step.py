#!/usr/bin/env python
"""ms per env.step launch at N envs (CUDA events, L2 flush between launches); no parity checks."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import prng
from rsr_mjx_b200.envs import AirbotPlayBase
kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
for N in [int(x) for x in (sys.argv[2:] or ["1024", "8192"])]:
    env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
    st = env.reset(prng.split(prng.PRNGKey(0), N))
    g = torch.Generator("cuda").manual_seed(0)
    a = torch.rand(60, N, 5, device="cuda", generator=g) * 2 - 1
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for t in range(10):
        env.step(st, a[t])
    ms = 0.0
    for t in range(10, 60):
        flush.fill_(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(st, a[t]); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    print(f" N={N}: {ms / 50:.3f} ms/step -> {N / (ms / 50) * 1e3:.3e} env-steps/s; status {sorted(set(st._buf['status'].cpu().tolist()))}")
