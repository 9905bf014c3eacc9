#!/usr/bin/env python
"""Secondary measurements: T-shape / cube_env env-steps/s at 8192 envs,
RSR loss kernel latency (fwd and fwd+bwd), friction sweep latency.  One JSON object on stdout."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import prng, rsr_loss, rsr_pipeline as RP, airbot_spec as A
from rsr_mjx_b200.envs import AirbotPlayBase

def time_ms(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

out = {}
for kind in ("sf", "cube", "T"):
    for N in (1024, 8192):
        env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
        st = env.reset(prng.split(prng.PRNGKey(0), N))
        act = torch.rand(N, 5, device="cuda") * 2 - 1
        ms = time_ms(lambda: env.step(st, act), n=40)
        out[f"{kind}_env_steps_per_s_N{N}"] = N / ms * 1e3
        out[f"{kind}_ms_per_step_N{N}"] = ms
# RSR loss: M=10, D=51, 50 reference rows, batch 128 / 1280 (rsr_pipeline.py:286-306)
g = np.random.default_rng(0)
for Nb in (128, 1280, 2560):
    grid = torch.from_numpy(g.uniform(-3, 3, (10, 51)).astype(np.float32)).cuda()
    ref = torch.from_numpy(g.normal(0, 1, (50, 51)).astype(np.float32)).cuda()
    x = torch.from_numpy(g.normal(0, 1, (Nb, 51)).astype(np.float32)).cuda()
    refd = rsr_loss.evaluate_kde(ref, grid, 0.1)
    data = rsr_loss.RSRData(torch.tensor(0.5), refd, ref, grid, 0.1)
    out[f"rsr_loss_fwd_us_Nb{Nb}"] = 1e3 * time_ms(lambda: rsr_loss.compute_rsr_loss(x[:, :23], x[:, 23:28], x[:, 28:], data), n=200)
    xg = x.clone().requires_grad_(True)
    def fb():
        xg.grad = None
        l, _ = rsr_loss.compute_rsr_loss(xg[:, :23], xg[:, 23:28], xg[:, 28:], data)
        l.backward()
    out[f"rsr_loss_fwd_bwd_us_Nb{Nb}"] = 1e3 * time_ms(fb, n=200)
# friction sweep: 64 params x 15 samples, one launch (rsr_env_params_tuning.py:92-94)
m = A.load_model("sf")
obs = np.tile(np.zeros(23, np.float32), (15, 1)); 
one = AirbotPlayBase("sf", num_envs=1, episode_length=0)
s0 = one.reset(prng.PRNGKey(0)[None])
obs[:] = s0.obs.cpu().numpy()
obs[:, :6] += g.uniform(-0.02, 0.02, (15, 6)).astype(np.float32)
act = g.uniform(-1, 1, (15, 5)).astype(np.float32)
sweep = RP.FrictionSweep("sf", obs, act, obs, num_params=64)
p = torch.linspace(0.08, 4.0, 64)
out["friction_sweep_64x15_ms"] = time_ms(lambda: sweep.loss(p), n=50)
print(json.dumps(out))
