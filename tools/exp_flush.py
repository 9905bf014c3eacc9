import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import sharding, _lib
from rsr_mjx_b200.envs import AirbotPlayBase
import bench
N=8192
env = AirbotPlayBase("sf", num_envs=N, episode_length=1200)
st = env.reset(sharding.shard_keys(0, N, 0, 1))
gen = torch.Generator(device="cuda").manual_seed(1)
actions = torch.rand(64, N, 5, device="cuda", generator=gen) * 2 - 1
def workload():
    d = env.physics_step_debug(st._buf["data"].clone())
    ncon, nefc, niter, ls = d[:,480], d[:,481], d[:,482], d[:,-1]
    age = st._buf["info"][:, _lib.INFO["STEPS"]]
    return "age %.0f ncon %.2f (max %d) nefc %.1f niter %.2f (max %d) ls %.1f (max %d)" % (age.mean(), ncon.mean(), ncon.max(), nefc.mean(), niter.mean(), niter.max(), ls.mean(), ls.max())
def run(n, off=0):
    evs=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for t in range(n):
        evs[t][0].record(); env.step(st, actions[(t+off)%64]); evs[t][1].record()
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a,b in evs])
for blk in range(12):
    ms=run(100, blk*100)
    print("from reset", blk, "%.3f"%ms.mean(), workload())
bench.stagger_episode_phases(env, st, gen)
for blk in range(30):
    ms=run(100, blk*100)
    print("after stagger", blk, "%.3f"%ms.mean(), workload())
