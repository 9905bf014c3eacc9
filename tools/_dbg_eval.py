import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
from rsr_mjx_b200 import ppo, prng, rsr_pipeline
from rsr_mjx_b200.envs import AirbotPlayBase
env = AirbotPlayBase("sf", num_envs=512, episode_length=1200)
S, A, S1r, S1p, S1c = bench.synthetic_rsr_files(env.observation_size, env.action_size, rollout=bench.random_action_rollout("sf", "cuda"))
past = rsr_pipeline.build_policy_rsr_data(S, A, S1r, S1p, S1c, bandwidth=0.1)
print("divergence", float(past.divergence), "refdens", past.reference_density.tolist())
seen = []
mk, (norm, net) = rsr_pipeline.policy_params_training(
    env, past_states=S, past_actions=A, past_next_states_real=S1r, past_next_states_sim=S1p, current_next_states_sim=S1c,
    num_envs=512, batch_size=128, num_timesteps=10**9, num_evals=2, max_training_steps=2, bandwidth=0.1, run_evals=False,
    progress_fn=lambda n, mm: seen.append((n, dict(mm))))
print("param absmax", {n_: float(p.abs().max()) for n_, p in net.policy.named_parameters()})
ev = AirbotPlayBase("sf", num_envs=128, episode_length=1200)
st = ev.reset(prng.split(prng.PRNGKey(9), 128))
pol = mk(deterministic=True)
tot = torch.zeros(128, device="cuda")
bad = None
for t in range(1200):
    a = pol(st.obs)
    if not torch.isfinite(a).all() and bad is None:
        bad = ("action", t, int((~torch.isfinite(a)).any(-1).sum()), "obs finite", bool(torch.isfinite(st.obs).all()))
    st = ev.step(st, a)
    if not torch.isfinite(st.reward).all() and bad is None:
        bad = ("reward", t, int((~torch.isfinite(st.reward)).sum()), "action absmax", float(a.abs().max()), "status", torch.bincount(st._buf["status"].flatten() & 7, minlength=8).tolist())
    tot += st.reward
print("manual eval mean", float(tot.mean()), "nonfinite envs", int((~torch.isfinite(tot)).sum()), "first bad", bad)
