#!/usr/bin/env python
"""The two reference entry scripts end to end on synthetic data:

  1. test/rsr_env_params_tuning.py — friction system identification from real transitions
     (here: one batched sweep launch per step instead of `grad(loss_fn)` through `env.step`);
  2. test/rsr_policy_training.py — load the six text tables, build the RSR statistics, train PPO with the RSR term.

"Real" data = the sf env with the cube friction of sf.xml (1.22) + N(0, 1e-3^2) sensor noise; "past sim" = friction 0.4
(rsr_env_params_tuning.py:85); "current sim" = the tuned friction.  All rollouts replay one fixed action sequence.

usage: python examples/rsr_end_to_end.py [out_dir] [ppo_training_steps]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from rsr_mjx_b200 import datasets, ppo, prng, rsr_pipeline
from rsr_mjx_b200.envs import AirbotPlayBase

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/rsr_example"
ppo_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
os.makedirs(out_dir, exist_ok=True)
T = 51
REAL_FRICTION, PAST_SIM_FRICTION = 1.22, 0.4


def rollout(friction, actions):
    """51 observations of one bare env (no episode wrapper) under `actions`, cube geom friction = `friction`"""
    env = AirbotPlayBase("sf", num_envs=1, episode_length=0)
    gf = torch.from_numpy(env.model.geom_friction.astype(np.float32)).cuda()[None].clone()
    gf[:, -1, :] = friction
    env.set_per_env(geom_friction=gf)
    st = env.reset(prng.PRNGKey(0)[None])
    obs = [st.obs[0, :23].cpu().numpy()]
    for a in actions:
        env.step(st, torch.from_numpy(a).cuda()[None])
        obs.append(st.obs[0, :23].cpu().numpy())
    return np.stack(obs).astype(np.float32)


g = np.random.default_rng(2)
actions = g.uniform(-1, 1, (T - 1, 5)).astype(np.float32)
real_obs = rollout(REAL_FRICTION, actions) + np.random.default_rng(3).normal(0, 1e-3, (T, 23)).astype(np.float32)
past_sim_obs = rollout(PAST_SIM_FRICTION, actions)

# ---- 1. friction system identification (rsr_env_params_tuning.py:83-120)
init = PAST_SIM_FRICTION
env = AirbotPlayBase("sf", num_envs=1, episode_length=0)
n = 15
tuned, log = rsr_pipeline.env_params_tuning(env, 6, init, init * 0.2, init * 10.0, real_obs[:n], actions[:n], real_obs[1:n + 1],
                                            log_path=os.path.join(out_dir, "log.txt"))
print(f"tuned cube friction: {float(tuned):.4f} (real {REAL_FRICTION}, start {init}); loss {log['loss'][0]:.5f} -> {log['loss'][-1]:.5f}")

# ---- 2. datasets on disk, then policy training with the RSR term (rsr_policy_training.py:149-260)
current_sim_obs = rollout(float(tuned), actions)
datasets.write_rsr_datasets(os.path.join(out_dir, "data"), real_obs, actions, past_sim_obs, current_sim_obs, current_sim_obs, actions)
S, A, S1_real, S1_past, S1_cur = datasets.load_rsr_datasets(os.path.join(out_dir, "data"), verbose=True)
train_env = AirbotPlayBase("sf", num_envs=512, episode_length=1200)
seen = []
make_policy, params = rsr_pipeline.policy_params_training(
    train_env, past_states=S, past_actions=A, past_next_states_real=S1_real, past_next_states_sim=S1_past,
    current_next_states_sim=S1_cur, num_envs=512, batch_size=128, num_timesteps=10**9, num_evals=ppo_steps,
    max_training_steps=ppo_steps, progress_fn=lambda steps, m: seen.append((steps, m)))
for steps, m in seen:  # the first call is the evaluation before training (no training/* keys yet)
    print(f"env-steps {steps}: eval/episode_reward {m['eval/episode_reward']:.3f} +- {m['eval/episode_reward_std']:.3f}"
          + (f" | total_loss {m['training/total_loss']:.4f} sim2real_loss {m['training/sim2real_loss']:.3e} "
             f"sps {m['training/sps']:.0f}" if "training/sps" in m else ""))
ckpt = os.path.join(out_dir, "policy.pt")
ppo.save_params(ckpt, params, extra={"tuned_friction": float(tuned)})
(norm, net), meta = ppo.load_params(ckpt)
act = ppo.make_inference_fn((norm, net))(deterministic=True)(train_env.reset(prng.split(prng.PRNGKey(5), 512)).obs)
print("checkpoint", ckpt, "->", tuple(act.shape), meta["extra"])
