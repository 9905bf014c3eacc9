#!/usr/bin/env python
"""PPO on BASELINE config 3 for its full budget (ppo_train/airbot_training/train.py:45-55: 15 M env-steps, 1024 envs,
episode_length 1200, unroll 10, 32 x 256 minibatches, 8 updates per batch, lr 1e-4, entropy 2e-2, gamma 0.96, reward scaling
0.1, observation normalisation, domain randomisation) with the reference's evaluation schedule: does the policy the
hand-written update trains actually improve?  One JSON line: env-steps, eval/episode_reward (mean, std over 128 eval
envs), wall time."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import domain_randomize as DR, ppo, prng
from rsr_mjx_b200.envs import AirbotPlayBase

steps = int(float(sys.argv[1])) if len(sys.argv) > 1 else 15_000_000
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 6
env = AirbotPlayBase("cube", num_envs=1024, episode_length=1200, randomization_fn=DR.domain_randomize,
                     randomization_rng=prng.split(prng.PRNGKey(1), 1024))
curve, t0 = [], time.time()
ppo.train(env, num_timesteps=steps, episode_length=1200, num_envs=1024, learning_rate=1e-4, entropy_cost=2e-2, discounting=0.96,
          unroll_length=10, batch_size=256, num_minibatches=32, num_updates_per_batch=8, num_evals=evals,
          normalize_observations=True, reward_scaling=0.1,
          progress_fn=lambda n, m: curve.append({"env_steps": int(n), "eval_episode_reward": m.get("eval/episode_reward"),
                                                 "eval_episode_reward_std": m.get("eval/episode_reward_std"),
                                                 "training_sps": m.get("training/sps"), "wall_s": time.time() - t0}))
status = env_status = None
print(json.dumps({"metric": "ppo_learning_curve", "config": "BASELINE config 3 (cube_env + DR, 1024 envs), full 15 M-step budget",
                  "total_wall_s": time.time() - t0, "curve": curve}))
