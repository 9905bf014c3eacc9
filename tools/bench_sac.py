#!/usr/bin/env python
"""SAC env-steps/sec at the reference's RSR defaults (test/rsr_policy_training.py:60-68: 512 envs, batch 128,
min_replay 10_000, max_replay 200_000; brax SAC defaults otherwise).  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import sac
from rsr_mjx_b200.envs import AirbotPlayBase

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
graph = (sys.argv[2] != "eager") if len(sys.argv) > 2 else True
env = AirbotPlayBase("sf", num_envs=512, episode_length=1200)
seen = []
sac.train(env, num_timesteps=10**9, episode_length=1200, num_envs=512, batch_size=128, min_replay_size=10_000,
          max_replay_size=200_000, num_evals=5, max_training_steps=steps, use_cuda_graph=graph, run_evals=False,
          progress_fn=lambda n, m: seen.append(m["training/sps"]))
# max_training_steps cuts the run short of an epoch: one progress call at the end
print(json.dumps({"metric": "sac_train_env_steps_per_sec", "n_gpus": 1, "training_steps": steps, "cuda_graph": graph,
                  "sps": seen[-1], "config": "sf env, 512 envs, batch 128, 1 grad update per actor step"}))
