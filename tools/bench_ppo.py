#!/usr/bin/env python
"""PPO train steps/sec at the reference's baseline hyper-parameters (ppo_train/airbot_training/train.py:45-55:
1024 envs, unroll 10, 32 minibatches x 256, 8 updates per batch, lr 1e-4, gamma 0.96, entropy 2e-2, reward scale
0.1, obs-norm, domain randomisation) — `training/sps` of RSR/train.py:378-385 = env-steps consumed / wall time of
a training step.  Under torchrun the 1024 envs are split across ranks (global batch fixed, like RSR/train.py:208-235).
One JSON line on rank 0.
usage: bench_ppo.py [training steps] [graph|eager] [rsr]     ("rsr": with the RSR term on bench.py's synthetic tables)
RSRX_PPO_PROFILE_STEP=k brackets training step k with cudaProfilerStart/Stop (`ncu --profile-from-start off` then lists
exactly the launches of one steady training step)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from rsr_mjx_b200 import domain_randomize as DR, ppo, prng
from rsr_mjx_b200.envs import AirbotPlayBase

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
graph = (sys.argv[2] != "eager") if len(sys.argv) > 2 else True
N = 1024 // world
env = AirbotPlayBase("cube", num_envs=N, episode_length=1200, device=f"cuda:{local}", randomization_fn=DR.domain_randomize,
                     randomization_rng=prng.split(prng.PRNGKey(1), N))
sps, split = [], []
past = None
if len(sys.argv) > 3 and sys.argv[3] == "rsr":
    import bench
    from rsr_mjx_b200 import rsr_pipeline as RP
    past = RP.build_policy_rsr_data(*bench.synthetic_rsr_files(env.observation_size, env.action_size,
                                                               rollout=bench.random_action_rollout("cube", f"cuda:{local}")),
                                    bandwidth=bench.RSR_BENCH_BANDWIDTH, device=f"cuda:{local}")
prof = int(os.environ.get("RSRX_PPO_PROFILE_STEP", "-1"))


def on_step(n, m):
    sps.append(m["training/sps"])
    split.append((m["training/collect_s"], m["training/update_s"], m.get("training/sim2real_loss")))
    if len(sps) == prof:
        torch.cuda.profiler.start()
    elif len(sps) == prof + 1:
        torch.cuda.profiler.stop()


ppo.train(env, num_timesteps=10**9, episode_length=1200, num_envs=N, past_data=past, learning_rate=1e-4, entropy_cost=2e-2, discounting=0.96,
          unroll_length=10, batch_size=256 // world, num_minibatches=32, num_updates_per_batch=8, num_evals=steps,
          normalize_observations=True, reward_scaling=0.1, use_cuda_graph=graph, max_training_steps=steps,
          run_evals=False,
          training_step_fn=on_step)
if rank == 0:
    print(json.dumps({"metric": "ppo_train_env_steps_per_sec", "n_gpus": world, "training_steps": steps, "cuda_graph": graph,
                      "sps_per_training_step": sps, "sps_steady": sum(sps[1:]) / max(len(sps) - 1, 1), "collect_s_update_s_sim2real": split,
                      "rsr_term": past is not None,
                      "config": "cube_env + DR, 1024 envs, unroll 10, 32 x 256 minibatches, 8 updates/batch"}))
if world > 1:
    dist.destroy_process_group()
