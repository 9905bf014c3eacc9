"""SAC trainer with the RSR actor term.

What it mirrors
  RSR/sac_train.py:28-123   `train(...)`: signature (minus the brax plumbing arguments), `rsr_loss_scale < 0` error,
                            past_data=None / scale 0 -> plain SAC
  RSR/sac_losses.py:23-130  alpha / twin-Q critic / actor losses; the actor adds `rsr.compute_rsr_loss` on the
                            post-tanh action of the policy being optimised
  brax==0.12.1 (un-vendored dependency, restated): `brax/training/agents/sac/train.py` — one actor step of all
      envs per training step, uniform replay buffer, `grad_updates_per_step` SGD steps on freshly sampled batches, in the
      order alpha -> critic (with the OLD alpha) -> actor (with the OLD q) -> polyak(tau) of the NEW q; prefill of
      ceil(min_replay_size / num_envs) actor steps with the initial policy; Adam(lr) for policy and q, Adam(3e-4) for
      log_alpha (init 0), target entropy -0.5 * action_size; `make_sac_networks`: policy MLP (256, 256) -> 2A and two
      Q MLPs (256, 256) -> 1 on concat(obs, action), relu, lecun-uniform kernels; running-statistics observation
      normaliser shared by policy and q.

The env is the batched CUDA `AirbotPlayBase` (already Vmap + Episode + AutoReset [+ DR]); networks, replay buffer and
losses are torch on the same device; the SGD step can be captured in a CUDA graph.  Multi-process: every rank steps its
env shard and samples its own buffer, gradients are averaged with one flat all-reduce per network update.
"""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import rsr_loss as rsr
from .ppo import MLP, Evaluator, NormalTanh, RunningStatistics, _flat_allreduce_mean


class _ReluMLP(MLP):
    """brax make_sac_networks uses relu (PPO: swish)"""
    activation = "relu"


class SACNetworks(nn.Module):
    """brax make_sac_networks: tanh-normal policy and `n_critics = 2` Q heads."""

    def __init__(self, obs_size: int, action_size: int, hidden=(256, 256)):
        super().__init__()
        self.policy = _ReluMLP([obs_size, *hidden, 2 * action_size])
        self.q1 = _ReluMLP([obs_size + action_size, *hidden, 1])
        self.q2 = _ReluMLP([obs_size + action_size, *hidden, 1])
        self.action_size = action_size

    def q(self, obs_n, action):
        x = torch.cat([obs_n, action], dim=-1)
        return torch.cat([self.q1(x), self.q2(x)], dim=-1)  # [..., 2]


class ReplayBuffer:
    """brax UniformSamplingQueue: a device-resident ring of flat transitions, uniform sampling with replacement."""

    def __init__(self, capacity: int, width: int, device):
        if capacity <= 0:
            raise ValueError(f"replay capacity must be positive, got {capacity}")
        self.data = torch.zeros(capacity, width, device=device)
        self.capacity, self.size, self.pos = int(capacity), 0, 0

    def insert(self, rows: torch.Tensor) -> None:
        n = rows.shape[0]
        if n > self.capacity:
            rows, n = rows[-self.capacity:], self.capacity
        end = self.pos + n
        if end <= self.capacity:
            self.data[self.pos:end] = rows
        else:
            k = self.capacity - self.pos
            self.data[self.pos:] = rows[:k]
            self.data[:end - self.capacity] = rows[k:]
        self.pos = end % self.capacity
        self.size = min(self.size + n, self.capacity)

    def sample(self, n: int, generator=None) -> torch.Tensor:
        if self.size == 0:
            raise RuntimeError("cannot sample an empty replay buffer")
        idx = torch.randint(0, self.size, (n,), device=self.data.device, generator=generator)
        return self.data[idx]


def _fields(obs_size: int, action_size: int) -> Dict[str, slice]:
    o, a = obs_size, action_size
    return dict(observation=slice(0, o), action=slice(o, o + a), reward=slice(o + a, o + a + 1),
                discount=slice(o + a + 1, o + a + 2), truncation=slice(o + a + 2, o + a + 3),
                next_observation=slice(o + a + 3, 2 * o + a + 3))


def alpha_loss(log_alpha, net: SACNetworks, normalize, tr, noise, target_entropy: float):
    """RSR/sac_losses.py:38-53 (SAC eq. 18)"""
    with torch.no_grad():
        logits = net.policy(normalize(tr["observation"]))
        loc, scale = NormalTanh.params(logits)
        log_prob = NormalTanh.log_prob(logits, loc + scale * noise)
    return torch.mean(torch.exp(log_alpha) * (-log_prob - target_entropy))


def critic_loss(net: SACNetworks, target: SACNetworks, normalize, alpha, tr, noise, reward_scaling: float,
                discounting: float):
    """RSR/sac_losses.py:55-96: twin-Q Bellman error against min target-Q minus alpha * log pi, truncated steps masked"""
    obs_n = normalize(tr["observation"])
    old_q = net.q(obs_n, tr["action"])
    with torch.no_grad():
        next_n = normalize(tr["next_observation"])
        logits = net.policy(next_n)
        loc, scale = NormalTanh.params(logits)
        raw = loc + scale * noise
        next_log_prob = NormalTanh.log_prob(logits, raw)
        next_q = target.q(next_n, torch.tanh(raw))
        next_value = next_q.min(dim=-1).values - alpha * next_log_prob
        target_q = tr["reward"].squeeze(-1) * reward_scaling + tr["discount"].squeeze(-1) * discounting * next_value
    q_error = (old_q - target_q.unsqueeze(-1)) * (1 - tr["truncation"])
    return 0.5 * torch.mean(q_error * q_error)


def actor_loss(net: SACNetworks, q_net: SACNetworks, normalize, alpha, tr, noise, past_data=None,
               rsr_loss_scale: float = 1.0):
    """RSR/sac_losses.py:98-128: alpha * log pi - min Q, plus the RSR penalty on (obs, tanh(raw), next_obs)"""
    obs_n = normalize(tr["observation"])
    logits = net.policy(obs_n)
    loc, scale = NormalTanh.params(logits)
    raw = loc + scale * noise
    log_prob = NormalTanh.log_prob(logits, raw)
    action = torch.tanh(raw)
    q_action = q_net.q(obs_n, action)
    base = torch.mean(alpha * log_prob - q_action.min(dim=-1).values)
    sim2real, distance = rsr.compute_rsr_loss(tr["observation"], action, tr["next_observation"], past_data,
                                              loss_scale=rsr_loss_scale)
    return base + sim2real, sim2real, distance


def train(environment, num_timesteps: int, episode_length: int, past_data: Any = None, action_repeat: int = 1,
          num_envs: int = 1, num_eval_envs: int = 128, learning_rate: float = 1e-4, discounting: float = 0.9,
          seed: int = 0, batch_size: int = 256, num_evals: int = 1, normalize_observations: bool = False,
          reward_scaling: float = 1.0, tau: float = 0.005, min_replay_size: int = 0,
          max_replay_size: Optional[int] = None, grad_updates_per_step: int = 1, deterministic_eval: bool = False,
          progress_fn: Callable[[int, Dict[str, float]], None] = lambda *a: None, rsr_loss_scale: float = 1.0,
          hidden_layer_sizes=(256, 256), use_cuda_graph: bool = True, allow_tf32: bool = True,
          max_training_steps: Optional[int] = None, eval_env=None, run_evals: bool = True,
          network_factory: Any = None, randomization_fn: Optional[Callable] = None,
          restore_checkpoint_path: Optional[str] = None, sgd_probe_fn: Optional[Callable[[str, Dict[str, Any]], None]] = None,
          **brax_plumbing):
    """SAC training (RSR/sac_train.py:28).  `environment`: an `AirbotPlayBase` with `num_envs` envs on this rank.
    `network_factory` (functools.partial over make_sac_networks: `hidden_layer_sizes` is used), `randomization_fn`
    (installed on the env) are honoured; `restore_checkpoint_path` raises like the reference ("Brax 0.12.1 SAC cannot
    resume complete training state", RSR/rsr_pipeline.py:399-403); unknown keywords raise (train_args.py).
    Returns (make_policy, (normalizer, networks), metrics)."""
    from . import train_args
    if rsr_loss_scale < 0:
        raise ValueError(f"rsr_loss_scale must be non-negative, got {rsr_loss_scale}")
    env = environment
    train_args.reject_unknown(brax_plumbing, "sac.train")
    if restore_checkpoint_path:
        raise ValueError('Brax 0.12.1 SAC cannot resume complete training state; use checkpoint_logdir to save '
                         'inference checkpoints instead')
    hidden_layer_sizes = train_args.hidden_sizes(network_factory, dict(hidden_layer_sizes=hidden_layer_sizes))["hidden_layer_sizes"]
    train_args.apply_randomization(env, randomization_fn, seed)
    past_data = rsr.prepare_rsr_data(past_data, env.device)
    if env.num_envs != num_envs:
        raise ValueError(f"environment has {env.num_envs} envs, num_envs={num_envs}")
    if env.episode_length != episode_length:
        raise ValueError("environment.episode_length differs from episode_length (the env is already wrapped)")
    if min_replay_size >= num_timesteps:
        raise ValueError("No training will happen because min_replay_size >= num_timesteps")
    if past_data is None or rsr_loss_scale == 0:
        past_data, rsr_loss_scale = None, 0.0
    torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    dev = env.device
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if max_replay_size is None:
        max_replay_size = num_timesteps
    env_steps_per_actor_step = action_repeat * num_envs * world
    num_prefill_actor_steps = -(-min_replay_size // (num_envs * world))
    num_prefill_env_steps = num_prefill_actor_steps * env_steps_per_actor_step
    num_evals_after_init = max(num_evals - 1, 1)
    steps_per_epoch = -(-(num_timesteps - num_prefill_env_steps) // (num_evals_after_init * env_steps_per_actor_step))
    total_steps = steps_per_epoch * num_evals_after_init
    if max_training_steps is not None:
        total_steps = min(total_steps, max_training_steps)

    obs_size, act_size = env.observation_size, env.action_size
    torch.manual_seed(seed)
    net = SACNetworks(obs_size, act_size, tuple(hidden_layer_sizes)).to(dev)
    target = SACNetworks(obs_size, act_size, tuple(hidden_layer_sizes)).to(dev)
    target.load_state_dict(net.state_dict())
    for p in target.parameters():
        p.requires_grad_(False)
    log_alpha = torch.zeros((), device=dev, requires_grad=True)
    q_params = list(net.q1.parameters()) + list(net.q2.parameters())
    pol_params = list(net.policy.parameters())
    cap = bool(use_cuda_graph)
    opt_alpha = torch.optim.Adam([log_alpha], lr=3e-4, eps=1e-8, capturable=cap, fused=True)
    opt_q = torch.optim.Adam(q_params, lr=learning_rate, eps=1e-8, capturable=cap, fused=True)
    opt_pi = torch.optim.Adam(pol_params, lr=learning_rate, eps=1e-8, capturable=cap, fused=True)
    norm = RunningStatistics(obs_size, dev)
    normalize = norm.normalize if normalize_observations else (lambda x: x)
    gen = torch.Generator(device=dev).manual_seed(seed * 7919 + rank + 1)
    target_entropy = -0.5 * act_size
    fields = _fields(obs_size, act_size)
    width = 2 * obs_size + act_size + 3
    buffer = ReplayBuffer(max(max_replay_size // world, 1), width, dev)

    from . import sharding
    state = env.reset(sharding.shard_keys(seed, num_envs, rank, world))
    row = torch.empty(num_envs, width, device=dev)

    @torch.no_grad()
    def actor_step():
        obs = state.obs[:, :obs_size]
        row[:, fields["observation"]] = obs
        logits = net.policy(normalize(obs))
        action = torch.tanh(NormalTanh.sample_raw(logits, gen))
        row[:, fields["action"]] = action
        env.step(state, action)
        row[:, fields["reward"]] = state.reward[:, None]
        row[:, fields["discount"]] = 1 - state.done[:, None]
        row[:, fields["truncation"]] = state.info["truncation"][:, None]
        row[:, fields["next_observation"]] = state.obs[:, :obs_size]
        if normalize_observations:
            norm.update(row[:, fields["observation"]])
        buffer.insert(row)

    static = torch.empty(batch_size, width, device=dev)
    noise = torch.empty(3, batch_size, act_size, device=dev)
    metrics: Dict[str, torch.Tensor] = {}

    def view(t):
        return {k: t[:, s] for k, s in fields.items()}

    def grads_alpha_q():
        tr = view(static)
        opt_alpha.zero_grad(set_to_none=True)
        la = alpha_loss(log_alpha, net, normalize, tr, noise[0], target_entropy)
        la.backward()
        opt_q.zero_grad(set_to_none=True)
        lq = critic_loss(net, target, normalize, alpha_static, tr, noise[1], reward_scaling, discounting)  # the OLD alpha
        lq.backward(inputs=q_params)
        return la, lq, alpha_static

    def grads_actor(alpha):
        tr = view(static)
        opt_pi.zero_grad(set_to_none=True)
        # all three gradients are taken at the OLD parameters and the optimisers step afterwards, which is brax's
        # alpha -> critic -> actor sequence (each of its updates reads `training_state`, not the fresh values)
        lp, s2r, distance = actor_loss(net, net, normalize, alpha, tr, noise[2], past_data, rsr_loss_scale)
        lp.backward(inputs=pol_params)
        return lp, s2r, distance

    target_params = list(target.q1.parameters()) + list(target.q2.parameters())

    @torch.no_grad()
    def finish():
        # polyak step of the target critics, two multi-tensor launches instead of two per tensor (same arithmetic)
        torch._foreach_mul_(target_params, 1 - tau)
        torch._foreach_add_(target_params, q_params, alpha=tau)

    alpha_static = torch.ones((), device=dev)
    actor_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def sgd_eager():
        # The three gradients are all taken at the OLD parameters (brax's update order), so the actor branch (policy
        # forward, both critics, backward through them into the policy: about half of the step's launches) does not
        # depend on the alpha / critic branch: it runs on its own stream, forked from and joined back into the current
        # one (two parallel branches of the graph when captured).  The parameter sets the two branches write gradients
        # for are disjoint.
        with torch.no_grad():
            alpha_static.copy_(torch.exp(log_alpha.detach()))
        if actor_stream is None:
            la, lq, _ = grads_alpha_q()
            lp, s2r, distance = grads_actor(alpha_static)
        else:
            main = torch.cuda.current_stream(dev)
            actor_stream.wait_stream(main)
            with torch.cuda.stream(actor_stream):
                lp, s2r, distance = grads_actor(alpha_static)
            la, lq, _ = grads_alpha_q()
            main.wait_stream(actor_stream)
        return dict(alpha_loss=la, critic_loss=lq, actor_loss=lp, sim2real_loss=s2r, rsr_distribution_distance=distance,
                    alpha=alpha_static)

    graph_a = graph_b = graph_c = None
    if use_cuda_graph:
        for p in q_params + pol_params + [log_alpha]:
            p.grad = torch.zeros_like(p)
        saved = {k: v.detach().clone() for k, v in net.state_dict().items()}
        static.zero_(); noise.zero_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                sgd_eager(); opt_alpha.step(); opt_q.step(); opt_pi.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        net.load_state_dict(saved)
        with torch.no_grad():
            log_alpha.zero_()
        for o in (opt_alpha, opt_q, opt_pi):
            for st_ in o.state.values():
                for v in st_.values():
                    if torch.is_tensor(v):
                        v.zero_()
        graph_a = torch.cuda.CUDAGraph()   # forward/backward of the three losses
        with torch.cuda.graph(graph_a):
            metrics = sgd_eager()
        graph_b = torch.cuda.CUDAGraph()   # optimiser steps, polyak
        with torch.cuda.graph(graph_b):
            opt_alpha.step(); opt_q.step(); opt_pi.step()
            finish()

    def sgd_step():
        nonlocal metrics
        static.copy_(buffer.sample(batch_size, gen))
        noise.normal_(generator=gen)
        if sgd_probe_fn is not None:  # tests: everything the three losses of this step are computed from
            sgd_probe_fn("before", dict(static=static, noise=noise, log_alpha=log_alpha, net=net, target=target, norm=norm,
                                        fields=fields))
        if graph_a is not None:
            graph_a.replay()
            _flat_allreduce_mean(q_params + pol_params + [log_alpha])
            graph_b.replay()
        else:
            metrics = sgd_eager()
            _flat_allreduce_mean(q_params + pol_params + [log_alpha])
            opt_alpha.step(); opt_q.step(); opt_pi.step()
            finish()
        if sgd_probe_fn is not None:
            sgd_probe_fn("after", dict(metrics=metrics, log_alpha=log_alpha, net=net, target=target))

    def make_policy(deterministic: bool = deterministic_eval):
        @torch.no_grad()
        def policy(obs, generator=None):
            logits = net.policy(normalize(obs[..., :obs_size]))
            return NormalTanh.mode(logits) if deterministic else torch.tanh(NormalTanh.sample_raw(logits, generator))
        return policy

    # brax sac.train: evaluation before training (num_evals > 1) and after every epoch, on rank 0
    evaluator = None
    if run_evals and rank == 0:
        if eval_env is None:
            from . import prng
            rfn = getattr(env, "_randomization_fn", None)
            eval_env = env.clone(num_eval_envs, randomization_fn=rfn,
                                 randomization_rng=prng.split(prng.PRNGKey(seed + 2), num_eval_envs) if rfn else None)
        evaluator = Evaluator(eval_env, lambda: make_policy(deterministic_eval), num_eval_envs, episode_length,
                              action_repeat, seed)
    metrics_out: Dict[str, Any] = {}
    if evaluator is not None and num_evals > 1:
        metrics_out = evaluator.run_evaluation({})
        progress_fn(0, metrics_out)

    for _ in range(num_prefill_actor_steps):
        actor_step()
    env_steps = num_prefill_env_steps
    t_start = time.time()
    eval_every = max(steps_per_epoch, 1)
    t0, steps_since = time.time(), 0
    for it in range(total_steps):
        actor_step()
        for _ in range(grad_updates_per_step):
            sgd_step()
        env_steps += env_steps_per_actor_step
        steps_since += 1
        if (it + 1) % eval_every == 0 or it + 1 == total_steps:
            torch.cuda.synchronize(dev)
            dt = time.time() - t0
            tm = {f"training/{k}": float(v.detach()) for k, v in metrics.items()}
            tm["training/sps"] = steps_since * env_steps_per_actor_step / dt
            tm["training/walltime"] = time.time() - t_start
            tm["training/reward_mean"] = float(state.reward.mean())
            tm["training/replay_size"] = float(buffer.size)
            metrics_out = evaluator.run_evaluation(tm) if evaluator is not None else tm
            if rank == 0:
                progress_fn(env_steps, metrics_out)
            t0, steps_since = time.time(), 0

    return make_policy, (norm, net), metrics_out
