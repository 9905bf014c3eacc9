"""PPO with the RSR transition-distribution term, on the batched CUDA env.

Mirrors the reference's fork of brax PPO:
  RSR/train.py:76-503   `train(environment, num_timesteps, episode_length, past_data, ...)`
  RSR/losses.py:39-205  `compute_gae`, `compute_ppo_loss` (+ `rsr.compute_rsr_loss` on `mode(logits)`)
with the same hyper-parameter names, data flow (unroll -> [B*M, T] batch -> normaliser update ->
num_updates_per_batch x shuffled minibatches) and metrics keys.  Differences, all on the caller side of the hot
path: torch instead of jax; the env is already the wrapped, batched stack (no `envs.training.wrap`); data
parallelism is one process per GPU with envs sharded across ranks, the gradient `pmean`
(`gradients.gradient_update_fn(..., pmap_axis_name)`, RSR/train.py:261-262) is ONE `all_reduce` of a flat gradient
buffer per minibatch step and the normaliser `psum` (train.py:333-336) one small all-reduce per training step.
The launch-bound minibatch step (two tiny MLPs) can be captured in a CUDA graph (`use_cuda_graph`).

[upstream-recall] brax 0.12.1 details restated here: MLP with swish activations and lecun-uniform kernels,
policy (32,)*4 / value (256,)*5 by default, NormalTanhDistribution(min_std=0.001) with entropy evaluated at one
sample, running_statistics normaliser (std clipped to [1e-6, 1e6]).
"""
from __future__ import annotations

import math
import time
from typing import Any, Callable, Dict, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import rsr_loss as rsr


# ------------------------------------------------------------------------ networks
_ACT = {"none": 0, "silu": 1, "relu": 2}
_FUSED_MIN_ROWS = 512  # below this torch's own reduction is as fast (SAC batches of 128-256 rows)


class _LinearActFn(torch.autograd.Function):
    """y = act(x W^T + b) on CUDA with the backward's activation derivative and bias gradient in one launch
    (`rsrx_act_bias_backward`) instead of an elementwise kernel plus a dim-0 reduction per layer."""

    @staticmethod
    def forward(ctx, x, weight, bias, act: int, workspace):
        z = torch.addmm(bias, x, weight.t())
        y = F.silu(z) if act == 1 else (F.relu(z) if act == 2 else z)
        ctx.save_for_backward(x, weight, z if act else x.new_empty(0), workspace)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import _lib
        x, weight, z, workspace = ctx.saved_tensors
        gy = gy.contiguous()
        rows, cols = gy.shape
        gz = torch.empty_like(gy) if ctx.act else gy
        db = torch.empty(cols, device=gy.device, dtype=gy.dtype)
        with torch.cuda.device(gy.device):
            _lib.check(_lib.lib().rsrx_act_bias_backward(
                gy.data_ptr(), z.data_ptr() if ctx.act else None, rows, cols, ctx.act, gz.data_ptr() if ctx.act else None,
                db.data_ptr(), workspace.data_ptr(), torch.cuda.current_stream(gy.device).cuda_stream),
                "rsrx_act_bias_backward")
        gx = gz @ weight if ctx.needs_input_grad[0] else None
        gw = gz.t() @ x
        return gx, gw, db, None, None


def linear_act(x, layer: nn.Linear, act: str, workspaces: dict):
    """act(layer(x)).  CUDA float32 tensors that need gradients go through the fused backward; everything else
    (CPU tensors in the host tests, no-grad inference, small batches) through plain torch."""
    rows = x.numel() // max(x.shape[-1], 1)
    if (x.device.type != "cuda" or rows < _FUSED_MIN_ROWS or not torch.is_grad_enabled()
            or not (layer.weight.requires_grad or x.requires_grad)):
        z = layer(x)
        return F.silu(z) if act == "silu" else (F.relu(z) if act == "relu" else z)
    from . import _lib
    x2 = x.reshape(-1, x.shape[-1])
    key = (x2.shape[0], layer.out_features, id(layer))
    ws = workspaces.get(key)
    if ws is None or ws.device != x.device:
        n = _lib.lib().rsrx_act_bias_backward_workspace(x2.shape[0], layer.out_features)
        ws = workspaces[key] = torch.zeros(n, device=x.device)
    y = _LinearActFn.apply(x2, layer.weight, layer.bias, _ACT[act], ws)
    return y.reshape(*x.shape[:-1], layer.out_features)


class MLP(nn.Module):
    activation = "silu"

    def __init__(self, sizes: Sequence[int]):
        super().__init__()
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))
        for l in self.layers:  # flax lecun_uniform kernel, zero bias
            bound = math.sqrt(3.0 / l.in_features)
            nn.init.uniform_(l.weight, -bound, bound)
            nn.init.zeros_(l.bias)
        self._ws: dict = {}

    def forward(self, x):
        for i, l in enumerate(self.layers):
            x = linear_act(x, l, self.activation if i + 1 < len(self.layers) else "none", self._ws)
        return x


class PPONetworks(nn.Module):
    def __init__(self, obs_size: int, action_size: int, policy_hidden=(32,) * 4, value_hidden=(256,) * 5):
        super().__init__()
        self.policy = MLP([obs_size, *policy_hidden, 2 * action_size])
        self.value = MLP([obs_size, *value_hidden, 1])
        self.action_size = action_size


class NormalTanh:
    """brax NormalTanhDistribution: raw ~ N(loc, softplus(s) + min_std), action = tanh(raw)."""
    min_std = 0.001

    @staticmethod
    def params(logits):
        loc, s = torch.chunk(logits, 2, dim=-1)
        return loc, F.softplus(s) + NormalTanh.min_std

    @staticmethod
    def _log_det_jac(x):
        return 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))

    @classmethod
    def sample_raw(cls, logits, gen=None):
        loc, scale = cls.params(logits)
        return loc + scale * torch.randn(loc.shape, device=loc.device, dtype=loc.dtype, generator=gen)

    @classmethod
    def log_prob(cls, logits, raw):
        loc, scale = cls.params(logits)
        lp = -0.5 * ((raw - loc) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(scale)
        return (lp - cls._log_det_jac(raw)).sum(-1)

    @classmethod
    def entropy(cls, logits, noise):
        loc, scale = cls.params(logits)
        ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale)
        return (ent + cls._log_det_jac(loc + scale * noise)).sum(-1)

    @classmethod
    def mode(cls, logits):
        return torch.tanh(cls.params(logits)[0])

    @staticmethod
    def act(logits, noise, raw_out=None, action_out=None, log_prob_out=None):
        """sample_no_postprocessing + postprocess + log_prob for the actor step in one CUDA launch
        (`rsrx_tanh_normal_act`): returns (raw, action, log_prob); the `_out` tensors are written in place."""
        from . import _lib
        N, A2 = logits.shape
        A = A2 // 2
        raw = raw_out if raw_out is not None else torch.empty(N, A, device=logits.device)
        action = action_out if action_out is not None else torch.empty(N, A, device=logits.device)
        lp = log_prob_out if log_prob_out is not None else torch.empty(N, device=logits.device)
        for t in (logits, noise, raw, action, lp):
            if t.device.type != "cuda" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("NormalTanh.act needs contiguous float32 CUDA tensors")
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().rsrx_tanh_normal_act(logits.data_ptr(), noise.data_ptr(), N, A, raw.data_ptr(), action.data_ptr(),
                                                       lp.data_ptr(), torch.cuda.current_stream(logits.device).cuda_stream),
                       "rsrx_tanh_normal_act")
        return raw, action, lp


# ---------------------------------------------------------------------- normaliser
class RunningStatistics:
    """brax.training.acme.running_statistics (count / mean / summed variance / std)."""

    def __init__(self, size: int, device):
        self.count = torch.zeros((), device=device)
        self.mean = torch.zeros(size, device=device)
        self.summed_variance = torch.zeros(size, device=device)
        self.std = torch.ones(size, device=device)

    def update(self, batch: torch.Tensor, std_min=1e-6, std_max=1e6):
        x = batch.reshape(-1, batch.shape[-1]).float()
        n = torch.tensor(float(x.shape[0]), device=x.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(n)
        count = self.count + n
        diff = x - self.mean
        s1 = diff.sum(0)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(s1)
        mean = self.mean + s1 / count
        s2 = (diff * (x - mean)).sum(0)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(s2)
        # in place: a captured CUDA graph keeps reading these tensors
        self.summed_variance.add_(s2)
        self.mean.copy_(mean)
        self.count.copy_(count)
        self.std.copy_(torch.sqrt(torch.clamp(self.summed_variance, min=0) / count).clamp(std_min, std_max))

    def normalize(self, x):
        return (x - self.mean) / self.std


# ----------------------------------------------------------------------- checkpoints
def save_params(path, params, *, normalize_observations: bool = True, extra: Optional[Dict[str, Any]] = None) -> None:
    """Writes `(normalizer, networks)` as returned by `train` to one file.  Takes the place of the Orbax checkpoint of
    test/rsr_policy_training.py:213-222 (`policy_params_fn`); the layout is a plain dict of tensors and ints so it can
    be read without this package."""
    norm, net = params
    pol_sizes = [net.policy.layers[0].in_features] + [l.out_features for l in net.policy.layers]
    val_sizes = [net.value.layers[0].in_features] + [l.out_features for l in net.value.layers]
    torch.save({"format": "rsr_mjx_b200.ppo/1", "policy_sizes": pol_sizes, "value_sizes": val_sizes,
                "normalize_observations": bool(normalize_observations),
                "normalizer": {k: getattr(norm, k).detach().cpu() for k in ("count", "mean", "summed_variance", "std")},
                "networks": {k: v.detach().cpu() for k, v in net.state_dict().items()}, "extra": dict(extra or {})}, path)


def load_params(path, device="cuda"):
    """Inverse of `save_params`: returns `((normalizer, networks), meta)` on `device`."""
    ck = torch.load(path, map_location="cpu", weights_only=True)
    if ck.get("format") != "rsr_mjx_b200.ppo/1":
        raise ValueError(f"{path}: not a rsr_mjx_b200 PPO checkpoint (format={ck.get('format')!r})")
    pol, val = ck["policy_sizes"], ck["value_sizes"]
    net = PPONetworks(pol[0], pol[-1] // 2, tuple(pol[1:-1]), tuple(val[1:-1]))
    net.load_state_dict(ck["networks"])
    net = net.to(device)
    norm = RunningStatistics(pol[0], device)
    for k, v in ck["normalizer"].items():
        getattr(norm, k).copy_(v.to(device))
    return (norm, net), {"normalize_observations": ck["normalize_observations"], "extra": ck["extra"]}


def make_inference_fn(params, normalize_observations: bool = True):
    """`make_policy(deterministic)` over saved parameters (ppo_inference.py:49-69: restore, then act)."""
    norm, net = params
    normalize = norm.normalize if normalize_observations else (lambda x: x)

    def make_policy(deterministic: bool = False):
        @torch.no_grad()
        def policy(obs, generator=None):
            logits = net.policy(normalize(obs))
            return NormalTanh.mode(logits) if deterministic else torch.tanh(NormalTanh.sample_raw(logits, generator))
        return policy
    return make_policy


# ----------------------------------------------------------------------- GAE / loss
def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float = 1.0, discount: float = 0.99):
    """RSR/losses.py:39-95; all inputs [T, B], bootstrap_value [B]. Returns (vs, advantages), detached."""
    truncation_mask = 1 - truncation
    values_t_plus_1 = torch.cat([values[1:], bootstrap_value[None]], 0)
    deltas = (rewards + discount * (1 - termination) * values_t_plus_1 - values) * truncation_mask
    acc = torch.zeros_like(bootstrap_value)
    out = []
    for t in range(truncation.shape[0] - 1, -1, -1):
        acc = deltas[t] + discount * (1 - termination[t]) * truncation_mask[t] * lambda_ * acc
        out.append(acc)
    vs_minus_v_xs = torch.stack(out[::-1], 0)
    vs = vs_minus_v_xs + values
    vs_t_plus_1 = torch.cat([vs[1:], bootstrap_value[None]], 0)
    advantages = (rewards + discount * (1 - termination) * vs_t_plus_1 - values) * truncation_mask
    return vs.detach(), advantages.detach()


def compute_ppo_loss(net: PPONetworks, normalize: Callable, data: Dict[str, torch.Tensor], noise: torch.Tensor,
                     past_data: Any = None, entropy_cost: float = 1e-4, discounting: float = 0.9,
                     reward_scaling: float = 1.0, gae_lambda: float = 0.95, clipping_epsilon: float = 0.3,
                     normalize_advantage: bool = True, rsr_loss_scale: float = 1.0):
    """RSR/losses.py:98-205.  `data` leaves have leading dims [B, T]."""
    d = {k: v.transpose(0, 1) for k, v in data.items()}  # time first
    obs_n = normalize(d["observation"])
    policy_logits = net.policy(obs_n)
    baseline = net.value(obs_n).squeeze(-1)
    bootstrap_value = net.value(normalize(d["next_observation"][-1])).squeeze(-1)
    rewards = d["reward"] * reward_scaling
    truncation = d["truncation"]
    termination = (1 - d["discount"]) * (1 - truncation)
    target_lp = NormalTanh.log_prob(policy_logits, d["raw_action"])
    vs, advantages = compute_gae(truncation, termination, rewards, baseline, bootstrap_value, gae_lambda, discounting)
    if normalize_advantage:
        advantages = (advantages - advantages.mean()) / (advantages.std(unbiased=False) + 1e-8)
    rho_s = torch.exp(target_lp - d["log_prob"])
    policy_loss = -torch.mean(torch.minimum(rho_s * advantages,
                                            torch.clamp(rho_s, 1 - clipping_epsilon, 1 + clipping_epsilon) * advantages))
    v_error = vs - baseline
    v_loss = torch.mean(v_error * v_error) * 0.5 * 0.5
    entropy_loss = entropy_cost * -torch.mean(NormalTanh.entropy(policy_logits, noise))
    task_loss = policy_loss + v_loss + entropy_loss
    # the RSR term uses the action of the policy being optimised (raw, unnormalised observations)
    sim2real_loss, distance = rsr.compute_rsr_loss(d["observation"], NormalTanh.mode(policy_logits), d["next_observation"],
                                                   past_data, loss_scale=rsr_loss_scale)
    total = task_loss + sim2real_loss
    return total, {"total_loss": total, "task_loss": task_loss, "policy_loss": policy_loss, "v_loss": v_loss,
                   "entropy_loss": entropy_loss, "sim2real_loss": sim2real_loss, "rsr_distribution_distance": distance}


class _PPOHeadFn(torch.autograd.Function):
    """Everything between the network outputs and the task loss in one launch (`rsrx_ppo_head`, csrc/rsrx_ppo.cuh).
    Inputs batch-major [B, T, ...]; returns (task_loss, stats[4] = task/policy/value/entropy losses, detached)."""

    @staticmethod
    def forward(ctx, logits, baseline, bootstrap, raw_action, log_prob, reward, discount, truncation, noise, hyper):
        from . import _lib
        if logits.device.type != "cuda":
            raise RuntimeError("the fused PPO head runs only on CUDA tensors (no CPU fallback)")
        B, T, A2 = logits.shape
        A = A2 // 2
        f = lambda x: x.detach().float().contiguous()
        logits_c, baseline_c, args = f(logits), f(baseline), [f(v) for v in (bootstrap, raw_action, log_prob, reward,
                                                                           discount, truncation, noise)]
        if tuple(baseline_c.shape) != (B, T) or tuple(args[1].shape) != (B, T, A) or tuple(args[6].shape) != (B, T, A):
            raise ValueError("rsrx_ppo_head: expected baseline [B,T], raw_action/noise [B,T,A], logits [B,T,2A]")
        dev = logits.device
        ws = torch.empty(2 * B * T, device=dev)
        out = torch.empty(4, device=dev)
        g_logits = torch.empty(B, T, A2, device=dev)
        g_base = torch.empty(B, T, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsrx_ppo_head(
                logits_c.data_ptr(), baseline_c.data_ptr(), *[a.data_ptr() for a in args], B, T, A,
                float(hyper["reward_scaling"]), float(hyper["discounting"]), float(hyper["gae_lambda"]),
                float(hyper["clipping_epsilon"]), float(hyper["entropy_cost"]), int(bool(hyper["normalize_advantage"])),
                ws.data_ptr(), out.data_ptr(), g_logits.data_ptr(), g_base.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream), "rsrx_ppo_head")
        ctx.save_for_backward(g_logits, g_base)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        g_logits, g_base = ctx.saved_tensors
        return (g_logits * g_loss, g_base * g_loss) + (None,) * 8


def compute_ppo_loss_fused(net: PPONetworks, normalize: Callable, data: Dict[str, torch.Tensor], noise: torch.Tensor,
                           past_data: Any = None, entropy_cost: float = 1e-4, discounting: float = 0.9,
                           reward_scaling: float = 1.0, gae_lambda: float = 0.95, clipping_epsilon: float = 0.3,
                           normalize_advantage: bool = True, rsr_loss_scale: float = 1.0):
    """`compute_ppo_loss` with the loss head in one CUDA launch.  `data` leaves [B, T, ...], `noise` [B, T, A].
    The value network sees the T observations and the bootstrap observation of every sequence in one batch."""
    obs = data["observation"]
    B, T, O = obs.shape
    obs_n = normalize(obs)
    policy_logits = net.policy(obs_n)
    values = net.value(torch.cat([obs_n, normalize(data["next_observation"][:, -1:])], dim=1)).squeeze(-1)  # [B, T + 1]
    baseline, bootstrap = values[:, :T], values[:, T].detach()
    hyper = dict(reward_scaling=reward_scaling, discounting=discounting, gae_lambda=gae_lambda,
                 clipping_epsilon=clipping_epsilon, entropy_cost=entropy_cost, normalize_advantage=normalize_advantage)
    task_loss, stats = _PPOHeadFn.apply(policy_logits, baseline, bootstrap, data["raw_action"], data["log_prob"],
                                        data["reward"], data["discount"], data["truncation"], noise, hyper)
    sim2real_loss, distance = rsr.compute_rsr_loss(obs, NormalTanh.mode(policy_logits), data["next_observation"],
                                                   past_data, loss_scale=rsr_loss_scale)
    total = task_loss + sim2real_loss
    return total, {"total_loss": total, "task_loss": task_loss, "policy_loss": stats[1], "v_loss": stats[2],
                   "entropy_loss": stats[3], "sim2real_loss": sim2real_loss, "rsr_distribution_distance": distance}


# ---------------------------------------------------------------------- evaluation
class Evaluator:
    """brax.training.acting.Evaluator + envs.training.EvalWrapper (RSR/train.py:441-447, :452, :482): every eval env
    runs one episode; reward and every env metric are summed while the episode is active (up to and including the step
    that ends it), `episode_steps` is the wrapper's step counter at that step."""

    def __init__(self, eval_env, make_policy: Callable, num_eval_envs: int, episode_length: int, action_repeat: int = 1,
                 seed: int = 0):
        if eval_env.num_envs != num_eval_envs:
            raise ValueError(f"eval env has {eval_env.num_envs} envs, num_eval_envs={num_eval_envs}")
        self.env, self.make_policy = eval_env, make_policy
        self.num_eval_envs, self.steps = num_eval_envs, episode_length // action_repeat
        self._steps_per_unroll = episode_length * num_eval_envs
        self._eval_walltime = 0.0
        self._seed, self._calls = seed, 0
        self._gen = torch.Generator(device=eval_env.device).manual_seed(seed * 104729 + 17)

    @torch.no_grad()
    def run_evaluation(self, training_metrics: Dict[str, float], aggregate_episodes: bool = True) -> Dict[str, Any]:
        from . import prng
        env, N = self.env, self.num_eval_envs
        t0 = time.time()
        key = prng.split(prng.PRNGKey(self._seed + 7919), self._calls + 2)[-1]  # a fresh unroll key per call
        self._calls += 1
        state = env.reset(prng.split(key, N))
        policy = self.make_policy()
        active = torch.ones(N, device=env.device)
        ep = {"reward": torch.zeros(N, device=env.device)}
        ep.update({k: torch.zeros(N, device=env.device) for k in state.metrics})
        ep_steps = torch.zeros(N, device=env.device)
        for _ in range(self.steps):
            env.step(state, policy(state.obs, self._gen))
            ep["reward"] += state.reward * active
            for k, v in state.metrics.items():
                ep[k] += v * active
            ep_steps = torch.where(active > 0, state.info["steps"].float(), ep_steps)
            active = active * (1 - state.done)
        torch.cuda.synchronize(env.device)
        dt = time.time() - t0
        self._eval_walltime += dt
        metrics: Dict[str, Any] = {}
        for suffix, fn in (("", torch.mean), ("_std", lambda v: torch.std(v, unbiased=False))):
            for k, v in ep.items():
                metrics[f"eval/episode_{k}{suffix}"] = float(fn(v)) if aggregate_episodes else v.cpu().numpy()
        metrics["eval/avg_episode_length"] = float(ep_steps.mean())
        metrics["eval/epoch_eval_time"] = dt
        metrics["eval/sps"] = self._steps_per_unroll / dt
        return {"eval/walltime": self._eval_walltime, **training_metrics, **metrics}


# ----------------------------------------------------------------------- training
def _flat_allreduce_mean(params):
    """gradient pmean (RSR/train.py:261-262): one all-reduce of a flat buffer; flatten, scale and scatter-back are one
    kernel each (multi-tensor copy), not one per parameter"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(dist.get_world_size())
    torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))


def train(environment, num_timesteps: int, episode_length: int, past_data: Any = None, action_repeat: int = 1,
          num_envs: int = 1, num_eval_envs: int = 128, learning_rate: float = 1e-4, entropy_cost: float = 1e-4,
          discounting: float = 0.9, seed: int = 0, unroll_length: int = 10, batch_size: int = 32,
          num_minibatches: int = 16, num_updates_per_batch: int = 2, num_evals: int = 1,
          normalize_observations: bool = False, reward_scaling: float = 1.0, clipping_epsilon: float = 0.3,
          gae_lambda: float = 0.95, rsr_loss_scale: float = 1.0, normalize_advantage: bool = True,
          policy_hidden=(32,) * 4, value_hidden=(256,) * 5,
          progress_fn: Callable[[int, Dict[str, float]], None] = lambda *a: None,
          use_cuda_graph: bool = True, fused_head: bool = True, allow_tf32: bool = True, graph_collect: bool = True,
          epoch_graph: bool = True, tensor_core_value: bool = True,
          max_training_steps: Optional[int] = None, num_resets_per_eval: int = 0, deterministic_eval: bool = False,
          eval_env=None, policy_params_fn: Callable[..., None] = lambda *a: None, run_evals: bool = True,
          training_step_fn: Optional[Callable[[int, Dict[str, float]], None]] = None,
          network_factory: Any = None, randomization_fn: Optional[Callable] = None,
          restore_checkpoint_path: Optional[str] = None, **brax_plumbing):
    """PPO training (RSR/train.py:76).  `environment` is an `AirbotPlayBase` with `num_envs` envs on this rank
    (under torch.distributed every rank passes its shard; `num_envs` is the per-rank count here).

    Epoch structure of the reference (RSR/train.py:185-196, :449-492): `max(num_evals - 1, 1)` epochs of
    `num_training_steps_per_epoch` training steps; rank 0 evaluates `num_eval_envs` single episodes before the first
    epoch (if num_evals > 1) and after every epoch and calls `progress_fn(env_steps, {eval/..., training/...})` and
    `policy_params_fn(env_steps, make_policy, params)`.  `run_evals=False` skips the evaluator (benchmarks),
    `training_step_fn(step, training_metrics)` is called after every training step.
    `tensor_core_value` (default, with `fused_head` on CUDA): the value network's forward / backward run on the
    hand-written tcgen05 linear kernels (fused_mlp.py) instead of torch autograd + cuBLAS.
    Reference arguments that configure the run are honoured or raise, never dropped (train_args.py):
    `network_factory` (a functools.partial over make_ppo_networks: its *_hidden_layer_sizes are used),
    `randomization_fn` (installed on the env with the keys RSR/train.py:212-217 derives from `seed`),
    `restore_checkpoint_path` (own torch file or a brax `model.save_params` pickle; like the reference, a path that does
    not exist is skipped; an Orbax directory raises with the conversion recipe — checkpoints.py).
    Returns (make_policy, (normalizer, networks), metrics)."""
    import os
    from . import checkpoints, train_args
    env = environment
    train_args.reject_unknown(brax_plumbing, "ppo.train")
    _sizes = train_args.hidden_sizes(network_factory, dict(policy_hidden_layer_sizes=policy_hidden,
                                                           value_hidden_layer_sizes=value_hidden))
    policy_hidden, value_hidden = _sizes["policy_hidden_layer_sizes"], _sizes["value_hidden_layer_sizes"]
    train_args.apply_randomization(env, randomization_fn, seed)
    past_data = rsr.prepare_rsr_data(past_data, env.device)
    graph_collect_enabled = bool(graph_collect and use_cuda_graph)
    # the two MLPs run their matmuls on the tensor cores in TF32, the precision jax gives float32 `dot` on NVIDIA GPUs
    # by default (the reference never raises `jax_default_matmul_precision`); everything else stays fp32
    torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    if env.num_envs != num_envs:
        raise ValueError(f"environment has {env.num_envs} envs, num_envs={num_envs}")
    if env.episode_length != episode_length:
        raise ValueError("environment.episode_length differs from episode_length (the env is already wrapped)")
    assert batch_size * num_minibatches % num_envs == 0
    dev = env.device
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    env_step_per_training_step = batch_size * unroll_length * num_minibatches * action_repeat * world
    num_evals_after_init = max(num_evals - 1, 1)
    steps_per_epoch = int(math.ceil(num_timesteps / (num_evals_after_init * env_step_per_training_step
                                                      * max(num_resets_per_eval, 1)))) * max(num_resets_per_eval, 1)
    num_training_steps = steps_per_epoch * num_evals_after_init
    if max_training_steps is not None:
        num_training_steps = min(num_training_steps, max_training_steps)

    torch.manual_seed(seed)  # networks identical on every rank (key_policy / key_value are global in the reference)
    net = PPONetworks(env.observation_size, env.action_size, policy_hidden, value_hidden).to(dev)
    params = list(net.parameters())
    opt = torch.optim.Adam(params, lr=learning_rate, eps=1e-8, capturable=bool(use_cuda_graph), fused=True)
    norm = RunningStatistics(env.observation_size, dev)
    if restore_checkpoint_path is not None and os.path.exists(str(restore_checkpoint_path)):  # RSR/train.py:410-422
        r_norm, r_net = checkpoints.restore(str(restore_checkpoint_path), "ppo", dev)
        net.load_state_dict(r_net.state_dict())
        for k in ("count", "mean", "summed_variance", "std"):
            getattr(norm, k).copy_(getattr(r_norm, k))
    normalize = norm.normalize if normalize_observations else (lambda x: x)
    gen = torch.Generator(device=dev).manual_seed(seed * 7919 + rank + 1)

    from . import sharding
    state = env.reset(sharding.shard_keys(seed, num_envs, rank, world))
    T, n_unrolls = unroll_length, batch_size * num_minibatches // num_envs
    B = batch_size * num_minibatches
    obs_size, act_size = env.observation_size, env.action_size
    buf = {k: torch.empty(n_unrolls, T, num_envs, *s, device=dev) for k, s in
           dict(observation=(obs_size,), next_observation=(obs_size,), raw_action=(act_size,), log_prob=(), reward=(),
                discount=(), truncation=()).items()}

    act_noise = torch.empty(n_unrolls, T, num_envs, act_size, device=dev)  # N(0,1) of the behaviour policy, per unroll step
    action_buf = torch.empty(num_envs, act_size, device=dev)

    # the unrolls re-laid out as [B = n_unrolls * N sequences, T, ...]: fixed buffers, so that the SGD epoch graph can
    # bake their addresses in
    batch_static = {k: torch.empty(B, T, *v.shape[3:], device=dev) for k, v in buf.items()}

    # actor-step policy forward: one warp-per-row launch instead of ~10 (csrc/rsrx_mlp.cuh) when the policy fits (<= 32 wide)
    from . import fused_mlp as _fm
    actor_mlp = None
    if tensor_core_value and fused_head and dev.type == "cuda" and _fm.warp_supported(net.policy, dev):
        actor_mlp = _fm.WarpMLP(net.policy, num_envs, dev)

    copy_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    @torch.no_grad()
    def collect_body():
        # Critical path of an actor step: policy forward (observation normalised inside the launch) -> action head ->
        # env.step.  The copies of the step's results into the rollout buffers run on a second stream (a parallel branch
        # of the graph when captured) while the next action is being computed; env.step waits for them, since it
        # overwrites what they read.
        main = torch.cuda.current_stream(dev) if copy_stream is not None else None

        def store_results(u, t):  # of the env.step that was just queued, + the observation the next actor step sees
            buf["next_observation"][u, t].copy_(state.obs)
            buf["reward"][u, t].copy_(state.reward)
            buf["discount"][u, t].copy_(1 - state.done)
            buf["truncation"][u, t].copy_(state.info["truncation"])

        prev = None
        for u in range(n_unrolls):
            for t in range(T):
                obs = state.obs
                if copy_stream is not None:
                    copy_stream.wait_stream(main)
                    with torch.cuda.stream(copy_stream):
                        if prev is not None:
                            store_results(*prev)
                        buf["observation"][u, t].copy_(obs)
                else:
                    buf["observation"][u, t].copy_(obs)
                if actor_mlp is not None:
                    logits = (actor_mlp.forward(obs, norm.mean, norm.std) if normalize_observations
                              else actor_mlp.forward(obs))
                else:
                    logits = net.policy(normalize(obs))
                # raw action, its log-prob (straight into the rollout buffers) and the tanh action: one launch
                NormalTanh.act(logits, act_noise[u, t], buf["raw_action"][u, t], action_buf, buf["log_prob"][u, t])
                if copy_stream is not None:
                    main.wait_stream(copy_stream)
                env.step(state, action_buf)
                if copy_stream is not None:
                    prev = (u, t)
                else:
                    store_results(u, t)
        if copy_stream is not None:
            store_results(*prev)
        # [n_unrolls, T, N, ...] -> [B = n_unrolls * N, T, ...]
        for k, v in buf.items():
            batch_static[k].view(n_unrolls, num_envs, T, *v.shape[3:]).copy_(v.permute(0, 2, 1, *range(3, v.dim())))

    collect_graph = None

    @torch.no_grad()
    def collect():
        """n_unrolls x T actor steps into `buf` (one CUDA-graph replay when captured: the ~45 small launches around
        every env.step then run back to back)"""
        act_noise.normal_(generator=gen)
        if collect_graph is not None:
            collect_graph.replay()
        else:
            collect_body()
        return batch_static

    mb = B // num_minibatches
    static = {k: torch.empty(mb, T, *v.shape[3:], device=dev) for k, v in buf.items()}
    # entropy-estimate noise: [B, T, A] for the fused head, time-major for the eager reference path
    static_noise = torch.empty(*((mb, T) if fused_head else (T, mb)), act_size, device=dev)
    loss_fn = compute_ppo_loss_fused if fused_head else compute_ppo_loss
    loss_kw = dict(past_data=past_data, entropy_cost=entropy_cost, discounting=discounting, reward_scaling=reward_scaling,
                   gae_lambda=gae_lambda, clipping_epsilon=clipping_epsilon, normalize_advantage=normalize_advantage,
                   rsr_loss_scale=rsr_loss_scale)
    last_metrics: Dict[str, torch.Tensor] = {}

    def fwd_bwd_autograd(noise=None):
        opt.zero_grad(set_to_none=True)
        loss, metrics = loss_fn(net, normalize, static, static_noise if noise is None else noise, **loss_kw)
        loss.backward()
        return metrics

    # ---- value network on the tensor-core kernels: no autograd through it.  Rows of its batch: the mb * T observations
    # (sequence-major) followed by the mb bootstrap observations, so that baseline / bootstrap / their gradients are
    # contiguous views.
    from . import fused_mlp
    use_tc = bool(tensor_core_value and fused_head and fused_mlp.supported(net.value, dev))
    vtc = ptc = None
    if use_tc:
        rows_v = mb * T + mb
        vtc = fused_mlp.TensorCoreMLP(net.value, rows_v, dev)
        if fused_mlp.warp_supported(net.policy, dev):  # the reference's policy (32,)*4: one launch per direction
            ptc = fused_mlp.WarpMLP(net.policy, mb * T, dev)
        # one-launch Adam over all ~22 tensors (same update rule as torch.optim.Adam / optax.adam)
        # (its step also keeps the transposed weight copies of the value network current: the dgrad GEMM's TMA operand)
        opt = fused_mlp.FusedAdam(params, lr=learning_rate, eps=1e-8, transposed=vtc.WT)
        vtc.refresh_transposed_weights()
        vtc.wt_fresh = True
        head_ws = torch.empty(2 * mb * T, device=dev)
        head_out = torch.empty(4, device=dev)
        g_logits_buf = torch.empty(mb, T, 2 * act_size, device=dev)
        g_values_buf = torch.zeros(rows_v, device=dev)  # the bootstrap rows keep a zero gradient (vs is stop-gradient)
        hyper = (float(reward_scaling), float(discounting), float(gae_lambda), float(clipping_epsilon), float(entropy_cost),
                 int(bool(normalize_advantage)))
        obs_n_buf = torch.empty(mb * T, obs_size, device=dev)
        policy_stream = torch.cuda.Stream(device=dev)
        zero_scalar = torch.zeros((), device=dev)
        rsr_term = None
        if ptc is not None and past_data is not None and rsr_loss_scale != 0:
            rsr_term = rsr.PolicyTerm(past_data, mb * T, obs_size, act_size, rsr_loss_scale, dev)
            g_total_buf = torch.empty(mb, T, 2 * act_size, device=dev)
        # the normaliser's tensors are updated in place, so the prep kernel can keep reading them; identity when off
        prep_mean = norm.mean if normalize_observations else torch.zeros(obs_size, device=dev)
        prep_std = norm.std if normalize_observations else torch.ones(obs_size, device=dev)

    def fwd_bwd_tc(noise=None):
        from . import _lib
        nz = static_noise if noise is None else noise
        opt.zero_grad(set_to_none=True)
        vtc.attach_grads()
        obs = static["observation"]
        # normalised policy / value inputs (+ the padded, transposed copies the tensor-core layers read): one launch
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsrx_ppo_prep(obs.data_ptr(), static["next_observation"].data_ptr(), prep_mean.data_ptr(),
                                                prep_std.data_ptr(), mb, T, obs_size, obs_n_buf.data_ptr(), vtc.x_pad.data_ptr(),
                                                vtc.k0p, vtc.xT.data_ptr(), vtc.ldt, torch.cuda.current_stream(dev).cuda_stream),
                       "rsrx_ppo_prep")
        obs_n = obs_n_buf.view(mb, T, obs_size)
        main = torch.cuda.current_stream(dev)
        # The policy branch (its forward, the RSR term, its backward) and the value branch (five tensor-core layers each way)
        # only meet at the loss head: the policy branch runs on its own stream, forked from and joined back into the current
        # one (two parallel branches of the graph when captured), so the critical path of a minibatch step is
        # prep -> value forward -> head -> value data-gradients -> reduce -> Adam.
        pol = policy_stream if ptc is not None else main
        if ptc is not None:
            ptc.attach_grads()
            pol.wait_stream(main)
            with torch.cuda.stream(pol):
                logits = ptc.forward(obs_n_buf).view(mb, T, 2 * act_size)
                if rsr_term is not None:
                    # the RSR term acts on mode(logits) (RSR/losses.py:186-195): pack + KDE + gradient kernels now, its
                    # gradient w.r.t. the logits joins the head's below
                    rsr_term.forward(obs, ptc.out, static["next_observation"])
                    sim2real_loss, distance = rsr_term.loss, rsr_term.distance
                else:
                    sim2real_loss = distance = zero_scalar
        else:
            logits = net.policy(obs_n)  # autograd fallback for policies wider than 32
        with torch.no_grad():
            values = vtc.forward(None)
            if ptc is not None:
                main.wait_stream(pol)  # the head needs the logits
            lg = logits.detach()
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().rsrx_ppo_head(
                    lg.data_ptr(), values.data_ptr(), values[mb * T:].data_ptr(), static["raw_action"].data_ptr(),
                    static["log_prob"].data_ptr(), static["reward"].data_ptr(), static["discount"].data_ptr(),
                    static["truncation"].data_ptr(), nz.data_ptr(), mb, T, act_size, *hyper, head_ws.data_ptr(),
                    head_out.data_ptr(), g_logits_buf.data_ptr(), g_values_buf.data_ptr(),
                    main.cuda_stream), "rsrx_ppo_head")
            if ptc is not None:
                pol.wait_stream(main)  # the head's logit gradients
                with torch.cuda.stream(pol):
                    if rsr_term is not None:
                        rsr_term.add_logit_grad(g_logits_buf, g_total_buf)
                    ptc.backward(g_total_buf if rsr_term is not None else g_logits_buf)
            vtc.backward(g_values_buf)
            if ptc is not None:
                main.wait_stream(pol)
        if ptc is None:
            sim2real_loss, distance = rsr.compute_rsr_loss(obs, NormalTanh.mode(logits), static["next_observation"], past_data,
                                                           loss_scale=rsr_loss_scale)
            if sim2real_loss.requires_grad:
                torch.autograd.backward([logits, sim2real_loss], [g_logits_buf, None])
            else:
                torch.autograd.backward([logits], [g_logits_buf])
        task_loss = head_out[0]
        return {"total_loss": task_loss + sim2real_loss.detach(), "task_loss": task_loss, "policy_loss": head_out[1],
                "v_loss": head_out[2], "entropy_loss": head_out[3], "sim2real_loss": sim2real_loss.detach(),
                "rsr_distribution_distance": distance.detach()}

    fwd_bwd = fwd_bwd_tc if use_tc else fwd_bwd_autograd

    # One graph for the whole SGD phase of a training step: num_updates_per_batch x num_minibatches minibatch steps, each
    # = gather + forward/backward + gradient all-reduce (NCCL is graph-capturable) + Adam, replayed with fresh
    # permutations and entropy noise written into fixed buffers beforehand.  One launch from the host instead of
    # ~1000 (the update was launch-bound: 0.43 ms per minibatch step for ~40 small kernels).
    n_mb_steps = num_updates_per_batch * num_minibatches
    epoch_graph_enabled = bool(epoch_graph and use_cuda_graph)
    perm_all = torch.empty(num_updates_per_batch, B, dtype=torch.int64, device=dev)
    noise_all = torch.empty(n_mb_steps, *static_noise.shape, device=dev)  # entropy noise of a whole SGD phase, one draw

    graph_bwd = graph_opt = None
    if use_cuda_graph:
        # warm-up on a side stream (allocator, cuBLAS handles, Adam state), roll the warm-up's parameter / optimizer
        # changes back, then capture forward+backward and the Adam step as two graphs; the (NCCL) gradient
        # all-reduce runs between them
        for p in params:
            p.grad = torch.zeros_like(p)
        for k in static:
            static[k].zero_()
        static["discount"].fill_(1.0)
        static_noise.zero_()
        saved = {k: v.detach().clone() for k, v in net.state_dict().items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                fwd_bwd()
                opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        net.load_state_dict(saved)
        for st_ in opt.state.values():
            for v in st_.values():
                if torch.is_tensor(v):
                    v.zero_()
        if not epoch_graph_enabled:
            graph_bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_bwd):
                last_metrics = fwd_bwd()
            graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_opt):
                opt.step()
        if graph_collect_enabled:
            # capture only records: the env state is not advanced here; env.step is a plain launch on the capture stream
            collect_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(collect_graph):
                collect_body()

    import ctypes as C
    from . import _lib
    _keys = list(static)
    _nf = len(_keys)
    _dst = (C.c_void_p * _nf)(*[static[k].data_ptr() for k in _keys])
    _widths = (C.c_int32 * _nf)(*[static[k][0].numel() for k in _keys])

    def gather_minibatch(batch, idx):
        # all seven fields of the minibatch in one gather launch (rows = sequences of T transitions)
        src = (C.c_void_p * _nf)(*[batch[k].data_ptr() for k in _keys])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsrx_gather_rows(src, _dst, _widths, _nf, idx.data_ptr(), mb,
                                                   torch.cuda.current_stream(dev).cuda_stream), "rsrx_gather_rows")

    sgd_epoch_graph = None
    if epoch_graph_enabled:
        perm_all.copy_(torch.arange(B, device=dev).expand(num_updates_per_batch, B))
        noise_all.zero_()
        for v in batch_static.values():
            v.zero_()
        batch_static["discount"].fill_(1.0)
        saved = {k: v.detach().clone() for k, v in net.state_dict().items()}
        sgd_epoch_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(sgd_epoch_graph):
            k_ = 0
            for u_ in range(num_updates_per_batch):
                for m_ in range(num_minibatches):
                    gather_minibatch(batch_static, perm_all[u_, m_ * mb:(m_ + 1) * mb])
                    last_metrics = fwd_bwd(noise_all[k_])
                    _flat_allreduce_mean(params)
                    opt.step()
                    k_ += 1
        # capture only records; make sure nothing of the set-up leaks into the parameters / optimiser state
        net.load_state_dict(saved)
        for st_ in opt.state.values():
            for v in st_.values():
                if torch.is_tensor(v):
                    v.zero_()

    def sgd_epoch(batch):
        """the SGD phase of one training step (RSR/train.py:279-299): num_updates_per_batch shuffles x num_minibatches"""
        for u_ in range(num_updates_per_batch):
            perm_all[u_].copy_(torch.randperm(B, device=dev, generator=gen))
        noise_all.normal_(generator=gen)
        if sgd_epoch_graph is not None:
            sgd_epoch_graph.replay()
            return
        for u_ in range(num_updates_per_batch):
            for m_ in range(num_minibatches):
                static_noise.copy_(noise_all[u_ * num_minibatches + m_])
                minibatch_step(batch, perm_all[u_, m_ * mb:(m_ + 1) * mb])

    def minibatch_step(batch, idx):
        nonlocal last_metrics
        gather_minibatch(batch, idx)
        if graph_bwd is not None:
            graph_bwd.replay()
            _flat_allreduce_mean(params)
            graph_opt.replay()
        else:
            last_metrics = fwd_bwd()
            _flat_allreduce_mean(params)
            opt.step()

    def make_policy(deterministic: bool = False):
        @torch.no_grad()
        def policy(obs, generator=None):
            logits = net.policy(normalize(obs))
            return NormalTanh.mode(logits) if deterministic else torch.tanh(NormalTanh.sample_raw(logits, generator))
        return policy

    evaluator = None
    if run_evals and rank == 0:
        if eval_env is None:
            from . import prng
            rfn = getattr(env, "_randomization_fn", None)
            eval_env = env.clone(num_eval_envs, randomization_fn=rfn,
                                 randomization_rng=prng.split(prng.PRNGKey(seed + 2), num_eval_envs) if rfn else None)
        evaluator = Evaluator(eval_env, lambda: make_policy(deterministic_eval), num_eval_envs, episode_length,
                              action_repeat, seed)

    metrics_out: Dict[str, float] = {}
    final_metrics: Dict[str, Any] = {}
    if evaluator is not None and num_evals > 1:
        final_metrics = evaluator.run_evaluation({})
        progress_fn(0, final_metrics)
    t_start = time.time()
    env_steps = 0
    for it in range(num_training_steps):
        t0 = time.time()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        data = collect()
        ev[1].record()
        if normalize_observations:
            norm.update(data["observation"])
        sgd_epoch(data)
        ev[2].record()
        torch.cuda.synchronize(dev)
        dt = time.time() - t0
        env_steps += env_step_per_training_step
        metrics_out = {f"training/{k}": float(v.detach()) for k, v in last_metrics.items()}
        metrics_out["training/sps"] = env_step_per_training_step / dt
        metrics_out["training/walltime"] = time.time() - t_start
        metrics_out["training/collect_s"] = ev[0].elapsed_time(ev[1]) * 1e-3  # device time of the unrolls
        metrics_out["training/update_s"] = ev[1].elapsed_time(ev[2]) * 1e-3   # ... of the SGD epochs
        metrics_out["training/reward_mean"] = float(data["reward"].mean())
        if training_step_fn is not None:
            training_step_fn(it, metrics_out)
        if num_resets_per_eval > 0 and (it + 1) % max(steps_per_epoch // num_resets_per_eval, 1) == 0:
            fresh = env.reset(sharding.shard_keys(seed + 1 + it, num_envs, rank, world))  # RSR/train.py:473-478
            for k_, v_ in state._buf.items():  # in place: a captured collect graph keeps reading these buffers
                v_.copy_(fresh._buf[k_])
        if (it + 1) % steps_per_epoch == 0 or it + 1 == num_training_steps:
            final_metrics = evaluator.run_evaluation(metrics_out) if evaluator is not None else dict(metrics_out)
            if rank == 0:
                progress_fn(env_steps, final_metrics)
                policy_params_fn(env_steps, make_policy, (norm, net))

    return make_policy, (norm, net), final_metrics or metrics_out
