"""RSR pipeline entry points with the reference's signatures
(RSR/rsr_pipeline.py): `env_params_tuning`, `build_policy_rsr_data`,
`policy_params_training`.

env_params_tuning — the reference tunes one scalar (the cube geom's friction,
written to all three coefficients of the LAST geom, rsr_pipeline.py:125-136) by
Adam on `grad(loss_fn)` through `env.step`, one tiny device dispatch per sample
(15 samples x 1000 Adam steps).  Here the same loss
    L(p) = sum_i | w . (obs_pred_i(p) - obs_true_i) |       (rsr_pipeline.py:119-123,146-162)
is evaluated for a whole grid of candidate p in ONE batched launch (P x S envs,
one bare-env step each, per-env geom_friction) and the interval is zoomed around
the minimiser — a forward sweep instead of reverse-mode AD (which MJX's
while_loop solver does not support anyway).
"""
from __future__ import annotations

import time
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import prng, rsr_loss
from .envs import AirbotPlayBase

# compute_error weights, rsr_pipeline.py:120
ERROR_WEIGHTS = (1, 1, 1, 1, 1, 1, 10, 10, 10, 0, 0, 0, 10, 10, 10, 10, 10, 0, 0, 0, 0, 0, 0)


def _scalar(x) -> float:
    if isinstance(x, dict):
        if len(x) != 1:
            raise ValueError("env_params_tuning sweeps exactly one scalar parameter")
        x = next(iter(x.values()))
    return float(np.asarray(x if not torch.is_tensor(x) else x.cpu()).reshape(-1)[0])


class FrictionSweep:
    """Batched evaluation of the env_params_tuning loss for many friction values."""

    def __init__(self, kind: str, obs, actions, next_obs_true, num_params: int = 64, device="cuda", **env_kwargs):
        obs = np.asarray(obs, np.float32)
        actions = np.asarray(actions, np.float32)
        next_obs_true = np.asarray(next_obs_true, np.float32)
        if obs.ndim != 2 or actions.ndim != 2 or next_obs_true.shape != obs.shape or actions.shape[0] != obs.shape[0]:
            raise ValueError("obs[S,23], actions[S,5], next_obs_true[S,23] expected")
        self.S, self.P = obs.shape[0], int(num_params)
        N = self.S * self.P
        # episode_length=0 -> the bare env, as the reference steps `init_env` without wrappers
        self.env = AirbotPlayBase(kind, num_envs=N, episode_length=0, device=device, **env_kwargs)
        env, m, L = self.env, self.env.model, self.env.layout
        one = AirbotPlayBase(kind, num_envs=1, episode_length=0, device=device, **env_kwargs)
        # obs2state (rsr_pipeline.py:75-98): reset(PRNGKey(0)), one step with zero action; per sample overwrite
        # joint angles, cube position in qpos and the (stale) xpos[cube]; info/metrics come from state_1.
        s0 = one.reset(prng.PRNGKey(0)[None])
        row0 = s0._buf["data"].clone()
        s1 = one.step(s0, torch.zeros(1, m.nu, device=device))
        info1 = s1._buf["info"].clone()
        data = row0.repeat(N, 1)
        o = torch.from_numpy(obs).to(device).repeat(self.P, 1)  # env index = p * S + i
        jid = torch.as_tensor(env.joint_id, device=device, dtype=torch.long)
        data[:, L.qpos + jid] = o[:, 0:6]
        cq = env._box_qposadr  # the reference writes qpos[cube_id+2 : cube_id+5] = 15:18 (same address)
        data[:, L.qpos + cq:L.qpos + cq + 3] = o[:, 12:15]
        data[:, L.xpos + 3 * env.cube_id:L.xpos + 3 * env.cube_id + 3] = o[:, 12:15]
        self._data0 = data
        self._info0 = info1.repeat(N, 1)
        self._actions = torch.from_numpy(actions).to(device).repeat(self.P, 1).contiguous()
        self._true = torch.from_numpy(next_obs_true).to(device).repeat(self.P, 1)
        self._w = torch.tensor(ERROR_WEIGHTS, dtype=torch.float32, device=device)
        self._buf = env._alloc()
        self._gf = torch.from_numpy(m.geom_friction.astype(np.float32)).to(device).repeat(N, 1, 1).contiguous()

    def loss(self, params: torch.Tensor) -> torch.Tensor:
        """params[P] -> loss[P] (one launch)."""
        env = self.env
        p = torch.as_tensor(params, dtype=torch.float32, device=self._gf.device).reshape(self.P)
        self._gf[:, -1, :] = p.repeat_interleave(self.S)[:, None]
        env.set_per_env(geom_friction=self._gf)
        b = self._buf
        b["data"].copy_(self._data0)
        b["info"].copy_(self._info0)
        b["done"].zero_()
        env.step_raw(b, self._actions.data_ptr())
        err = (b["obs"][:, :23] - self._true) @ self._w
        return err.abs().reshape(self.P, self.S).sum(1)


def env_params_tuning(init_env: AirbotPlayBase, num_steps: int, init_env_params, env_params_min, env_params_max,
                      obs: Any, actions: Any, next_obs_true: Any, log_path: Optional[str] = None,
                      num_params: int = 64, zoom: float = 0.25):
    """Tune the cube friction to reproduce `next_obs_true` (reference signature,
    rsr_pipeline.py:49-56).  `num_steps` sweeps of `num_params` candidates each; every
    sweep keeps the best candidate and shrinks the interval to `zoom` of its width.

    Returns (tuned_env_params, train_log) like the reference."""
    lo, hi = _scalar(env_params_min), _scalar(env_params_max)
    if not lo < hi:
        raise ValueError("env_params_min must be below env_params_max")
    p0 = min(max(_scalar(init_env_params), lo), hi)
    sweep = FrictionSweep(init_env.kind, obs, actions, next_obs_true, num_params=num_params, device=init_env.device,
                          **init_env._params)
    log: Dict[str, list] = {"time_cost": [], "loss": [], "params": []}
    best_p, best_l = p0, float("inf")
    a, b = lo, hi
    for i in range(max(int(num_steps), 1)):
        t0 = time.time()
        cand = torch.linspace(a, b, sweep.P, device=init_env.device)
        cand[0] = best_p if i else p0  # always re-evaluate the incumbent
        losses = sweep.loss(cand)
        k = int(torch.argmin(losses))
        if float(losses[k]) <= best_l:
            best_p, best_l = float(cand[k]), float(losses[k])
        width = (b - a) * zoom
        a, b = max(lo, best_p - width / 2), min(hi, best_p + width / 2)
        dt = time.time() - t0
        line = f"step {i}: {dt:.2f}s. params = {best_p}. loss = {best_l}."
        log["time_cost"].append(dt)
        log["loss"].append(best_l)
        log["params"].append(best_p)
        if log_path:
            with open(log_path, "a") as f:
                f.write(line + "\n")
    tuned = {k: best_p for k in init_env_params} if isinstance(init_env_params, dict) else best_p
    return tuned, log


def build_policy_rsr_data(past_states: Any, past_actions: Any, past_next_states_real: Any, past_next_states_sim: Any,
                          current_next_states_sim: Any, num_samples: int = 10, min_val: float = -3.0,
                          max_val: float = 3.0, bandwidth: float = 0.1, seed: int = 0, device="cuda") -> rsr_loss.RSRData:
    """Builds the fixed RSR statistics shared by PPO and SAC (rsr_pipeline.py:209-271)."""
    arrays = tuple(torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.float32)
                   for v in (past_states, past_actions, past_next_states_real, past_next_states_sim,
                             current_next_states_sim))
    past_states, past_actions, past_next_states_real, past_next_states_sim, current_next_states_sim = arrays
    if any(v.ndim != 2 for v in arrays):
        raise ValueError(f'all RSR datasets must be rank 2, got {tuple(tuple(v.shape) for v in arrays)}')
    counts = {v.shape[0] for v in arrays}
    if len(counts) != 1:
        raise ValueError(f'RSR datasets must have equal lengths, got {tuple(tuple(v.shape) for v in arrays)}')
    if not counts or next(iter(counts)) == 0:
        raise ValueError('RSR datasets must not be empty')
    if past_next_states_real.shape[1] != past_states.shape[1]:
        raise ValueError('real next-state width must match state width')
    if past_next_states_sim.shape[1] != past_states.shape[1]:
        raise ValueError('previous sim next-state width must match state width')
    if current_next_states_sim.shape[1] != past_states.shape[1]:
        raise ValueError('current sim next-state width must match state width')
    real = torch.hstack([past_states, past_actions, past_next_states_real])
    prev = torch.hstack([past_states, past_actions, past_next_states_sim])
    cur = torch.hstack([past_states, past_actions, current_next_states_sim])
    return rsr_loss.build_rsr_data(real, prev, cur, num_samples=num_samples, min_value=min_val, max_value=max_val,
                                   bandwidth=bandwidth, seed=seed, device=device)


def policy_params_training(env, restore_checkpoint_path: Optional[str] = None, policy_params_fn=None, network_factory=None,
                           progress_fn=None, past_states: Any = None, past_actions: Any = None,
                           past_next_states_real: Any = None, past_next_states_sim: Any = None,
                           current_next_states_sim: Any = None, algorithm: str = "ppo", num_samples: int = 10,
                           min_val: float = -3.0, max_val: float = 3.0, bandwidth: float = 0.1, rsr_loss_scale: float = 1.0,
                           num_timesteps: int = 5_000_000, num_evals: int = 10, reward_scaling: float = 0.1,
                           episode_length: int = 1200, normalize_observations: bool = True, action_repeat: int = 1,
                           discounting: float = 0.96, learning_rate: float = 1e-4, num_envs: int = 512, batch_size: int = 128,
                           seed: int = 0, num_eval_envs: int = 128, deterministic_eval: bool = False,
                           max_devices_per_host: Optional[int] = None, unroll_length: int = 10, num_minibatches: int = 32,
                           num_updates_per_batch: int = 8, entropy_cost: float = 2e-2, num_resets_per_eval: int = 0,
                           **sac_and_extra_options):
    """Trains an RSR policy (reference signature and defaults, rsr_pipeline.py:274-319).

    `env` is a batched `AirbotPlayBase` whose `num_envs` / `episode_length` match the arguments.
    `network_factory` and `restore_checkpoint_path` go to the trainer (ppo.train / sac.train docstrings: hidden sizes
    from the partial's keywords; own torch file or brax pickle, an Orbax directory raises with the conversion recipe;
    SAC refuses to resume, like the reference).  ``algorithm='sac'`` takes tau / min_replay_size / max_replay_size /
    grad_updates_per_step through the keyword options, like the reference (rsr_pipeline.py:396-426); with
    ``algorithm='ppo'`` those four are accepted and unused, as in the reference.  Any other unknown keyword raises.
    Returns ``(make_inference_fn, (normalizer, networks))``."""
    from . import ppo
    if rsr_loss_scale < 0:
        raise ValueError(f'rsr_loss_scale must be non-negative, got {rsr_loss_scale}')
    required = (past_states, past_actions, past_next_states_real, past_next_states_sim, current_next_states_sim)
    if any(v is None for v in required):
        raise ValueError('all five RSR policy datasets are required')
    algorithm = algorithm.strip().lower()
    if algorithm not in ("ppo", "sac"):
        raise ValueError(f'unsupported algorithm {algorithm!r}; expected "ppo" or "sac"')
    _sac_only = ("tau", "min_replay_size", "max_replay_size", "grad_updates_per_step", "hidden_layer_sizes")
    _common = ("use_cuda_graph", "allow_tf32", "max_training_steps", "eval_env", "run_evals", "randomization_fn",
               "checkpoint_logdir", "wrap_env_fn", "wrap_env")
    _ppo_only = ("fused_head", "policy_hidden", "value_hidden", "clipping_epsilon", "gae_lambda", "normalize_advantage",
                 "graph_collect", "training_step_fn")
    for k in sac_and_extra_options:
        if k not in _sac_only + _common + _ppo_only:
            raise TypeError(f"policy_params_training() got an unexpected keyword argument {k!r}")
    if restore_checkpoint_path:
        import os
        if algorithm == "sac":  # rsr_pipeline.py:399-403
            raise ValueError('Brax 0.12.1 SAC cannot resume complete training state; use checkpoint_logdir to save '
                             'inference checkpoints instead')
        if os.path.isdir(str(restore_checkpoint_path)):
            from . import checkpoints
            checkpoints.load_brax_params(str(restore_checkpoint_path))  # raises: Orbax directories cannot be read here
    past_data = build_policy_rsr_data(past_states, past_actions, past_next_states_real, past_next_states_sim,
                                      current_next_states_sim, num_samples=num_samples, min_val=min_val, max_val=max_val,
                                      bandwidth=bandwidth, seed=seed, device=env.device)
    if algorithm == "sac":
        from . import sac
        opts = {k: v for k, v in sac_and_extra_options.items() if k in _sac_only + _common}
        make_inference_fn, params, _ = sac.train(
            environment=env, past_data=past_data, num_timesteps=num_timesteps, num_evals=num_evals,
            num_eval_envs=num_eval_envs, reward_scaling=reward_scaling, episode_length=episode_length,
            normalize_observations=normalize_observations, action_repeat=action_repeat, discounting=discounting,
            learning_rate=learning_rate, num_envs=num_envs, batch_size=batch_size, deterministic_eval=deterministic_eval,
            progress_fn=progress_fn or (lambda *a: None), rsr_loss_scale=rsr_loss_scale, seed=seed,
            network_factory=network_factory, restore_checkpoint_path=restore_checkpoint_path, **opts)
        return make_inference_fn, params
    make_inference_fn, params, _ = ppo.train(
        environment=env, past_data=past_data, num_timesteps=num_timesteps, num_evals=num_evals, num_eval_envs=num_eval_envs,
        reward_scaling=reward_scaling, episode_length=episode_length, normalize_observations=normalize_observations,
        action_repeat=action_repeat, unroll_length=unroll_length, num_minibatches=num_minibatches,
        num_updates_per_batch=num_updates_per_batch, discounting=discounting, learning_rate=learning_rate,
        entropy_cost=entropy_cost, num_envs=num_envs, batch_size=batch_size, progress_fn=progress_fn or (lambda *a: None),
        rsr_loss_scale=rsr_loss_scale, seed=seed, num_resets_per_eval=num_resets_per_eval,
        deterministic_eval=deterministic_eval, policy_params_fn=policy_params_fn or (lambda *a: None),
        network_factory=network_factory, restore_checkpoint_path=restore_checkpoint_path,
        **{k: v for k, v in sac_and_extra_options.items() if k in _common + _ppo_only})
    return make_inference_fn, params
