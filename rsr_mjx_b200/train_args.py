"""Arguments of the reference trainers that configure the RUN rather than the maths (`network_factory`,
`randomization_fn`, `restore_checkpoint_path`, brax plumbing) — shared by `ppo.train`, `sac.train` and
`rsr_pipeline.policy_params_training`.  Nothing is dropped silently: an argument is honoured, or raises."""
from __future__ import annotations

import functools
from typing import Any, Dict, Sequence

import numpy as np

from . import prng

# brax plumbing without an equivalent here; accepted only at their no-op values
_NOOP_ONLY = {"max_devices_per_host": (None,), "wrap_env": (False, None), "wrap_env_fn": (None,), "checkpoint_logdir": (None,),
              "init_params": (None,)}


def reject_unknown(kwargs: Dict[str, Any], fn_name: str) -> None:
    """Reference keyword arguments this trainer cannot honour raise instead of being ignored."""
    for k, v in kwargs.items():
        if k in _NOOP_ONLY:
            if not any(v is ok or v == ok for ok in _NOOP_ONLY[k]):
                raise NotImplementedError(
                    f"{fn_name}({k}={v!r}): not supported — the env is already the wrapped, batched stack, devices are "
                    "one process per GPU (torchrun), checkpoints are written by `policy_params_fn` + ppo.save_params")
        else:
            raise TypeError(f"{fn_name}() got an unexpected keyword argument {k!r}")


def hidden_sizes(network_factory, defaults: Dict[str, Sequence[int]]) -> Dict[str, tuple]:
    """The reference passes `functools.partial(ppo_networks.make_ppo_networks, policy_hidden_layer_sizes=...,
    value_hidden_layer_sizes=...)` / `partial(sac_networks.make_sac_networks, hidden_layer_sizes=...)`
    (test/rsr_policy_training.py:261-271, ppo_train/airbot_training/train.py:40-44).  brax is not importable here, so the
    callable itself cannot be run; what it configures — the hidden layer sizes — is read from the partial's keywords (or
    from a plain dict).  Anything else a factory could change (activation, a custom module) is rejected."""
    out = {k: tuple(v) for k, v in defaults.items()}
    if network_factory is None:
        return out
    if isinstance(network_factory, dict):
        kw = dict(network_factory)
    elif isinstance(network_factory, functools.partial):
        if network_factory.args:
            raise NotImplementedError("network_factory: positional partial arguments are not supported")
        kw = dict(network_factory.keywords)
    else:
        raise NotImplementedError(
            "network_factory must be a functools.partial over brax's make_ppo_networks / make_sac_networks (its "
            "*_hidden_layer_sizes keywords are honoured) or a dict of those keywords; arbitrary network builders "
            "cannot be run without brax/flax")
    for k, v in kw.items():
        if k not in out:
            raise NotImplementedError(f"network_factory keyword {k!r} is not supported (supported: {sorted(out)})")
        out[k] = tuple(int(x) for x in v)
    return out


def randomization_keys(seed: int, n: int) -> np.ndarray:
    """The keys RSR/train.py:196-217 hands to `randomization_fn`: key = PRNGKey(seed); _, local = split(key);
    local = fold_in(local, process_id = 0); _, key_env, _ = split(local, 3); split(key_env, n).  Every device (here:
    rank) gets the same keys, as in the reference."""
    key = prng.PRNGKey(seed)
    local = prng.split(key, 2)[1]
    local = prng.fold_in(local, 0)
    key_env = prng.split(local, 3)[1]
    return prng.split(key_env, n)


def apply_randomization(env, randomization_fn, seed: int) -> None:
    """`randomization_fn` given to the trainer (the reference wraps the env with it inside `train`): installs the
    per-env model leaves on the already-batched env.  An env that was constructed with a randomisation keeps it, and
    passing a DIFFERENT function on top raises."""
    if randomization_fn is None:
        return
    have = getattr(env, "_randomization_fn", None)
    if have is not None:
        if have is not randomization_fn:
            raise ValueError("the environment was built with a different randomization_fn than the one passed to train()")
        return
    env.randomize(randomization_fn, randomization_keys(seed, env.num_envs))
