"""Import of policies trained with the reference stack.

What reference users hold on disk
  * `brax.io.model.save_params(path, params)` files (ppo_train/airbot_training/train.py:91-92, train_sac.py:76-78,
    go2_training/learning/train_jax_sac.py:218-222; RSR/rsr_pipeline.py:399-403 `checkpoint_logdir`): a PICKLE of
    `(normalizer_params, policy_params[, value_params])` — `RunningStatisticsState(count, mean, summed_variance, std)`,
    flax parameter dicts `{'params': {'hidden_i': {'kernel' [in, out], 'bias' [out]}}}`;
  * Orbax `PyTreeCheckpointer` DIRECTORIES (test/rsr_policy_training.py:213-222,
    real_robot_inference/.../ppo_inference.py:47-69 `restore_checkpoint_path`).

What this module reads
  * the pickle files, WITHOUT importing brax / flax / jax (none is installable here) and without executing arbitrary
    pickled callables: a restricted unpickler that rebuilds NumPy arrays (also jax `Array` pickles, which wrap a NumPy
    payload), maps containers (flax FrozenDict, dataclass-like records such as RunningStatisticsState / PPONetworkParams)
    to plain dicts, and rejects everything else;
  * this package's own single-file torch checkpoints (`ppo.save_params`).
What it writes: `save_brax_params` — the same pickle layout from a policy trained here (NumPy leaves, the normaliser as a
`brax.training.acme.running_statistics.RunningStatisticsState` global reference), so the reference's inference code
(`model.load_params` + `make_inference_fn`, ppo_inference.py:49-69 with a params file) can load it where brax is installed.
What it cannot read: Orbax directories (OCDBT / zarr via tensorstore — not present in this image).  Convert them once
where orbax is installed:  `brax.io.model.save_params(out, ocp.PyTreeCheckpointer().restore(ckpt_dir))`.
"""
from __future__ import annotations

import io
import os
import pickle
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np
import torch


class _Record(dict):
    """stand-in for a pickled dataclass / struct / namedtuple-like object: its fields as a dict"""

    def __init__(self, *args, **kwargs):
        super().__init__(**kwargs)
        if args:
            self["_args"] = list(args)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict, slots)
            self.update(state[0] or {})
            self.update(state[1])
        else:
            self["_state"] = state


def _record_class(name):
    return type(name, (_Record,), {"_pickled_name": name})


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """jax._src.array._reconstruct_array: the payload is a pickled NumPy array"""
    a = fun(*args)
    a.__setstate__(arr_state)
    return np.asarray(a)


_NUMPY_OK = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
             ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
             ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.numeric", "_frombuffer"),
             ("numpy._core.numeric", "_frombuffer"), ("numpy", "float32"), ("numpy", "float64"), ("numpy", "int32"),
             ("numpy", "int64"), ("numpy", "bool_"), ("numpy", "uint32")}


class _BraxUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _NUMPY_OK:
            return super().find_class(module, name)
        if module in ("collections",) and name == "OrderedDict":
            return dict
        if module.startswith("jax") and name == "_reconstruct_array":
            return _reconstruct_jax_array
        if module.startswith(("flax.core.frozen_dict",)) and name == "FrozenDict":
            return _Record
        if module.startswith(("brax.", "flax.", "jax.", "jaxlib.", "optax.", "ml_collections.")) or module == "__main__":
            if name[:1].isupper():  # a class: RunningStatisticsState, PPONetworkParams, NestedMeanStd, ...
                return _record_class(name)
        raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}: only NumPy arrays and brax/flax parameter "
                                     "containers are accepted")


def load_brax_params(path_or_bytes) -> Any:
    """The pytree of a `brax.io.model.save_params` file with NumPy leaves (containers become dicts / tuples / lists)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        f = io.BytesIO(path_or_bytes)
    else:
        if os.path.isdir(path_or_bytes):
            raise NotImplementedError(
                f"{path_or_bytes} is a directory — an Orbax PyTreeCheckpointer checkpoint.  Orbax / tensorstore are not "
                "available here; convert it once where they are: brax.io.model.save_params(out, "
                "ocp.PyTreeCheckpointer().restore(ckpt_dir)) and pass the resulting file.")
        f = open(path_or_bytes, "rb")
    with f:
        return _BraxUnpickler(f).load()


def _mlp_layers(tree) -> List[Tuple[np.ndarray, np.ndarray]]:
    """flax MLP parameters {'params': {'hidden_0': {'kernel','bias'}, ...}} -> [(kernel [in,out], bias [out])] in order"""
    p = tree["params"] if isinstance(tree, dict) and "params" in tree else tree
    if not isinstance(p, dict):
        raise ValueError("expected a flax parameter dict")
    if len(p) == 1 and not any(k.startswith("hidden_") for k in p):  # e.g. {'MLP_0': {...}}
        p = next(iter(p.values()))
    names = sorted((k for k in p if k.startswith("hidden_")), key=lambda k: int(k.split("_")[1]))
    if not names:
        raise ValueError(f"no hidden_i layers in {list(p)}")
    return [(np.asarray(p[k]["kernel"], np.float32), np.asarray(p[k]["bias"], np.float32)) for k in names]


def _fill_mlp(mlp, layers, what):
    if len(layers) != len(mlp.layers):
        raise ValueError(f"{what}: checkpoint has {len(layers)} layers, network has {len(mlp.layers)}")
    with torch.no_grad():
        for l, (k, b) in zip(mlp.layers, layers):
            if tuple(k.shape) != (l.in_features, l.out_features):
                raise ValueError(f"{what}: kernel {k.shape} does not fit layer ({l.in_features}, {l.out_features})")
            l.weight.copy_(torch.from_numpy(k.T.copy()))
            l.bias.copy_(torch.from_numpy(b))


def _normalizer(tree, size, device):
    from .ppo import RunningStatistics
    norm = RunningStatistics(size, device)
    if tree is None:
        return norm
    t = tree if isinstance(tree, dict) else dict(zip(("count", "mean", "summed_variance", "std"), tree))
    for k in ("count", "mean", "summed_variance", "std"):
        v = t[k]
        if isinstance(v, dict) and "hi" in v and "lo" in v:  # brax.training.types.UInt64(hi, lo) step counter
            v = float(int(np.asarray(v["hi"])) * 2.0 ** 32 + int(np.asarray(v["lo"])))
        v = np.asarray(v, np.float32)
        getattr(norm, k).copy_(torch.from_numpy(v.reshape(getattr(norm, k).shape)).to(device))
    return norm


def _split(tree) -> Sequence[Any]:
    if isinstance(tree, dict) and "_args" in tree:
        return tree["_args"]
    if isinstance(tree, (tuple, list)):
        return list(tree)
    raise ValueError("expected (normalizer_params, policy_params[, value_params])")


def ppo_params_from_brax(tree, device="cuda"):
    """(normalizer_params, policy_params, value_params) or (normalizer_params, PPONetworkParams(policy, value)) ->
    `(RunningStatistics, PPONetworks)` as `ppo.train` returns them; layer sizes are read from the kernels."""
    from .ppo import PPONetworks
    parts = _split(tree)
    norm_t, rest = parts[0], parts[1:]
    if len(rest) == 1 and isinstance(rest[0], dict) and "policy" in rest[0]:
        pol_t, val_t = rest[0]["policy"], rest[0].get("value")
    else:
        pol_t, val_t = rest[0], (rest[1] if len(rest) > 1 else None)
    pol = _mlp_layers(pol_t)
    obs, act2 = pol[0][0].shape[0], pol[-1][0].shape[1]
    val = _mlp_layers(val_t) if val_t is not None else None
    net = PPONetworks(obs, act2 // 2, tuple(k.shape[1] for k, _ in pol[:-1]),
                      tuple(k.shape[1] for k, _ in val[:-1]) if val else (256,) * 5)
    _fill_mlp(net.policy, pol, "policy")
    if val:
        _fill_mlp(net.value, val, "value")
    return _normalizer(norm_t, obs, device), net.to(device)


def sac_params_from_brax(tree, device="cuda"):
    """(normalizer_params, policy_params[, q_params]) of brax SAC -> `(RunningStatistics, SACNetworks)`.  brax's q
    network stacks the critics along the output axis of every layer (`n_critics = 2` parallel MLPs named
    hidden_0..; or a list of two MLPs): both layouts are accepted; without q parameters the critics keep their init."""
    from .sac import SACNetworks
    parts = _split(tree)
    pol = _mlp_layers(parts[1])
    obs, act2 = pol[0][0].shape[0], pol[-1][0].shape[1]
    net = SACNetworks(obs, act2 // 2, tuple(k.shape[1] for k, _ in pol[:-1]))
    _fill_mlp(net.policy, pol, "policy")
    if len(parts) > 2 and parts[2] is not None:
        q = parts[2]["params"] if "params" in parts[2] else parts[2]
        heads = sorted(k for k in q if isinstance(q[k], dict) and any(n.startswith("hidden_") for n in q[k]))
        if len(heads) >= 2:
            _fill_mlp(net.q1, _mlp_layers(q[heads[0]]), "q1")
            _fill_mlp(net.q2, _mlp_layers(q[heads[1]]), "q2")
    return _normalizer(parts[0], obs, device), net.to(device)


def _flax_mlp(mlp) -> Dict[str, Any]:
    """torch MLP -> {'params': {'hidden_i': {'kernel' [in, out], 'bias' [out]}}} with NumPy leaves"""
    return {"params": {f"hidden_{i}": {"kernel": l.weight.detach().t().contiguous().cpu().numpy().astype(np.float32),
                                       "bias": l.bias.detach().cpu().numpy().astype(np.float32)}
                       for i, l in enumerate(mlp.layers)}}


def save_brax_params(path: str, params, algorithm: str = "ppo") -> None:
    """Writes `(normalizer, networks)` as returned by `ppo.train` / `sac.train` in the layout of
    `brax.io.model.save_params`: a pickle of `(RunningStatisticsState, policy_params, value_params)` (PPO) or
    `(RunningStatisticsState, policy_params, {'params': {'q1': ..., 'q2': ...}})` (SAC; brax's own q-network naming
    depends on its version, the policy is what inference needs).  The normaliser is pickled as a reference to
    `brax.training.acme.running_statistics.RunningStatisticsState` (fields count / mean / summed_variance / std), so
    `brax.io.model.load_params` rebuilds the real class where brax is installed; this package reads the file back
    through `load_brax_params` without brax."""
    import sys
    import types as _types
    norm, net = params
    mod_name, cls_name = "brax.training.acme.running_statistics", "RunningStatisticsState"
    cls = type(cls_name, (), {"__module__": mod_name})
    state = cls()
    state.__dict__.update(count=np.float32(float(norm.count)), mean=norm.mean.detach().cpu().numpy().astype(np.float32),
                          summed_variance=norm.summed_variance.detach().cpu().numpy().astype(np.float32),
                          std=norm.std.detach().cpu().numpy().astype(np.float32))
    if algorithm == "ppo":
        tree = (state, _flax_mlp(net.policy), _flax_mlp(net.value))
    elif algorithm == "sac":
        tree = (state, _flax_mlp(net.policy), {"params": {"q1": _flax_mlp(net.q1)["params"], "q2": _flax_mlp(net.q2)["params"]}})
    else:
        raise ValueError(f"unsupported algorithm: {algorithm}")
    # pickle verifies a class by importing its module: stand in for the (absent) brax modules while dumping
    added = []
    try:
        parts = mod_name.split(".")
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            if name not in sys.modules:
                sys.modules[name] = _types.ModuleType(name)
                added.append(name)
        had = getattr(sys.modules[mod_name], cls_name, None)
        setattr(sys.modules[mod_name], cls_name, cls)
        try:
            with open(path, "wb") as f:
                pickle.dump(tree, f, protocol=4)
        finally:
            if had is None:
                delattr(sys.modules[mod_name], cls_name)
            else:
                setattr(sys.modules[mod_name], cls_name, had)
    finally:
        for name in added:
            sys.modules.pop(name, None)


def restore(path: str, algorithm: str = "ppo", device="cuda"):
    """`restore_checkpoint_path` of the trainers: own torch file, brax pickle, or a clear error for an Orbax directory."""
    from . import ppo
    if os.path.isdir(path):
        load_brax_params(path)  # raises the NotImplementedError that explains the conversion
    with open(path, "rb") as f:
        head = f.read(4)
    if head[:2] == b"PK":  # torch zip container
        if algorithm != "ppo":
            raise ValueError("own-format checkpoints exist for PPO only")
        return ppo.load_params(path, device)[0]
    tree = load_brax_params(path)
    return ppo_params_from_brax(tree, device) if algorithm == "ppo" else sac_params_from_brax(tree, device)
