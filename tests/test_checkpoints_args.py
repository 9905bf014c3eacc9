"""Reference trainer arguments are honoured or raise (nothing silently dropped), and policies saved by the
reference stack (`brax.io.model.save_params` pickles) are importable without brax / flax / jax."""
import dataclasses
import functools
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from rsr_mjx_b200 import checkpoints, prng, train_args


def _fake_brax_pickle(with_value=True, nested=False, seed=0):
    """A pickle with the module / class names brax writes: RunningStatisticsState + flax param dicts (+ jax-style array
    wrappers).  The fake modules exist only while dumping; loading must work without them."""
    rng = np.random.default_rng(seed)
    mods = {}
    for name in ("brax", "brax.training", "brax.training.acme", "brax.training.acme.running_statistics",
                 "brax.training.agents", "brax.training.agents.ppo", "brax.training.agents.ppo.losses", "flax", "flax.core",
                 "flax.core.frozen_dict"):
        mods[name] = types.ModuleType(name)

    @dataclasses.dataclass
    class RunningStatisticsState:
        count: np.ndarray
        mean: np.ndarray
        summed_variance: np.ndarray
        std: np.ndarray
    RunningStatisticsState.__module__ = "brax.training.acme.running_statistics"
    RunningStatisticsState.__qualname__ = "RunningStatisticsState"
    mods["brax.training.acme.running_statistics"].RunningStatisticsState = RunningStatisticsState

    @dataclasses.dataclass
    class PPONetworkParams:
        policy: dict
        value: dict
    PPONetworkParams.__module__ = "brax.training.agents.ppo.losses"
    PPONetworkParams.__qualname__ = "PPONetworkParams"
    mods["brax.training.agents.ppo.losses"].PPONetworkParams = PPONetworkParams

    def mlp(sizes):
        return {"params": {f"hidden_{i}": {"kernel": rng.normal(size=(a, b)).astype(np.float32),
                                           "bias": rng.normal(size=(b,)).astype(np.float32)}
                           for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]))}}
    norm = RunningStatisticsState(np.float32(1234.0), rng.normal(size=23).astype(np.float32),
                                  rng.uniform(1, 2, 23).astype(np.float32), rng.uniform(0.5, 1.5, 23).astype(np.float32))
    pol, val = mlp([23, 32, 32, 32, 32, 10]), mlp([23, 64, 64, 1])
    tree = (norm, PPONetworkParams(pol, val)) if nested else ((norm, pol, val) if with_value else (norm, pol))
    sys.modules.update(mods)
    try:
        blob = pickle.dumps(tree)
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    return blob, norm, pol, val


@pytest.mark.parametrize("nested", [False, True])
def test_brax_ppo_pickle_import(nested):
    blob, norm, pol, val = _fake_brax_pickle(nested=nested)
    assert "brax" not in sys.modules
    tree = checkpoints.load_brax_params(blob)
    r_norm, net = checkpoints.ppo_params_from_brax(tree, device="cpu")
    assert [l.out_features for l in net.policy.layers] == [32, 32, 32, 32, 10]
    assert [l.out_features for l in net.value.layers] == [64, 64, 1]
    np.testing.assert_array_equal(r_norm.mean.numpy(), norm.mean)
    np.testing.assert_array_equal(r_norm.std.numpy(), norm.std)
    assert float(r_norm.count) == 1234.0
    # the imported network computes what flax's MLP (swish, kernel [in, out]) computes
    x = np.random.default_rng(1).normal(size=(5, 23)).astype(np.float32)
    h = x
    for i in range(5):
        p = pol["params"][f"hidden_{i}"]
        h = h @ p["kernel"] + p["bias"]
        if i < 4:
            h = h / (1 + np.exp(-h))
    np.testing.assert_allclose(net.policy(torch.from_numpy(x)).detach().numpy(), h, rtol=1e-5, atol=1e-5)


def test_brax_sac_pickle_import_and_restore_dispatch(tmp_path):
    blob, norm, pol, _ = _fake_brax_pickle(with_value=False)
    f = tmp_path / "sac_params"
    f.write_bytes(blob)
    r_norm, net = checkpoints.restore(str(f), "sac", device="cpu")
    assert [l.out_features for l in net.policy.layers] == [32, 32, 32, 32, 10]
    d = tmp_path / "orbax_ckpt"
    d.mkdir()
    with pytest.raises(NotImplementedError, match="Orbax"):
        checkpoints.restore(str(d), "ppo", device="cpu")


def test_brax_pickle_export_round_trip(tmp_path):
    """save_brax_params writes the layout `brax.io.model.load_params` expects — the normaliser as a global reference to
    brax.training.acme.running_statistics.RunningStatisticsState, flax-shaped parameter dicts with [in, out] kernels — and
    leaves no fake module behind; our own loader reads it back to identical networks.  A step counter stored as
    brax's UInt64(hi, lo) is accepted on import."""
    import pickletools
    from rsr_mjx_b200 import ppo
    torch.manual_seed(0)
    norm = ppo.RunningStatistics(23, "cpu")
    norm.update(torch.randn(100, 23))
    net = ppo.PPONetworks(23, 5)
    f = tmp_path / "ppo_params"
    checkpoints.save_brax_params(str(f), (norm, net), "ppo")
    assert not any(m.startswith("brax") for m in sys.modules)
    strings = [a for op, a, _ in pickletools.genops(f.read_bytes()) if op.name in ("SHORT_BINUNICODE", "BINUNICODE")]
    assert "brax.training.acme.running_statistics" in strings and "RunningStatisticsState" in strings
    r_norm, r_net = checkpoints.restore(str(f), "ppo", device="cpu")
    for a, b in zip(net.parameters(), r_net.parameters()):
        assert torch.equal(a, b)
    for k in ("count", "mean", "summed_variance", "std"):
        assert torch.equal(getattr(norm, k), getattr(r_norm, k)), k
    # UInt64 step counter of newer brax versions
    tree = checkpoints.load_brax_params(str(f))
    tree[0]["count"] = {"hi": np.uint32(0), "lo": np.uint32(100)}
    assert float(checkpoints.ppo_params_from_brax(tree, device="cpu")[0].count) == 100.0
    from rsr_mjx_b200 import sac
    snet = sac.SACNetworks(23, 5, (32, 32))
    checkpoints.save_brax_params(str(tmp_path / "sac_params"), (norm, snet), "sac")
    _, r_snet = checkpoints.restore(str(tmp_path / "sac_params"), "sac", device="cpu")
    for a, b in zip(snet.parameters(), r_snet.parameters()):
        assert torch.equal(a, b)


def test_unpickler_refuses_code():
    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned",))
    with pytest.raises(pickle.UnpicklingError):
        checkpoints.load_brax_params(pickle.dumps(Evil()))


def test_network_factory_keywords():
    d = dict(policy_hidden_layer_sizes=(32,) * 4, value_hidden_layer_sizes=(256,) * 5)
    assert train_args.hidden_sizes(None, d) == d
    f = functools.partial(lambda **k: None, policy_hidden_layer_sizes=(32, 32, 32, 32), value_hidden_layer_sizes=(32, 32, 32, 32))
    assert train_args.hidden_sizes(f, d)["value_hidden_layer_sizes"] == (32, 32, 32, 32)
    with pytest.raises(NotImplementedError):
        train_args.hidden_sizes(lambda *a, **k: None, d)
    with pytest.raises(NotImplementedError):
        train_args.hidden_sizes(functools.partial(lambda **k: None, activation="relu"), d)


def test_unknown_and_plumbing_kwargs():
    train_args.reject_unknown(dict(max_devices_per_host=None, wrap_env=False), "ppo.train")
    with pytest.raises(TypeError):
        train_args.reject_unknown(dict(entropy_costs=1.0), "ppo.train")
    with pytest.raises(NotImplementedError):
        train_args.reject_unknown(dict(max_devices_per_host=2), "ppo.train")


def test_randomization_keys_and_apply():
    k = train_args.randomization_keys(0, 8)
    assert k.shape == (8, 2) and k.dtype == np.uint32 and len({tuple(r) for r in k}) == 8
    assert not np.array_equal(k, train_args.randomization_keys(1, 8))
    np.testing.assert_array_equal(prng.fold_in(prng.PRNGKey(3), 0).shape, (2,))

    class Env:
        num_envs, _randomization_fn, got = 8, None, None

        def randomize(self, fn, rng):
            self.got, self._randomization_fn = rng, fn
    e, fn = Env(), (lambda sys_, rng: (sys_, None))
    train_args.apply_randomization(e, fn, 0)
    np.testing.assert_array_equal(e.got, k)
    train_args.apply_randomization(e, fn, 0)  # same fn again: keeps the installed randomisation
    with pytest.raises(ValueError):
        train_args.apply_randomization(e, lambda s, r: (s, None), 0)
