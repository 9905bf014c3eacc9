"""Host-side PPO pieces against NumPy transcriptions of the reference formulas
(RSR/losses.py:39-95 GAE; brax NormalTanhDistribution), CPU only."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rsr_mjx_b200 import ppo


def _np_gae(truncation, termination, rewards, values, bootstrap, lam, disc):
    T = rewards.shape[0]
    mask = 1 - truncation
    v1 = np.concatenate([values[1:], bootstrap[None]], 0)
    deltas = (rewards + disc * (1 - termination) * v1 - values) * mask
    acc = np.zeros_like(bootstrap)
    out = np.zeros_like(values)
    for t in reversed(range(T)):
        acc = deltas[t] + disc * (1 - termination[t]) * mask[t] * lam * acc
        out[t] = acc
    vs = out + values
    vs1 = np.concatenate([vs[1:], bootstrap[None]], 0)
    adv = (rewards + disc * (1 - termination) * vs1 - values) * mask
    return vs, adv


def test_gae_matches_reference_formula():
    rng = np.random.default_rng(0)
    T, B = 10, 7
    trunc = (rng.random((T, B)) < 0.1).astype(np.float64)
    term = (rng.random((T, B)) < 0.1).astype(np.float64) * (1 - trunc)
    r, v, bs = rng.normal(size=(T, B)), rng.normal(size=(T, B)), rng.normal(size=B)
    vs, adv = ppo.compute_gae(*[torch.from_numpy(x) for x in (trunc, term, r, v, bs)], lambda_=0.95, discount=0.96)
    vs_ref, adv_ref = _np_gae(trunc, term, r, v, bs, 0.95, 0.96)
    np.testing.assert_allclose(vs.numpy(), vs_ref, rtol=1e-12)
    np.testing.assert_allclose(adv.numpy(), adv_ref, rtol=1e-12)


def test_normal_tanh_distribution():
    torch.manual_seed(0)
    logits = torch.randn(5, 10, dtype=torch.float64)
    raw = torch.randn(5, 5, dtype=torch.float64)
    loc, scale = ppo.NormalTanh.params(logits)
    assert (scale > 0.001).all()
    base = torch.distributions.Normal(loc, scale)
    ref = (base.log_prob(raw) - torch.log(1 - torch.tanh(raw) ** 2)).sum(-1)
    np.testing.assert_allclose(ppo.NormalTanh.log_prob(logits, raw).numpy(), ref.numpy(), rtol=1e-9)
    noise = torch.randn(5, 5, dtype=torch.float64)
    x = loc + scale * noise
    ent_ref = (base.entropy() + torch.log(1 - torch.tanh(x) ** 2)).sum(-1)
    np.testing.assert_allclose(ppo.NormalTanh.entropy(logits, noise).numpy(), ent_ref.numpy(), rtol=1e-8)
    np.testing.assert_allclose(ppo.NormalTanh.mode(logits).numpy(), torch.tanh(loc).numpy())


def test_running_statistics_matches_numpy():
    rs = ppo.RunningStatistics(3, "cpu")
    rng = np.random.default_rng(1)
    chunks = [rng.normal(2.0, 3.0, (50, 4, 3)) for _ in range(4)]
    for c in chunks:
        rs.update(torch.from_numpy(c))
    allx = np.concatenate([c.reshape(-1, 3) for c in chunks])
    np.testing.assert_allclose(rs.mean.numpy(), allx.mean(0), rtol=1e-5)
    np.testing.assert_allclose(rs.std.numpy(), allx.std(0), rtol=1e-4)
    np.testing.assert_allclose(rs.normalize(torch.from_numpy(allx[:5]).float()).numpy(), (allx[:5] - allx.mean(0)) / allx.std(0), rtol=1e-3, atol=1e-4)


def test_ppo_loss_runs_on_cpu_without_rsr_term():
    torch.manual_seed(0)
    net = ppo.PPONetworks(23, 5)
    assert sum(p.numel() for p in net.policy.parameters()) == 23 * 32 + 32 + 3 * (32 * 32 + 32) + 32 * 10 + 10
    B, T = 12, 10
    data = dict(observation=torch.randn(B, T, 23), next_observation=torch.randn(B, T, 23), raw_action=torch.randn(B, T, 5),
                log_prob=torch.randn(B, T), reward=torch.randn(B, T), discount=torch.ones(B, T), truncation=torch.zeros(B, T))
    loss, m = ppo.compute_ppo_loss(net, lambda x: x, data, torch.randn(T, B, 5), past_data=None, entropy_cost=2e-2,
                                   discounting=0.96, reward_scaling=0.1)
    loss.backward()
    assert torch.isfinite(loss) and m["sim2real_loss"].item() == 0.0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert set(m) == {"total_loss", "task_loss", "policy_loss", "v_loss", "entropy_loss", "sim2real_loss", "rsr_distribution_distance"}


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = ppo.PPONetworks(6, 2, (8, 8), (8,))
        x = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 + rank))
        (net.policy(x).sum() + net.value(x).sum()).backward()
        local = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
        ppo._flat_allreduce_mean(list(net.parameters()))
        q.put((rank, local.numpy(), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_pmean_world2():
    world, port = 2, 29617 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=90) for _ in range(world)], key=lambda r: r[0])
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    mean = (res[0][1] + res[1][1]) / 2
    assert not np.allclose(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][2], mean, rtol=1e-6)
    np.testing.assert_allclose(res[1][2], mean, rtol=1e-6)


def test_policy_params_training_validation_matches_reference():
    """rsr_pipeline.py:335-347 — checked before any device work"""
    from rsr_mjx_b200 import rsr_pipeline as RP
    with pytest.raises(ValueError, match="rsr_loss_scale must be non-negative"):
        RP.policy_params_training(None, rsr_loss_scale=-1.0)
    with pytest.raises(ValueError, match="all five RSR policy datasets are required"):
        RP.policy_params_training(None, past_states=np.zeros((3, 23)))
    z = np.zeros((3, 23))
    with pytest.raises(ValueError, match="unsupported algorithm 'td3'"):
        RP.policy_params_training(None, past_states=z, past_actions=np.zeros((3, 5)), past_next_states_real=z,
                                  past_next_states_sim=z, current_next_states_sim=z, algorithm="TD3")
    kw = dict(past_states=z, past_actions=np.zeros((3, 5)), past_next_states_real=z, past_next_states_sim=z,
              current_next_states_sim=z)
    with pytest.raises(NotImplementedError, match="Orbax"):  # a directory = an Orbax PyTreeCheckpointer checkpoint
        RP.policy_params_training(None, restore_checkpoint_path=os.path.dirname(os.path.abspath(__file__)), **kw)
    with pytest.raises(ValueError, match="SAC cannot resume"):  # rsr_pipeline.py:399-403
        RP.policy_params_training(None, algorithm="sac", restore_checkpoint_path="/x", **kw)
    with pytest.raises(TypeError, match="unexpected keyword argument 'entropy_costs'"):  # nothing is dropped silently
        RP.policy_params_training(None, entropy_costs=1.0, **kw)
    with pytest.raises(ValueError, match="RSR datasets must have equal lengths"):
        RP.build_policy_rsr_data(z, np.zeros((4, 5)), z, z, z)
    with pytest.raises(ValueError, match="real next-state width must match state width"):
        RP.build_policy_rsr_data(z, np.zeros((3, 5)), np.zeros((3, 22)), z, z)
    with pytest.raises(ValueError, match="all RSR datasets must be rank 2"):
        RP.build_policy_rsr_data(np.zeros(3), np.zeros((3, 5)), z, z, z)


def test_checkpoint_round_trip(tmp_path):
    """save_params / load_params / make_inference_fn (the Orbax slot of test/rsr_policy_training.py:213-222)"""
    torch.manual_seed(3)
    net = ppo.PPONetworks(23, 5, (32,) * 4, (64,) * 2)
    norm = ppo.RunningStatistics(23, "cpu")
    norm.update(torch.randn(100, 23) * 2 + 1)
    path = tmp_path / "policy.pt"
    ppo.save_params(path, (norm, net), extra={"step": 7})
    (norm2, net2), meta = ppo.load_params(path, device="cpu")
    assert meta["normalize_observations"] is True and meta["extra"] == {"step": 7}
    for k, v in net.state_dict().items():
        assert torch.equal(v, net2.state_dict()[k])
    for k in ("count", "mean", "summed_variance", "std"):
        assert torch.equal(getattr(norm, k), getattr(norm2, k))
    obs = torch.randn(9, 23)
    a1 = ppo.make_inference_fn((norm, net))(deterministic=True)(obs)
    a2 = ppo.make_inference_fn((norm2, net2))(deterministic=True)(obs)
    assert a1.shape == (9, 5) and torch.equal(a1, a2) and a1.abs().max() <= 1
    bad = tmp_path / "bad.pt"
    torch.save({"format": "other"}, bad)
    with pytest.raises(ValueError, match="not a rsr_mjx_b200 PPO checkpoint"):
        ppo.load_params(bad, device="cpu")
