"""SAC pieces that run without a GPU: replay ring, the three losses of RSR/sac_losses.py against a
numpy restatement, update order bookkeeping."""
import math

import numpy as np
import pytest
import torch

from rsr_mjx_b200 import sac
from rsr_mjx_b200.ppo import NormalTanh


def test_replay_ring_wraps_and_samples_only_valid_rows():
    buf = sac.ReplayBuffer(10, 3, "cpu")
    with pytest.raises(RuntimeError):
        buf.sample(4)
    rows = torch.arange(36, dtype=torch.float32).reshape(12, 3)
    buf.insert(rows[:4])
    assert buf.size == 4 and buf.pos == 4
    s = buf.sample(64, torch.Generator().manual_seed(0))
    assert set(s[:, 0].tolist()) <= {0.0, 3.0, 6.0, 9.0}
    buf.insert(rows[4:12])  # wraps: rows 10, 11 overwrite slots 0, 1
    assert buf.size == 10 and buf.pos == 2
    assert buf.data[0, 0] == 30 and buf.data[1, 0] == 33 and buf.data[2, 0] == 6 and buf.data[9, 0] == 27
    big = torch.ones(25, 3)
    buf.insert(big)     # more rows than capacity: the newest `capacity` survive
    assert buf.size == 10 and (buf.data == 1).all()
    with pytest.raises(ValueError):
        sac.ReplayBuffer(0, 3, "cpu")


def _np_log_prob(logits, raw):
    A = logits.shape[-1] // 2
    loc, s = logits[..., :A], logits[..., A:]
    scale = np.log1p(np.exp(s)) + 0.001
    lp = -0.5 * ((raw - loc) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - np.log(scale)
    ldj = 2.0 * (math.log(2.0) - raw - np.log1p(np.exp(-2.0 * raw)))
    return (lp - ldj).sum(-1), loc, scale


def test_losses_match_numpy_restatement():
    torch.manual_seed(0)
    O, A, B = 7, 3, 32
    net, target = sac.SACNetworks(O, A, (16, 16)).double(), sac.SACNetworks(O, A, (16, 16)).double()
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    tr = dict(observation=r(B, O), action=torch.tanh(r(B, A)), reward=r(B, 1), discount=(torch.rand(B, 1, generator=g) > 0.2).double(),
              truncation=(torch.rand(B, 1, generator=g) > 0.8).double(), next_observation=r(B, O))
    noise = r(3, B, A)
    ident = lambda x: x
    log_alpha = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    alpha = float(torch.exp(log_alpha.detach()))
    n = lambda t: t.detach().numpy()

    def q_np(netx, obs, act):
        x = torch.cat([obs, act], -1)
        return np.concatenate([n(netx.q1(x)), n(netx.q2(x))], -1)

    # alpha loss: mean(alpha * stop_grad(-log_prob - target_entropy)), target_entropy = -A / 2
    la = sac.alpha_loss(log_alpha, net, ident, tr, noise[0], -0.5 * A)
    logits = n(net.policy(tr["observation"]))
    lp, loc, scale = _np_log_prob(logits, logits[:, :A] + (np.log1p(np.exp(logits[:, A:])) + 0.001) * n(noise[0]))
    assert float(la.detach()) == pytest.approx(np.mean(alpha * (-lp + 0.5 * A)), rel=1e-10)
    la.backward()
    assert float(log_alpha.grad) == pytest.approx(float(la.detach()), rel=1e-10)  # d/d log_alpha of exp(log_alpha) * c

    # critic loss
    lq = sac.critic_loss(net, target, ident, torch.tensor(alpha, dtype=torch.float64), tr, noise[1], 0.1, 0.96)
    nlogits = n(net.policy(tr["next_observation"]))
    raw = nlogits[:, :A] + (np.log1p(np.exp(nlogits[:, A:])) + 0.001) * n(noise[1])
    nlp, _, _ = _np_log_prob(nlogits, raw)
    next_q = q_np(target, tr["next_observation"], torch.tensor(np.tanh(raw)))
    next_v = next_q.min(-1) - alpha * nlp
    tq = n(tr["reward"])[:, 0] * 0.1 + n(tr["discount"])[:, 0] * 0.96 * next_v
    err = (q_np(net, tr["observation"], tr["action"]) - tq[:, None]) * (1 - n(tr["truncation"]))
    assert float(lq) == pytest.approx(0.5 * np.mean(err ** 2), rel=1e-10)

    # actor loss (no RSR data): mean(alpha * log_prob - min_q(obs, tanh(raw)))
    lp_t, s2r, dist = sac.actor_loss(net, net, ident, torch.tensor(alpha, dtype=torch.float64), tr, noise[2], None, 1.0)
    raw = logits[:, :A] + (np.log1p(np.exp(logits[:, A:])) + 0.001) * n(noise[2])
    alp, _, _ = _np_log_prob(logits, raw)
    qa = q_np(net, tr["observation"], torch.tensor(np.tanh(raw)))
    assert float(lp_t) == pytest.approx(np.mean(alpha * alp - qa.min(-1)), rel=1e-10)
    assert float(s2r) == 0.0 and float(dist) == 0.0
    # the actor gradient reaches the policy only when restricted to it (q parameters are left alone)
    net.zero_grad()
    lp_t.backward(inputs=list(net.policy.parameters()))
    assert all(p.grad is None or float(p.grad.abs().sum()) == 0 for p in list(net.q1.parameters()) + list(net.q2.parameters()))
    assert sum(float(p.grad.abs().sum()) for p in net.policy.parameters()) > 0


def test_train_argument_errors():
    class Env:
        num_envs, episode_length, device = 4, 100, "cpu"
    with pytest.raises(ValueError, match="rsr_loss_scale must be non-negative"):
        sac.train(Env(), 1000, 100, num_envs=4, rsr_loss_scale=-1.0)
    with pytest.raises(ValueError, match="environment has 4 envs"):
        sac.train(Env(), 1000, 100, num_envs=8)
    with pytest.raises(ValueError, match="min_replay_size >= num_timesteps"):
        sac.train(Env(), 1000, 100, num_envs=4, min_replay_size=1000)
