"""SAC trainer on the CUDA env: runs through `policy_params_training(algorithm='sac')`, the RSR term
reaches the actor gradient only, and the CUDA-graphed SGD step equals the eager one."""
import numpy as np
import pytest
import torch

from rsr_mjx_b200 import prng, rsr_loss, rsr_pipeline as RP, sac
from rsr_mjx_b200.envs import AirbotPlayBase

pytestmark = pytest.mark.gpu


def _run(graph, past, steps=6, **kw):
    env = AirbotPlayBase("sf", num_envs=64, episode_length=1200)
    seen = []
    mk, (norm, net), m = sac.train(env, num_timesteps=10**9, episode_length=1200, past_data=past, num_envs=64,
                                   learning_rate=1e-3, discounting=0.96, batch_size=64, num_evals=2,
                                   normalize_observations=True, reward_scaling=0.1, min_replay_size=256,
                                   max_replay_size=4096, grad_updates_per_step=2, hidden_layer_sizes=(64, 64),
                                   use_cuda_graph=graph, max_training_steps=steps,
                                   progress_fn=lambda n, mm: seen.append((n, dict(mm))), **kw)
    return mk, norm, net, m, seen


def _rsr_data():
    g = np.random.default_rng(0)
    real = g.normal(0, 0.5, (50, 51)).astype(np.float32)
    return rsr_loss.build_rsr_data(real, real + 0.05, real + 0.02, num_samples=10, min_value=-1, max_value=1, bandwidth=0.5)


def test_sac_trains_and_reports():
    mk, norm, net, m, seen = _run(True, None)
    assert len(seen) >= 1 and seen[-1][0] == 256 + 6 * 64  # prefill (4 actor steps) + 6 training steps
    for k in ("training/sps", "training/alpha_loss", "training/critic_loss", "training/actor_loss", "training/alpha",
              "training/replay_size"):
        assert k in m and np.isfinite(m[k]), k
    assert m["training/replay_size"] == 256 + 6 * 64 and m["training/sim2real_loss"] == 0.0
    assert float(norm.count) == (4 + 6) * 64
    a = mk(deterministic=True)(torch.zeros(5, 23, device="cuda"))
    assert a.shape == (5, 5) and (a.abs() <= 1).all()


def test_cuda_graph_step_equals_eager():
    _, _, net_g, mg, _ = _run(True, None)
    _, _, net_e, me, _ = _run(False, None)
    for (n1, p1), (n2, p2) in zip(net_g.named_parameters(), net_e.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=2e-3, atol=2e-5, msg=n1)
    assert mg["training/critic_loss"] == pytest.approx(me["training/critic_loss"], rel=1e-2)


def test_rsr_term_reaches_the_actor_only():
    torch.manual_seed(0)
    net = sac.SACNetworks(23, 5, (32, 32)).cuda()
    g = torch.Generator("cuda").manual_seed(1)
    B = 48
    tr = dict(observation=torch.randn(B, 23, device="cuda", generator=g) * 0.3,
              next_observation=torch.randn(B, 23, device="cuda", generator=g) * 0.3)
    noise = torch.randn(B, 5, device="cuda", generator=g)
    alpha = torch.tensor(0.5, device="cuda")

    def grad(past, scale):
        net.zero_grad()
        loss, s2r, _ = sac.actor_loss(net, net, lambda x: x, alpha, tr, noise, past, scale)
        loss.backward(inputs=list(net.policy.parameters()))
        return torch.cat([p.grad.reshape(-1) for p in net.policy.parameters()]).clone(), float(s2r)
    g0, s0 = grad(None, 1.0)
    g1, s1 = grad(_rsr_data(), 5.0)
    assert s0 == 0.0 and s1 != 0.0 and (g1 - g0).abs().max().item() > 0
    assert all(p.grad is None for p in list(net.q1.parameters()) + list(net.q2.parameters()))


def test_policy_params_training_sac_end_to_end():
    """RSR/rsr_pipeline.py:396-426 with the default ALGORITHM = 'sac' of test/rsr_policy_training.py:60"""
    env = AirbotPlayBase("sf", num_envs=64, episode_length=1200)
    st = env.reset(prng.split(prng.PRNGKey(3), 64))
    gen = torch.Generator("cuda").manual_seed(2)
    A = torch.rand(64, 5, device="cuda", generator=gen) * 2 - 1
    S = st.obs.clone()
    env.step(st, A)
    S2 = st.obs.clone()
    S, A, S2 = (v[:50].cpu().numpy().astype(np.float32) for v in (S, A, S2))
    seen = []
    make_policy, (norm, net) = RP.policy_params_training(
        env, past_states=S, past_actions=A, past_next_states_real=S2 + 0.01, past_next_states_sim=S2 + 0.03,
        current_next_states_sim=S2, algorithm="SAC ", num_envs=64, batch_size=64, num_timesteps=10**9, num_evals=2,
        min_replay_size=128, max_replay_size=2048, grad_updates_per_step=1, max_training_steps=4,
        progress_fn=lambda n, m: seen.append(m), bandwidth=0.5, min_val=-1.0, max_val=1.5)
    assert seen and np.isfinite(seen[-1]["training/actor_loss"]) and seen[-1]["training/sim2real_loss"] != 0
    assert make_policy()(env.reset(prng.split(prng.PRNGKey(0), 64)).obs).shape == (64, 5)
    with pytest.raises(ValueError, match="unsupported algorithm"):
        RP.policy_params_training(env, past_states=S, past_actions=A, past_next_states_real=S2, past_next_states_sim=S2,
                                  current_next_states_sim=S2, algorithm="td3")


def test_cuda_graphed_step_matches_numpy_restatement():
    """The SGD step the trainer actually runs (CUDA-graphed, on the batches it sampled from its replay ring) against a
    float64 NumPy restatement of RSR/sac_losses.py:38-128 evaluated on the same batch, noise and pre-update parameters:
    alpha / critic / actor losses within 1e-4 (fp32 matmuls: allow_tf32=False here), and the polyak step of the target
    critics (tau) exact."""
    import math
    snaps, results = [], []

    def mlp_np(mlp, x, act_relu=True):
        ls = list(mlp.layers)
        for i, l in enumerate(ls):
            x = x @ l["weight"].T + l["bias"]
            if i + 1 < len(ls) and act_relu:
                x = np.maximum(x, 0)
        return x

    class NP:  # parameter snapshot of an MLP as float64 numpy
        def __init__(self, mlp):
            self.layers = [dict(weight=l.weight.detach().double().cpu().numpy(), bias=l.bias.detach().double().cpu().numpy())
                           for l in mlp.layers]

    def probe(phase, c):
        if phase == "before":
            n = c["norm"]
            snaps.append(dict(static=c["static"].double().cpu().numpy(), noise=c["noise"].double().cpu().numpy(),
                              log_alpha=float(c["log_alpha"].detach()), fields=c["fields"],
                              mean=n.mean.double().cpu().numpy(), std=n.std.double().cpu().numpy(),
                              pol=NP(c["net"].policy), q1=NP(c["net"].q1), q2=NP(c["net"].q2), t1=NP(c["target"].q1),
                              t2=NP(c["target"].q2)))
        else:
            torch.cuda.synchronize()
            results.append(dict({k: float(v) for k, v in c["metrics"].items()}, t1=NP(c["target"].q1), q1=NP(c["net"].q1)))

    env = AirbotPlayBase("sf", num_envs=64, episode_length=1200)
    tau = 0.01
    sac.train(env, num_timesteps=10**9, episode_length=1200, past_data=None, num_envs=64, learning_rate=1e-3, discounting=0.96,
              batch_size=64, num_evals=2, normalize_observations=True, reward_scaling=0.1, min_replay_size=256,
              max_replay_size=4096, grad_updates_per_step=1, hidden_layer_sizes=(64, 64), use_cuda_graph=True,
              allow_tf32=False, tau=tau, max_training_steps=5, run_evals=False, sgd_probe_fn=probe)
    assert len(snaps) == len(results) == 5
    A = 5

    def lp_np(logits, raw):
        loc, s_ = logits[:, :A], logits[:, A:]
        scale = np.log1p(np.exp(s_)) + 0.001
        lp = -0.5 * ((raw - loc) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - np.log(scale)
        return (lp - 2.0 * (math.log(2.0) - raw - np.log1p(np.exp(-2.0 * raw)))).sum(-1), loc, scale

    for s_, r in zip(snaps, results):
        f = s_["fields"]
        tr = {k: s_["static"][:, sl] for k, sl in f.items()}
        nz = lambda x: (x - s_["mean"]) / s_["std"]
        alpha = math.exp(s_["log_alpha"])
        logits = mlp_np(s_["pol"], nz(tr["observation"]))
        loc, scale = logits[:, :A], np.log1p(np.exp(logits[:, A:])) + 0.001
        lp0, _, _ = lp_np(logits, loc + scale * s_["noise"][0])
        alpha_loss = np.mean(alpha * (-lp0 + 0.5 * A))
        nlogits = mlp_np(s_["pol"], nz(tr["next_observation"]))
        raw = nlogits[:, :A] + (np.log1p(np.exp(nlogits[:, A:])) + 0.001) * s_["noise"][1]
        nlp, _, _ = lp_np(nlogits, raw)
        xq = np.concatenate([nz(tr["next_observation"]), np.tanh(raw)], -1)
        next_v = np.minimum(mlp_np(s_["t1"], xq)[:, 0], mlp_np(s_["t2"], xq)[:, 0]) - alpha * nlp
        tq = tr["reward"][:, 0] * 0.1 + tr["discount"][:, 0] * 0.96 * next_v
        xo = np.concatenate([nz(tr["observation"]), tr["action"]], -1)
        err = (np.stack([mlp_np(s_["q1"], xo)[:, 0], mlp_np(s_["q2"], xo)[:, 0]], -1) - tq[:, None]) * (1 - tr["truncation"])
        critic_loss = 0.5 * np.mean(err ** 2)
        raw2 = loc + scale * s_["noise"][2]
        alp, _, _ = lp_np(logits, raw2)
        xa = np.concatenate([nz(tr["observation"]), np.tanh(raw2)], -1)
        actor_loss = np.mean(alpha * alp - np.minimum(mlp_np(s_["q1"], xa)[:, 0], mlp_np(s_["q2"], xa)[:, 0]))
        assert r["alpha_loss"] == pytest.approx(alpha_loss, rel=1e-4, abs=1e-5)
        assert r["critic_loss"] == pytest.approx(critic_loss, rel=1e-4, abs=1e-6)
        assert r["actor_loss"] == pytest.approx(actor_loss, rel=1e-4, abs=1e-5)
        assert r["alpha"] == pytest.approx(alpha, rel=1e-6)
        # polyak: target <- (1 - tau) target + tau * NEW q
        for lt0, lt1, lq1 in zip(s_["t1"].layers, r["t1"].layers, r["q1"].layers):
            np.testing.assert_allclose(lt1["weight"], (1 - tau) * lt0["weight"] + tau * lq1["weight"], rtol=1e-5, atol=1e-7)
