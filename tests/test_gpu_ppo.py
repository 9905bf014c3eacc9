"""PPO trainer on the CUDA env (caller of the hot path): runs, learns finite numbers, the RSR
term reaches the policy gradient, and the CUDA-graphed minibatch step equals the eager one."""
import numpy as np
import pytest
import torch

from rsr_mjx_b200 import _lib, domain_randomize as DR, ppo, prng, rsr_loss
from rsr_mjx_b200.envs import AirbotPlayBase

pytestmark = pytest.mark.gpu


def _rsr_data():
    g = np.random.default_rng(0)
    real = g.normal(0, 0.5, (50, 51)).astype(np.float32)
    return rsr_loss.build_rsr_data(real, real + 0.05, real + 0.02, num_samples=10, min_value=-1, max_value=1, bandwidth=0.5)


def _run(graph, past, steps=2, dr=False):
    kw = {}
    if dr:
        kw = dict(randomization_fn=DR.domain_randomize, randomization_rng=prng.split(prng.PRNGKey(3), 128))
    env = AirbotPlayBase("cube" if dr else "sf", num_envs=128, episode_length=1200, **kw)
    seen = []
    mk, (norm, net), metrics = ppo.train(env, num_timesteps=10**9, episode_length=1200, past_data=past, num_envs=128,
                                         learning_rate=1e-3, entropy_cost=2e-2, discounting=0.96, unroll_length=5,
                                         batch_size=16, num_minibatches=8, num_updates_per_batch=2, num_evals=steps,
                                         normalize_observations=True, reward_scaling=0.1, rsr_loss_scale=1.0,
                                         use_cuda_graph=graph, max_training_steps=steps,
                                         progress_fn=lambda n, m: seen.append((n, dict(m))))
    return mk, norm, net, metrics, seen, env


def test_ppo_trains_and_reports_sps():
    mk, norm, net, metrics, seen, env = _run(True, _rsr_data(), steps=2, dr=True)
    assert len(seen) == 2 and seen[-1][0] == 2 * 16 * 5 * 8
    for k in ("training/sps", "training/total_loss", "training/policy_loss", "training/v_loss", "training/entropy_loss",
              "training/sim2real_loss", "training/rsr_distribution_distance"):
        assert k in metrics and np.isfinite(metrics[k]), k
    assert metrics["training/sps"] > 0 and metrics["training/sim2real_loss"] != 0.0
    assert float(norm.count) == 2 * 128 * 5
    pol = mk(deterministic=True)
    a = pol(torch.zeros(4, 23, device="cuda"))
    assert a.shape == (4, 5) and (a.abs() <= 1).all()


def test_cuda_graph_step_equals_eager():
    past = _rsr_data()
    _, _, net_g, mg, _, _ = _run(True, past, steps=1)
    _, _, net_e, me, _, _ = _run(False, past, steps=1)
    for (n1, p1), (n2, p2) in zip(net_g.named_parameters(), net_e.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=2e-4, atol=2e-6, msg=n1)
    assert mg["training/total_loss"] == pytest.approx(me["training/total_loss"], rel=1e-3)


def test_rsr_term_changes_the_policy_gradient():
    torch.manual_seed(0)
    net = ppo.PPONetworks(23, 5).cuda()
    B, T = 32, 5
    g = torch.Generator("cuda").manual_seed(1)
    data = dict(observation=torch.randn(B, T, 23, device="cuda", generator=g) * 0.3,
                next_observation=torch.randn(B, T, 23, device="cuda", generator=g) * 0.3,
                raw_action=torch.randn(B, T, 5, device="cuda", generator=g), log_prob=torch.zeros(B, T, device="cuda"),
                reward=torch.zeros(B, T, device="cuda"), discount=torch.ones(B, T, device="cuda"),
                truncation=torch.zeros(B, T, device="cuda"))
    noise = torch.zeros(T, B, 5, device="cuda")

    def grad(past, scale):
        net.zero_grad()
        loss, m = ppo.compute_ppo_loss(net, lambda x: x, data, noise, past_data=past, rsr_loss_scale=scale)
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in net.policy.parameters()]).clone(), m
    g0, m0 = grad(None, 1.0)
    g1, m1 = grad(_rsr_data(), 5.0)
    assert m0["sim2real_loss"].item() == 0.0 and m1["sim2real_loss"].item() != 0.0
    assert (g1 - g0).abs().max().item() > 0


def test_policy_params_training_end_to_end():
    """RSR/rsr_pipeline.py:274-436 with the reference's dataset layout (23-d states, 5-d actions)"""
    from rsr_mjx_b200 import rsr_pipeline as RP
    # "real" transitions: 50 (s, a, s') triples rolled out of the env itself, so the policy's own transitions land
    # inside the support of the reference KDE (rows far from every grid point drop out of the logsumexp exactly)
    env = AirbotPlayBase("sf", num_envs=64, episode_length=1200)
    st = env.reset(prng.split(prng.PRNGKey(3), 64))
    gen = torch.Generator("cuda").manual_seed(2)
    Aa = (torch.rand(64, 5, device="cuda", generator=gen) * 2 - 1)
    S = st.obs.clone()
    env.step(st, Aa)
    S2 = st.obs.clone()
    S, Aa, S2 = (v[:50].cpu().numpy().astype(np.float32) for v in (S, Aa, S2))
    seen = []
    make_policy, (norm, net) = RP.policy_params_training(
        env, past_states=S, past_actions=Aa, past_next_states_real=S2 + 0.01, past_next_states_sim=S2 + 0.03,
        current_next_states_sim=S2, num_envs=64, batch_size=8, num_minibatches=8, unroll_length=4,
        num_updates_per_batch=1, num_timesteps=10**9, num_evals=1, progress_fn=lambda n, m: seen.append(m),
        max_training_steps=1, bandwidth=0.5, min_val=-1.0, max_val=1.5)
    assert len(seen) == 1 and np.isfinite(seen[0]["training/total_loss"]) and seen[0]["training/sim2real_loss"] != 0
    a = make_policy()(env.reset(prng.split(prng.PRNGKey(0), 64)).obs)
    assert a.shape == (64, 5)


@pytest.mark.parametrize("normalize_advantage,B", [(True, 64), (False, 64), (True, 13), (True, 3)])
def test_fused_head_matches_torch_reference(normalize_advantage, B):
    """csrc/rsrx_ppo.cuh (GAE + tanh-normal log-prob + clipped surrogate + value + entropy, fwd and bwd in one launch)
    against the plain torch fp32 restatement of RSR/losses.py (`compute_ppo_loss`): losses 1e-5 rel, grads 1e-4 rel.
    The head runs on an 8-CTA cluster, sequences split across the CTAs: B = 13 leaves the last CTA empty, B = 3 five."""
    torch.manual_seed(1)
    T = 10
    net = ppo.PPONetworks(23, 5).cuda()
    g = torch.Generator("cuda").manual_seed(4)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    data = dict(observation=r(B, T, 23) * 0.5, next_observation=r(B, T, 23) * 0.5, raw_action=r(B, T, 5),
                log_prob=r(B, T) * 0.3 - 3.0, reward=r(B, T),
                discount=(torch.rand(B, T, device="cuda", generator=g) > 0.1).float(),
                truncation=(torch.rand(B, T, device="cuda", generator=g) > 0.9).float())
    data["log_prob"] = ppo.NormalTanh.log_prob(net.policy(data["observation"]), data["raw_action"]).detach() + r(B, T) * 0.4
    noise = r(B, T, 5)
    kw = dict(entropy_cost=2e-2, discounting=0.96, reward_scaling=0.1, gae_lambda=0.95, clipping_epsilon=0.3,
              normalize_advantage=normalize_advantage, rsr_loss_scale=0.0)

    def grads(fn, nz):
        net.zero_grad()
        loss, m = fn(net, lambda x: x, data, nz, **kw)
        loss.backward()
        return loss.item(), {k: float(v) for k, v in m.items()}, torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    l_ref, m_ref, g_ref = grads(ppo.compute_ppo_loss, noise.transpose(0, 1).contiguous())
    l_fus, m_fus, g_fus = grads(ppo.compute_ppo_loss_fused, noise)
    assert l_fus == pytest.approx(l_ref, rel=1e-5, abs=1e-6)
    for k in ("policy_loss", "v_loss", "entropy_loss"):
        assert m_fus[k] == pytest.approx(m_ref[k], rel=1e-5, abs=1e-6), k
    # some transitions sit outside the clip range (the behaviour log-prob was perturbed): both branches exercised
    assert (g_fus - g_ref).abs().max().item() <= 1e-4 * g_ref.abs().max().item() + 1e-7


def test_fused_head_training_equals_torch_head_training():
    def run(fused):
        env = AirbotPlayBase("sf", num_envs=128, episode_length=1200)
        _, (norm, net), m = ppo.train(env, num_timesteps=10**9, episode_length=1200, num_envs=128, learning_rate=1e-3,
                                      entropy_cost=2e-2, discounting=0.96, unroll_length=5, batch_size=16, num_minibatches=8,
                                      num_updates_per_batch=1, num_evals=1, normalize_observations=True, reward_scaling=0.1,
                                      use_cuda_graph=True, fused_head=fused, max_training_steps=1)
        return net, m
    net_f, m_f = run(True)
    net_t, m_t = run(False)
    # same rollouts; the entropy noise is laid out differently ([B,T,A] vs [T,B,A]), so only statistical agreement
    assert m_f["training/v_loss"] == pytest.approx(m_t["training/v_loss"], rel=0.2)
    assert np.isfinite(m_f["training/total_loss"]) and np.isfinite(m_t["training/total_loss"])


@pytest.mark.parametrize("act", ["silu", "relu", "none"])
@pytest.mark.parametrize("rows,cin,cout", [(2816, 23, 256), (2560, 32, 10), (641, 7, 33)])
def test_linear_act_fused_backward_matches_torch(act, rows, cin, cout):
    """csrc/rsrx_ppo.cuh::act_bias_backward_kernel (activation derivative + deterministic bias gradient) behind
    `ppo.linear_act` against plain torch autograd in fp32 (TF32 off): 1e-5 relative"""
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator("cuda").manual_seed(rows + cout)
        layer = torch.nn.Linear(cin, cout).cuda()
        x = torch.randn(rows, cin, device="cuda", generator=g, requires_grad=True)
        w = torch.randn(rows, cout, device="cuda", generator=g)
        ws = {}

        def run(fused):
            layer.zero_grad(); x.grad = None
            if fused:
                y = ppo.linear_act(x, layer, act, ws)
            else:
                z = layer(x)
                y = torch.nn.functional.silu(z) if act == "silu" else (torch.relu(z) if act == "relu" else z)
            (y * w).sum().backward()
            return y.detach().clone(), x.grad.clone(), layer.weight.grad.clone(), layer.bias.grad.clone()
        ref = run(False)
        a = run(True)
        b = run(True)  # second call reuses the workspace: the ticket counter must have been reset
        for name, r_, a_, b_ in zip(("y", "dx", "dW", "db"), ref, a, b):
            tol = 1e-5 * float(r_.abs().max()) + 1e-6
            assert float((a_ - r_).abs().max()) <= tol, name
            assert torch.equal(a_, b_), name  # deterministic
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def test_evaluator_matches_eval_wrapper_semantics():
    """acting.Evaluator / EvalWrapper (RSR/train.py:441-447): reward and metrics summed while the episode is active,
    one episode per eval env, `eval/*` keys the reference's progress_fn reads (test/rsr_policy_training.py:247-248)"""
    env = AirbotPlayBase("sf", num_envs=32, episode_length=50)
    zero_policy = lambda: (lambda obs, generator=None: torch.zeros(obs.shape[0], 5, device="cuda"))
    ev = ppo.Evaluator(env, zero_policy, 32, 50, 1, seed=0)
    m = ev.run_evaluation({"training/x": 1.0})
    for k in ("eval/episode_reward", "eval/episode_reward_std", "eval/avg_episode_length", "eval/epoch_eval_time", "eval/sps",
              "eval/walltime", "training/x"):
        assert k in m, k
    assert m["eval/avg_episode_length"] == 50.0  # nothing terminates early under the zero action: truncation at 50
    # the same episode by hand
    from rsr_mjx_b200 import prng
    key = prng.split(prng.PRNGKey(7919), 2)[-1]
    st = env.reset(prng.split(key, 32))
    total = torch.zeros(32, device="cuda")
    for _ in range(50):
        env.step(st, torch.zeros(32, 5, device="cuda"))
        total += st.reward
    assert m["eval/episode_reward"] == pytest.approx(float(total.mean()), rel=1e-6)
    per_env = ev.run_evaluation({}, aggregate_episodes=False)
    assert per_env["eval/episode_reward"].shape == (32,)
    with pytest.raises(ValueError):
        ppo.Evaluator(env, zero_policy, 16, 50)


def test_train_epochs_and_callbacks_follow_the_reference():
    """num_evals=3 -> eval before training + after each of the 2 epochs; policy_params_fn after every epoch"""
    env = AirbotPlayBase("sf", num_envs=64, episode_length=40)
    seen, saved = [], []
    per_step = 8 * 5 * 8  # batch_size * unroll_length * num_minibatches
    ppo.train(env, num_timesteps=4 * per_step, episode_length=40, num_envs=64, unroll_length=5, batch_size=8, num_minibatches=8,
              num_updates_per_batch=1, num_evals=3, num_eval_envs=16, progress_fn=lambda n, m: seen.append((n, m)),
              policy_params_fn=lambda n, mk, params: saved.append(n))
    assert [n for n, _ in seen] == [0, 2 * per_step, 4 * per_step] and saved == [2 * per_step, 4 * per_step]
    assert "training/sps" not in seen[0][1] and "eval/episode_reward" in seen[0][1]
    assert "training/sps" in seen[-1][1] and "eval/episode_reward_std" in seen[-1][1]


def test_gather_rows_matches_torch_indexing():
    import ctypes as C
    g = torch.Generator("cuda").manual_seed(0)
    srcs = [torch.randn(100, 10, 23, device="cuda", generator=g), torch.randn(100, 10, device="cuda", generator=g),
            torch.randn(100, 10, 5, device="cuda", generator=g)]
    idx = torch.randperm(100, device="cuda", generator=g)[10:47]
    dsts = [torch.zeros(37, *s.shape[1:], device="cuda") for s in srcs]
    n = len(srcs)
    _lib.check(_lib.lib().rsrx_gather_rows((C.c_void_p * n)(*[s.data_ptr() for s in srcs]),
                                           (C.c_void_p * n)(*[d.data_ptr() for d in dsts]),
                                           (C.c_int32 * n)(*[s[0].numel() for s in srcs]), n, idx.data_ptr(), 37,
                                           torch.cuda.current_stream().cuda_stream), "rsrx_gather_rows")
    torch.cuda.synchronize()
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s[idx])
    with pytest.raises(RuntimeError):
        _lib.check(_lib.lib().rsrx_gather_rows(None, None, None, 1, idx.data_ptr(), 1, None), "rsrx_gather_rows")


def test_minibatch_prep_matches_torch():
    """rsrx_ppo_prep: normalised policy input, padded value input (observations then bootstrap observations) and its
    transpose in one launch == the torch ops it replaces, bit for bit ((x - mean) * (1 / std) vs (x - mean) / std: 1 ulp)"""
    g = torch.Generator("cuda").manual_seed(3)
    mb, T, O, ldp = 37, 10, 23, 32
    obs = torch.randn(mb, T, O, device="cuda", generator=g)
    nxt = torch.randn(mb, T, O, device="cuda", generator=g)
    mean = torch.randn(O, device="cuda", generator=g) * 0.2
    std = torch.rand(O, device="cuda", generator=g) + 0.5
    rows = mb * T + mb
    ldt = (rows + 3) // 4 * 4
    obs_n = torch.full((mb * T, O), 7.0, device="cuda")
    x_pad = torch.full((rows, ldp), 7.0, device="cuda")
    xT = torch.zeros(ldp, ldt, device="cuda")
    _lib.check(_lib.lib().rsrx_ppo_prep(obs.data_ptr(), nxt.data_ptr(), mean.data_ptr(), std.data_ptr(), mb, T, O, obs_n.data_ptr(),
                                        x_pad.data_ptr(), ldp, xT.data_ptr(), ldt, torch.cuda.current_stream().cuda_stream), "prep")
    torch.cuda.synchronize()
    ref = torch.cat([((obs - mean) / std).reshape(mb * T, O), (nxt[:, -1] - mean) / std], 0)
    torch.testing.assert_close(obs_n, ref[:mb * T], rtol=2e-7, atol=1e-7)
    torch.testing.assert_close(x_pad[:, :O], ref, rtol=2e-7, atol=1e-7)
    assert (x_pad[:, O:] == 0).all()
    assert torch.equal(xT[:, :rows], x_pad.t())
    with pytest.raises(RuntimeError):
        _lib.check(_lib.lib().rsrx_ppo_prep(obs.data_ptr(), nxt.data_ptr(), mean.data_ptr(), std.data_ptr(), mb, T, O, obs_n.data_ptr(),
                                            x_pad.data_ptr(), 16, xT.data_ptr(), ldt, None), "prep")


def test_tanh_normal_act_matches_torch():
    g = torch.Generator("cuda").manual_seed(5)
    logits = torch.randn(1000, 10, device="cuda", generator=g) * 2
    noise = torch.randn(1000, 5, device="cuda", generator=g)
    raw, action, lp = ppo.NormalTanh.act(logits, noise)
    loc, scale = ppo.NormalTanh.params(logits)
    raw_ref = loc + scale * noise
    torch.testing.assert_close(raw, raw_ref, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(action, torch.tanh(raw_ref), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(lp, ppo.NormalTanh.log_prob(logits, raw_ref), rtol=1e-5, atol=2e-5)
    with pytest.raises(ValueError):
        ppo.NormalTanh.act(logits.cpu(), noise.cpu())


# ---- round 2: the value network on the hand-written tcgen05 kernels ---------------------------------------------------
def test_tensor_core_value_net_matches_torch_autograd():
    """fused_mlp.TensorCoreMLP (tcgen05 TF32 GEMMs with fused bias / swish / swish' / bias-gradient epilogues) against
    torch autograd on the same MLP in full fp32: values and every parameter gradient within TF32 accuracy."""
    from rsr_mjx_b200 import fused_mlp
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    mlp = ppo.MLP([23, 256, 256, 256, 256, 256, 1]).cuda()
    for l in mlp.layers:
        torch.nn.init.normal_(l.bias, 0, 0.1)
    rows = 2816
    x = torch.randn(rows, 23, device="cuda")
    g = torch.randn(rows, device="cuda") / rows
    tc = fused_mlp.TensorCoreMLP(mlp, rows, "cuda")
    v = tc.forward(x)
    tc.attach_grads()
    tc.backward(g)
    torch.cuda.synchronize()
    got = {n: p.grad.clone() for n, p in mlp.named_parameters()}
    for p in mlp.parameters():
        p.grad = None
    ref = ppo.MLP.forward(mlp, x).squeeze(-1) if False else None
    h = x
    for i, l in enumerate(mlp.layers):
        h = torch.nn.functional.linear(h, l.weight, l.bias)
        if i + 1 < len(mlp.layers):
            h = torch.nn.functional.silu(h)
    ref = h.squeeze(-1)
    ref.backward(g)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    # TF32 inputs (10-bit mantissa) through six layers: ~1e-2 of the output / gradient scale against full fp32
    assert rel(v, ref.detach()) <= 1.5e-2
    for n, p in mlp.named_parameters():
        assert rel(got[n], p.grad) <= 1.5e-2, (n, rel(got[n], p.grad))
    # ragged rows (not a multiple of 128 or 256) and relu (the SAC critics)
    class R(ppo.MLP):
        activation = "relu"
    mlp2 = R([28, 64, 64, 1]).cuda()
    x2, g2 = torch.randn(300, 28, device="cuda"), torch.randn(300, device="cuda")
    tc2 = fused_mlp.TensorCoreMLP(mlp2, 300, "cuda")
    v2 = tc2.forward(x2)
    tc2.attach_grads()
    tc2.backward(g2)
    got2 = {n: p.grad.clone() for n, p in mlp2.named_parameters()}
    for p in mlp2.parameters():
        p.grad = None
    h = x2
    for i, l in enumerate(mlp2.layers):
        h = torch.nn.functional.linear(h, l.weight, l.bias)
        if i + 1 < len(mlp2.layers):
            h = torch.relu(h)
    h.squeeze(-1).backward(g2)
    # relu: a pre-activation within TF32 noise of 0 flips its derivative, so single entries move more; the gradient as a
    # whole (relative L2) stays at TF32 accuracy
    rel2 = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    assert rel(v2, h.squeeze(-1).detach()) <= 1e-2
    for n, p in mlp2.named_parameters():
        # (a fraction p of flipped derivatives moves the relative L2 error by ~sqrt(p): 6e-4 of the entries -> 2.5 %)
        assert rel2(got2[n], p.grad) <= 5e-2, (n, rel2(got2[n], p.grad))


def test_tensor_core_value_path_trains_like_the_autograd_path():
    """one training step with the value network on the tcgen05 kernels vs on torch autograd (both TF32): same parameters
    up to TF32 rounding of the two implementations, same losses"""
    def run(tc):
        env = AirbotPlayBase("sf", num_envs=128, episode_length=1200)
        _, (norm, net), m = ppo.train(env, num_timesteps=10**9, episode_length=1200, past_data=_rsr_data(), num_envs=128,
                                      learning_rate=1e-3, entropy_cost=2e-2, discounting=0.96, unroll_length=5, batch_size=16,
                                      num_minibatches=8, num_updates_per_batch=2, num_evals=1, normalize_observations=True,
                                      reward_scaling=0.1, max_training_steps=1, run_evals=False, tensor_core_value=tc)
        return net, m
    net_a, ma = run(True)
    net_b, mb_ = run(False)
    for k in ("training/policy_loss", "training/v_loss", "training/entropy_loss", "training/sim2real_loss"):
        # 16 Adam steps at lr 1e-3 apart on two TF32 implementations (truncated vs rounded inputs): a few per cent
        assert ma[k] == pytest.approx(mb_[k], rel=0.1, abs=1e-5), k
    for (n1, p1), (n2, p2) in zip(net_a.named_parameters(), net_b.named_parameters()):
        # Adam normalises the step: every weight moves by ~lr = 1e-3 per step whatever its gradient's size, so a weight
        # whose tiny gradient changes sign between the two TF32 implementations ends up to 16 * 2 * lr apart; on average the
        # two runs must stay together
        d = (p1 - p2).abs()
        assert float(d.max()) <= 3.5e-2 and float(d.mean()) <= 2e-3, (n1, float(d.max()), float(d.mean()))


def test_warp_policy_net_matches_torch_autograd():
    """fused_mlp.WarpMLP (one warp per row, lane = neuron; csrc/rsrx_mlp.cuh) against torch autograd in fp32: logits and
    every parameter gradient to fp32 accuracy (no tensor cores involved), for the reference policy 23 -> 32 x 4 -> 10 and a
    ragged relu net"""
    from rsr_mjx_b200 import fused_mlp
    torch.manual_seed(1)
    torch.backends.cuda.matmul.allow_tf32 = False
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))

    class R(ppo.MLP):
        activation = "relu"
    for cls, sizes, rows in ((ppo.MLP, [23, 32, 32, 32, 32, 10], 2560), (R, [28, 17, 32, 6], 333), (ppo.MLP, [23, 10], 40)):
        mlp = cls(sizes).cuda()
        for l in mlp.layers:
            torch.nn.init.normal_(l.bias, 0, 0.1)
        x = torch.randn(rows, sizes[0], device="cuda")
        g = torch.randn(rows, sizes[-1], device="cuda")
        wm = fused_mlp.WarpMLP(mlp, rows, "cuda")
        out = wm.forward(x).clone()
        wm.attach_grads()
        wm.backward(g)
        torch.cuda.synchronize()
        got = {n: p.grad.clone() for n, p in mlp.named_parameters()}
        for p in mlp.parameters():
            p.grad = None
        h = x
        for i, l in enumerate(mlp.layers):
            h = torch.nn.functional.linear(h, l.weight, l.bias)
            if i + 1 < len(mlp.layers):
                h = torch.nn.functional.silu(h) if cls is ppo.MLP else torch.relu(h)
        h.backward(g)
        assert rel(out, h.detach()) <= 1e-5, sizes
        for n, p in mlp.named_parameters():
            assert rel(got[n], p.grad) <= 1e-4, (sizes, n, rel(got[n], p.grad))


def test_warp_policy_forward_with_in_kernel_normalisation():
    """the actor step's form: rows of a padded observation buffer, (x - mean) / std applied inside the launch ==
    forward of the separately normalised, contiguous input, bit for bit; backward after it is refused"""
    from rsr_mjx_b200 import fused_mlp
    torch.manual_seed(4)
    mlp = ppo.MLP([23, 32, 32, 32, 32, 10]).cuda()
    rows = 1024
    padded = torch.randn(rows, 24, device="cuda")
    x = padded[:, :23]
    mean, std = torch.randn(23, device="cuda") * 0.3, torch.rand(23, device="cuda") + 0.5
    wm = fused_mlp.WarpMLP(mlp, rows, "cuda")
    a = wm.forward(x, mean, std).clone()
    with pytest.raises(RuntimeError):
        wm.backward(torch.zeros(rows, 10, device="cuda"))
    b = wm.forward(((x - mean) / std).contiguous()).clone()
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        wm.forward(x, mean, None)


def test_fused_adam_matches_torch_adam():
    from rsr_mjx_b200 import fused_mlp
    torch.manual_seed(2)
    shapes = [(32, 23), (32,), (256, 256), (256,), (1, 256), (1,)]
    pa = [torch.randn(*s, device="cuda") for s in shapes]
    pb = [p.clone().requires_grad_(True) for p in pa]
    fa = fused_mlp.FusedAdam(pa, lr=1e-3, eps=1e-8)
    tb = torch.optim.Adam(pb, lr=1e-3, eps=1e-8)
    g = torch.cuda.CUDAGraph()
    grads = [torch.zeros_like(p) for p in pa]
    for p, gr in zip(pa, grads):
        p.grad = gr
    fa.step()                      # warm-up outside the graph (restored below)
    for p, q in zip(pa, pb):
        p.copy_(q.detach())
    fa.reset_state()
    with torch.cuda.graph(g):
        fa.step()
    for it in range(7):
        for gr, q in zip(grads, pb):
            gr.normal_()
            q.grad = gr.clone()
        g.replay()                 # the captured step keeps counting: bias correction of step it + 1
        tb.step()
    torch.cuda.synchronize()
    assert fa.steps_taken == 7
    for p, q in zip(pa, pb):
        torch.testing.assert_close(p, q.detach(), rtol=2e-5, atol=2e-6)
