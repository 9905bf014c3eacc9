"""CUDA RSR loss (forward + backward) against the NumPy oracle.
Tolerance: 1e-4 relative on density / distance / loss (float32 kernel vs float64
oracle), 1e-3 relative on the gradient (vs float64 finite differences)."""
import numpy as np
import pytest
import torch

from oracle import rsr_loss_ref as R
from rsr_mjx_b200 import _lib, rsr_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    _lib.build()


def _case(M, D, Nref, Nb, h, seed, spread=0.3):
    rng = np.random.default_rng(seed)
    grid = rng.uniform(-1, 1, (M, D)).astype(np.float32)
    ref = (rng.normal(0, spread, (Nref, D))).astype(np.float32)
    batch = (rng.normal(0, spread, (Nb, D))).astype(np.float32)
    return grid, ref, batch


@pytest.mark.parametrize("M,D,N,h", [(10, 51, 50, 0.1), (10, 108, 50, 0.1), (7, 5, 33, 0.5), (64, 51, 1330, 0.3), (1, 3, 1, 0.2)])
def test_kde_matches_oracle(M, D, N, h):
    grid, data, _ = _case(M, D, N, 1, h, seed=M + D)
    p = rsr_loss.evaluate_kde(torch.from_numpy(data).cuda(), torch.from_numpy(grid).cuda(), h).cpu().numpy()
    ref = R.evaluate_kde(data, grid, h)
    np.testing.assert_allclose(p, ref, rtol=1e-4, atol=1e-7)
    assert p.sum() == pytest.approx(1.0, rel=1e-5)


@pytest.mark.parametrize("M,D,Nref,Nb,h", [(10, 51, 50, 128, 0.1), (10, 51, 50, 1280, 0.1), (10, 108, 50, 512, 0.1), (6, 9, 12, 20, 0.4)])
def test_loss_forward_matches_oracle(M, D, Nref, Nb, h):
    grid, ref, batch = _case(M, D, Nref, Nb, h, seed=7)
    refd = R.evaluate_kde(ref, grid, h)
    obs_n = (D - 5) // 2
    obs, act, nxt = batch[:, :obs_n], batch[:, obs_n:obs_n + (D - 2 * obs_n)], batch[:, D - obs_n:]
    data = rsr_loss.RSRData(torch.tensor(0.37), torch.from_numpy(refd.astype(np.float32)).cuda(),
                            torch.from_numpy(ref).cuda(), torch.from_numpy(grid).cuda(), h)
    loss, dist = rsr_loss.compute_rsr_loss(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(),
                                           torch.from_numpy(nxt).cuda(), data, loss_scale=1.5)
    l_ref, d_ref = R.compute_rsr_loss(obs, act, nxt, ref, refd, grid, h, 0.37, 1.5)
    assert dist.item() == pytest.approx(d_ref, rel=1e-4, abs=1e-7)
    assert loss.item() == pytest.approx(l_ref, rel=1e-4, abs=1e-7)
    # legacy tuple formats and the zero short-circuit (rsr_loss.py:94-119,140-142)
    l5, _ = rsr_loss.compute_rsr_loss(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(),
                                      torch.from_numpy(nxt).cuda(), tuple(data), loss_scale=1.5)
    assert l5.item() == pytest.approx(loss.item())
    z, _ = rsr_loss.compute_rsr_loss(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(),
                                     torch.from_numpy(nxt).cuda(), None)
    assert z.item() == 0.0


def test_loss_gradient_matches_finite_differences():
    M, D, Nref, Nb, h = 6, 9, 12, 10, 0.4
    grid, ref, batch = _case(M, D, Nref, Nb, h, seed=3)
    refd = R.evaluate_kde(ref, grid, h)
    fd = R.loss_grad_fd(batch, ref.astype(np.float64), refd, grid.astype(np.float64), h, 0.8, 2.0)
    obs = torch.from_numpy(batch[:, :3]).cuda().requires_grad_(True)
    act = torch.from_numpy(batch[:, 3:6]).cuda().requires_grad_(True)
    nxt = torch.from_numpy(batch[:, 6:]).cuda().requires_grad_(True)
    data = rsr_loss.RSRData(torch.tensor(0.8), torch.from_numpy(refd.astype(np.float32)).cuda(),
                            torch.from_numpy(ref).cuda(), torch.from_numpy(grid).cuda(), h)
    loss, dist = rsr_loss.compute_rsr_loss(obs, act, nxt, data, loss_scale=2.0)
    loss.backward()
    got = torch.cat([obs.grad, act.grad, nxt.grad], 1).cpu().numpy()
    scale = np.abs(fd).max()
    assert scale > 1e-3
    np.testing.assert_allclose(got, fd, atol=1e-3 * scale)


def test_realistic_size_gradient_is_consistent():
    """D = 51, bandwidth 0.1 (reference defaults): directional derivative vs float64 loss difference"""
    M, D, Nref, Nb, h = 10, 51, 50, 256, 0.1
    rng = np.random.default_rng(5)
    grid = rng.uniform(-3, 3, (M, D)).astype(np.float32)
    centre = grid[rng.integers(0, M, Nref + Nb)] + rng.normal(0, 0.05, (Nref + Nb, D))
    ref, batch = centre[:Nref].astype(np.float32), centre[Nref:].astype(np.float32)
    refd = R.evaluate_kde(ref, grid, h)
    x = torch.from_numpy(batch).cuda().requires_grad_(True)
    data = rsr_loss.RSRData(torch.tensor(1.0), torch.from_numpy(refd.astype(np.float32)).cuda(),
                            torch.from_numpy(ref).cuda(), torch.from_numpy(grid).cuda(), h)
    loss, _ = rsr_loss.compute_rsr_loss(x[:, :23], x[:, 23:28], x[:, 28:], data, loss_scale=1.0)
    loss.backward()
    g = x.grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(g).all()
    direction = rng.normal(0, 1, batch.shape)
    eps = 1e-5

    def f(b):
        aug = np.concatenate([ref.astype(np.float64), b], 0)
        return R.wasserstein_distance(R.evaluate_kde(aug, grid.astype(np.float64), h), refd)

    fd = (f(batch + eps * direction) - f(batch - eps * direction)) / (2 * eps)
    assert np.sum(g * direction) == pytest.approx(fd, rel=2e-3, abs=1e-6)


def test_validation_errors_match_reference_messages():
    with pytest.raises(ValueError, match="real_data must be rank 2"):
        rsr_loss.build_rsr_data(torch.zeros(3), torch.zeros(3), torch.zeros(3))
    with pytest.raises(ValueError, match="previous_sim_data must match real_data"):
        rsr_loss.build_rsr_data(torch.zeros(4, 3), torch.zeros(5, 3), torch.zeros(4, 3))
    with pytest.raises(ValueError, match="num_samples must be positive"):
        rsr_loss.build_rsr_data(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 3), num_samples=0)
    with pytest.raises(ValueError, match="bandwidth must be positive"):
        rsr_loss.build_rsr_data(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 3), bandwidth=0.0)
    d = rsr_loss.build_rsr_data(torch.randn(20, 7), torch.randn(20, 7), torch.randn(20, 7), num_samples=5, min_value=-1, max_value=1, bandwidth=0.5)
    assert d.grid.shape == (5, 7) and d.reference_density.shape == (5,) and float(d.divergence) >= 0
    with pytest.raises(ValueError, match="online transition width does not match"):
        rsr_loss.compute_rsr_loss(torch.zeros(4, 3, device="cuda"), torch.zeros(4, 1, device="cuda"), torch.zeros(4, 2, device="cuda"), d)
    with pytest.raises(TypeError):
        rsr_loss.compute_rsr_loss(torch.zeros(4, 3, device="cuda"), torch.zeros(4, 1, device="cuda"), torch.zeros(4, 3, device="cuda"), 3.0)


def test_policy_term_equals_compute_rsr_loss_with_autograd():
    """rsr_loss.PolicyTerm (pack -> KDE/Wasserstein kernels -> chain rule through tanh, no torch in between: what the PPO
    update runs) == compute_rsr_loss(obs, tanh(loc), next_obs) + autograd w.r.t. the logits, bit for bit on the loss and
    to 1 ulp-level on the gradient (same kernels underneath; only the tanh derivative is formed differently)."""
    g = torch.Generator("cuda").manual_seed(7)
    O, A, rows = 23, 5, 640
    D = 2 * O + A
    real = torch.randn(50, D, device="cuda", generator=g) * 0.4
    past = rsr_loss.build_rsr_data(real, real + 0.05, real + 0.02, num_samples=10, min_value=-1.0, max_value=1.0, bandwidth=0.7)
    obs = torch.randn(rows, O, device="cuda", generator=g) * 0.4
    nxt = torch.randn(rows, O, device="cuda", generator=g) * 0.4
    logits = torch.randn(rows, 2 * A, device="cuda", generator=g)
    g_head = torch.randn(rows, 2 * A, device="cuda", generator=g) * 1e-3
    leaf = logits.clone().requires_grad_(True)
    loss, dist = rsr_loss.compute_rsr_loss(obs, torch.tanh(leaf[:, :A]), nxt, past, loss_scale=3.0)
    (g_ref,) = torch.autograd.grad(loss, leaf)
    term = rsr_loss.PolicyTerm(past, rows, O, A, 3.0, "cuda")
    term.forward(obs, logits, nxt)
    out = torch.empty_like(g_head)
    term.add_logit_grad(g_head, out)
    torch.cuda.synchronize()
    assert float(term.loss) == float(loss) and float(term.distance) == float(dist) and float(loss) != 0.0
    assert (g_ref[:, A:] == 0).all() and torch.equal(out[:, A:], g_head[:, A:])
    torch.testing.assert_close(out[:, :A] - g_head[:, :A], g_ref[:, :A], rtol=1e-4, atol=1e-9)
    assert g_ref.abs().max() > 0
    with pytest.raises(ValueError):
        rsr_loss.PolicyTerm(past, rows, O + 1, A, 3.0, "cuda")
