"""NumPy oracle of the RSR loss vs closed forms, and the host-side validation
rules of RSR/rsr_loss.py:55-70,158-162 (mirrored by rsr_mjx_b200.rsr_loss)."""
import numpy as np
import pytest

from oracle import rsr_loss_ref as R


def test_kde_single_point_is_softmax_of_distances():
    grid = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    data = np.array([[0.0, 0.0]])
    h = 0.5
    p = R.evaluate_kde(data, grid, h)
    logits = -np.array([0.0, 1.0, 4.0]) / (2 * h * h)
    e = np.exp(logits - logits.max())
    np.testing.assert_allclose(p, e / e.sum(), rtol=1e-12)
    assert p.sum() == pytest.approx(1.0)


def test_wasserstein_and_kl_closed_forms():
    p, q = np.array([0.5, 0.5, 0.0]), np.array([0.0, 0.5, 0.5])
    assert R.wasserstein_distance(p, q) == pytest.approx(0.5 + 0.5 + 0.0)
    assert R.wasserstein_distance(p, p) == 0
    assert R.kl_divergence(p, p) == pytest.approx(0.0, abs=1e-12)
    assert R.kl_divergence(np.array([1.0, 0.0]), np.array([0.5, 0.5])) == pytest.approx(np.log(2.0), rel=1e-6)


def test_loss_prepends_reference_data_and_scales():
    rng = np.random.default_rng(0)
    D, M = 7, 5
    grid = rng.uniform(-1, 1, (M, D))
    ref = rng.normal(0, 0.3, (6, D))
    obs, act, nxt = rng.normal(0, 0.3, (4, 3)), rng.normal(0, 0.3, (4, 1)), rng.normal(0, 0.3, (4, 3))
    refd = R.evaluate_kde(ref, grid, 0.4)
    loss, dist = R.compute_rsr_loss(obs, act, nxt, ref, refd, grid, 0.4, divergence=0.7, loss_scale=2.0)
    aug = np.concatenate([ref, np.concatenate([obs, act, nxt], 1)], 0)
    d2 = R.wasserstein_distance(R.evaluate_kde(aug, grid, 0.4), refd)
    assert dist == pytest.approx(d2) and loss == pytest.approx(2.0 * 0.7 * d2)
    # leading batch dims are flattened
    loss2, _ = R.compute_rsr_loss(obs.reshape(2, 2, 3), act.reshape(2, 2, 1), nxt.reshape(2, 2, 3), ref, refd, grid, 0.4, 0.7, 2.0)
    assert loss2 == pytest.approx(loss)


def test_float32_underflow_case_stays_finite():
    # bandwidth 0.1, D = 51: direct exponentiation underflows; the logsumexp/softmax form must not
    rng = np.random.default_rng(1)
    grid = rng.uniform(-3, 3, (10, 51)).astype(np.float32)
    data = rng.normal(0, 1, (50, 51)).astype(np.float32)
    p = R.evaluate_kde(data, grid, 0.1, dtype=np.float32)
    assert np.isfinite(p).all() and p.sum() == pytest.approx(1.0, rel=1e-5)
