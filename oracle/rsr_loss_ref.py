"""NumPy restatement of the reference's RSR distribution loss (TEST INFRASTRUCTURE).

Follows RSR/dataset_processor.py:17-43 (evaluate_kde, kl_divergence,
wasserstein_distance) and RSR/rsr_loss.py:43-91,122-175 line by line; the
reference is pure jnp, so this restatement is exact up to the float type
(`dtype=np.float32` reproduces jnp's default precision, float64 is the check)."""
from __future__ import annotations

import numpy as np


def _logsumexp(a, axis):
    mx = np.max(a, axis=axis, keepdims=True)
    return (mx + np.log(np.sum(np.exp(a - mx), axis=axis, keepdims=True))).squeeze(axis)


def evaluate_kde(data, grid, bandwidth=0.1, dtype=np.float64):
    data = np.asarray(data, dtype)
    grid = np.asarray(grid, dtype)
    diffs = grid[:, None, :] - data[None, :, :]                       # (M, N, D)
    log_kernel_vals = -np.sum(diffs ** 2, axis=-1) / dtype(2 * bandwidth ** 2)
    log_pdf = _logsumexp(log_kernel_vals, -1) - np.log(dtype(data.shape[0]))
    z = log_pdf - log_pdf.max()
    e = np.exp(z)
    return e / e.sum()


def kl_divergence(p, q):
    return np.sum(p * np.log((p + 1e-10) / (q + 1e-10)))


def wasserstein_distance(p, q):
    return np.sum(np.abs(np.cumsum(p) - np.cumsum(q)))


def compute_rsr_loss(observations, policy_actions, next_observations, reference_data, reference_density, grid,
                     bandwidth, divergence, loss_scale=1.0, dtype=np.float64):
    obs = np.reshape(observations, (-1, observations.shape[-1]))
    act = np.reshape(policy_actions, (-1, policy_actions.shape[-1]))
    nxt = np.reshape(next_observations, (-1, next_observations.shape[-1]))
    current = np.concatenate([obs, act, nxt], axis=-1)
    augmented = np.concatenate([reference_data, current], axis=0)
    dens = evaluate_kde(augmented, grid, bandwidth, dtype)
    distance = wasserstein_distance(dens, np.asarray(reference_density, dtype))
    return dtype(loss_scale) * dtype(divergence) * distance, distance


def loss_grad_fd(batch, reference_data, reference_density, grid, bandwidth, divergence, loss_scale, eps=1e-6):
    """central finite differences of loss w.r.t. every entry of `batch` (float64)"""
    batch = np.asarray(batch, np.float64)
    g = np.zeros_like(batch)

    def f(b):
        aug = np.concatenate([reference_data, b], axis=0)
        d = evaluate_kde(aug, grid, bandwidth)
        return loss_scale * divergence * wasserstein_distance(d, reference_density)

    for i in range(batch.shape[0]):
        for j in range(batch.shape[1]):
            bp, bm = batch.copy(), batch.copy()
            bp[i, j] += eps
            bm[i, j] -= eps
            g[i, j] = (f(bp) - f(bm)) / (2 * eps)
    return g
