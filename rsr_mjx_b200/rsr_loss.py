"""RSR distribution loss on the fused CUDA kernel — same names, argument meaning
and error behaviour as the reference's RSR/rsr_loss.py (RSRData, make_grid,
build_rsr_data, compute_rsr_loss) and RSR/dataset_processor.py (evaluate_kde,
kl_divergence, wasserstein_distance).  Tensors are torch CUDA float32;
`compute_rsr_loss` is differentiable w.r.t. observations / policy_actions /
next_observations (the reference's policy gradient flows through
`policy_actions` only, RSR/losses.py:186-193)."""
from __future__ import annotations

import ctypes as C
from typing import Any, NamedTuple, Optional

import numpy as np
import torch

from . import _lib, prng


class RSRData(NamedTuple):
    """Precomputed real/sim distribution statistics used during training."""
    divergence: torch.Tensor
    reference_density: torch.Tensor
    reference_data: torch.Tensor
    grid: torch.Tensor
    bandwidth: float


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(x, device=None):
    t = torch.as_tensor(x, dtype=torch.float32, device=device)
    return t.contiguous()


def make_grid(num_samples: int, dimension: int, min_value: float = -3.0, max_value: float = 3.0, seed: int = 0,
              device="cuda") -> torch.Tensor:
    """U(min,max)^{M x D} from jax.random.uniform(PRNGKey(seed)) (rsr_loss.py:27-40),
    drawn with the NumPy threefry restatement."""
    g = prng.uniform(prng.PRNGKey(seed), (num_samples, dimension), np.float32(min_value), np.float32(max_value))
    return torch.from_numpy(g).to(device)


def evaluate_kde(data, grid, bandwidth: float = 0.1) -> torch.Tensor:
    grid = _f32c(grid)
    data = _f32c(data, grid.device)
    if grid.device.type != "cuda":
        raise RuntimeError("evaluate_kde runs only on CUDA tensors (no CPU fallback)")
    M, D = grid.shape
    if data.ndim != 2 or data.shape[1] != D:
        raise ValueError(f"data must be (N, {D}), got {tuple(data.shape)}")
    out = torch.empty(M, dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        _lib.check(_lib.lib().rsrx_kde(grid.data_ptr(), M, D, data.data_ptr(), data.shape[0], float(bandwidth),
                                       out.data_ptr(), _stream(grid.device)), "rsrx_kde")
    return out


def kl_divergence(p, q):
    return torch.sum(p * torch.log((p + 1e-10) / (q + 1e-10)))


def wasserstein_distance(p, q):
    return torch.sum(torch.abs(torch.cumsum(p, 0) - torch.cumsum(q, 0)))


def build_rsr_data(real_data, previous_sim_data, current_sim_data, *, num_samples: int = 10, min_value: float = -3.0,
                   max_value: float = 3.0, bandwidth: float = 0.1, seed: int = 0, device="cuda") -> RSRData:
    """Precomputes the fixed part of the RSR distribution objective (rsr_loss.py:43-91)."""
    real_data = _f32c(real_data, device)
    previous_sim_data = _f32c(previous_sim_data, device)
    current_sim_data = _f32c(current_sim_data, device)
    if real_data.ndim != 2:
        raise ValueError(f'real_data must be rank 2, got shape {tuple(real_data.shape)}')
    if previous_sim_data.shape != real_data.shape:
        raise ValueError('previous_sim_data must match real_data: '
                         f'{tuple(previous_sim_data.shape)} != {tuple(real_data.shape)}')
    if current_sim_data.shape != real_data.shape:
        raise ValueError('current_sim_data must match real_data: '
                         f'{tuple(current_sim_data.shape)} != {tuple(real_data.shape)}')
    if num_samples <= 0:
        raise ValueError(f'num_samples must be positive, got {num_samples}')
    if bandwidth <= 0:
        raise ValueError(f'bandwidth must be positive, got {bandwidth}')
    grid = make_grid(num_samples, real_data.shape[-1], min_value=min_value, max_value=max_value, seed=seed,
                     device=real_data.device)
    real_density = evaluate_kde(real_data, grid, bandwidth)
    previous_sim_density = evaluate_kde(previous_sim_data, grid, bandwidth)
    reference_density = evaluate_kde(current_sim_data, grid, bandwidth)
    # a host-side 0-d tensor: a constant of the objective, read on every loss call without a device sync
    divergence = kl_divergence(real_density, previous_sim_density).cpu()
    return RSRData(divergence=divergence, reference_density=reference_density, reference_data=current_sim_data,
                   grid=grid, bandwidth=bandwidth)


def _as_rsr_data(past_data: Any) -> RSRData:
    """Accepts RSRData and the legacy 3-/5-tuple formats (rsr_loss.py:94-119)."""
    if isinstance(past_data, RSRData):
        return past_data
    if not isinstance(past_data, (tuple, list)):
        raise TypeError('past_data must be RSRData or a tuple/list')
    if len(past_data) == 5:
        return RSRData(*past_data)
    if len(past_data) != 3:
        raise ValueError('legacy past_data must contain (KLD, density, reference_data)')
    divergence, reference_density, reference_data = past_data
    grid = make_grid(int(reference_density.shape[0]), int(reference_data.shape[-1]),
                     device=torch.as_tensor(reference_data).device if torch.is_tensor(reference_data) else "cuda")
    return RSRData(divergence=divergence, reference_density=reference_density, reference_data=reference_data,
                   grid=grid, bandwidth=0.1)


def prepare_rsr_data(past_data: Any, device) -> Any:
    """Normalises `past_data` ONCE (the trainers call this before warm-up and CUDA-graph capture): `divergence` becomes a
    host-side 0-d tensor and `bandwidth` a float, grid / reference arrays contiguous float32 on `device`.  A
    reference-style RSRData or legacy tuple whose KLD lives on the device would otherwise force a device sync on every
    loss call (and fail outright inside a stream capture)."""
    if past_data is None:
        return None
    r = _as_rsr_data(past_data)
    div = r.divergence
    div = div.detach().cpu().float().reshape(()) if torch.is_tensor(div) else torch.tensor(float(np.asarray(div)))
    bw = float(r.bandwidth.item() if torch.is_tensor(r.bandwidth) else np.asarray(r.bandwidth))
    return RSRData(divergence=div, reference_density=_f32c(r.reference_density, device),
                   reference_data=_f32c(r.reference_data, device), grid=_f32c(r.grid, device), bandwidth=bw)


class _RSRLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, batch, grid, reference_data, reference_density, bandwidth, divergence, loss_scale):
        M, D = grid.shape
        Nb, Nref = batch.shape[0], reference_data.shape[0]
        out = torch.empty(2, dtype=torch.float32, device=batch.device)
        need_grad = batch.requires_grad
        grad = torch.empty_like(batch) if need_grad else None
        with torch.cuda.device(batch.device):
            _lib.check(_lib.lib().rsrx_rsr_loss(
                grid.data_ptr(), M, D, reference_data.data_ptr(), Nref, batch.data_ptr(), Nb,
                reference_density.data_ptr(), float(bandwidth), float(divergence), float(loss_scale), None,
                out.data_ptr(), grad.data_ptr() if need_grad else None, _stream(batch.device)), "rsrx_rsr_loss")
        ctx.save_for_backward(grad if need_grad else torch.empty(0, device=batch.device))
        ctx.loss_scale, ctx.divergence = float(loss_scale), float(divergence)
        return out[0].clone(), out[1].clone()

    @staticmethod
    def backward(ctx, g_loss, g_dist):
        (grad,) = ctx.saved_tensors
        if grad.numel() == 0:
            return (None,) * 7
        g = grad * g_loss
        if g_dist is not None and ctx.loss_scale * ctx.divergence != 0.0:
            g = g + grad * (g_dist / (ctx.loss_scale * ctx.divergence))
        return g, None, None, None, None, None, None


class PolicyTerm:
    """The RSR term of the PPO loss (RSR/losses.py:186-195) on fixed buffers, without autograd or torch glue:
    `forward(obs, logits, next_obs)` packs [obs | tanh(loc) | next_obs], runs the KDE + Wasserstein kernels and keeps
    d loss / d transition; `add_logit_grad(g_in, g_out)` writes g_in + d loss / d logits.  `loss` / `distance` are views of
    a device buffer (valid after forward, overwritten by the next one).  Same arithmetic as
    `compute_rsr_loss(obs, NormalTanh.mode(logits), next_obs, past_data, loss_scale=...)` + autograd."""

    def __init__(self, past_data: Any, rows: int, obs_size: int, act_size: int, loss_scale: float, device):
        self.rsr = prepare_rsr_data(past_data, device)
        self.rows, self.O, self.A, self.scale = int(rows), int(obs_size), int(act_size), float(loss_scale)
        D = 2 * self.O + self.A
        if self.rsr.grid.shape[1] != D or self.rsr.reference_data.shape[1] != D:
            raise ValueError(f'online transition width does not match RSR reference data: {D} != {self.rsr.reference_data.shape[1]}')
        self.dev = torch.device(device)
        self.transition = torch.empty(self.rows, D, device=self.dev)
        self.grad_transition = torch.empty(self.rows, D, device=self.dev)
        self.out = torch.zeros(2, device=self.dev)
        self.loss, self.distance = self.out[0], self.out[1]
        self.divergence = float(self.rsr.divergence)

    def forward(self, obs: torch.Tensor, logits: torch.Tensor, next_obs: torch.Tensor) -> None:
        r = self.rsr
        for t, w in ((obs, self.O), (logits, 2 * self.A), (next_obs, self.O)):
            if t.numel() != self.rows * w or not t.is_contiguous():
                raise ValueError("PolicyTerm.forward: contiguous [rows, width] inputs expected")
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().rsrx_rsr_policy_term(
                r.grid.data_ptr(), r.grid.shape[0], r.reference_data.data_ptr(), r.reference_data.shape[0],
                r.reference_density.data_ptr(), float(r.bandwidth), self.divergence, self.scale, obs.data_ptr(), logits.data_ptr(),
                next_obs.data_ptr(), self.rows, self.O, self.A, self.transition.data_ptr(), self.grad_transition.data_ptr(),
                self.out.data_ptr(), _stream(self.dev)), "rsrx_rsr_policy_term")

    def add_logit_grad(self, g_in: Optional[torch.Tensor], g_out: torch.Tensor) -> None:
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().rsrx_rsr_logit_grad(
                self.transition.data_ptr(), self.grad_transition.data_ptr(), None if g_in is None else g_in.data_ptr(), self.rows,
                self.O, self.A, g_out.data_ptr(), _stream(self.dev)), "rsrx_rsr_logit_grad")


def compute_rsr_loss(observations, policy_actions, next_observations, past_data: Any, *, loss_scale: float = 1.0):
    """RSR transition-distribution penalty (rsr_loss.py:122-175).
    Returns (scaled_loss, distribution_distance)."""
    if past_data is None or loss_scale == 0.0:
        zero = torch.zeros((), dtype=observations.dtype, device=observations.device)
        return zero, zero
    rsr = _as_rsr_data(past_data)
    observation_size = observations.shape[-1]
    action_size = policy_actions.shape[-1]
    next_observation_size = next_observations.shape[-1]
    current = torch.cat([observations.reshape(-1, observation_size), policy_actions.reshape(-1, action_size),
                         next_observations.reshape(-1, next_observation_size)], dim=-1).float().contiguous()
    ref = _f32c(rsr.reference_data, current.device)
    if current.shape[-1] != ref.shape[-1]:
        raise ValueError('online transition width does not match RSR reference data: '
                         f'{current.shape[-1]} != {ref.shape[-1]}')
    if current.device.type != "cuda":
        raise RuntimeError("compute_rsr_loss runs only on CUDA tensors (no CPU fallback)")
    divergence = float(rsr.divergence)
    loss, dist = _RSRLossFn.apply(current, _f32c(rsr.grid, current.device), ref,
                                  _f32c(rsr.reference_density, current.device), float(rsr.bandwidth), divergence,
                                  float(loss_scale))
    return loss, dist
