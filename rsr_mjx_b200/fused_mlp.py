"""Value-network forward / backward on the hand-written tensor-core kernels (csrc/rsrx_gemm.cuh).

The value MLP of the PPO trainer (brax make_ppo_networks: obs -> 256 x 5 -> 1, swish; RSR/train.py, RSR/losses.py:128-131)
is the one real contraction on the training path: 2816 rows x (23->256, 4 x 256->256, 256->1) forward and backward on every
minibatch step, 256 times per training step.  Through torch it is ~40 launches per step (addmm, SiLU, SiLU', bias
reductions, three cuBLAS GEMMs per layer).  Here every hidden layer is ONE tcgen05 launch per direction:

    forward   Z_l = H_{l-1} W_l^T + b_l, H_l = act(Z_l)                      rsrx_linear_forward   (bias + act fused)
    dgrad     dZ_{l-1} = (dZ_l W_l) * act'(Z_{l-1}), colsum partials         rsrx_linear_dgrad     (act' + bias grad fused)
    wgrad     dW_l partials = dZ_l^T H_{l-1} over 256-row slices             rsrx_linear_wgrad     (split-K, deterministic)
    scalar output layer: torch.addmv forward, rsrx_value_head_backward (fused with act' of the last hidden layer)
    one rsrx_reduce_partials launch finishes every bias / weight gradient straight into the parameters' .grad buffers.

Parameters stay those of the `MLP` module (checkpoints, optimiser, all-reduce unchanged); this class only owns activation
and gradient scratch.  Inputs are fp32 read as TF32 by the tensor core (the precision torch / jax use for fp32 matmuls on
NVIDIA GPUs by default), accumulation fp32.  No autograd inside: `backward(g)` is called with d loss / d values.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib

_ACT = {"none": 0, "silu": 1, "relu": 2}


def supported(mlp, device) -> bool:
    """hidden widths multiples of 4 (16-byte rows), scalar output, CUDA"""
    layers = list(mlp.layers)
    return (torch.device(device).type == "cuda" and len(layers) >= 2 and layers[-1].out_features == 1
            and all(l.out_features % 4 == 0 for l in layers[:-1]))


class TensorCoreMLP:
    ROWS_PER_SPLIT = 256

    def __init__(self, mlp, rows: int, device):
        if not supported(mlp, device):
            raise ValueError("TensorCoreMLP needs a CUDA MLP with a scalar output and hidden widths that are multiples of 4")
        self.mlp, self.rows, self.dev = mlp, int(rows), torch.device(device)
        self.act = _ACT[mlp.activation]
        self.layers = list(mlp.layers)
        self.hidden = self.layers[:-1]
        self.k0 = self.hidden[0].in_features
        self.k0p = (self.k0 + 31) // 32 * 32
        z = lambda *s: torch.zeros(*s, device=self.dev)
        M = self.rows
        self.ldt = (M + 3) // 4 * 4   # leading dimension of the transposed copies ([width][rows]: contraction-contiguous for wgrad)
        self.x_pad = z(M, self.k0p)
        self.xT = z(self.k0p, self.ldt)
        self.w0_pad = z(self.hidden[0].out_features, self.k0p)
        self.Z = [z(M, l.out_features) for l in self.hidden]
        self.H = [z(M, l.out_features) for l in self.hidden]
        self.HT = [z(l.out_features, self.ldt) for l in self.hidden[:-1]]  # inputs of layers 1 .. (the last H only feeds the head)
        # one gradient buffer (+ transposed copy) per hidden layer rather than a ping-pong pair: the weight-gradient GEMM of
        # layer l runs on a side stream while the main stream already computes dZ of layer l - 2
        self.dZ = [z(M, l.out_features) for l in self.hidden]
        self.dZT = [z(l.out_features, self.ldt) for l in self.hidden]
        self._side = None
        # transposed weight copies [in][out] of layers 1 .. for the dgrad GEMM's TMA; kept current by FusedAdam (wT_of), or
        # refreshed at the start of forward() when another optimiser is used
        self.WT = {l.weight: z(l.in_features, l.out_features) for l in self.hidden[1:]}
        self.wt_fresh = False
        self.mtiles = (M + 127) // 128
        self.splits = (M + self.ROWS_PER_SPLIT - 1) // self.ROWS_PER_SPLIT
        self.colsum = [z(self.mtiles, l.out_features) for l in self.hidden]
        self.wpart = [z(self.splits, l.out_features, self.k0p if i == 0 else l.in_features) for i, l in enumerate(self.hidden)]
        n_last = self.hidden[-1].out_features
        self.dw_out_part, self.db_out_part = z(self.mtiles, n_last), z(self.mtiles)
        self.w0_grad_pad = z(self.hidden[0].out_features, self.k0p)
        # persistent gradient buffers, handed to the parameters as .grad
        self.grads = {p: torch.zeros_like(p) for l in self.layers for p in (l.weight, l.bias)}
        self._reduce_args = None

    def attach_grads(self) -> None:
        """point every value-network parameter's .grad at its persistent buffer (after zero_grad(set_to_none=True))"""
        for p, g in self.grads.items():
            p.grad = g

    def refresh_transposed_weights(self) -> None:
        for w, wt in self.WT.items():
            wt.copy_(w.detach().t())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [rows, in_features] (CUDA float32) -> values [rows].  x = None: the caller has already written the padded input
        and its transpose into self.x_pad / self.xT (rsrx_ppo_prep does, in the trainer)."""
        L, M, s = _lib.lib(), self.rows, self._stream()
        if x is not None:
            if tuple(x.shape) != (M, self.k0):
                raise ValueError(f"expected input [{M}, {self.k0}], got {tuple(x.shape)}")
            self.x_pad[:, :self.k0].copy_(x)
            self.xT[:self.k0, :M].copy_(x.t())
        if not self.wt_fresh:
            self.refresh_transposed_weights()
        self.w0_pad[:, :self.k0].copy_(self.hidden[0].weight)
        h, ldh, k = self.x_pad, self.k0p, self.k0p
        for i, l in enumerate(self.hidden):
            w = self.w0_pad if i == 0 else l.weight
            n = l.out_features
            ht = self.HT[i].data_ptr() if i < len(self.HT) else None
            _lib.check(L.rsrx_linear_forward(h.data_ptr(), ldh, w.data_ptr(), k, l.bias.data_ptr(), M, n, k, self.act,
                                             self.Z[i].data_ptr(), self.H[i].data_ptr(), n, ht, self.ldt, s), "rsrx_linear_forward")
            h, ldh, k = self.H[i], n, n
        out = self.layers[-1]
        return torch.addmv(out.bias.expand(M), h, out.weight[0])

    # ----------------------------------------------------------------- backward
    @torch.no_grad()
    def backward(self, g: torch.Tensor, overlap: bool = True) -> None:
        """g = d loss / d values [rows]; fills the .grad buffers of every value-network parameter.
        overlap: the weight-gradient GEMMs (dW_l = dZ_l^T X_{l-1}) go to a side stream and run beside the data-gradient
        chain (dZ_{l-1} = dZ_l W_l * act'), which is the critical path; both are forked from and joined back into the
        current stream (inside a stream capture this becomes two parallel branches of the graph)."""
        L, M = _lib.lib(), self.rows
        main = torch.cuda.current_stream(self.dev)
        if overlap and self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        side = self._side if overlap else main
        s, s2 = C.c_void_p(main.cuda_stream), C.c_void_p(side.cuda_stream)
        g = g.contiguous()
        out, nh = self.layers[-1], len(self.hidden)
        n = self.hidden[-1].out_features
        _lib.check(L.rsrx_value_head_backward(g.data_ptr(), out.weight.data_ptr(), self.Z[-1].data_ptr(), self.H[-1].data_ptr(),
                                              M, n, n, self.act, self.dZ[-1].data_ptr(), self.colsum[-1].data_ptr(),
                                              self.dw_out_part.data_ptr(), self.db_out_part.data_ptr(),
                                              self.dZT[-1].data_ptr(), self.ldt, s), "rsrx_value_head_backward")
        for i in range(nh - 1, -1, -1):
            l = self.hidden[i]
            n, kin = l.out_features, (self.k0p if i == 0 else l.in_features)
            xt = self.xT if i == 0 else self.HT[i - 1]
            if overlap:
                side.wait_stream(main)  # dZ_i / dZT_i are complete
            # dW_l partials = dZ_l^T X_{l-1}: both operands from their transposed copies (contraction-contiguous, TMA)
            _lib.check(L.rsrx_linear_wgrad(self.dZT[i].data_ptr(), self.ldt, xt.data_ptr(), self.ldt, 1, M, n, kin,
                                           self.ROWS_PER_SPLIT, self.wpart[i].data_ptr(), kin, s2), "rsrx_linear_wgrad")
            if i > 0:
                wt = self.WT[l.weight]
                _lib.check(L.rsrx_linear_dgrad(self.dZ[i].data_ptr(), n, l.weight.data_ptr(), kin, wt.data_ptr(), n,
                                               self.Z[i - 1].data_ptr(), M, kin, n, self.act, self.dZ[i - 1].data_ptr(), kin,
                                               self.colsum[i - 1].data_ptr(), self.dZT[i - 1].data_ptr(), self.ldt, s),
                           "rsrx_linear_dgrad")
        if overlap:
            main.wait_stream(side)
        if self._reduce_args is None:
            ins, outs, ns, Ss, strides = [], [], [], [], []
            for i, l in enumerate(self.hidden):
                n, kin = l.out_features, (self.k0p if i == 0 else l.in_features)
                ins.append(self.colsum[i]); outs.append(self.grads[l.bias]); ns.append(n); Ss.append(self.mtiles); strides.append(n)
                ins.append(self.wpart[i]); outs.append(self.w0_grad_pad if i == 0 else self.grads[l.weight])
                ns.append(n * kin); Ss.append(self.splits); strides.append(n * kin)
            nl = self.hidden[-1].out_features
            ins.append(self.dw_out_part); outs.append(self.grads[out.weight]); ns.append(nl); Ss.append(self.mtiles); strides.append(nl)
            ins.append(self.db_out_part); outs.append(self.grads[out.bias]); ns.append(1); Ss.append(self.mtiles); strides.append(1)
            k = len(ins)
            self._reduce_keep = (ins, outs)
            self._reduce_args = ((C.c_void_p * k)(*[t.data_ptr() for t in ins]), (C.c_void_p * k)(*[t.data_ptr() for t in outs]),
                                 (C.c_int32 * k)(*ns), (C.c_int32 * k)(*Ss), (C.c_int64 * k)(*strides), k)
        a = self._reduce_args
        _lib.check(L.rsrx_reduce_partials(a[0], a[1], a[2], a[3], a[4], a[5], s), "rsrx_reduce_partials")
        self.grads[self.hidden[0].weight].copy_(self.w0_grad_pad[:, :self.k0])


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps) for a fixed list of parameters as ONE launch per step (`rsrx_adam_step`); the
    step count lives on the device, so the step can be captured in a CUDA graph.  `grad_scale` multiplies the gradients
    (1 / world_size after a sum all-reduce).  Same update rule as torch / optax.adam (bias-corrected first and second
    moments, eps outside the square root)."""

    def __init__(self, params, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0, transposed=None):
        """transposed: {parameter: tensor [cols, rows]} — 2-D parameters whose transposed copy the step keeps current
        (TensorCoreMLP.WT: the weight operand of the dgrad GEMM)"""
        self.params: List[torch.Tensor] = [p for p in params]
        if not self.params or len(self.params) > 32:
            raise ValueError("FusedAdam takes 1..32 parameter tensors")
        if any(p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous() for p in self.params):
            raise ValueError("FusedAdam needs contiguous float32 CUDA parameters")
        self.lr, self.betas, self.eps, self.grad_scale = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(grad_scale)
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.ticket = torch.zeros(1, dtype=torch.int64, device=self.params[0].device)
        k = len(self.params)
        self._p = (C.c_void_p * k)(*[p.data_ptr() for p in self.params])
        self._m = (C.c_void_p * k)(*[t.data_ptr() for t in self.exp_avg])
        self._v = (C.c_void_p * k)(*[t.data_ptr() for t in self.exp_avg_sq])
        self._n = (C.c_int32 * k)(*[p.numel() for p in self.params])
        tr = dict(transposed or {})
        for p, t in tr.items():
            if p.dim() != 2 or tuple(t.shape) != (p.shape[1], p.shape[0]) or not t.is_contiguous():
                raise ValueError("FusedAdam: a transposed copy must be a contiguous [cols, rows] tensor of a 2-D parameter")
        self._keep_t = tr
        self._pt = (C.c_void_p * k)(*[(tr[p].data_ptr() if p in tr else None) for p in self.params])
        self._cols = (C.c_int32 * k)(*[(p.shape[1] if p in tr else 1) for p in self.params])

    @property
    def state(self):
        """torch-optimizer-shaped view of the state (the trainers zero it after their warm-up)"""
        return {i: {"exp_avg": m, "exp_avg_sq": v, "ticket": self.ticket} for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))}

    @property
    def steps_taken(self) -> int:
        return int(self.ticket.item()) & 0xFFFFFFFF

    def reset_state(self) -> None:
        for t in self.exp_avg + self.exp_avg_sq:
            t.zero_()
        self.ticket.zero_()

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, grads=None) -> None:
        gs = grads if grads is not None else [p.grad for p in self.params]
        if any(g is None for g in gs):
            raise RuntimeError("FusedAdam.step: a parameter has no gradient")
        gs = [g if g.is_contiguous() else g.contiguous() for g in gs]
        k = len(self.params)
        gp = (C.c_void_p * k)(*[g.data_ptr() for g in gs])
        dev = self.params[0].device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsrx_adam_step(self._p, gp, self._m, self._v, self._n, self._pt, self._cols, k, self.lr, self.betas[0], self.betas[1],
                                                 self.eps, self.grad_scale, self.ticket.data_ptr(),
                                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rsrx_adam_step")


def warp_supported(mlp, device) -> bool:
    layers = list(mlp.layers)
    return (torch.device(device).type == "cuda" and 1 <= len(layers) <= 8
            and all(l.in_features <= 32 and l.out_features <= 32 for l in layers))


class WarpMLP:
    """Policy-network forward / backward on the warp-per-row kernels (csrc/rsrx_mlp.cuh): every width <= 32, one launch
    per direction instead of ~45 torch launches.  Parameters stay the module's; `backward(g)` fills persistent .grad
    buffers (one partial per CTA, summed in order by rsrx_reduce_partials: deterministic).  No autograd inside."""

    def __init__(self, mlp, rows: int, device):
        if not warp_supported(mlp, device):
            raise ValueError("WarpMLP needs a CUDA MLP with <= 8 layers and every width <= 32")
        self.mlp, self.rows, self.dev = mlp, int(rows), torch.device(device)
        self.layers = list(mlp.layers)
        self.act = _ACT[mlp.activation]
        nl = len(self.layers)
        self.widths = [self.layers[0].in_features] + [l.out_features for l in self.layers]
        self._W = (C.c_void_p * nl)(*[l.weight.data_ptr() for l in self.layers])
        self._b = (C.c_void_p * nl)(*[l.bias.data_ptr() for l in self.layers])
        self._w = (C.c_int32 * (nl + 1))(*self.widths)
        self.zs = torch.zeros(max(nl - 1, 1), self.rows, 32, device=self.dev)
        self.out = torch.zeros(self.rows, self.widths[-1], device=self.dev)
        self.total = sum(l.weight.numel() + l.bias.numel() for l in self.layers)
        self.ctas = int(_lib.lib().rsrx_small_mlp_backward_ctas(self.rows))
        self.partials = torch.zeros(self.ctas, self.total, device=self.dev)
        self.grads = {p: torch.zeros_like(p) for l in self.layers for p in (l.weight, l.bias)}
        ins, outs, ns, off = [], [], [], 0
        for l in self.layers:
            for p in (l.weight, l.bias):
                ins.append(self.partials.data_ptr() + 4 * off); outs.append(self.grads[p].data_ptr()); ns.append(p.numel())
                off += p.numel()
        k = len(ins)
        self._reduce = ((C.c_void_p * k)(*ins), (C.c_void_p * k)(*outs), (C.c_int32 * k)(*ns), (C.c_int32 * k)(*([self.ctas] * k)),
                        (C.c_int64 * k)(*([self.total] * k)), k)
        self._x = None

    def attach_grads(self) -> None:
        for p, g in self.grads.items():
            p.grad = g

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [rows, in_features], rows `x.stride(0)` floats apart (a padded observation buffer is fine).  mean / std: the
        input is normalised inside the launch, (x - mean) / std (forward-only use: the actor step)."""
        if tuple(x.shape) != (self.rows, self.widths[0]) or x.stride(1) != 1:
            raise ValueError(f"expected an input [{self.rows}, {self.widths[0]}] with contiguous rows")
        if (mean is None) != (std is None):
            raise ValueError("mean and std go together")
        self._x = x if mean is None else None  # backward() needs the layer-0 input as the kernel saw it
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().rsrx_small_mlp_forward(self._W, self._b, self._w, len(self.layers), self.act, x.data_ptr(),
                                                         x.stride(0), self.rows, self.zs.data_ptr(), self.out.data_ptr(),
                                                         self.widths[-1], None if mean is None else mean.data_ptr(),
                                                         None if std is None else std.data_ptr(), self._stream()),
                       "rsrx_small_mlp_forward")
        return self.out

    @torch.no_grad()
    def backward(self, g: torch.Tensor) -> None:
        g = g.reshape(self.rows, self.widths[-1])
        if not g.is_contiguous():
            g = g.contiguous()
        if self._x is None:
            raise RuntimeError("WarpMLP.backward after a forward with in-kernel normalisation: pass the normalised input instead")
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            _lib.check(L.rsrx_small_mlp_backward(self._W, self._b, self._w, len(self.layers), self.act, self._x.data_ptr(),
                                                 self._x.stride(0), self.rows, self.zs.data_ptr(), g.data_ptr(), self.widths[-1],
                                                 self.partials.data_ptr(), self._stream()), "rsrx_small_mlp_backward")
            a = self._reduce
            _lib.check(L.rsrx_reduce_partials(a[0], a[1], a[2], a[3], a[4], a[5], self._stream()), "rsrx_reduce_partials")
