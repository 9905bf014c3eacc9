#!/usr/bin/env python
"""tcgen05 linear-layer kernels (csrc/rsrx_gemm.cuh) against torch fp32 matmul: forward / dgrad / wgrad, the value-net
shapes of BASELINE config 3 plus ragged ones.  Prints max relative errors and kernel times."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import _lib
L = _lib.lib()
torch.backends.cuda.matmul.allow_tf32 = False
s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator("cuda").manual_seed(0)
R = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
out = {}

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-9))

def timeit(fn, n=20):
    """us per call inside a CUDA graph of n calls (no host launch overhead)"""
    fn(); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(n): fn()
    g_.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g_.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3

for (M, N, K) in [(2816, 256, 256), (2816, 256, 32), (300, 64, 64), (128, 8, 32), (2816, 32, 256)]:
    x, w, b = R(M, K), R(N, K) * 0.1, R(N)
    z, y = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    ldt = (M + 3) // 4 * 4
    yT = torch.zeros(N, ldt, device="cuda")
    fwd = lambda: _lib.check(L.rsrx_linear_forward(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), M, N, K, 1, z.data_ptr(), y.data_ptr(), N, yT.data_ptr(), ldt, s()))
    fwd(); torch.cuda.synchronize()
    zr = x @ w.t() + b
    out[f"fwd_{M}x{N}x{K}"] = dict(err_z=rel(z, zr), err_y=rel(y, torch.nn.functional.silu(zr)), err_yT=rel(yT[:, :M], y.t()), us=timeit(fwd),
                                   torch_us=timeit(lambda: torch.nn.functional.silu(torch.addmm(b, x, w.t()))))
    # dgrad: dzprev[M, K] = (dz[M, N] w[N, K]) * silu'(zprev)
    dz, zp = R(M, N), R(M, K)
    dzp = torch.zeros(M, K, device="cuda"); nblk = (M + 127) // 128
    cs = torch.zeros(nblk, K, device="cuda")
    dzpT = torch.zeros(K, ldt, device="cuda")
    wT = w.t().contiguous()
    dg = lambda: _lib.check(L.rsrx_linear_dgrad(dz.data_ptr(), N, w.data_ptr(), K, None, 0, zp.data_ptr(), M, K, N, 1, dzp.data_ptr(), K, cs.data_ptr(), None, 0, s()))
    dgt = lambda: _lib.check(L.rsrx_linear_dgrad(dz.data_ptr(), N, None, 0, wT.data_ptr(), N, zp.data_ptr(), M, K, N, 1, dzp.data_ptr(), K, cs.data_ptr(), dzpT.data_ptr(), ldt, s()))
    if N % 4 == 0:
        dg(); torch.cuda.synchronize()
        sg = torch.sigmoid(zp)
        ref = (dz @ w) * (sg * (1 + zp * (1 - sg)))
        out[f"dgrad_{M}x{N}x{K}"] = dict(err=rel(dzp, ref), err_colsum=rel(cs.sum(0), ref.sum(0)), us=timeit(dg))
        dzp.zero_(); dgt(); torch.cuda.synchronize()
        out[f"dgrad_tma_{M}x{N}x{K}"] = dict(err=rel(dzp, ref), err_T=rel(dzpT[:, :M], dzp.t()), us=timeit(dgt))
        # wgrad: dw[N, K] = dz^T x over row slices of 256
        rps = 256
        S = (M + rps - 1) // rps
        part = torch.zeros(S, N, K, device="cuda")
        wg = lambda: _lib.check(L.rsrx_linear_wgrad(dz.data_ptr(), N, x.data_ptr(), K, 0, M, N, K, rps, part.data_ptr(), K, s()))
        dzT_, xT_ = torch.zeros(N, ldt, device="cuda"), torch.zeros(K, ldt, device="cuda")
        dzT_[:, :M] = dz.t(); xT_[:, :M] = x.t()
        wgt = lambda: _lib.check(L.rsrx_linear_wgrad(dzT_.data_ptr(), ldt, xT_.data_ptr(), ldt, 1, M, N, K, rps, part.data_ptr(), K, s()))
        wg(); torch.cuda.synchronize()
        dw = torch.zeros(N, K, device="cuda")
        ins = (C.c_void_p * 1)(part.data_ptr()); outs = (C.c_void_p * 1)(dw.data_ptr())
        ns = (C.c_int32 * 1)(N * K); Ss = (C.c_int32 * 1)(S); st = (C.c_int64 * 1)(N * K)
        rd = lambda: _lib.check(L.rsrx_reduce_partials(ins, outs, ns, Ss, st, 1, s()))
        rd(); torch.cuda.synchronize()
        out[f"wgrad_{M}x{N}x{K}"] = dict(err=rel(dw, dz.t() @ x), us=timeit(wg), reduce_us=timeit(rd), torch_us=timeit(lambda: dz.t() @ x))
        part.zero_(); wgt(); rd(); torch.cuda.synchronize()
        out[f"wgrad_tma_{M}x{N}x{K}"] = dict(err=rel(dw, dz.t() @ x), us=timeit(wgt))
print(json.dumps(out, indent=1))
