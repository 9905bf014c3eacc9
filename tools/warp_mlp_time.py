import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import ppo, fused_mlp
def timeit(fn, n=20):
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
for rows in (320, 1024, 2560):
    mlp = ppo.MLP([23, 32, 32, 32, 32, 10]).cuda()
    x = torch.randn(rows, 23, device="cuda"); g = torch.randn(rows, 10, device="cuda")
    wm = fused_mlp.WarpMLP(mlp, rows, "cuda")
    wm.forward(x); wm.attach_grads()
    print(rows, "fwd us %.1f" % timeit(lambda: wm.forward(x)), "bwd+reduce us %.1f" % timeit(lambda: wm.backward(g)), "ctas", wm.ctas)
    opt = fused_mlp.FusedAdam(list(mlp.parameters()), lr=1e-3)
    print("   adam (10 tensors) us %.1f" % timeit(lambda: opt.step()))
