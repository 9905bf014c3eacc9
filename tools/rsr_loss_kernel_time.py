#!/usr/bin/env python
"""RSR loss kernels through the C-ABI alone (no torch ops in the loop): µs per rsrx_rsr_loss call (forward only / forward +
gradient) at the batch sizes of BASELINE config 4.  `ncu --metrics gpu__time_duration.sum` of this script gives
profiles/r2_rsr_loss_ncu.csv."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import _lib, rsr_loss
L = _lib.lib()
g = np.random.default_rng(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
for Nb in (128, 1280, 2560):
    grid = torch.from_numpy(g.uniform(-3, 3, (10, 51)).astype(np.float32)).cuda()
    ref = torch.from_numpy(g.normal(0, 1, (50, 51)).astype(np.float32)).cuda()
    x = torch.from_numpy(g.normal(0, 1, (Nb, 51)).astype(np.float32)).cuda()
    refd = rsr_loss.evaluate_kde(ref, grid, 0.1)
    o = torch.empty(2, device="cuda"); gr = torch.empty_like(x)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for with_grad in (False, True):
        def call():
            _lib.check(L.rsrx_rsr_loss(grid.data_ptr(), 10, 51, ref.data_ptr(), 50, x.data_ptr(), Nb, refd.data_ptr(), 0.1, 0.5, 1.0,
                                       None, o.data_ptr(), gr.data_ptr() if with_grad else None, s))
        for _ in range(10): call()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): call()
        b.record(); torch.cuda.synchronize()
        out[f"rsr_loss_{'fwd_bwd' if with_grad else 'fwd'}_us_Nb{Nb}"] = a.elapsed_time(b) / reps * 1e3
print(json.dumps(out))
