"""Per-phase timestamps (globaltimer, CTA 0) of the tensor-core GEMM launches the PPO value network makes:
RSRX_GEMM_STAMPS=<device pointer> makes gemm_tf32_kernel record 8 stamps.  usage: python tools/gemm_debug.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
stamps = torch.zeros(8, dtype=torch.int64, device="cuda")
os.environ["RSRX_GEMM_STAMPS"] = str(stamps.data_ptr())
from rsr_mjx_b200 import _lib
L = _lib.lib()
s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator("cuda").manual_seed(0)
R = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
names = ["start", "alloc+init", "stage issued", "staged+sync", "mma issued", "mma done", "tmem->smem", "epilogue"]
def show(tag, ms):
    torch.cuda.synchronize()
    t = stamps.cpu().tolist()
    print(f"{tag:28s} launch {ms * 1e3:6.2f} us |", " ".join(f"{n}:{(t[i]-t[0])/1e3:.2f}" for i, n in enumerate(names)), "(us since CTA 0 start)")
def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
M, N, K = 2816, 256, 256
ldt = M
x, w, b = R(M, K), R(N, K) * 0.1, R(N)
wT = w.t().contiguous()
z, y, yT = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda"), torch.zeros(N, ldt, device="cuda")
show("forward 2816x256x256", timed(lambda: _lib.check(L.rsrx_linear_forward(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), M, N, K, 1, z.data_ptr(), y.data_ptr(), N, yT.data_ptr(), ldt, s()))))
show("forward, no transposed copy", timed(lambda: _lib.check(L.rsrx_linear_forward(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), M, N, K, 1, z.data_ptr(), y.data_ptr(), N, None, 0, s()))))
dz, zp = R(M, N), R(M, K)
dzp, dzpT = torch.zeros(M, K, device="cuda"), torch.zeros(K, ldt, device="cuda")
cs = torch.zeros((M + 127) // 128, K, device="cuda")
show("dgrad", timed(lambda: _lib.check(L.rsrx_linear_dgrad(dz.data_ptr(), N, w.data_ptr(), K, wT.data_ptr(), N, zp.data_ptr(), M, K, N, 1, dzp.data_ptr(), K, cs.data_ptr(), dzpT.data_ptr(), ldt, s()))))
dzT, xT = dz.t().contiguous(), x.t().contiguous()
S = (M + 255) // 256
part = torch.zeros(S, N, K, device="cuda")
show("wgrad (transposed inputs)", timed(lambda: _lib.check(L.rsrx_linear_wgrad(dzT.data_ptr(), ldt, xT.data_ptr(), ldt, 1, M, N, K, 256, part.data_ptr(), K, s()))))
