import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import _lib
L = _lib.lib()
s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
stamps = torch.zeros(8, dtype=torch.int64, device="cuda")
os.environ["RSRX_GEMM_STAMPS"] = str(stamps.data_ptr())
g = torch.Generator("cuda").manual_seed(0)
R = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
names = ["start", "alloc+init", "stage issued", "staged+sync", "mma issued", "mma done", "tmem->smem", "epilogue"]
def show(tag):
    torch.cuda.synchronize()
    t = stamps.cpu().tolist()
    print(tag, " ".join(f"{n}:{(t[i]-t[0])/1e3:.2f}" for i, n in enumerate(names)), "(us since start)")
for (M, N, K) in [(2816, 256, 256), (128, 8, 32), (2816, 256, 32)]:
    x, w, b = R(M, K), R(N, K) * 0.1, R(N)
    z, y = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    for rep in range(3):
        _lib.check(L.rsrx_linear_forward(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), M, N, K, 1, z.data_ptr(), y.data_ptr(), N, s()))
    show(f"fwd {M}x{N}x{K}:")
    if N % 4 == 0:
        dz, zp = R(M, N), R(M, K)
        dzp = torch.zeros(M, K, device="cuda"); cs = torch.zeros((M + 127) // 128, K, device="cuda")
        for rep in range(3):
            _lib.check(L.rsrx_linear_dgrad(dz.data_ptr(), N, w.data_ptr(), K, zp.data_ptr(), M, K, N, 1, dzp.data_ptr(), K, cs.data_ptr(), s()))
        show(f"dgrad {M}x{N}x{K}:")
        S = (M + 255) // 256
        part = torch.zeros(S, N, K, device="cuda")
        for rep in range(3):
            _lib.check(L.rsrx_linear_wgrad(dz.data_ptr(), N, x.data_ptr(), K, M, N, K, 256, part.data_ptr(), K, s()))
        show(f"wgrad {M}x{N}x{K}:")
