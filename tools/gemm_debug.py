import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rsr_mjx_b200 import _lib
L = _lib.lib()
torch.set_printoptions(linewidth=250, precision=0, sci_mode=False)
s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = 128, 64, 64   # dz [M,N], w [N,K] -> dzp [M,K]
w = (torch.arange(N, device="cuda")[:, None] * 100 + torch.arange(K, device="cuda")[None, :] + 1).float()
dump = torch.full((64 * 32,), -1.0, device="cuda")
for cfg in ["0,0,0,0"]:
    os.environ["RSRX_GEMM_DBG"] = cfg + f",{dump.data_ptr()}"
    print("=== LBO,SBO,KSTEP,IDESC_XOR =", cfg)
    for sel in (0, 9):
        dz = torch.zeros(M, N, device="cuda"); dz[:, sel] = 1.0
        dzp = torch.full((M, K), -7.0, device="cuda")
        _lib.check(L.rsrx_linear_dgrad(dz.data_ptr(), N, w.data_ptr(), K, None, M, K, N, 0, dzp.data_ptr(), K, None, s()))
        torch.cuda.synchronize()
        print(f" sel {sel}: expect {sel*100+1}..{sel*100+64}:", dzp[3, :24].tolist(), "...", dzp[3, 60:64].tolist())
print("staged B tile image (first 3 core matrices = 96 floats):")
print(dump[:96].reshape(-1, 4))
print("floats 512..544:", dump[512:544].tolist())
