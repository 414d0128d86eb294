"""scratch: one training GEMM launch for ncu: argv = mode (fwd | dact | dw) K N   (dw: dz^T[N][M] h^T[K][M] -> dW[N][K], reduction over the M = 131072 rows)
      (fwd: x[M][K] W[N][K] -> z, y, yT;  dact: dz[M][K] WT[N][K] z_prev[M][N] -> dz_prev, dz_prev^T, db)"""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
L = nn._lib(); dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
M, K, N = 131072, int(sys.argv[2]) if len(sys.argv) > 2 else 657, int(sys.argv[3]) if len(sys.argv) > 3 else 2048
p64 = lambda n: (n + 63) // 64 * 64
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(M, p64(K), device=dev).to(torch.bfloat16); W = torch.randn(N, p64(K), device=dev).to(torch.bfloat16); b = torch.zeros(N, device=dev)
y = torch.zeros(M, p64(N), device=dev, dtype=torch.bfloat16); z = torch.randn(M, N, device=dev); yT = torch.zeros(N, p64(M), device=dev, dtype=torch.bfloat16)
db = torch.zeros(N, device=dev)
if mode == "dw":
    dzT = torch.randn(N, M, device=dev).to(torch.bfloat16); hT = torch.randn(K, M, device=dev).to(torch.bfloat16); dW = torch.empty(N, K, device=dev)
for _ in range(3):
    if mode == "dw":
        assert L.uhc_linear_forward_tc(nn._p(dzT), nn._p(hT), None, None, nn._p(dW), N, K, M, 0, 0, st) == 0
        continue
    if mode == "fwd":
        rc = L.uhc_linear_forward_tc_train_t(nn._p(x), nn._p(W), nn._p(b), nn._p(y), nn._p(yT), p64(M), nn._p(z), M, N, p64(K), p64(N), 1, st)
    else:
        rc = L.uhc_linear_dx_dact_tc(nn._p(x), nn._p(W), nn._p(z), nn._p(y), nn._p(yT), nn._p(db), M, N, p64(K), p64(N), p64(M), 1, st)
    assert rc == 0
torch.cuda.synchronize()
