"""scratch: one training-forward GEMM launch (layer 0 shape by default) for ncu"""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
L = nn._lib(); dev = torch.device("cuda", 0)
M, K, N = 131072, int(sys.argv[1]) if len(sys.argv) > 1 else 657, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
p64 = lambda n: (n + 63) // 64 * 64
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(M, p64(K), device=dev).to(torch.bfloat16); W = torch.randn(N, p64(K), device=dev).to(torch.bfloat16); b = torch.zeros(N, device=dev)
y = torch.zeros(M, p64(N), device=dev, dtype=torch.bfloat16); z = torch.empty(M, N, device=dev)
for _ in range(3):
    L.uhc_linear_forward_tc_train(nn._p(x), nn._p(W), nn._p(b), nn._p(y), None, nn._p(z), M, N, p64(K), p64(N), 1, st)
torch.cuda.synchronize()
