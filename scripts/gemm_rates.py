"""scratch: every kernel of one net's training step (forward with transposed activations, dW, fused dX + activation backward, head dact) at N = 131072 rows,
timed in isolation with CUDA events; `legacy` as argv[2] times the unfused kernels (separate dX, dact, transposes) instead"""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
L = nn._lib()
dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
legacy = len(sys.argv) > 2 and sys.argv[2] == "legacy"
dims = [657, 2048, 1024, 512, 105]
p64 = lambda n: (n + 63) // 64 * 64
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def bench(fn, flops, reps=5):
    for _ in range(2): assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, flops / ms / 1e9
tot = 0.0
for i in range(4):
    K, N = dims[i], dims[i + 1]
    Kp, Np, Mp = p64(K), p64(N), p64(M)
    x = torch.randn(M, Kp, device=dev).to(torch.bfloat16); W = torch.randn(N, Kp, device=dev).to(torch.bfloat16); b = torch.zeros(N, device=dev)
    y = torch.zeros(M, Np, device=dev, dtype=torch.bfloat16); z = torch.empty(M, N, device=dev); out = torch.empty(M, N, device=dev)
    yT = torch.zeros(N, Mp, device=dev, dtype=torch.bfloat16)
    last = i == 3
    if last or legacy:
        f = lambda: L.uhc_linear_forward_tc_train(nn._p(x), nn._p(W), nn._p(b), None if last else nn._p(y), nn._p(out) if last else None, None if last else nn._p(z), M, N, Kp, 0 if last else Np, 0 if last else 1, st)
    else:
        f = lambda: L.uhc_linear_forward_tc_train_t(nn._p(x), nn._p(W), nn._p(b), nn._p(y), nn._p(yT), Mp, nn._p(z), M, N, Kp, Np, 1, st)
    ms, tf = bench(f, 2.0 * M * N * K); tot += ms
    print(f"forward  L{i} {K:5d}->{N:5d}: {ms:7.3f} ms {tf:7.1f} TFLOP/s")
    dz = torch.randn(M, Np, device=dev).to(torch.bfloat16); dzT = torch.randn(N, Mp, device=dev).to(torch.bfloat16); hT = torch.randn(K, Mp, device=dev).to(torch.bfloat16)
    dW = torch.empty(N, K, device=dev)
    f = lambda: L.uhc_linear_forward_tc(nn._p(dzT), nn._p(hT), None, None, nn._p(dW), N, K, Mp, 0, 0, st)
    ms, tf = bench(f, 2.0 * M * N * K); tot += ms
    print(f"dW       L{i} [{N}x{M}]x[{M}x{K}]: {ms:7.3f} ms {tf:7.1f} TFLOP/s")
    if i > 0:
        WT = torch.randn(K, Np, device=dev).to(torch.bfloat16); dh = torch.empty(M, K, device=dev)
        if legacy:
            f = lambda: L.uhc_linear_forward_tc(nn._p(dz), nn._p(WT), None, None, nn._p(dh), M, K, Np, 0, 0, st)
        else:
            zp = torch.randn(M, K, device=dev); dzp = torch.zeros(M, Kp, device=dev, dtype=torch.bfloat16); dzpT = torch.zeros(K, Mp, device=dev, dtype=torch.bfloat16); dbp = torch.zeros(K, device=dev)
            f = lambda: L.uhc_linear_dx_dact_tc(nn._p(dz), nn._p(WT), nn._p(zp), nn._p(dzp), nn._p(dzpT), nn._p(dbp), M, K, Np, Kp, Mp, 1, st)
        ms, tf = bench(f, 2.0 * M * N * K); tot += ms
        print(f"dX{'' if legacy else '+dact'} L{i} [{M}x{N}]x[{N}x{K}]: {ms:7.3f} ms {tf:7.1f} TFLOP/s")
    if last or legacy:
        dhh = torch.randn(M, N, device=dev); zz = torch.randn(M, N, device=dev); db = torch.zeros(N, device=dev)
        f = lambda: L.uhc_dact_bf16(nn._p(dhh), None if last else nn._p(zz), nn._p(dz), nn._p(dzT), nn._p(db), M, N, Np, Mp, 1, st)
        ms, _ = bench(f, 1.0); tot += ms
        print(f"dact     L{i} {M}x{N}: {ms:7.3f} ms  ({(M*N*(8+2+2))/ms/1e6:.0f} GB/s)")
    if i > 0 and legacy:
        h = torch.randn(M, Kp, device=dev).to(torch.bfloat16)
        f = lambda: L.uhc_transpose_bf16(nn._p(h), nn._p(hT), M, K, Kp, Mp, st)
        ms, _ = bench(f, 1.0); tot += ms
        print(f"transp   L{i} {M}x{K}: {ms:7.3f} ms  ({(M*K*4)/ms/1e6:.0f} GB/s)")
print(f"sum for one net: {tot:.2f} ms (an epoch runs two nets)")
