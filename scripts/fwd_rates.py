"""scratch: the rollout-time policy forward (uhc_linear_forward_tc chain, bf16 activations, fp32 head) at M rows: per layer and whole chain"""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
L = nn._lib(); dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dims = [657, 2048, 1024, 512, 105]
p64 = lambda n: (n + 63) // 64 * 64
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
acts = [torch.randn(M, p64(dims[0]), device=dev).to(torch.bfloat16)] + [torch.zeros(M, p64(d), device=dev, dtype=torch.bfloat16) for d in dims[1:-1]]
Ws = [torch.randn(dims[i + 1], p64(dims[i]), device=dev).to(torch.bfloat16) for i in range(4)]
bs = [torch.zeros(dims[i + 1], device=dev) for i in range(4)]
out = torch.empty(M, dims[-1], device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def layer(i):
    last = i == 3
    return L.uhc_linear_forward_tc(nn._p(acts[i]), nn._p(Ws[i]), nn._p(bs[i]), None if last else nn._p(acts[i + 1]), nn._p(out) if last else None, M, dims[i + 1], p64(dims[i]), 0 if last else p64(dims[i + 1]), 0 if last else 1, st)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
for i in range(4):
    assert layer(i) == 0
    ms = timeit(lambda: layer(i))
    print(f"forward L{i} {dims[i]:5d}->{dims[i+1]:5d} M={M}: {ms*1e3:8.1f} us  {2.0*M*dims[i]*dims[i+1]/ms/1e9:7.1f} TFLOP/s")
ms = timeit(lambda: [layer(i) for i in range(4)])
fl = sum(2.0 * M * dims[i] * dims[i + 1] for i in range(4))
print(f"chain M={M}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (L2 flushed before every timing)")
