"""Scratch timing of the PPO update path (config 3 shape: N = T x E transitions)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pol = nn.MLPNet(657, (2048, 1024, 512), 105, "gelu", head_name="action_mean", seed=1)
val = nn.MLPNet(657, (2048, 1024, 512), 1, "gelu", head_name="value_head", seed=2)
log_std = torch.full((105,), -2.3, device="cuda")
x = torch.randn(N, 657, device="cuda").clamp(-5, 5)
a = torch.randn(N, 105, device="cuda") * 0.1
adv = torch.randn(N, device="cuda"); ret = torch.rand(N, device="cuda"); exps = torch.ones(N, device="cuda")
op, ov = nn.Adam(pol.params(), 5e-5), nn.Adam(val.params(), 3e-4)
flop = 6 * (4024530 + 3971073) * N
for tc in (False, True):
    nn.ppo_update(pol, val, log_std, op, ov, x, a, ret, adv, exps, epochs=1, use_tc=tc)
    torch.cuda.synchronize(); t0 = time.time()
    nn.ppo_update(pol, val, log_std, op, ov, x, a, ret, adv, exps, epochs=epochs, use_tc=tc)
    torch.cuda.synchronize(); dt = (time.time() - t0) / epochs
    print(f"N={N} use_tc={tc} per-epoch {dt*1e3:.1f} ms  -> {flop/dt/1e12:.1f} TFLOP/s effective; 10 epochs = {dt*10:.2f} s")
# tensor-core forward throughput
for M in (4096, 32768):
    xs = x[:M].contiguous(); pol.forward_tc(xs); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20): pol.forward_tc(xs)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 20
    print(f"policy forward_tc M={M}: {ms:.3f} ms -> {2*4024530*M/ms/1e9:.1f} TFLOP/s")
