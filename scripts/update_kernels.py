"""Scratch: one tensor-core PPO epoch at N transitions (for a per-kernel ncu duration list)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
pol = nn.MLPNet(657, (2048, 1024, 512), 105, "gelu", head_name="action_mean", seed=1)
val = nn.MLPNet(657, (2048, 1024, 512), 1, "gelu", head_name="value_head", seed=2)
log_std = torch.full((105,), -2.3, device="cuda")
x = torch.randn(N, 657, device="cuda").clamp(-5, 5)
a = torch.randn(N, 105, device="cuda") * 0.1
adv = torch.randn(N, device="cuda"); ret = torch.rand(N, device="cuda"); exps = torch.ones(N, device="cuda")
op, ov = nn.Adam(pol.params(), 5e-5), nn.Adam(val.params(), 3e-4)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    nn.ppo_update(pol, val, log_std, op, ov, x, a, ret, adv, exps, epochs=1, use_tc=True)
torch.cuda.synchronize()
