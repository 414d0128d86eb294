"""Scratch: tensor-core policy forward at M rows (for ncu captures of k_linear_tc)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import nn
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
pol = nn.MLPNet(657, (2048, 1024, 512), 105, "gelu", head_name="action_mean", seed=1)
x = torch.randn(M, 657, device="cuda").clamp(-5, 5)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    pol.forward_tc(x)
torch.cuda.synchronize()
