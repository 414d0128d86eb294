"""Scratch timing of the env-step kernel (not the bench contract)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
z = np.load(os.path.join(g, "expert_sway.npz")); ex = {k: z[k] for k in z.files}
so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
eng = Engine(E); eng.load_clips([ex], [so])
rng = np.random.RandomState(1)
starts = rng.randint(0, 40, E).astype(np.int32)
eng.reset(start=starts)
acts = torch.tensor(rng.normal(0, 0.1, (E, 105)), dtype=torch.float32, device="cuda")
for _ in range(3): eng.step(acts)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps): eng.step(acts)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
it = np.array([eng.get_state(e)["newton_iters"] for e in range(0, E, max(1, E // 64))])
nc = np.array([eng.get_state(e)["ncon"] for e in range(0, E, max(1, E // 64))])
print(f"E={E} ms/step={ms:.3f} env-steps/s={E / ms * 1e3:.0f} newton iters/step mean {it.mean():.1f} max {it.max()} ncon mean {nc.mean():.1f} fail frac {eng.fail.float().mean().item():.2f}")
