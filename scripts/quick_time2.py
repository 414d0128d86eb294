"""Scratch: env-step kernel timing under bench-like conditions (auto reset, optional L2 flush between steps)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200.engine import Engine
from uhc_b200 import motion_lib
E, steps = 4096, 20
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
g = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
if "synth" in mode:
    ex = motion_lib.synthetic_clip(320, np.random.default_rng(1)); so = np.zeros(17)
else:
    z = np.load(os.path.join(g, "expert_sway.npz")); ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
eng = Engine(E, auto_reset=1 if "auto" in mode else 0)
eng.load_clips([ex], [so])
rng = np.random.RandomState(1)
eng.reset(start=rng.randint(0, 40, E).astype(np.int32))
acts = torch.tensor(rng.normal(0, 0.1, (E, 105)), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if "flush" in mode else None
for _ in range(3): eng.step(acts)
torch.cuda.synchronize()
tot = 0.0
for _ in range(steps):
    if flush is not None: flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(acts); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"{mode}: {tot / steps:.3f} ms/step  fail frac {eng.fail.float().mean().item():.3f}")
