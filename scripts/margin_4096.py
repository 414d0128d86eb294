"""scratch: worst deviations of the 4096-env launch test (tests/test_gpu_product_paths.py) for the library in UHC_B200_SO"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from oracle import oracle as O
from uhc_b200.engine import Engine
KEYS = ("qpos", "qvel", "wbpos", "wbquat", "bquat", "bangvel", "ee_wpos", "com")
g = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
z = np.load(os.path.join(g, "expert_sway.npz")); ex = {k: z[k] for k in z.files}
so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
for seed in (21, 22, 23):
    E, T = 4096, 10
    rng = np.random.RandomState(seed)
    starts = rng.randint(0, 70, E).astype(np.int32)
    eng = Engine(E); eng.load_clips([ex], [so])
    obs = eng.reset(start=starts).cpu().numpy().copy()
    ids = np.unique(np.concatenate([rng.choice(E, 60, replace=False), [0, 6, 7, E - 1]])).astype(np.int32)
    om = O.Model(); envs = []
    for e in ids:
        oe = O.Env(om, {k: ex[k][starts[e]:] for k in KEYS}, so); oe.reset(); envs.append(oe)
    alive = np.ones(len(ids), bool); wq = []
    for t in range(T):
        a = rng.normal(0, 0.1, (E, 105)).astype(np.float32); a[:, 69:75] *= 0.3
        eng.step(torch.tensor(a, device="cuda"))
        st = eng.get_states(ids)
        for i, e in enumerate(ids):
            if not alive[i]: continue
            oo, ro, done, info = envs[i].step(a[e].astype(np.float64))
            wq.append(np.abs(st["qpos"][i] - envs[i].d.qpos).max())
            alive[i] = not done
    wq = np.array(wq)
    print(os.path.basename(os.environ.get("UHC_B200_SO", "default")), "seed", seed, "worst_q %.3e  p99 %.3e  median %.3e" % (wq.max(), np.quantile(wq, 0.99), np.median(wq)))
    eng.close()
