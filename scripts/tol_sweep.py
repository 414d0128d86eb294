"""Scratch: Newton tolerance vs parity (golden traces) and throughput."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200.engine import Engine
G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
def load(tag):
    z = np.load(f"{G}/expert_{tag}.npz"); ex = {k: z[k] for k in z.files}
    return ex, np.concatenate([ex["beta"][0], [ex["gender"][0]]])
for tol in [float(x) for x in sys.argv[1:]] or [1e-5, 1e-4, 1e-3]:
    line = f"tol {tol:.0e}:"
    for tag, act in (("sway", "zero"), ("sway", "noise"), ("kick", "noise")):
        g = np.load(f"{G}/env_{tag}_{act}.npz"); ex, so = load(tag)
        eng = Engine(1, newton_tol=tol); eng.load_clips([ex], [so]); eng.reset()
        err, its, T = 0.0, [], len(g["reward"])
        term = np.nonzero(g["fail"] | g["done"] if "done" in g.files else g["fail"])[0]
        stop = (term[0] if len(term) else T)
        for t in range(min(T, max(stop, 1))):
            eng.step(torch.tensor(g["action"][t][None], dtype=torch.float32, device="cuda"))
            st = eng.get_state(0); its.append(st["newton_iters"])
            err = max(err, np.abs(st["qpos"] - g["qpos"][t]).max())
        line += f"  {tag}_{act}[{stop}] qpos err {err:.1e} iters/step {np.mean(its):.1f};"
        eng.close()
    E = 4096
    ex, so = load("sway")
    eng = Engine(E, newton_tol=tol); eng.load_clips([ex], [so])
    rng = np.random.RandomState(1)
    eng.reset(start=rng.randint(0, 40, E).astype(np.int32))
    acts = torch.tensor(rng.normal(0, 0.1, (E, 105)), dtype=torch.float32, device="cuda")
    for _ in range(3): eng.step(acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.step(acts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(line + f"  E=4096 {E / ms * 1e3 / 1e3:.0f}k env-steps/s", flush=True)
    eng.close()
