import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200.engine import Engine
G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
tag, act = sys.argv[1], sys.argv[2]
g = np.load(f"{G}/env_{tag}_{act}.npz"); z = np.load(f"{G}/expert_{tag}.npz"); ex = {k: z[k] for k in z.files}
so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
eng = Engine(1); eng.load_clips([ex], [so]); eng.reset()
for t in range(len(g["reward"])):
    o, r, ci, f, e, p = eng.step(torch.tensor(g["action"][t][None], dtype=torch.float32, device="cuda"))
    st = eng.get_state(0); ob = o.cpu().numpy()[0]
    d = np.abs(ob - g["obs"][t]); i = int(d.argmax())
    print(t, "qpos %.2e qvel %.2e obs %.2e @%d (gold %.3f) fail %d/%d newton %d ncon %d" % (np.abs(st["qpos"] - g["qpos"][t]).max(), np.abs(st["qvel"] - g["qvel"][t]).max(), d.max(), i, g["obs"][t][i], int(f[0]), g["fail"][t], st["newton_iters"], st["ncon"]))
    if t > 34: break
