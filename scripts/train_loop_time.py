"""Config 3 of BASELINE.json: 4096 envs, full PPO loop (rollout + GAE + 10-epoch update), synthetic AMASS-shaped clips."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import motion_lib
from uhc_b200.agent import BatchedAgent
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(1)
clips = [motion_lib.synthetic_clip(int(rng.integers(69, 376)), rng) for _ in range(10)]   # take5_test_small: 10 clips, 69-375 frames
ag = BatchedAgent(E, clips, None, seed=1)
ag.optimize_policy(T)
out = []
for i in range(iters):
    log = ag.optimize_policy(T)
    out.append(log)
    print(f"iter {i}: sample {log['sample_time']:.3f}s ({log['num_steps']/log['sample_time']:.0f} env-steps/s) update {log['update_time']:.3f}s "
          f"avg_r {log['avg_reward']:.4f} eps_len {log['avg_episode_len']:.1f} fail {log['fail_rate']:.2f} vloss {log['value_loss']:.4f}")
tot = np.mean([l['sample_time'] + l['update_time'] for l in out])
print(json.dumps({"config": "4096 envs full PPO loop, T=%d (N=%d), 10 epochs" % (T, T * E), "iter_s": tot, "env_steps_per_s_full_loop": T * E / tot}))
