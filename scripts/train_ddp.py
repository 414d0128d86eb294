"""Config 4 of BASELINE.json in miniature: env-sharded PPO over N GPUs (one process per GPU, torchrun), mixed synthetic body shapes,
one NCCL all-reduce of the flat gradients per optimisation step.  Checks that every rank holds identical weights afterwards.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/train_ddp.py [envs_per_gpu] [T] [iters] [variant]
variant: default (uhc_implicit_shape shape) | explicit (315-wide actions) | mcp (PolicyMCP actor, obs v1, relu, no meta-PD, reactive starts)
"""
import os, sys, json, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from uhc_b200 import motion_lib
from uhc_b200.agent import BatchedAgent, make_nccl_grad_sync
from uhc_b200.model import HumanoidModel, NB

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(7)                                  # the same clips / shapes on every rank, different env seeds
clips = [motion_lib.synthetic_clip(int(rng.integers(80, 200)), rng) for _ in range(8)]
base = HumanoidModel()
variants = [base] + [HumanoidModel(scale=np.full(NB, s)) for s in (0.92, 1.08)]
clip_models = [i % 3 for i in range(len(clips))]
shapes = [np.concatenate([np.full(16, 0.1 * m), [0.0]]) for m in clip_models]
variant = sys.argv[4] if len(sys.argv) > 4 else "default"
extra = {"default": {}, "explicit": dict(rfc_mode="explicit"),
         "mcp": dict(actor_type="mcp", obs_v=1, meta_pd=0, htype="relu", policy_hsize=(512, 256), value_hsize=(512, 256), reactive_v=1)}[variant]
ag = BatchedAgent(E, clips, shapes, device=local, seed=1, rank=rank, world=world, model=base, variants=variants, clip_models=clip_models,
                  grad_sync=make_nccl_grad_sync(world) if world > 1 else None, **extra)
ag.optimize_policy(T)
torch.cuda.synchronize()
t0 = time.time()
for i in range(iters):
    log = ag.optimize_policy(T)
torch.cuda.synchronize()
dt = (time.time() - t0) / iters
w = torch.cat([ag.policy.flat, ag.value.flat])
chk = torch.stack([w.double().sum(), w.double().abs().sum(), ag.running_state.stats.sum(), ag.running_state.stats.abs().sum()])     # weights AND the observation normaliser
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.allclose(lo, hi, rtol=0, atol=0))
else:
    same = True
if rank == 0:
    print(json.dumps({"world": world, "envs_per_gpu": E, "T": T, "iter_s": dt, "env_steps_per_s_full_loop": world * E * T / dt,
                      "variant": variant, "weights_and_running_state_identical_across_ranks": same, "avg_reward": log["avg_reward"], "value_loss": log["value_loss"]}))
if world > 1:
    dist.destroy_process_group()
