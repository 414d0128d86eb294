#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched humanoid-imitation rollout (BASELINE.json metric) on N B200s.

A "step" is one lock-step control step of every environment: observation normaliser -> policy MLP forward (tcgen05) ->
Gaussian sample -> fused physics(15 substeps)+task kernel -> transition written to the HBM rollout buffer -> re-seeding of
finished episodes.  Workload at N=1 = BASELINE.json configs[1]: 4096 SMPL-neutral humanoids imitating one AMASS-shaped clip,
policy rollout only.  N>1: weak scaling, 4096 envs per GPU, no data-path collective (the rollout has none).

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference          # the reference's CPU path: oracle port (fp64 C) on all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ENVS_PER_GPU = 4096
CLIP_FRAMES = 320                      # length of '0-ACCAD_Male2General_c3d_A2- Sway_poses', the clip configs[1] names
BYTES_PER_ENV_STEP = 6396              # algorithmic HBM bytes of the fused physics/task step, SURVEY.md section 8(d)
METRIC, UNIT = "env_steps_per_s", "env-steps/s"


def make_clip(seed=1):
    from uhc_b200 import motion_lib
    ex = motion_lib.synthetic_clip(CLIP_FRAMES, np.random.default_rng(seed))
    shape = np.zeros(17)
    return ex, shape


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    seed, budget_s = args
    from oracle import oracle as O
    ex, shape = make_clip()
    rng = np.random.RandomState(seed)
    env = O.Env(O.Model(), ex, shape)
    env.reset()
    n, t0 = 0, time.time()
    while time.time() - t0 < budget_s:
        a = rng.normal(0, 0.1, 105)
        a[69:75] *= 0.3
        _, _, done, _ = env.step(a)
        n += 1
        if done:
            s = rng.randint(0, CLIP_FRAMES - 5)
            sl = {k: np.asarray(ex[k])[s:] for k in ("qpos", "qvel", "wbpos", "wbquat", "bquat", "bangvel", "ee_wpos", "com")}
            env.load_expert(sl, shape)
            env.reset()
    return n, time.time() - t0


def usable_cores():
    """host threads this process may actually use: min(os.cpu_count, scheduler affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read().split()[0])
                    n = min(n, max(1, q // period))
        except Exception:
            pass
    return max(1, n)


def cpu_baseline(budget_s=12.0, cores=None):
    """The reference's per-env CPU path restated (oracle/uhc_oracle.c: PD + 15 substeps + obs + reward), one process per
    host core, bounded sample."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    cores = cores or usable_cores()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i + 1, budget_s) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return dict(value=steps / wall, unit=UNIT, cores=cores, kind="port",
                sample=f"{steps} env-steps of the same workload (noise actions, re-seeded on termination) in {wall:.1f} s on {cores} processes; "
                       "fp64 C restatement of the MuJoCo+Python path, NOT MuJoCo itself")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    per = max(2.0, min(20.0, 6.0 * (args.steps + args.warmup) / 23.0))
    cb = cpu_baseline(budget_s=per)
    line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * ENVS_PER_GPU * args.gpus / cb["value"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=workload_config(args.gpus), cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), wall_s=time.time() - t0)
    print(json.dumps(line))


def workload_config(n):
    return {"workload": f"{ENVS_PER_GPU} SMPL-neutral humanoids per GPU imitating one {CLIP_FRAMES}-frame AMASS-shaped clip, policy rollout only "
                        "(obs normaliser + 657-2048-1024-512-105 gelu policy + physics/task step), uhc_implicit_shape hyper-parameters",
            "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n, "parallelism": f"env-sharded x{n}, no data-path collective",
            "l2": "flushed (256 MiB write) between timed steps", "policy_gemm": "tcgen05 bf16 operands, fp32 accumulate",
            "physics": "fp32, 15 substeps/step, primal Newton contact solve"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms from before the warm-up; stop(t0, t1) keeps the samples whose
    timestamp falls inside the timed region [t0, t1] (wall clock), or the nearest ones when the region is shorter than a sample."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        self.p.wait()
        self.f.seek(0)
        import datetime
        rows = []
        for line in self.f.read().strip().splitlines():
            r = [x.strip() for x in line.split(",")]
            if len(r) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(r[1]), float(r[2]), float(r[3]), r[4:8]))
            except Exception:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        sel = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        where = "inside the timed region"
        if not sel:
            mid = 0.5 * ((t0 or rows[-1][0]) + (t1 or rows[-1][0]))
            sel = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            where = "nearest to the timed region (region shorter than the sampling period)"
        reasons = set()
        for r in sel:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in sel])), "sm_max_mhz": sel[0][2], "power_w_max": max(r[3] for r in sel),
                "samples": len(sel), "where": where, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from uhc_b200.agent import BatchedAgent, RolloutBuffer
    ex, shape = make_clip()
    E, K, W = ENVS_PER_GPU, args.steps, max(args.warmup, 3)
    agent = BatchedAgent(E, [ex], [shape], device=local, seed=1, rank=rank, world=world)
    agent.reset_envs()
    buf = RolloutBuffer(max(K, 4), E, agent.dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=agent.dev)
    clocks = ClockSampler(local) if rank == 0 else None
    for k in range(W):
        agent.step_once(buf, k % buf.T)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    # per-kernel timing of the dominant kernel: wrap engine.step
    raw_step = agent.engine.step
    cur = [0]

    def timed_step(a, torque_out=None, reward_out=None):
        kev[cur[0]][0].record()
        out = raw_step(a, torque_out, reward_out)
        kev[cur[0]][1].record()
        return out
    agent.engine.step = timed_step
    l0, n0 = agent.engine.kernel_launches, agent.nn_launches
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for k in range(K):
        flush.fill_(k & 0xFF)                       # L2 flush, outside the timed region
        cur[0] = k
        ev[k][0].record()
        agent.step_once(buf, k % buf.T)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    agent.engine.step = raw_step
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    launches = (agent.engine.kernel_launches - l0) + (agent.nn_launches - n0)
    tt = torch.tensor([total_ms], device=agent.dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = E * world * K / (total_ms * 1e-3)

    # end to end through the host-buffer API: obs (pinned host) -> device policy -> actions to host -> uhc_env_step_host -> obs/reward to host
    Ke = max(3, min(K, 20))
    obs_h = torch.empty(E, 657, dtype=torch.float32).pin_memory()
    act_h = torch.empty(E, 105, dtype=torch.float32).pin_memory()
    rew_h, pct_h = np.empty(E, np.float32), np.empty(E, np.float32)
    ci_h, fail_h, end_h = np.empty((E, 5), np.float32), np.empty(E, np.int32), np.empty(E, np.int32)
    obs_h.copy_(agent.obs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    def e2e_step():
        od = obs_h.to(agent.dev, non_blocking=True)
        _, a, _ = agent.policy_step(od, True, None, True)
        act_h.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        agent.engine.step_host(act_h.numpy(), obs_h.numpy(), rew_h, ci_h, fail_h, end_h, pct_h)
        done = np.nonzero(fail_h | end_h)[0]
        if len(done) and not agent.auto_reset:
            agent.reset_envs(done.astype(np.int32))
            torch.cuda.synchronize()
            obs_h[torch.as_tensor(done)] = agent.obs[torch.as_tensor(done, device=agent.dev)].cpu()
    for _ in range(min(W, 3)):      # untimed warm-up of the host path (staging buffers, page faults)
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=agent.dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = E * world * Ke / float(te.item())
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_PER_ENV_STEP * E / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "env_step_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    cb = cpu_baseline(budget_s=10.0) if world == 1 else None
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=total_ms / K, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=workload_config(world),
                roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                              kernel="k_env_step<float,7>", kernel_ms=kern_ms, kernel_share_of_step=kern_ms / (total_ms / K),
                              peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                              note="algorithmic bytes 6396 B/env-step (SURVEY 8d); the step is latency/ALU bound, not HBM bound -- see DESIGN.md"),
                e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=E * (657 + 105) * 4, d2h_bytes_per_step=E * (105 + 657 + 1 + 5 + 1 + 1 + 1) * 4, steps=Ke),
                gpu_launches=launches, clocks=clk)
    if cb:
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
