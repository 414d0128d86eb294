"""Batched rollout + PPO driver on the engine.

Replaces the reference's fork-per-iteration CPU workers (uhc/agents/agent_copycat.py:496-605, khrylib/rl/agents/agent.py:42-126)
by E device-resident environments stepped in lock-step: per control step one obs-normaliser pass, one policy forward
(tensor cores), one Gaussian sample, one fused physics/task kernel, and the transition lands in a time-major [T][E] buffer
in HBM.  Episode semantics follow Appendix C of SURVEY.md: an env that fails or reaches the end of its clip slice gets
mask 0 and is re-seeded with a freshly sampled (clip, start) slice (dataset_amass_single.py:172-253); rollouts cut mid-episode
are bootstrapped with V(s_T) (the reference never truncates, so this is the one semantic addition, documented in DESIGN.md).
Multi-GPU: envs are sharded, one process per GPU; `grad_sync` all-reduces the flat gradients once per optimisation step.
"""
import time

import numpy as np

from . import nn
from .engine import ACT_DIM, OBS_DIM, Engine


def ewma(x, alpha=0.05):
    """uhc/utils/math_utils.py:25-29"""
    avg = float(x[0])
    for i in x[1:]:
        avg = alpha * float(i) + (1 - alpha) * avg
    return avg


def failure_weights(success_hist, sampling_temp=0.2, sampling_freq=0.5):
    """Per-clip sampling weights of the training loop (DatasetAMASSSingle.sample_seq with freq_dict, dataset_amass_single.py:183-186):
    with probability sampling_freq the clip is drawn from p ~ exp(-ewma(success history) / temp) (0 for an empty history), else
    uniformly over the clips.  success_hist: one list of 0/1 outcomes per clip (freq_dict[k][:, 0] == 1).  Returns the mixture
    weights w = sampling_freq * p + (1 - sampling_freq) / C for uhc_set_clip_weights."""
    s = np.array([ewma(np.asarray(h, dtype=np.float64) == 1) if len(h) > 0 else 0.0 for h in success_hist])
    p = np.exp(-s / sampling_temp)
    p = p / p.sum()
    return (sampling_freq * p + (1.0 - sampling_freq) / len(p)).astype(np.float32)


class ClipSampler:
    """DatasetAMASSSingle.sample_seq / get_sample_from_key (dataset_amass_single.py:172-253): uniform clip choice (or
    failure-weighted when freq stats are given), start ~ U[0, len - t_min), slice length min(t_max, len - start)."""

    def __init__(self, clip_lens, t_min=5, t_max=300, seed=0):
        self.lens, self.t_min, self.t_max = np.asarray(clip_lens), t_min, t_max
        self.rng = np.random.RandomState(seed)
        self.probs = None

    def set_failure_weights(self, success_ewma, temp=0.1):
        p = np.exp(-np.asarray(success_ewma) / temp)
        self.probs = p / p.sum()

    def sample(self, n, freq=0.5):
        if self.probs is not None:
            pick_w = self.rng.binomial(1, freq, n).astype(bool)
            clip = np.where(pick_w, self.rng.choice(len(self.lens), n, p=self.probs), self.rng.randint(0, len(self.lens), n))
        else:
            clip = self.rng.randint(0, len(self.lens), n)
        L = self.lens[clip]
        start = (self.rng.random_sample(n) * np.maximum(L - self.t_min, 1)).astype(np.int64)
        length = np.minimum(self.t_max if self.t_max > 0 else L, L - start)
        return clip.astype(np.int32), start.astype(np.int32), length.astype(np.int32)


class RolloutBuffer:
    def __init__(self, T, E, device):
        import torch
        f = dict(device=device, dtype=torch.float32)
        self.T, self.E = T, E
        self.states = torch.empty(T, E, OBS_DIM, **f)
        self.actions = torch.empty(T, E, ACT_DIM, **f)
        self.rewards, self.masks, self.exps, self.logp = (torch.empty(T, E, **f) for _ in range(4))
        self.last_obs = torch.empty(E, OBS_DIM, **f)
        self.last_alive = torch.empty(E, **f)
        self.fails = torch.zeros(T, E, device=device, dtype=torch.int32)

    # TrajBatch-compatible flat views (khrylib/rl/core/trajbatch.py)
    def flat(self, name):
        t = getattr(self, name)
        return t.reshape(self.T * self.E, *t.shape[2:])


class BatchedAgent:
    def __init__(self, num_envs, clips, shapes=None, device=0, seed=1, precision=32, policy_hsize=(2048, 1024, 512),
                 value_hsize=(2048, 1024, 512), htype="gelu", log_std=-2.3, policy_lr=5e-5, value_lr=3e-4, gamma=0.95, tau=0.95,
                 clip_epsilon=0.2, num_optim_epoch=10, grad_clip=40.0, t_min=5, t_max=300, noise_rate=1.0, rank=0, world=1,
                 grad_sync=None, model=None, update_tc=True, variants=None, clip_models=None, **env_cfg):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.E, self.seed, self.rank, self.world = num_envs, seed, rank, world
        self.auto_reset = bool(env_cfg.pop("auto_reset", True))
        self.engine = Engine(num_envs, model=model, device=device, precision=precision, variants=variants, auto_reset=int(self.auto_reset), t_min=t_min, t_max=t_max,
                             reset_seed=seed * 7919 + rank * 104729 + 1, **env_cfg)
        self.engine.load_clips(clips, shapes, clip_models)   # clip_models: body-shape variant per clip (the reference rebuilds the robot per clip)
        self.sampler = ClipSampler(self.engine.clip_len, t_min, t_max, seed=seed * 9973 + rank)
        self.policy = nn.MLPNet(OBS_DIM, policy_hsize, ACT_DIM, htype, device=self.dev, head_name="action_mean", seed=seed)
        self.value = nn.MLPNet(OBS_DIM, value_hsize, 1, htype, device=self.dev, head_name="value_head", seed=seed + 1)
        self.log_std = torch.full((ACT_DIM,), float(log_std), device=self.dev, dtype=torch.float32)
        self.running_state = nn.ZFilter(OBS_DIM, clip=5.0, device=self.dev)
        self.opt_p, self.opt_v = nn.Adam(self.policy.params(), policy_lr), nn.Adam(self.value.params(), value_lr)
        self.gamma, self.tau, self.clip_epsilon, self.epochs, self.grad_clip = gamma, tau, clip_epsilon, num_optim_epoch, grad_clip
        self.noise_rate, self.grad_sync, self.update_tc = noise_rate, grad_sync, update_tc
        self.global_step = 0
        self.obs = None
        self.ep_len = torch.zeros(num_envs, device=self.dev)
        self.ep_ret = torch.zeros(num_envs, device=self.dev)
        self.nn_launches = 0
        self._notdone = torch.zeros(num_envs, device=self.dev, dtype=torch.bool)

    # ---- env.reset for a subset with freshly sampled clip slices
    def reset_envs(self, ids=None):
        ids = np.arange(self.E, dtype=np.int32) if ids is None else np.asarray(ids, dtype=np.int32)
        clip, start, length = self.sampler.sample(len(ids))
        self.obs = self.engine.reset(ids, clip, start, length)
        return self.obs

    def policy_step(self, obs, update_filter=True, mean_action=None, use_tc=True, out_state=None, out_action=None, out_logp=None):
        """running_state -> policy -> sample.  Returns (normalised state, action, logp); the out_* tensors (rows of the rollout
        buffer) are written in place by the kernels."""
        s = self.running_state(obs, update=update_filter, out=out_state)
        mean = self.policy.forward_tc(s) if use_tc else self.policy.forward(s)
        a, lp = nn.gaussian_sample(mean, self.log_std, self.seed * 1000003 + self.rank, self.global_step, mean_action, out_action, out_logp)
        self.nn_launches += (3 if update_filter else 1) + (1 + len(self.policy.W) if use_tc else len(self.policy.W)) + 1
        return s, a, lp

    def step_once(self, buf, k, use_tc=True):
        """one lock-step control step of every env: normalise, policy, sample, physics+task kernel, buffer write, re-seed ended episodes."""
        t = self.torch
        mean_action = None
        if self.noise_rate < 1.0:
            mean_action = (t.rand(self.E, device=self.dev) < (1.0 - self.noise_rate)).to(t.uint8)
        s, a, lp = self.policy_step(self.obs, True, mean_action, use_tc, buf.states[k], buf.actions[k], buf.logp[k])
        if mean_action is not None:
            buf.exps[k].copy_(1.0 - mean_action.float())
        elif not getattr(buf, "_exps_ones", False):
            buf.exps.fill_(1.0)
            buf._exps_ones = True
        obs, rew, cinfo, fail, end, pct = self.engine.step(a, reward_out=buf.rewards[k])
        done = (fail | end) != 0
        t.logical_not(done, out=self._notdone)
        buf.masks[k].copy_(self._notdone)
        self.global_step += 1
        buf.fails[k].copy_(fail)
        if self.auto_reset:
            return            # finished episodes were re-seeded inside the step kernel (no host round trip)
        ids = done.nonzero().flatten()
        if ids.numel():
            self.reset_envs(ids.cpu().numpy().astype(np.int32))

    def sample(self, T, buf=None, use_tc=True):
        """agent.sample(): T lock-step control steps of all envs.  Returns (buffer, log)."""
        t = self.torch
        if self.obs is None:
            self.reset_envs()
        buf = buf or RolloutBuffer(T, self.E, self.dev)
        t0 = time.time()
        len0, ret0 = self.ep_len.clone(), self.ep_ret.clone()
        for k in range(T):
            self.step_once(buf, k, use_tc)
        buf.last_obs.copy_(self.obs)
        # episode statistics from the buffer (one sync at the end of the rollout): segment the [T][E] masks per env
        m, r = buf.masks[:T], buf.rewards[:T]
        done = m == 0
        n_eps = int(done.sum())
        run_len = t.zeros(self.E, device=self.dev); run_ret = t.zeros(self.E, device=self.dev)
        run_len += len0; run_ret += ret0
        tot_len = t.zeros((), device=self.dev); tot_ret = t.zeros((), device=self.dev)
        for k in range(T):
            run_len += 1; run_ret += r[k]
            d = done[k]
            tot_len += (run_len * d).sum(); tot_ret += (run_ret * d).sum()
            run_len *= ~d; run_ret *= ~d
        self.ep_len, self.ep_ret = run_len, run_ret
        n_fail = int(((buf.fails[:T] != 0) & done).sum())
        log = dict(num_steps=T * self.E, num_episodes=n_eps, avg_episode_len=float(tot_len) / max(n_eps, 1),
                   avg_episode_reward=float(tot_ret) / max(n_eps, 1), fail_rate=n_fail / max(n_eps, 1),
                   avg_reward=float(r.mean()), sample_time=time.time() - t0)
        return buf, log

    def update_params(self, buf):
        """AgentPG.update_params (agent_pg.py:39-56): V(s), GAE (+bootstrap), advantage normalisation, PPO epochs."""
        t = self.torch
        t0 = time.time()
        T, E = buf.T, buf.E
        states = buf.flat("states")
        values = self.value.forward(states).reshape(T, E)
        last_v = self.value.forward(self.running_state(buf.last_obs, update=False)).reshape(E)
        adv, ret = nn.gae(buf.rewards, buf.masks, values, last_v, self.gamma, self.tau, normalize=True)
        if self.grad_sync is not None:
            self._wrap_sync()
        losses = nn.ppo_update(self.policy, self.value, self.log_std, self.opt_p, self.opt_v, states, buf.flat("actions"), ret.reshape(-1),
                               adv.reshape(-1), buf.flat("exps"), self.clip_epsilon, self.epochs, self.grad_clip, use_tc=self.update_tc)
        t.cuda.synchronize()
        return dict(update_time=time.time() - t0, surr_loss=float(losses[0]), value_loss=float(losses[1]))

    def _wrap_sync(self):
        """one all-reduce of the flat gradient buffer per optimisation step."""
        if getattr(self, "_sync_wrapped", False):
            return
        sync = self.grad_sync
        for opt in (self.opt_p, self.opt_v):
            raw = opt.step

            def stepper(grads, max_norm=None, _raw=raw):
                _raw(sync(grads), max_norm=max_norm)
            opt.step = stepper
        self._sync_wrapped = True

    def optimize_policy(self, T):
        buf, log = self.sample(T)
        log.update(self.update_params(buf))
        return log

    # checkpoint in the reference's wire format (agent_copycat.py:190-201): policy_dict / value_dict / running_state
    def state_dicts(self):
        pd = self.policy.state_dict()
        pd["action_log_std"] = self.log_std.detach().cpu().reshape(1, -1)
        return {"policy_dict": pd, "value_dict": self.value.state_dict(),
                "running_state": {"n": self.running_state.n, "mean": self.running_state.mean, "std": self.running_state.std, "clip": 5.0}}

    def load_state_dicts(self, cp):
        t = self.torch
        self.policy.load_state_dict(cp["policy_dict"])
        self.value.load_state_dict(cp["value_dict"])
        if "action_log_std" in cp["policy_dict"]:
            self.log_std.copy_(t.as_tensor(np.asarray(cp["policy_dict"]["action_log_std"]), dtype=t.float32).reshape(-1))
        rs = cp.get("running_state")
        if rs is not None:
            if isinstance(rs, dict):
                self.running_state.load(rs["n"], rs["mean"], rs["std"])
            else:  # a pickled khrylib ZFilter
                self.running_state.load(rs.rs.n, rs.rs.mean, rs.rs.std)


def make_nccl_grad_sync(world):
    """flatten -> one torch.distributed all_reduce(sum) -> unflatten, averaged over ranks (full-batch mean semantics)."""
    import torch
    import torch.distributed as dist

    def sync(grads):
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        out, o = [], 0
        for g in grads:
            n = g.numel()
            out.append(flat[o:o + n].view_as(g))
            o += n
        return out
    return sync
