"""AgentCopycat with the reference's surface (uhc/agents/agent_copycat.py:51-605) on the batched B200 engine.

  AgentCopycat(cfg, dtype, device, training=True, checkpoint_epoch=0)
  .optimize_policy(epoch)   per_epoch_update -> sample -> update_params -> checkpoint / eval every save_n_epochs -> log   (:326-352)
  .sample(min_batch_size)   -> (batch, log)      batch: states / actions / masks / rewards / exps (device resident, numpy on demand)
  .update_params(batch)     GAE + PPO epochs (khrylib agent_pg.py:39-56, agent_ppo.py:16-51)
  .eval_policy(epoch, dump) deterministic roll-out of every clip, coverage / error statistics                               (:354-494)
  .save_checkpoint / .load_checkpoint   pickle {"policy_dict", "value_dict", "running_state"} at models/iter_%04d.p        (:190-260)

Differences that are inherent to the batched design are listed in DESIGN.md (lock-step horizon with value bootstrap, batched
ZFilter merge, clip sampling without the per-clip failure history unless eval statistics are available).
"""
import logging
import math
import os
import os.path as osp
import pickle
import time
from collections import defaultdict

import joblib
import numpy as np

from uhc.data_loaders.dataset_amass_single import DatasetAMASSSingle
from uhc.envs.humanoid_im import HumanoidEnv
from uhc.losses.reward_function import reward_func
from uhc_b200 import nn
from uhc_b200.agent import BatchedAgent, RolloutBuffer, make_nccl_grad_sync
from uhc_b200.model import HumanoidModel


class _Batch:
    """TrajBatch-compatible view (khrylib/rl/core/trajbatch.py:4-15) over the device rollout buffer."""

    def __init__(self, buf):
        self.buf = buf

    def __getattr__(self, name):
        if name in ("states", "actions", "rewards", "masks", "exps"):
            return self.buf.flat(name).cpu().numpy()
        raise AttributeError(name)


class _Log(dict):
    __getattr__ = dict.get


class AgentCopycat:
    def __init__(self, cfg, dtype, device, training=True, checkpoint_epoch=0):
        import torch
        self.cfg = self.cc_cfg = cfg
        self.dtype, self.device, self.training = dtype, device, training
        self.epoch, self.max_freq = 0, 50
        dev_index = device.index if getattr(device, "type", "cpu") == "cuda" and device.index is not None else int(getattr(cfg, "gpu_index", 0) or 0)
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 engine needs a CUDA device (the reference's CPU sampling path is replaced, not kept as a fallback)")
        self.model_tables = HumanoidModel()
        # data (setup_data_loader :128-134)
        self.data_loader = DatasetAMASSSingle(cfg.data_specs, data_mode="train", model=self.model_tables)
        self.test_data_loaders = [self.data_loader]
        if len(cfg.data_specs.get("test_file_path", [])) > 0:
            self.test_data_loaders.append(DatasetAMASSSingle(cfg.data_specs, data_mode="test", model=self.model_tables))
        self.freq_dict = {k: [] for k in self.data_loader.data_keys}
        rw = cfg.reward_weights or {}
        w = [rw.get(k, d) for k, d in (("w_p", 0.6), ("w_v", 0.1), ("w_e", 0.2), ("w_c", 0.1), ("w_vf", 0.0))]
        kk = [rw.get(k, d) for k, d in (("k_p", 2), ("k_v", 0.005), ("k_e", 20), ("k_c", 1000), ("k_vf", 1))]
        self.num_envs = int(cfg.get("num_envs", 4096))
        self.horizon = max(2, int(math.ceil(cfg.min_batch_size / self.num_envs)))
        world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        sync = None
        if world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("nccl")
            sync = make_nccl_grad_sync(world)
        assert cfg.obs_v == 2 and cfg.actor_type == "gauss" and cfg.reward_id in reward_func, \
            "the B200 engine implements obs_v 2 / gauss actor / world_rfc_implicit (SURVEY.md section 8f lists the other variants as next)"
        self.agent = BatchedAgent(
            self.num_envs, self.data_loader.experts, self.data_loader.shapes, device=dev_index, seed=cfg.seed, policy_hsize=cfg.policy_hsize,
            value_hsize=cfg.value_hsize, htype=cfg.policy_htype, log_std=cfg.log_std, policy_lr=cfg.policy_lr, value_lr=cfg.value_lr,
            gamma=cfg.gamma, tau=cfg.tau, clip_epsilon=cfg.clip_epsilon, num_optim_epoch=cfg.num_optim_epoch, grad_clip=40.0,
            t_min=cfg.data_specs.get("t_min", 90), t_max=cfg.data_specs.get("t_max", -1), rank=rank, world=world, grad_sync=sync,
            model=self.model_tables, base_rot=cfg.data_specs.get("base_rot", [0.7071, 0.7071, 0.0, 0.0]), rfc_scale=cfg.residual_force_scale,
            rfc_lim=cfg.residual_force_lim, rfc_rate=0.0 if cfg.rfc_decay else 1.0, body_diff_thresh=cfg.get("body_diff_thresh", 0.5),
            meta_pd=int(cfg.meta_pd), env_episode_len=cfg.env_episode_len, trail_steps=cfg.env_expert_trail_steps, w=w, k=kk)
        self.policy_net, self.value_net, self.running_state = self.agent.policy, self.agent.value, self.agent.running_state
        self.state_dim, self.action_dim = 657, 105
        self.expert_reward = reward_func[cfg.reward_id]
        self.env = None   # single-env facade, built lazily (eval_seq / visualisation code paths)
        self.logger = logging.getLogger(f"uhc_b200.{cfg.id}")
        if not self.logger.handlers:
            logging.basicConfig(level=logging.INFO, format="%(message)s")
        if checkpoint_epoch > 0:
            self.load_checkpoint(checkpoint_epoch)
            self.epoch = checkpoint_epoch

    # ---------------------------------------------------------------- schedules (:279-297)
    def per_epoch_update(self, epoch):
        cfg = self.cfg
        cfg.update_adaptive_params(epoch)
        self.agent.noise_rate = cfg.adp_noise_rate
        self.agent.opt_p.lr = cfg.adp_policy_lr
        if cfg.rfc_decay:
            rate = float(np.clip(1 - epoch / cfg.get("rfc_decay_max", 10000), 0, 1))
            self.agent.engine.set_cfg(rfc_rate=rate)
        if cfg.fix_std:
            self.agent.log_std.fill_(float(cfg.adp_log_std))

    # ---------------------------------------------------------------- sampling / update
    def sample(self, min_batch_size=None):
        T = self.horizon if min_batch_size is None else max(2, int(math.ceil(min_batch_size / self.num_envs)))
        buf, log = self.agent.sample(T)
        log = _Log(log)
        log.update(avg_c_reward=log["avg_reward"], avg_episode_c_reward=log["avg_episode_reward"])
        return _Batch(buf), log

    def update_params(self, batch):
        return self.agent.update_params(batch.buf)["update_time"]

    def optimize_policy(self, epoch, save_model=True):
        cfg = self.cfg
        self.epoch = epoch
        t0 = time.time()
        self.per_epoch_update(epoch)
        batch, log = self.sample(cfg.min_batch_size)
        t1 = time.time()
        self.update_params(batch)
        t2 = time.time()
        info = {"log": log, "T_sample": t1 - t0, "T_update": t2 - t1, "T_total": t2 - t0}
        if save_model and (self.epoch + 1) % cfg.save_n_epochs == 0:
            self.save_checkpoint(epoch)
            info["log_eval"] = self.eval_policy(epoch)
        self.log_train(info)
        return info

    def log_train(self, info):
        log = info["log"]
        self.logger.info(f"{self.cfg.id} | {self.epoch:4d} | T_s {info['T_sample']:.2f} T_u {info['T_update']:.2f} | steps {log['num_steps']} "
                         f"({log['num_steps'] / max(info['T_sample'], 1e-9):.0f}/s) | eps_len {log['avg_episode_len']:.1f} | avg_r {log['avg_reward']:.4f} "
                         f"| eps_r {log['avg_episode_reward']:.2f} | fail {log['fail_rate']:.2f}")
        if not getattr(self.cfg, "no_log", True):
            try:
                import wandb
                wandb.log({"rewards": log["avg_reward"], "eps_len": log["avg_episode_len"], "avg_rwd": log["avg_episode_reward"]}, step=self.epoch)
                if "log_eval" in info:
                    [wandb.log(t, step=self.epoch) for t in info["log_eval"]]
            except Exception:
                pass

    # ---------------------------------------------------------------- evaluation (:354-494), batched: one env per clip
    def eval_policy(self, epoch=0, dump=False):
        import torch
        cfg = self.cfg
        res_dicts = []
        for loader in self.test_data_loaders:
            ag = self.agent if loader is self.data_loader else None
            eng = self.agent.engine
            n = loader.get_len()
            if loader is not self.data_loader:
                eng.load_clips(loader.experts, loader.shapes)
            E = self.num_envs
            res = {}
            for c0 in range(0, n, E):
                ids = np.arange(min(E, n - c0), dtype=np.int32)
                clips = (c0 + ids).astype(np.int32)
                eng.set_cfg(**self._env_cfg(test=True))
                obs = eng.reset(ids, clips, 0, None)
                lens = eng.clip_len[clips]
                alive = np.ones(len(ids), bool); fail_any = np.zeros(len(ids), bool)
                rsum, jerr, rdist, cnt = (np.zeros(len(ids)) for _ in range(4))
                det = torch.ones(E, dtype=torch.uint8, device=obs.device)
                for t in range(int(lens.max()) - 1):
                    s = self.running_state(obs, update=False)
                    mean = self.policy_net.forward_tc(s)
                    a, _ = nn.gaussian_sample(mean, self.agent.log_std, 0, 0, det)
                    obs, rew, ci, fail, end, pct = eng.step(a)
                    f, e, r = fail.cpu().numpy()[ids] != 0, end.cpu().numpy()[ids] != 0, rew.cpu().numpy()[ids]
                    for i in np.nonzero(alive)[0]:
                        st = eng.get_state(int(i)) if (t % 10 == 0 or f[i] or e[i]) else None
                        if st is not None:
                            ex = loader.experts[clips[i]]
                            tt = min(st["cur_t"], ex["len"] - 1)
                            jerr[i] += np.linalg.norm(st["xpos"] - ex["wbpos"][tt].reshape(24, 3), axis=1).mean() * 1000
                            rdist[i] += np.linalg.norm(st["qpos"][:3] - ex["qpos"][tt][:3]); cnt[i] += 1
                        rsum[i] += r[i]
                        if f[i]:
                            fail_any[i] = True
                            if cfg.fail_safe:
                                ex = loader.experts[clips[i]]
                                tt = min(t + 1, ex["len"] - 1)
                                eng.set_state(int(i), ex["qpos"][tt], ex["qvel"][tt])
                            else:
                                alive[i] = False
                        if e[i]:
                            alive[i] = False
                    if not alive.any():
                        break
                for i in ids:
                    k = loader.data_keys[c0 + i]
                    res[k] = {"succ": [not fail_any[i]], "reward": rsum[i] / max(lens[i] - 1, 1), "mpjpe_g": jerr[i] / max(cnt[i], 1),
                              "root_dist": rdist[i] / max(cnt[i], 1), "percent": 1.0}
                    if k in self.freq_dict:
                        self.freq_dict[k] += [[res[k]["succ"][0], 0]] * (1 if res[k]["succ"][0] else 3)
                        self.freq_dict[k] = self.freq_dict[k][-self.max_freq:]
            if loader is not self.data_loader:
                eng.load_clips(self.data_loader.experts, self.data_loader.shapes)
            eng.set_cfg(**self._env_cfg(test=False))
            self.agent.obs = None
            metrics = {m: float(np.mean([np.mean(r[m]) for r in res.values()])) for m in ("succ", "reward", "mpjpe_g", "root_dist")}
            coverage = int(round(metrics["succ"] * n))
            self.logger.info(f"Coverage {loader.name} of {coverage} out of {n} | " + " \t".join(f"{k}: {v:.3f}" for k, v in metrics.items()))
            metrics.update(mean_coverage=coverage / n, num_coverage=coverage, all_coverage=n)
            del metrics["succ"]
            res_dicts.append({f"coverage_{loader.name}": metrics})
            if dump:
                path = osp.join(cfg.output_dir, f"{epoch}_{loader.name}_coverage_full.pkl")
                joblib.dump(res, path)
        return res_dicts

    def _env_cfg(self, test):
        cfg = self.cfg
        rw = cfg.reward_weights or {}
        return dict(base_rot=cfg.data_specs.get("base_rot", [0.7071, 0.7071, 0.0, 0.0]), rfc_scale=cfg.residual_force_scale, rfc_lim=cfg.residual_force_lim,
                    rfc_rate=0.0 if cfg.rfc_decay else 1.0, body_diff_thresh=cfg.get("body_diff_thresh_test" if test else "body_diff_thresh", 0.5),
                    meta_pd=int(cfg.meta_pd), env_episode_len=cfg.env_episode_len, trail_steps=cfg.env_expert_trail_steps, auto_reset=0 if test else 1,
                    w=[rw.get(k, d) for k, d in (("w_p", 0.6), ("w_v", 0.1), ("w_e", 0.2), ("w_c", 0.1), ("w_vf", 0.0))],
                    k=[rw.get(k, d) for k, d in (("k_p", 2), ("k_v", 0.005), ("k_e", 20), ("k_c", 1000), ("k_vf", 1))])

    def make_env(self, init_expert=None, mode="test"):
        """single-env facade (HumanoidEnv) for gym-style loops"""
        if init_expert is None:
            init_expert = self.data_loader.get_sample_from_key(self.data_loader.data_keys[0], full_sample=True)
        self.env = HumanoidEnv(self.cfg, init_expert, self.cfg.data_specs, mode=mode)
        return self.env

    # ---------------------------------------------------------------- checkpoints (:190-260)
    def save_checkpoint(self, epoch):
        cfg = self.cfg
        path = "%s/iter_%04d.p" % (cfg.model_dir, epoch + 1)
        pickle.dump(self.agent.state_dicts(), open(path, "wb"))
        joblib.dump(self.freq_dict, osp.join(cfg.result_dir, "freq_dict.pt"))
        return path

    def load_checkpoint(self, epoch):
        cfg = self.cfg
        path = "%s/iter_%04d.p" % (cfg.model_dir, epoch) if isinstance(epoch, int) else epoch
        self.logger.info("loading model from checkpoint: %s" % path)
        cp = pickle.load(open(path, "rb"))
        self.agent.load_state_dicts(cp)
        fd = osp.join(cfg.result_dir, "freq_dict.pt")
        if osp.exists(fd):
            self.freq_dict = joblib.load(fd)
