"""Config of the reference (uhc/utils/config_utils/copycat_config.py:12-179): same attributes, same defaults, same
piece-wise-linear adaptive schedules.  The MuJoCo XML asset is replaced by the compiled model tables of uhc_b200, so
`mujoco_model_file` resolves to the XML when it exists (reference checkout) and to the compiled asset otherwise."""
import os.path as osp

import numpy as np

from uhc.utils.config_utils.base_config import Base_Config, _config_roots


class Config(Base_Config):
    def __init__(self, mujoco_path="%s.xml", **kwargs):
        super().__init__(**kwargs)
        c = self.cfg_dict
        g = c.get
        self.gamma, self.tau = g("gamma", 0.95), g("tau", 0.95)
        self.policy_htype, self.policy_hsize = g("policy_htype", "relu"), g("policy_hsize", [300, 200])
        self.policy_optimizer, self.policy_lr = g("policy_optimizer", "Adam"), g("policy_lr", 5e-5)
        self.policy_momentum, self.policy_weightdecay = g("policy_momentum", 0.0), g("policy_weightdecay", 0.0)
        self.value_htype, self.value_hsize = g("value_htype", "relu"), g("value_hsize", [300, 200])
        self.value_optimizer, self.value_lr = g("value_optimizer", "Adam"), g("value_lr", 3e-4)
        self.value_momentum, self.value_weightdecay = g("value_momentum", 0.0), g("value_weightdecay", 0.0)
        self.adv_clip, self.clip_epsilon = g("adv_clip", np.inf), g("clip_epsilon", 0.2)
        self.log_std, self.fix_std = g("log_std", -2.3), g("fix_std", False)
        self.num_optim_epoch = g("num_optim_epoch", 10)
        self.min_batch_size = g("min_batch_size", 50000)
        self.mini_batch_size = g("mini_batch_size", self.min_batch_size)
        self.save_n_epochs = g("save_n_epochs", 100)
        self.reward_id, self.reward_weights = g("reward_id", "quat"), g("reward_weights", None)
        self.end_reward, self.actor_type = g("end_reward", False), g("actor_type", "gauss")
        if self.actor_type == "mcp":
            self.num_primitive, self.composer_dim = g("num_primitive", 8), g("composer_dim", [300, 200])
        self.adp_iter_cp = np.array(g("adp_iter_cp", [0]))
        pad = lambda a: np.pad(np.array(a, dtype=np.float64), (0, self.adp_iter_cp.size - len(a)), "edge")
        self.adp_noise_rate_cp = pad(g("adp_noise_rate_cp", [1.0]))
        self.adp_log_std_cp = pad(g("adp_log_std_cp", [self.log_std]))
        self.adp_policy_lr_cp = pad(g("adp_policy_lr_cp", [self.policy_lr]))
        self.adp_noise_rate = self.adp_log_std = self.adp_policy_lr = None
        self.mujoco_model_file = self.find_asset(mujoco_path % c["mujoco_model"])
        self.vis_model_file = self.find_asset(mujoco_path % c.get("vis_model", c["mujoco_model"]), required=False)
        self.env_start_first, self.env_init_noise = g("env_start_first", False), g("env_init_noise", 0.0)
        self.env_episode_len, self.env_term_body = g("env_episode_len", 200), g("env_term_body", "head")
        self.env_expert_trail_steps = g("env_expert_trail_steps", 0)
        self.obs_v, self.obs_type, self.obs_coord = g("obs_v", 0), g("obs_type", "full"), g("obs_coord", "root")
        self.obs_phase, self.obs_heading, self.obs_vel = g("obs_phase", True), g("obs_heading", False), g("obs_vel", "full")
        self.root_deheading, self.action_type, self.action_v = g("root_deheading", False), g("action_type", "position"), g("action_v", 0)
        self.reactive_v, self.no_root, self.reactive_rate = g("reactive_v", 0), g("no_root", False), g("reactive_rate", 0.3)
        self.sampling_temp, self.sampling_freq = g("sampling_temp", 0.2), g("sampling_freq", 0.75)
        self.residual_force, self.residual_force_scale = g("residual_force", False), g("residual_force_scale", 200.0)
        self.residual_force_lim, self.residual_force_mode = g("residual_force_lim", 100.0), g("residual_force_mode", "implicit")
        self.residual_force_bodies, self.residual_force_torque = g("residual_force_bodies", "all"), g("residual_force_torque", True)
        self.rfc_decay = g("rfc_decay", False)
        self.meta_pd, self.meta_pd_joint = g("meta_pd", False), g("meta_pd_joint", False)
        self.masterfoot, self.fail_safe = g("masterfoot", False), g("fail_safe", True)
        self.robot_cfg = g("robot", {})
        if len(self.robot_cfg) == 0:
            self.robot_cfg["model"] = "smpl"
            self.robot_cfg["mesh"] = "mesh" in str(self.mujoco_model_file)
        self.has_shape = g("has_shape", False)
        self.agent_name, self.model_name = g("agent_name", "agent_copycat"), g("model_name", "super_net")

    def update_adaptive_params(self, i_iter):
        cp = self.adp_iter_cp
        ind = np.where(i_iter >= cp)[0][-1]
        nind = ind + int(ind < len(cp) - 1)
        t = (i_iter - cp[ind]) / (cp[nind] - cp[ind]) if nind > ind else 0.0
        self.adp_noise_rate = self.adp_noise_rate_cp[ind] * (1 - t) + self.adp_noise_rate_cp[nind] * t
        self.adp_log_std = self.adp_log_std_cp[ind] * (1 - t) + self.adp_log_std_cp[nind] * t
        self.adp_policy_lr = self.adp_policy_lr_cp[ind] * (1 - t) + self.adp_policy_lr_cp[nind] * t

    def find_asset(self, asset_path, required=True):
        if osp.exists(asset_path):
            return asset_path
        for r in _config_roots(self.base_dir):
            full = osp.join(r, "assets/mujoco_models", osp.basename(asset_path))
            if osp.exists(full):
                return full
        from uhc_b200.model import ASSET
        name = osp.splitext(osp.basename(asset_path))[0]
        if name.startswith("humanoid_smpl_neutral_mesh") and osp.exists(ASSET):
            return ASSET   # compiled tables of that XML (tools/compile_model.py)
        if required:
            raise IOError("File %s does not exist" % asset_path)
        return None
