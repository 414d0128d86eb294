"""Base_Config of the reference (uhc/utils/config_utils/base_config.py:8-62): YAML by id (glob config/**/<id>.yml), result dirs,
attribute defaults, `.get`, `.update(args)`."""
import glob
import os
import os.path as osp

import yaml


def _config_roots(base_dir):
    roots = [base_dir or ""]
    ref = os.environ.get("UHC_REFERENCE")
    if ref:
        roots.append(ref)
    roots.append(osp.join(osp.dirname(osp.abspath(__file__)), "..", "..", ".."))   # the repo's own config/
    return roots


class Base_Config:
    def __init__(self, cfg_id, base_dir="", create_dirs=False, cfg_dict=None):
        self.id = cfg_id
        base_dir = base_dir if base_dir else ""
        self.base_dir = os.path.expanduser(base_dir)
        if cfg_dict is not None:
            cfg = cfg_dict
        else:
            files = []
            for r in _config_roots(self.base_dir):
                files = glob.glob(osp.join(r, f"config/**/{cfg_id}.yml"), recursive=True)
                if files:
                    break
            assert len(files) >= 1, f"config id {cfg_id} not found"
            cfg = yaml.safe_load(open(files[0], "r"))
        self.cfg_dict = cfg
        self.main_result_dir = osp.join(self.base_dir, "results")
        self.proj_name = proj_name = cfg.get("proj_name", "motion_im")
        self.cfg_dir = osp.join(self.main_result_dir, proj_name, cfg_id)
        self.model_dir = osp.join(self.cfg_dir, "models")
        self.output_dir = self.result_dir = osp.join(self.cfg_dir, "results")
        self.log_dir = osp.join(self.cfg_dir, "log")
        os.makedirs(self.model_dir, exist_ok=True)
        os.makedirs(self.output_dir, exist_ok=True)
        if create_dirs:
            os.makedirs(self.log_dir, exist_ok=True)
        self.seed = cfg.get("seed", 1)
        self.notes = cfg.get("notes", "exp notes")
        self.data_specs = cfg.get("data_specs", {})
        self.loss_specs = cfg.get("loss_specs", {})
        self.model_specs = cfg.get("model_specs", {})
        self.lr = cfg.get("lr", 3.0e-4)
        self.num_epoch = cfg.get("num_epoch", 100)
        self.num_epoch_fix = cfg.get("num_epoch_fix", 10)
        self.save_n_epochs = cfg.get("save_n_epochs", 20)
        self.eval_n_epochs = cfg.get("eval_n_epochs", 20)
        self.num_samples = self.data_specs.get("num_samples", 5000)
        self.batch_size = self.data_specs.get("batch_size", 5000)

    def get(self, key, default=None):
        return self.cfg_dict.get(key, default)

    def update(self, args):
        for k, v in vars(args).items():
            setattr(self, k, v)
