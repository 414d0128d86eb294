"""CPU: the drop-in `uhc` package surface: Config defaults / schedules, dataset schema and sampling."""
import os

import numpy as np

from tests.helpers import write_synthetic_pkl


def test_config_defaults_and_schedules(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from uhc.utils.config_utils.copycat_config import Config
    from uhc.utils.flags import flags
    cfg = Config(cfg_id="uhc_b200_default", create_dirs=False)
    assert cfg.policy_hsize == [2048, 1024, 512] and cfg.policy_htype == "gelu" and cfg.obs_v == 2 and cfg.meta_pd
    assert cfg.residual_force and cfg.residual_force_scale == 100 and cfg.residual_force_lim == 100.0       # copycat_config.py:106 default
    assert cfg.reward_id == "world_rfc_implicit" and cfg.env_term_body == "body" and cfg.fix_std and cfg.log_std == -2.3
    assert os.path.isdir(cfg.model_dir) and cfg.get("nonexistent", 7) == 7
    cfg.update_adaptive_params(10)
    assert cfg.adp_noise_rate == 1.0 and cfg.adp_policy_lr == 5e-5 and cfg.adp_log_std == -2.3

    import types
    cfg.update(types.SimpleNamespace(num_threads=3, no_log=True))
    assert cfg.num_threads == 3 and cfg.no_log is True
    flags.debug = True
    assert flags.debug


def test_adaptive_schedule_interpolates(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from uhc.utils.config_utils.copycat_config import Config
    import yaml
    base = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "config", "uhc_b200_default.yml")))
    base.update(adp_iter_cp=[0, 100], adp_noise_rate_cp=[1.0, 0.0], adp_policy_lr_cp=[5e-5, 1e-5])
    cfg = Config(cfg_id="x", cfg_dict=base)
    cfg.update_adaptive_params(50)
    assert abs(cfg.adp_noise_rate - 0.5) < 1e-12 and abs(cfg.adp_policy_lr - 3e-5) < 1e-12
    cfg.update_adaptive_params(500)
    assert cfg.adp_noise_rate == 0.0


def test_dataset_schema_and_sampling(tmp_path):
    from uhc.data_loaders.dataset_amass_single import DatasetAMASSSingle
    p = write_synthetic_pkl(str(tmp_path / "sample_data" / "clips.pkl"))
    ds = DatasetAMASSSingle({"file_path": p, "t_min": 5, "t_max": 60})
    assert ds.get_len() == 3 and len(ds.experts) == 3
    for ex, k in zip(ds.experts, ds.data_keys):
        T = ds.get_sample_len_from_key(k)
        assert ex["qpos"].shape == (T, 76) and ex["wbpos"].shape == (T, 72) and ex["bangvel"].shape == (T, 72)
    s = ds.sample_seq()
    assert s["pose_aa"].shape[0] <= 60 and s["beta"].shape[1] == 16 and s["seq_name"] in ds.data_keys
    full = ds.get_sample_from_key(ds.data_keys[1], full_sample=True)
    assert full["pose_aa"].shape[0] == ds.get_sample_len_from_key(ds.data_keys[1]) and full["gender"][0] == 1
