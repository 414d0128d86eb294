"""CPU: product motion library (uhc_b200/motion_lib.py) against expert tables produced by the reference's own
smpl_to_qpose + Humanoid.qpos_fk (tests/golden/expert_*.npz)."""
import os

import numpy as np
import pytest

from uhc_b200 import motion_lib as ML


@pytest.mark.parametrize("tag", ["sway", "kick"])
def test_expert_tables_match_reference(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"expert_{tag}.npz"))
    pose = z["pose_aa"].copy()
    pose[:, 66:] = 0                     # the golden keeps SMPL-H columns 66:72; smplh_to_smpl zeroes the hands (smpl_mujoco.py:533)
    ex = ML.make_expert(pose, z["trans"])
    assert np.abs(ex["qpos"][:, :3] - z["qpos"][:, :3]).max() < 1e-12
    assert np.abs(ex["qpos"][:, 7:] - z["qpos"][:, 7:]).max() < 3e-6   # the reference converts axis-angle in float32
    # root quaternion up to sign
    d = np.minimum(np.abs(ex["qpos"][:, 3:7] - z["qpos"][:, 3:7]).max(1), np.abs(ex["qpos"][:, 3:7] + z["qpos"][:, 3:7]).max(1))
    assert d.max() < 1e-6
    ex = ML.qpos_fk(z["qpos"])          # FK / finite differences on the reference's own qpos: tight tolerances
    for k, tol in (("qvel", 1e-8), ("wbpos", 1e-10), ("wbquat", 1e-10), ("bquat", 1e-10), ("body_com", 1e-10), ("bangvel", 1e-7),
                   ("ee_wpos", 1e-10), ("com", 1e-10)):
        assert np.abs(np.asarray(ex[k]).reshape(z[k].shape) - z[k]).max() < tol, k
    assert abs(ex["height_lb"] - float(z["height_lb"])) < 1e-12 and ex["len"] == int(z["length"])


def test_synthetic_clip_is_well_formed():
    ex = ML.synthetic_clip(120, np.random.default_rng(0))
    assert ex["qpos"].shape == (120, 76) and np.isfinite(ex["qvel"]).all() and np.abs(ex["qvel"]).max() <= 10.0
    assert np.abs(np.linalg.norm(ex["wbquat"].reshape(120, 24, 4), axis=-1) - 1).max() < 1e-6
