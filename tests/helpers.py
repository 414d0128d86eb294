"""Synthetic AMASS-style motion pickle with the reference's schema (sample_data/amass_copycat_take5_test_small.pkl)."""
import os

import joblib
import numpy as np


def write_synthetic_pkl(path, nclips=3, seed=0):
    rng = np.random.RandomState(seed)
    data = {}
    for c in range(nclips):
        T = int(rng.randint(40, 90))
        t = np.arange(T) / 30.0
        pose = np.zeros((T, 72))
        for j in range(1, 22):
            for a in range(3):
                pose[:, 3 * j + a] = rng.uniform(0, 0.25) * np.sin(2 * np.pi * rng.uniform(0.2, 1.0) * t + rng.uniform(0, 6.28))
        pose[:, :3] = [np.pi / 2, 0, 0] if False else pose[:, :3]
        pose[:, 0] = 1.2092                       # root orientation: SMPL (Y-up) pose seen upright in the Z-up world
        pose[:, 1] = 1.2092
        pose[:, 2] = 1.2092
        trans = np.stack([0.3 * t, 0.0 * t, 0.0 * t + 0.9 - 0.2233 + 0.05], 1)
        data[f"0-synthetic_{c}_poses"] = {"pose_aa": pose, "pose_6d": np.zeros((T, 24, 6)), "trans": trans, "beta": rng.normal(0, 1, 16),
                                          "gender": ["neutral", "male", "female"][c % 3], "seq_name": f"synthetic_{c}"}
    os.makedirs(os.path.dirname(path), exist_ok=True)
    joblib.dump(data, path)
    return path
