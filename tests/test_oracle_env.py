"""CPU: the C oracle's env-level logic (PD, RFC, obs v2, termination, reward) against traces of the REFERENCE's own
Python run over the same oracle physics (tools/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def load_expert(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"expert_{tag}.npz"))
    ex = {k: z[k] for k in z.files}
    shape_obs = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    return ex, shape_obs


@pytest.mark.parametrize("tag,act", [("sway", "zero"), ("sway", "noise"), ("kick", "noise")])
def test_env_trace_matches_reference_python(golden_dir, tag, act):
    g = np.load(os.path.join(golden_dir, f"env_{tag}_{act}.npz"))
    ex, so = load_expert(golden_dir, tag)
    env = O.Env(O.Model(), ex, so)
    obs0 = env.reset()
    np.testing.assert_allclose(obs0, g["obs0"], rtol=0, atol=1e-9)
    worst = 0.0
    for t in range(len(g["reward"])):
        obs, r, done, info = env.step(g["action"][t])
        np.testing.assert_allclose(env.torque, g["torque"][t], rtol=1e-7, atol=1e-6, err_msg=f"torque t={t}")
        np.testing.assert_allclose(env.d.qpos, g["qpos"][t], rtol=0, atol=1e-7, err_msg=f"qpos t={t}")
        np.testing.assert_allclose(env.d.qvel, g["qvel"][t], rtol=0, atol=1e-5, err_msg=f"qvel t={t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5, err_msg=f"obs t={t}")
        np.testing.assert_allclose(env.bquat, g["bquat"][t], atol=1e-8)
        np.testing.assert_allclose(env.prev_bquat, g["prev_bquat"][t], atol=1e-8)
        np.testing.assert_allclose(info["c_info"], g["c_info"][t], rtol=0, atol=1e-6, err_msg=f"c_info t={t}")
        assert abs(r - g["reward"][t]) < 1e-6
        assert abs(env.body_diff() - g["body_diff"][t]) < 1e-7
        assert info["fail"] == bool(g["fail"][t]) and info["end"] == bool(g["end"][t])
        assert abs(info["percent"] - g["percent"][t]) < 1e-12
        worst = max(worst, np.abs(env.d.qpos - g["qpos"][t]).max())
    assert worst < 1e-7


def test_explicit_residual_force_trace_matches_reference_python(golden_dir):
    """config/release/uhc_explicit.yml: per-body contact point / force / torque applied through mj_applyFT from the pose of the last forward
    pass (humanoid_im.py:1080-1132), action = 69 + 216 + 30, world_rfc_explicit reward (reward_function.py:253-341) -- the reference's own
    Python over the oracle physics (tools/make_golden.py explicit) against the C restatement."""
    g = np.load(os.path.join(golden_dir, "env_sway_explicit_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so)
    env.set_rfc_mode(True)
    assert env.action_dim == 315 == g["action"].shape[1] and int(g["vf_dim"]) == 216
    np.testing.assert_allclose(env.reset(), g["obs0"], rtol=0, atol=1e-9)
    first_fail = int(np.argmax(g["fail"])) if g["fail"].any() else len(g["reward"])
    assert first_fail >= 20
    for t in range(len(g["reward"])):
        obs, r, done, info = env.step(g["action"][t])
        tol = 1.0 if t < first_fail else 50.0                       # after the fall the trajectory is chaotic: round-off grows
        np.testing.assert_allclose(env.torque, g["torque"][t], rtol=1e-7, atol=1e-6 * tol, err_msg=f"torque t={t}")
        np.testing.assert_allclose(env.d.qpos, g["qpos"][t], rtol=0, atol=1e-7 * tol, err_msg=f"qpos t={t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5 * tol, err_msg=f"obs t={t}")
        np.testing.assert_allclose(info["c_info"], g["c_info"][t], rtol=0, atol=1e-6 * tol, err_msg=f"c_info t={t}")
        assert abs(r - g["reward"][t]) < 1e-6 * tol
        assert info["fail"] == bool(g["fail"][t]) and info["end"] == bool(g["end"][t])


def test_obs_v1_no_meta_pd_trace_matches_reference_python(golden_dir):
    """config/release/uhc_implicit.yml at the env level: obs_v 1 (get_full_obs_v1, humanoid_im.py:323-417: 784 dims = v2 without the shape vector plus
    the per-body COM blocks), no meta-PD (75-wide actions), the yaml's joint gains -- the reference's own Python over the oracle physics against the C restatement."""
    g = np.load(os.path.join(golden_dir, "env_sway_implicit_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so, meta_pd=0)
    env.set_obs_v(1)
    assert env.action_dim == 75 == g["action"].shape[1] and env.obs_dim == 784 == g["obs"].shape[1]
    np.testing.assert_allclose(env.reset(), g["obs0"], rtol=0, atol=1e-9)
    first_fail = int(np.argmax(g["fail"])) if g["fail"].any() else len(g["reward"])
    assert first_fail >= 20
    for t in range(len(g["reward"])):
        obs, r, done, info = env.step(g["action"][t])
        tol = 1.0 if t < first_fail else 50.0
        np.testing.assert_allclose(env.torque, g["torque"][t], rtol=1e-7, atol=1e-6 * tol, err_msg=f"torque t={t}")
        np.testing.assert_allclose(env.d.qpos, g["qpos"][t], rtol=0, atol=1e-7 * tol, err_msg=f"qpos t={t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5 * tol, err_msg=f"obs t={t}")
        assert abs(r - g["reward"][t]) < 1e-6 * tol
        assert info["fail"] == bool(g["fail"][t]) and info["end"] == bool(g["end"][t])


def test_obs_v3_no_shape_no_residual_force_trace_matches_reference_python(golden_dir):
    """config/meta_pd/copycat_35.yml at the env level: obs_v 3 (get_full_obs_v3, humanoid_im.py:505-513: five v2 blocks ten frames apart (no `skip` key: cc_cfg.get("skip", 10))), has_shape false
    (640-wide blocks), residual_force false (99-wide actions: 69 joint targets + 30 meta-PD; the reward's residual-force term is 0)."""
    g = np.load(os.path.join(golden_dir, "env_sway_obsv3_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so)
    env.set_rfc_mode("none"); env.set_has_shape(False); env.set_obs_v(3, fut_frames=5, skip=10)
    assert env.action_dim == 99 == g["action"].shape[1] and env.obs_dim == 3200 == g["obs"].shape[1]
    np.testing.assert_allclose(env.reset(), g["obs0"], rtol=0, atol=1e-9)
    for t in range(len(g["reward"])):
        obs, r, done, info = env.step(g["action"][t])
        np.testing.assert_allclose(env.d.qpos, g["qpos"][t], rtol=0, atol=1e-7, err_msg=f"qpos t={t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5, err_msg=f"obs t={t}")
        np.testing.assert_allclose(info["c_info"], g["c_info"][t], rtol=0, atol=1e-6)
        assert abs(r - g["reward"][t]) < 1e-6 and info["c_info"][4] == 0.0


@pytest.mark.parametrize("term", ["root", "Head"])
def test_env_term_body_root_and_head_flags_match_reference_python(golden_dir, term):
    """cfg.env_term_body "root" / "Head" (humanoid_im.py:1223-1226): the fail flags of the reference's own step() over the noise trajectory of env_sway_noise
    (root below expert["height_lb"] - 0.1 from step 26, head below expert["head_height_lb"] - 0.1 from step 23, the default body criterion from step 31)."""
    g = np.load(os.path.join(golden_dir, f"env_sway_term{term.lower()}_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so)
    env.set_term_body(term, head_body=int(g["head_idx"]) if term == "Head" else 13)
    env.reset()
    fails = []
    for t in range(len(g["fail"])):
        _, _, _, info = env.step(g["action"][t])
        fails.append(info["fail"])
        assert abs(env.d.qpos[2] - g["root_z"][t]) < 1e-7 and info["end"] == bool(g["end"][t])
    assert fails == [bool(f) for f in g["fail"]]
    assert 20 < int(np.argmax(fails)) < 31           # (a different step than the body-position criterion's 31)


def test_multiplicative_reward_matches_reference_python(golden_dir):
    """reward_id world_rfc_implicit_v1_mul (reward_function.py:174-250) over the env_sway_noise trajectory: the reference's own reward and c_info"""
    g = np.load(os.path.join(golden_dir, "env_sway_rewmul_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so)
    env.set_reward_mul(True)
    env.reset()
    for t in range(len(g["reward"])):
        _, r, _, info = env.step(g["action"][t])
        assert abs(r - g["reward"][t]) < 1e-7, t
        np.testing.assert_allclose(info["c_info"], g["c_info"][t], rtol=0, atol=1e-6)
    assert float(g["reward_weights"][4]) != 0.0 and g["reward"].min() < 0.2 < g["reward"].max()      # (w_vf = 0.05: the residual-force term is part of the product)


@pytest.mark.parametrize("v,dim", [(5, 653), (6, 401)])
def test_obs_v5_v6_traces_match_reference_python(golden_dir, v, dim):
    """obs_v 5 / 6 (get_full_obs_v5 humanoid_im.py:505-594, get_full_obs_v6 :596-666, the _new heading helpers of math_utils.py:142-207): the reference's own
    observations over the noise trajectory, bug for bug (v6 drops the x ROW of the transformed joint positions)"""
    g = np.load(os.path.join(golden_dir, f"env_sway_obsv{v}_noise.npz"))
    ex, so = load_expert(golden_dir, "sway")
    env = O.Env(O.Model(), ex, so)
    env.set_obs_v(v)
    assert env.obs_dim == dim == g["obs"].shape[1]
    obs0 = env.reset()
    np.testing.assert_allclose(obs0, g["obs0"], rtol=0, atol=1e-9)
    for t in range(len(g["reward"])):
        obs, r, done, info = env.step(g["action"][t])
        np.testing.assert_allclose(env.d.qpos, g["qpos"][t], rtol=0, atol=1e-7, err_msg=f"qpos t={t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5, err_msg=f"obs t={t}")
        assert abs(r - g["reward"][t]) < 1e-6
