"""CPU: the product kernel source compiled as a host lane-loop emulation (tests/emu) against the oracle / goldens.
This exercises every line of uhc_b200/csrc/sim_core.h + env_step.h without a GPU."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.emu.emu import Emu


def test_forward_dynamics_matches_oracle_fp64():
    om, d = O.Model(), O.Data()
    e = Emu(64)
    rng = np.random.default_rng(0)
    for case in range(4):
        q = om.qpos0.copy()
        q[2] = 3.0 if case == 0 else rng.uniform(0.85, 0.95)
        q[3:7] = [0.7071068, 0.7071068, 0, 0] + (rng.normal(size=4) * (1.0 if case == 0 else 0.05))
        q[3:7] /= np.linalg.norm(q[3:7])
        q[7:] = rng.uniform(-0.3, 0.3, 69)
        v = rng.normal(size=75) * 0.5
        tau, fapp = rng.normal(size=69) * 20, rng.normal(size=6) * 10
        d.qpos[:], d.qvel[:], d.ctrl[:] = q, v, tau
        d.qfrc_applied[:] = 0
        d.qfrc_applied[:6] = fapp
        d.qacc_warm[:] = 0
        O.forward(om, d)
        r = e.forward(q, v, tau, fapp)
        assert np.abs(r["C"] - d.C).max() < 1e-9
        assert np.abs(r["xpos"] - d.xpos.reshape(24, 3)).max() < 1e-13
        assert r["ncon"] == d.ncon
        assert np.abs(r["qacc"] - d.qacc).max() < 1e-6 * max(1.0, np.abs(d.qacc).max())


def test_forward_dynamics_many_contacts_fp64():
    """Leaning far forward close to the floor: 33..40 contacts, i.e. the second 32-contact chunk of the wrench prefix sums."""
    om, d = O.Model(), O.Data()
    e = Emu(64)
    q = om.qpos0.copy()
    q[2] = 0.16
    a = np.deg2rad(40.0) / 2
    q[3:7] = [np.cos(a), 0, np.sin(a), 0]
    v = np.random.default_rng(3).normal(size=75) * 0.3
    d.qpos[:], d.qvel[:], d.ctrl[:] = q, v, 0
    d.qfrc_applied[:] = 0
    d.qacc_warm[:] = 0
    O.forward(om, d)
    r = e.forward(q, v, np.zeros(69), np.zeros(6))
    assert 32 < d.ncon <= 40 and r["ncon"] == d.ncon
    assert np.abs(r["qacc"] - d.qacc).max() < 1e-6 * max(1.0, np.abs(d.qacc).max())


@pytest.mark.parametrize("seed", range(6))
def test_forward_dynamics_random_contact_states_fp64(seed):
    """Seeded random poses from standing to crouching / leaning (0 .. ~30 contacts), random velocities, torques, applied root wrench and
    warm start: the O(n) kernel pipeline (spatial RNE, centre-rooted block articulated-body solves, prefix-sum contact wrenches, Newton)
    must land on the oracle's dense solution."""
    om, d = O.Model(), O.Data()
    e = Emu(64)
    rng = np.random.default_rng(100 + seed)
    ncons = []
    for case in range(5):
        q = om.qpos0.copy()
        q[2] = rng.uniform(0.25, 0.95)
        tilt = rng.normal(size=3) * rng.uniform(0.0, 0.6)
        ang = np.linalg.norm(tilt) + 1e-12
        dq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * tilt / ang])
        b = np.array([0.7071068, 0.7071068, 0.0, 0.0])
        q[3:7] = [dq[0] * b[0] - dq[1:] @ b[1:], *(dq[0] * b[1:] + b[0] * dq[1:] + np.cross(dq[1:], b[1:]))]
        q[3:7] /= np.linalg.norm(q[3:7])
        q[7:] = rng.uniform(-0.6, 0.6, 69)
        v = rng.normal(size=75) * rng.uniform(0.0, 1.5)
        tau, fapp, aw = rng.normal(size=69) * 30, rng.normal(size=6) * 20, rng.normal(size=75) * 5
        d.qpos[:], d.qvel[:], d.ctrl[:] = q, v, tau
        d.qfrc_applied[:] = 0
        d.qfrc_applied[:6] = fapp
        d.qacc_warm[:] = aw
        O.forward(om, d)
        if d.ncon > 40:
            continue
        r = e.forward(q, v, tau, fapp, aw)
        ncons.append(d.ncon)
        assert r["ncon"] == d.ncon
        assert np.abs(r["C"] - d.C).max() < 1e-8 * max(1.0, np.abs(d.C).max())
        assert np.abs(r["qacc"] - d.qacc).max() < 2e-6 * max(1.0, np.abs(d.qacc).max()), (case, d.ncon, np.abs(r["qacc"] - d.qacc).max())
    assert len(ncons) >= 3


@pytest.mark.parametrize("prec,tol_q,tol_o", [(64, 1e-11, 1e-9), (32, 1e-4, 2e-3)])
def test_env_trace_matches_reference_golden(golden_dir, prec, tol_q, tol_o):
    g = np.load(os.path.join(golden_dir, "env_sway_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    e = Emu(prec)
    e.load_clips([ex], [so])
    obs0 = e.reset()
    assert np.abs(obs0 - g["obs0"]).max() < max(tol_o * 1e-2, 1e-12)
    for t in range(35):
        obs, r, done, info = e.step(g["action"][t])
        st, _ = e.state()
        assert np.abs(st[:76] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(obs - g["obs"][t]).max() < tol_o, t
        assert abs(r - g["reward"][t]) < tol_o
        assert info["fail"] == bool(g["fail"][t])


@pytest.mark.parametrize("prec,tol_q,tol_o", [(64, 1e-10, 1e-8), (32, 2e-4, 4e-3)])
def test_explicit_residual_force_trace_matches_reference_golden(golden_dir, prec, tol_q, tol_o):
    """residual_force_mode = explicit (config/release/uhc_explicit.yml): the kernel source (per-body wrenches about the stale root position,
    projected through the stale motion subspaces; action = 69 + 216 + 30; world_rfc_explicit reward) against the reference's own Python."""
    from uhc_b200.model import HumanoidModel
    g = np.load(os.path.join(golden_dir, "env_sway_explicit_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    import ctypes as C
    e = Emu(prec, rfc_mode=1, vf_slot=(C.c_int * 24)(*HumanoidModel().vf_slot()))
    e.load_clips([ex], [so])
    obs0 = e.reset()
    assert np.abs(obs0 - g["obs0"]).max() < max(tol_o * 1e-2, 1e-12)
    first_fail = int(np.argmax(g["fail"]))
    for t in range(first_fail - (0 if prec == 64 else 8)):      # fp32: the last steps before the fall amplify round-off (see tests/test_gpu_env.py)
        obs, r, done, info = e.step(g["action"][t])
        st, _ = e.state()
        assert np.abs(st[:76] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(obs - g["obs"][t]).max() < tol_o, t
        assert abs(r - g["reward"][t]) < tol_o and np.abs(info["c_info"] - g["c_info"][t]).max() < tol_o
        assert info["fail"] == bool(g["fail"][t])


@pytest.mark.parametrize("prec,tol_q,tol_o", [(64, 1e-10, 1e-8), (32, 2e-4, 4e-3)])
def test_obs_v1_no_meta_pd_trace_matches_reference_golden(golden_dir, prec, tol_q, tol_o):
    """config/release/uhc_implicit.yml at the env level (obs_v 1 = 784 dims with the per-body COM blocks, 75-wide actions without meta-PD): the
    kernel source against the reference's own Python."""
    g = np.load(os.path.join(golden_dir, "env_sway_implicit_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    e = Emu(prec, obs_v=1, meta_pd=0)
    e.load_clips([ex], [so])
    obs0 = e.reset()
    assert obs0.shape == (784,) and np.abs(obs0 - g["obs0"]).max() < max(tol_o * 1e-2, 1e-12)
    first_fail = int(np.argmax(g["fail"]))
    for t in range(first_fail - (0 if prec == 64 else 6)):
        obs, r, done, info = e.step(g["action"][t])
        st, _ = e.state()
        assert np.abs(st[:76] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(obs - g["obs"][t]).max() < tol_o, t
        assert abs(r - g["reward"][t]) < tol_o
        assert info["fail"] == bool(g["fail"][t])


@pytest.mark.parametrize("with_contacts", [False, True])
def test_joint_limit_rows_match_oracle_fp64(with_contacts):
    """Tightened hinge ranges (as smpl_robot.py:1087-1110 does per shape) with several joints past them, airborne and standing on the floor: the
    kernel's limit rows (unit-Jacobian soft rows folded into the solve's joint-space diagonal and the carried gradient) against the oracle's dense rows."""
    from uhc_b200.model import HumanoidModel
    z = np.load(O.MODEL_NPZ)
    jr = np.tile(np.array([[-0.4, 0.5]]), (69, 1))
    om, d = O.Model(tables={"jnt_range": jr}), O.Data()
    e = Emu(64, model=HumanoidModel(jnt_range=jr))
    rng = np.random.default_rng(17)
    nviol = []
    for case in range(6):
        q = om.qpos0.copy()
        q[2] = rng.uniform(0.88, 0.93) if with_contacts else 3.0
        q[3:7] = [0.7071068, 0.7071068, 0, 0]
        q[7:] = rng.uniform(-0.7, 0.8, 69) if not with_contacts else rng.uniform(-0.05, 0.05, 69)
        if with_contacts:
            idx = rng.choice(np.arange(24, 69), 6, replace=False)        # upper-body joints past the range, feet on the ground
            q[7 + idx] = rng.choice([-0.55, 0.65], 6)
        v = rng.normal(size=75) * 0.5
        tau, fapp, aw = rng.normal(size=69) * 20, rng.normal(size=6) * 10, rng.normal(size=75) * 3
        d.qpos[:], d.qvel[:], d.ctrl[:] = q, v, tau
        d.qfrc_applied[:] = 0; d.qfrc_applied[:6] = fapp; d.qacc_warm[:] = aw
        O.forward(om, d)
        if d.ncon > 40:
            continue
        assert (d.ncon > 0) == with_contacts
        r = e.forward(q, v, tau, fapp, aw)
        nviol.append(int(((q[7:] < jr[:, 0]) | (q[7:] > jr[:, 1])).sum()))
        assert r["ncon"] == d.ncon
        assert np.abs(r["qacc"] - d.qacc).max() < 2e-6 * max(1.0, np.abs(d.qacc).max()), (case, d.ncon, nviol[-1], np.abs(r["qacc"] - d.qacc).max())
    assert len(nviol) >= 4 and min(nviol) >= 4


@pytest.mark.parametrize("prec,tol_q,tol_o", [(64, 1e-10, 1e-8), (32, 1e-4, 2e-3)])
def test_obs_v3_no_shape_no_residual_force_matches_reference_golden(golden_dir, prec, tol_q, tol_o):
    """config/meta_pd/copycat_35.yml at the env level (obs_v 3 = 5 x 640, 99-wide actions without residual force): kernel source vs the reference's Python"""
    g = np.load(os.path.join(golden_dir, "env_sway_obsv3_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    e = Emu(prec, obs_v=3, fut_frames=5, fut_skip=10, no_shape=1, rfc_mode=2)
    e.load_clips([ex], [so])
    obs0 = e.reset()
    assert obs0.shape == (3200,) and np.abs(obs0 - g["obs0"]).max() < max(tol_o * 1e-2, 1e-12)
    for t in range(len(g["reward"])):
        obs, r, done, info = e.step(g["action"][t])
        st, _ = e.state()
        assert np.abs(st[:76] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(obs - g["obs"][t]).max() < tol_o, t
        assert abs(r - g["reward"][t]) < tol_o and info["c_info"][4] == 0.0


@pytest.mark.parametrize("term", ["root", "Head"])
def test_env_term_body_root_and_head_flags_match_reference_golden(golden_dir, term):
    """env_term_body root / Head in the kernel source (host emulation, fp64): the reference's own fail flags, the window minimum taken in the kernel"""
    g = np.load(os.path.join(golden_dir, f"env_sway_term{term.lower()}_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    e = Emu(64, term_body={"root": 1, "Head": 2}[term], head_body=int(g["head_idx"]) if term == "Head" else 13)
    e.load_clips([ex], [so])
    e.reset()
    fails = []
    for t in range(34):
        _, _, _, info = e.step(g["action"][t])
        fails.append(info["fail"])
    assert fails == [bool(f) for f in g["fail"][:34]] and any(fails)


@pytest.mark.parametrize("v,prec,tol_q,tol_o", [(5, 64, 1e-10, 1e-8), (6, 64, 1e-10, 1e-8), (5, 32, 2e-4, 4e-3), (6, 32, 2e-4, 4e-3)])
def test_obs_v5_v6_match_reference_golden(golden_dir, v, prec, tol_q, tol_o):
    """obs_v 5 / 6 in the kernel source (host emulation): the reference's own get_full_obs_v5 / get_full_obs_v6 over the noise trajectory"""
    g = np.load(os.path.join(golden_dir, f"env_sway_obsv{v}_noise.npz"))
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    e = Emu(prec, obs_v=v)
    assert e.obs_dim == g["obs"].shape[1]
    e.load_clips([ex], [so])
    obs0 = e.reset()
    assert np.abs(obs0 - g["obs0"]).max() < max(tol_o * 1e-2, 1e-12)
    for t in range(len(g["reward"])):
        obs, r, done, info = e.step(g["action"][t])
        st, _ = e.state()
        assert np.abs(st[:76] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(obs - g["obs"][t]).max() < tol_o, (t, int(np.abs(obs - g["obs"][t]).argmax()))
