"""GPU parity: CUDA env kernels (through the C ABI, uhc_b200.engine) against the reference-pinned golden traces and
against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expert(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"expert_{tag}.npz"))
    ex = {k: z[k] for k in z.files}
    return ex, np.concatenate([ex["beta"][0], [ex["gender"][0]]])


def _run_trace(golden_dir, tag, act, precision, E=1):
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, f"env_{tag}_{act}.npz"))
    ex, so = _expert(golden_dir, tag)
    eng = Engine(E, precision=precision)
    eng.load_clips([ex], [so])
    obs0 = eng.reset().cpu().numpy()
    out = dict(obs0=obs0, qpos=[], obs=[], reward=[], cinfo=[], fail=[], end=[], percent=[], torque=[])
    tq = torch.zeros(E, 15, 69, device="cuda")
    for t in range(len(g["reward"])):
        a = torch.tensor(np.tile(g["action"][t], (E, 1)), dtype=torch.float32, device="cuda")
        obs, rew, ci, fail, end, pct = eng.step(a, torque_out=tq)
        torch.cuda.synchronize()
        out["obs"].append(obs.cpu().numpy().copy()); out["reward"].append(rew.cpu().numpy().copy()); out["cinfo"].append(ci.cpu().numpy().copy())
        out["fail"].append(fail.cpu().numpy().copy()); out["end"].append(end.cpu().numpy().copy()); out["percent"].append(pct.cpu().numpy().copy())
        out["torque"].append(tq.cpu().numpy().copy())
        out["qpos"].append(np.stack([eng.get_state(e)["qpos"] for e in range(min(E, 2))]))
    eng.close()
    return g, {k: np.array(v) for k, v in out.items()}


@pytest.mark.parametrize("tag,act", [("sway", "noise"), ("kick", "noise")])
def test_fp64_kernels_match_reference_trace(golden_dir, tag, act):
    """Same algorithm in fp64 on the GPU: agreement to solver tolerance with the reference Python run on the oracle."""
    g, o = _run_trace(golden_dir, tag, act, 64)
    n = len(g["reward"])
    first_fail = int(np.argmax(g["fail"])) if g["fail"].any() else n
    lim = min(n, first_fail + 5)
    assert np.abs(o["obs0"][0] - g["obs0"]).max() < 1e-6          # obs are written as float32
    assert np.abs(o["qpos"][:lim, 0] - g["qpos"][:lim]).max() < 1e-8
    assert np.abs(o["torque"][:lim, 0] - g["torque"][:lim]).max() < 1e-3
    assert np.abs(o["reward"][:lim, 0] - g["reward"][:lim]).max() < 1e-6
    assert (o["fail"][:, 0].astype(bool) == g["fail"]).all() and (o["end"][:, 0].astype(bool) == g["end"]).all()


@pytest.mark.parametrize("tag,act", [("sway", "zero"), ("sway", "noise"), ("kick", "noise")])
def test_fp32_kernels_within_1e3_rad_after_60_steps(golden_dir, tag, act):
    """BASELINE north_star tolerance: joint qpos within 1e-3 rad after 60 steps (checked on every step up to the first
    termination; after a fall the tumbling ragdoll is chaotic and only the flags are compared)."""
    g, o = _run_trace(golden_dir, tag, act, 32, E=3)
    n = len(g["reward"])
    first_fail = int(np.argmax(g["fail"])) if g["fail"].any() else n
    lim = min(n, max(first_fail, 1))
    err = np.abs(o["qpos"][:lim, 0] - g["qpos"][:lim])
    assert err[:, 7:].max() < 1e-3 and err[:, :7].max() < 1e-3, err.max()
    # velocities of the light distal links jump at contact make/break events; the last steps before a fall are violent, so the
    # observation / reward comparison stops three steps before the first termination (qpos is compared on every step above)
    lo = max(lim - 3, 1)
    assert np.abs(o["obs"][:lo, 0] - g["obs"][:lo]).max() < 5e-3
    assert np.abs(o["reward"][:lo, 0] - g["reward"][:lo]).max() < 1e-3
    assert np.abs(o["cinfo"][:lo, 0] - g["c_info"][:lo]).max() < 2e-3
    assert np.abs(o["percent"][:, 0] - g["percent"]).max() < 1e-6
    assert (o["fail"][:lim, 0].astype(bool) == g["fail"][:lim]).all() and (o["end"][:, 0].astype(bool) == g["end"]).all()
    # identical inputs in different envs of the batch give bit-identical outputs
    assert (o["obs"][:, 0] == o["obs"][:, 2]).all() and (o["reward"][:, 0] == o["reward"][:, 1]).all()


def test_batched_envs_match_oracle_on_seeded_inputs(golden_dir):
    """64 envs with different start frames and seeded actions against the CPU oracle env, 12 steps."""
    import torch
    from oracle import oracle as O
    from uhc_b200.engine import Engine
    ex, so = _expert(golden_dir, "sway")
    E, T = 64, 12
    rng = np.random.RandomState(7)
    starts = rng.randint(0, 60, E).astype(np.int32)
    acts = rng.normal(0, 0.1, (T, E, 105)).astype(np.float32)
    acts[:, :, 69:75] *= 0.3
    eng = Engine(E)
    eng.load_clips([ex], [so])
    obs = eng.reset(start=starts).cpu().numpy().copy()
    om = O.Model()
    worst_q = worst_r = 0.0
    envs = []
    for e in range(0, E, 8):
        sl = {k: ex[k][starts[e]:] for k in ("qpos", "qvel", "wbpos", "wbquat", "bquat", "bangvel", "ee_wpos", "com")}
        oe = O.Env(om, sl, so)
        o0 = oe.reset()
        assert np.abs(o0 - obs[e]).max() < 1e-4
        envs.append((e, oe))
    for t in range(T):
        o, r, ci, f, en, p = eng.step(torch.tensor(acts[t], device="cuda"))
        r = r.cpu().numpy(); f = f.cpu().numpy()
        for e, oe in envs:
            _, ro, _, info = oe.step(acts[t, e].astype(np.float64))
            q = eng.get_state(e)["qpos"]
            worst_q = max(worst_q, np.abs(q - oe.d.qpos).max()); worst_r = max(worst_r, abs(ro - r[e]))
            assert bool(f[e]) == info["fail"]
    assert worst_q < 1e-4 and worst_r < 1e-4, (worst_q, worst_r)
    eng.close()


def test_mixed_body_shapes_match_oracle(golden_dir):
    """BASELINE configs[3]: envs with different body shapes in one batch (per-clip model variant, synthetic limb scaling
    U[0.85, 1.15] since the SMPL files are licence-gated) against the oracle built from the same scaled tables."""
    import torch
    from oracle import oracle as O
    from uhc_b200 import motion_lib
    from uhc_b200.engine import Engine
    from uhc_b200.model import HumanoidModel
    z = np.load(os.path.join(golden_dir, "expert_sway.npz"))
    pose = np.concatenate([z["pose_aa"][:, :66], np.zeros((len(z["pose_aa"]), 6))], 1)
    rng = np.random.RandomState(4)
    hm0 = HumanoidModel()
    hm1 = HumanoidModel(scale=rng.uniform(0.85, 1.15, 24))
    assert abs(hm1.mass.sum() - hm0.mass.sum()) > 0.5 and np.abs(hm1.invw - hm0.invw).max() > 1e-4
    exs = [motion_lib.make_expert(pose, z["trans"], m) for m in (hm0, hm1)]
    so = np.concatenate([z["beta"][0], [z["gender"][0]]])
    E, T = 8, 8
    eng = Engine(E, hm0, variants=[hm0, hm1])
    eng.load_clips(exs, [so, so], clip_models=[0, 1])
    clip = (np.arange(E) % 2).astype(np.int32)
    obs = eng.reset(np.arange(E, dtype=np.int32), clip, 0, None).cpu().numpy().copy()
    oms = [O.Model(), O.Model(tables=dict(body_offset=hm1.offset, body_mass=hm1.mass, body_ipos=hm1.ipos, body_inertia=hm1.inertia,
                                          body_invweight0=np.stack([hm1.invw, np.zeros(24)], 1), hull_vert=hm1.hull))]
    oes = [O.Env(oms[c], exs[c], so) for c in (0, 1)]
    for c in (0, 1):
        assert np.abs(oes[c].reset() - obs[c]).max() < 1e-4
    assert np.abs(obs[0] - obs[1]).max() > 1e-3                       # the two shapes really differ
    acts = rng.normal(0, 0.1, (T, 105)).astype(np.float32)
    for t in range(T):
        o, r, ci, f, en, p = eng.step(torch.tensor(np.tile(acts[t], (E, 1)), device="cuda"))
        r = r.cpu().numpy()
        for c in (0, 1):
            _, ro, _, info = oes[c].step(acts[t].astype(np.float64))
            for e in (c, c + 2, c + 6):
                assert np.abs(eng.get_state(e)["qpos"] - oes[c].d.qpos).max() < 1e-4
                assert abs(r[e] - ro) < 1e-4
    eng.close()


@pytest.mark.parametrize("kind", ["sitting", "airborne"])
def test_varying_contact_counts_match_oracle(kind):
    """BASELINE configs[4]-style synthetic clips (root lowered to a sitting height -> many body/floor contacts; hops -> none):
    contact counts from 0 to >20 per env, engine vs oracle on the same seeded actions until the first termination."""
    import torch
    from oracle import oracle as O
    from uhc_b200 import motion_lib
    from uhc_b200.engine import Engine
    ex = motion_lib.synthetic_clip(60, np.random.default_rng(5), kind=kind)
    eng = Engine(4)
    eng.load_clips([ex], None)
    obs = eng.reset().cpu().numpy().copy()
    oe = O.Env(O.Model(), ex, np.zeros(17))
    assert np.abs(oe.reset() - obs[0]).max() < 1e-4
    rng = np.random.RandomState(2)
    maxcon, steps = 0, 0
    for t in range(40):
        a = rng.normal(0, 0.1, 105).astype(np.float32)
        o, r, ci, f, en, p = eng.step(torch.tensor(np.tile(a, (4, 1)), device="cuda"))
        _, ro, done, info = oe.step(a.astype(np.float64))
        st = eng.get_state(1)
        maxcon = max(maxcon, st["ncon"], oe.d.ncon)
        assert oe.d.ncon <= 40, "oracle contact count exceeds the kernel's capacity: raise MAXCON"
        assert np.abs(st["qpos"] - oe.d.qpos).max() < 1e-3, (t, np.abs(st["qpos"] - oe.d.qpos).max())
        assert bool(f[1]) == info["fail"]
        steps += 1
        if info["fail"] or info["end"]:
            break
    assert steps >= 3
    if kind == "sitting":
        assert maxcon >= 12
    eng.close()


def test_more_than_32_contacts_match_oracle():
    """Leaning far forward close to the floor (36 contacts): exercises the second chunk of the contact-wrench prefix sums."""
    import torch
    from oracle import oracle as O
    from uhc_b200 import motion_lib
    from uhc_b200.engine import Engine
    ex = motion_lib.synthetic_clip(20, np.random.default_rng(5), kind="sitting")
    q = ex["qpos"][0].copy()
    q[7:] = 0
    q[2] = 0.16
    a = np.deg2rad(40.0) / 2
    q[3:7] = [np.cos(a), 0, np.sin(a), 0]
    eng = Engine(2, body_diff_thresh=100.0)
    eng.load_clips([ex], None)
    eng.reset(qpos=np.tile(q, (2, 1)), qvel=np.zeros((2, 75)))
    oe = O.Env(O.Model(), ex, np.zeros(17), body_diff_thresh=100.0)
    oe.reset(q, np.zeros(75))
    assert 32 < oe.d.ncon <= 40 and eng.get_state(1)["ncon"] == oe.d.ncon
    big = 0
    for t in range(3):
        act = np.zeros(105, np.float32)
        eng.step(torch.tensor(np.tile(act, (2, 1)), device="cuda"))
        oe.step(act.astype(np.float64))
        if oe.d.ncon > 40:
            break
        st = eng.get_state(1)
        big = max(big, st["ncon"])
        assert np.abs(st["qpos"] - oe.d.qpos).max() < 1e-3, (t, np.abs(st["qpos"] - oe.d.qpos).max())
    eng.close()


def test_work_sorted_slots_do_not_change_results(golden_dir, monkeypatch):
    """The step kernel assigns environments to warps by the previous step's solver iterations (k_order_envs); outputs are indexed
    by environment id, so a run with the identity assignment must give bit-identical results."""
    import torch
    from uhc_b200.engine import Engine
    z = np.load(os.path.join(golden_dir, "expert_kick.npz"))
    ex = {k: z[k] for k in z.files}
    so = np.concatenate([ex["beta"][0], [ex["gender"][0]]])
    E = 96
    rng = np.random.RandomState(3)
    starts = rng.randint(0, 30, E).astype(np.int32)
    acts = [torch.tensor(rng.normal(0, 0.3, (E, 105)), dtype=torch.float32, device="cuda") for _ in range(4)]
    outs = []
    for sort in ("1", "0"):
        monkeypatch.setenv("UHC_SORT_ENVS", sort)
        eng = Engine(E)
        eng.load_clips([ex], [so])
        eng.reset(start=starts)
        rec = []
        for a in acts:
            o, r, ci, f, en, p = eng.step(a)
            rec.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), f.cpu().numpy().copy()))
        outs.append(rec)
        eng.close()
    for (o1, r1, f1), (o0, r0, f0) in zip(*outs):
        assert np.array_equal(o1, o0) and np.array_equal(r1, r0) and np.array_equal(f1, f0)


def test_ragged_clips_partial_resets_and_clip_end(golden_dir):
    """Edge cases of the batched surface: an env count that does not fill the last CTA (13), clips of very different lengths
    (3 .. ~70 frames), slices with explicit (start, length), re-seeding a SUBSET of
    envs leaves the others untouched, and the `end` flag / percent at the clip end agree with the oracle env."""
    import torch
    from oracle import oracle as O
    from uhc_b200.engine import Engine
    ex, so = _expert(golden_dir, "sway")
    keys = ("qpos", "qvel", "wbpos", "wbquat", "bquat", "bangvel", "ee_wpos", "com")
    Tm = len(ex["qpos"])
    cuts = [(0, 3), (5, 9), (20, Tm - 20)]                              # (first frame, length) of three clips cut from the golden motion
    assert Tm - 20 > 55
    clips = [{k: ex[k][a:a + n] for k in keys} for a, n in cuts]
    E = 13
    eng = Engine(E)
    eng.load_clips(clips, [so] * len(clips))
    clip = np.array([0, 1, 2] * 4 + [1], np.int32)
    start = np.array([0, 2, 10, 1, 0, 30, 0, 4, 50, 0, 1, 0, 3], np.int32)
    length = np.array([cuts[c][1] - s for c, s in zip(clip, start)], np.int32)
    obs = eng.reset(clip=clip, start=start, length=length).cpu().numpy().copy()
    om = O.Model()
    envs = []
    for e in range(E):
        sl = {k: clips[clip[e]][k][start[e]:start[e] + length[e]] for k in keys}
        oe = O.Env(om, sl, so)
        assert np.abs(oe.reset() - obs[e]).max() < 1e-4, e
        envs.append(oe)
    rng = np.random.RandomState(11)
    alive = np.ones(E, bool)
    for t in range(8):
        a = rng.normal(0, 0.05, (E, 105)).astype(np.float32)
        o, r, ci, f, en, p = eng.step(torch.tensor(a, device="cuda"))
        f, en, p = f.cpu().numpy(), en.cpu().numpy(), p.cpu().numpy()
        for e in range(E):
            if not alive[e]:
                continue
            _, ro, done, info = envs[e].step(a[e].astype(np.float64))
            assert bool(f[e]) == info["fail"] and bool(en[e]) == info["end"], (t, e)
            assert abs(p[e] - info["percent"]) < 1e-6
            assert np.abs(eng.get_state(e)["qpos"] - envs[e].d.qpos).max() < 3e-3   # fp32 vs fp64 across contact switches mid-motion
            if done:
                alive[e] = False
        if t == 3:                                                       # re-seed three envs in the middle of the others' episodes
            ids = np.array([1, 6, 12], np.int32)
            before = {e: eng.get_state(e)["qpos"].copy() for e in range(E) if e not in ids}
            eng.reset(ids, clip=np.array([2, 2, 0], np.int32), start=np.array([0, 40, 0], np.int32))
            for e, q in before.items():
                assert np.array_equal(eng.get_state(e)["qpos"], q)
            for e, (c, s) in zip(ids, ((2, 0), (2, 40), (0, 0))):
                sl = {k: clips[c][k][s:] for k in keys}
                envs[e] = O.Env(om, sl, so)
                assert np.abs(envs[e].reset() - eng.obs[e].cpu().numpy()).max() < 1e-4
                alive[e] = True
    assert (~alive).sum() >= 3                                           # the 3-frame and 9-frame clips ended on the way
    with pytest.raises(RuntimeError):                                    # a slice that leaves its clip is an argument error, not a silent read
        eng.reset(np.array([0], np.int32), clip=np.array([1], np.int32), start=np.array([5], np.int32), length=np.array([9], np.int32))
    eng.close()


def _run_explicit(golden_dir, precision, E):
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "env_sway_explicit_noise.npz"))
    ex, so = _expert(golden_dir, "sway")
    eng = Engine(E, precision=precision, rfc_mode="explicit")
    assert eng.act_dim == 315 == g["action"].shape[1]
    eng.load_clips([ex], [so])
    obs0 = eng.reset().cpu().numpy()
    qpos, obs, rew, ci, fail = [], [], [], [], []
    for t in range(len(g["reward"])):
        a = torch.tensor(np.tile(g["action"][t], (E, 1)), dtype=torch.float32, device="cuda")
        o, r, c, f, en, pct = eng.step(a)
        torch.cuda.synchronize()
        qpos.append(eng.get_state(E - 1)["qpos"]); obs.append(o[E - 1].cpu().numpy().copy()); rew.append(float(r[E - 1])); ci.append(c[E - 1].cpu().numpy().copy())
        fail.append(int(f[E - 1]))
    eng.close()
    return g, obs0[E - 1], np.array(qpos), np.array(obs), np.array(rew), np.array(ci), np.array(fail)


@pytest.mark.parametrize("precision,tol_q,tol_o", [(64, 1e-8, 2e-5), (32, 1e-3, 5e-3)])
def test_explicit_residual_force_matches_reference_trace(golden_dir, precision, tol_q, tol_o):
    """residual_force_mode = explicit (config/release/uhc_explicit.yml, humanoid_im.py:1080-1132 + reward_function.py:253-341) through the C ABI:
    per-body contact point / force / torque from a 315-wide action row, applied with the Jacobian of the last forward pass."""
    g, obs0, qpos, obs, rew, ci, fail = _run_explicit(golden_dir, precision, 3)
    first_fail = int(np.argmax(g["fail"]))
    assert np.abs(obs0 - g["obs0"]).max() < 1e-5
    # the random per-body forces topple the humanoid at step ~32; the last steps before the fall amplify round-off (fp32: compared up to 8 steps before it)
    lim = first_fail if precision == 64 else first_fail - 8
    assert np.abs(qpos[:lim] - g["qpos"][:lim]).max() < tol_q
    assert np.abs(obs[:lim - 2] - g["obs"][:lim - 2]).max() < tol_o
    assert np.abs(rew[:lim - 2] - g["reward"][:lim - 2]).max() < tol_o
    assert np.abs(ci[:lim - 2] - g["c_info"][:lim - 2]).max() < tol_o
    assert (fail[:first_fail - 1] == 0).all() and fail[first_fail - 1:first_fail + 2].any()


@pytest.mark.parametrize("precision,tol_q", [(64, 1e-7), (32, 1e-3)])
def test_joint_limit_rows_match_oracle(golden_dir, precision, tol_q):
    """Hinge ranges tightened to +-0.3 rad (smpl_robot.py:1087-1110 tightens knee / ankle ranges per shape; the shipped neutral model has +-180 deg on every
    hinge): the sway clip then drives many joints past their limits, so the limit rows act in every substep.  Kernel against the oracle env on the
    same tables, 12 control steps."""
    import torch
    from oracle import oracle as O
    from uhc_b200.engine import Engine
    from uhc_b200.model import HumanoidModel
    ex, so = _expert(golden_dir, "sway")
    jr = np.tile(np.array([[-0.3, 0.3]]), (69, 1))
    eng = Engine(2, model=HumanoidModel(jnt_range=jr), precision=precision)
    eng.load_clips([ex], [so])
    obs0 = eng.reset().cpu().numpy()
    oe = O.Env(O.Model(tables={"jnt_range": jr}), ex, so)
    assert np.abs(oe.reset() - obs0[1]).max() < 1e-4
    assert ((ex["qpos"][0, 7:] < -0.3) | (ex["qpos"][0, 7:] > 0.3)).sum() >= 3          # the clip's first frame already violates several limits
    rng = np.random.RandomState(3)
    worst = 0.0
    for t in range(12):
        a = rng.normal(0, 0.1, 105); a[69:75] *= 0.3
        eng.step(torch.tensor(np.tile(a, (2, 1)), dtype=torch.float32, device="cuda"))
        torch.cuda.synchronize()
        oo, ro, done, info = oe.step(a)
        worst = max(worst, np.abs(eng.get_state(1)["qpos"] - oe.d.qpos).max())
        if done:
            break
    assert t >= 5 and worst < tol_q, (t, worst)
    eng.close()


@pytest.mark.parametrize("precision,tol_q,tol_o", [(64, 1e-8, 2e-5), (32, 1e-3, 5e-3)])
def test_obs_v3_no_shape_no_residual_force_matches_reference_trace(golden_dir, precision, tol_q, tol_o):
    """config/meta_pd/copycat_35.yml at the env level through the C ABI: obs_v 3 (five 640-wide v2 blocks ten frames apart, humanoid_im.py:505-513), has_shape
    false, residual_force false (99-wide actions, residual-force reward term 0)."""
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "env_sway_obsv3_noise.npz"))
    ex, so = _expert(golden_dir, "sway")
    eng = Engine(2, precision=precision, obs_v=3, fut_frames=5, fut_skip=10, has_shape=False, rfc_mode="none")
    assert eng.obs_dim == 3200 == g["obs"].shape[1] and eng.act_dim == 99 == g["action"].shape[1]
    eng.load_clips([ex], [so])
    obs0 = eng.reset().cpu().numpy()
    assert np.abs(obs0[1] - g["obs0"]).max() < 1e-5
    for t in range(len(g["reward"])):
        a = torch.tensor(np.tile(g["action"][t], (2, 1)), dtype=torch.float32, device="cuda")
        o, r, c, f, en, pct = eng.step(a)
        torch.cuda.synchronize()
        assert np.abs(eng.get_state(1)["qpos"] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(o[1].cpu().numpy() - g["obs"][t]).max() < tol_o, t
        assert abs(float(r[1]) - g["reward"][t]) < tol_o and float(c[1, 4]) == 0.0
    eng.close()


@pytest.mark.parametrize("term", ["root", "Head"])
def test_env_term_body_root_and_head_match_reference_flags(golden_dir, term):
    """env_term_body root / Head through the C ABI (fp64 kernels, so that the crossing step is the reference's): the fail flags of the reference's own step()"""
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, f"env_sway_term{term.lower()}_noise.npz"))
    ex, so = _expert(golden_dir, "sway")
    E = 3
    eng = Engine(E, precision=64, term_body=term, head_body=int(g["head_idx"]) if term == "Head" else 13)
    eng.load_clips([ex], [so])
    eng.reset()
    fails = []
    for t in range(len(g["fail"])):
        a = torch.tensor(np.tile(g["action"][t], (E, 1)), dtype=torch.float32, device="cuda")
        o, r, c, f, en, pct = eng.step(a)
        torch.cuda.synchronize()
        fails.append(bool(int(f[E - 1])))
        assert abs(eng.get_state(E - 1)["qpos"][2] - g["root_z"][t]) < 1e-6
    eng.close()
    assert fails == [bool(x) for x in g["fail"]]


@pytest.mark.parametrize("precision,tol", [(64, 1e-6), (32, 5e-3)])
def test_multiplicative_reward_matches_reference_trace(golden_dir, precision, tol):
    """reward_id world_rfc_implicit_v1_mul through the C ABI: the reference's reward / c_info over the noise trajectory (before the fall)"""
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "env_sway_rewmul_noise.npz"))
    ex, so = _expert(golden_dir, "sway")
    E = 2
    eng = Engine(E, precision=precision, reward_mul=True)
    eng.load_clips([ex], [so])
    eng.reset()
    for t in range(20):
        a = torch.tensor(np.tile(g["action"][t], (E, 1)), dtype=torch.float32, device="cuda")
        o, r, c, f, en, pct = eng.step(a)
        torch.cuda.synchronize()
        assert abs(float(r[E - 1]) - g["reward"][t]) < tol, t
        assert np.abs(c[E - 1].cpu().numpy() - g["c_info"][t]).max() < tol, t
    eng.close()


@pytest.mark.parametrize("v,precision,tol_q,tol_o", [(5, 64, 1e-8, 2e-5), (6, 64, 1e-8, 2e-5), (5, 32, 1e-3, 5e-3), (6, 32, 1e-3, 5e-3)])
def test_obs_v5_v6_match_reference_trace(golden_dir, v, precision, tol_q, tol_o):
    """obs_v 5 / 6 (get_full_obs_v5 / get_full_obs_v6, humanoid_im.py:505-666) through the C ABI against the reference's own observations"""
    import torch
    from uhc_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, f"env_sway_obsv{v}_noise.npz"))
    ex, so = _expert(golden_dir, "sway")
    E = 3
    eng = Engine(E, precision=precision, obs_v=v)
    assert eng.obs_dim == g["obs"].shape[1]
    eng.load_clips([ex], [so])
    obs0 = eng.reset().cpu().numpy()
    assert np.abs(obs0[E - 1] - g["obs0"]).max() < 1e-5
    for t in range(len(g["reward"])):
        a = torch.tensor(np.tile(g["action"][t], (E, 1)), dtype=torch.float32, device="cuda")
        o, r, c, f, en, pct = eng.step(a)
        torch.cuda.synchronize()
        assert np.abs(eng.get_state(E - 1)["qpos"] - g["qpos"][t]).max() < tol_q, t
        assert np.abs(o[E - 1].cpu().numpy() - g["obs"][t]).max() < tol_o, t
    eng.close()
