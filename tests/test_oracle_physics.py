"""CPU: physical invariants the L0 restatement is held to in lieu of MuJoCo."""
import os
import tempfile

import numpy as np

from oracle import oracle as O


def _random_state(m, rng, z=3.0):
    q = m.qpos0.copy()
    q[2] = z
    q[3:7] = rng.normal(size=4)
    q[3:7] /= np.linalg.norm(q[3:7])
    q[7:] = rng.uniform(-0.6, 0.6, 69)
    return q


def test_total_mass_and_mass_matrix_spd():
    m, d = O.Model(), O.Data()
    assert abs(m.z["body_mass"].sum() - 80.29) < 0.05          # 80.3 kg at density 1000
    d.qpos[:] = _random_state(m, np.random.default_rng(0))
    O.forward(m, d)
    M = d.M.reshape(75, 75)
    assert np.abs(M - M.T).max() < 1e-12
    assert np.linalg.eigvalsh(M).min() > 0.009                  # >= armature
    assert abs(M[0, 0] - m.z["body_mass"].sum()) < 1e-9


def test_free_fall_is_minus_g():
    m, d = O.Model(), O.Data()
    d.qpos[:] = _random_state(m, np.random.default_rng(1))
    O.forward(m, d)
    assert d.ncon == 0
    assert abs(d.qacc[2] + 9.81) < 1e-9
    assert np.abs(np.delete(d.qacc, 2)).max() < 1e-9


def test_energy_and_momentum_conservation_contact_free():
    z = dict(np.load(O.MODEL_NPZ))
    z["timestep"] = np.float64(1e-5)
    p = os.path.join(tempfile.mkdtemp(), "m.npz")
    np.savez(p, **z)
    m, d = O.Model(p), O.Data()
    rng = np.random.default_rng(2)
    d.qpos[:] = _random_state(m, rng)
    d.qvel[:] = rng.normal(size=75)
    e0, p0 = O.energy(m, d)
    n = 1500
    for _ in range(n):
        O.step(m, d)
    e1, p1 = O.energy(m, d)
    assert abs(e1 - e0) / abs(e0) < 2e-6                       # C(q,v) consistent with M(q)
    assert np.abs(p1[:2] - p0[:2]).max() < 1e-3
    assert abs((p1[2] - p0[2]) - (-9.81 * z["body_mass"].sum() * n * 1e-5)) < 1e-3


def test_resting_contact_supports_weight_without_deep_penetration():
    import joblib  # noqa: F401  (not needed: qpos below is a plain array)
    m, d = O.Model(), O.Data()
    q = m.qpos0.copy()
    q[2] = 0.97
    q[3:7] = [0.7071068, 0.7071068, 0, 0]                      # Y-up rest pose -> Z-up world (base_rot)
    d.qpos[:] = q
    for _ in range(1800):
        O.step(m, d)
    assert d.ncon > 0
    f = d.efc_force[:4 * d.ncon].sum()
    assert abs(f - 9.81 * m.z["body_mass"].sum()) / (9.81 * 80.29) < 0.05
    assert d.con_dist[:d.ncon].min() > -0.02
    assert np.abs(d.qvel[:3]).max() < 0.05                     # undamped limp limbs may still jiggle; the root rests


def test_contact_solution_is_consistent_with_its_forces():
    """At the converged constraint solution the constraint force M (a - a_smooth) equals J^T f with f >= 0.  Without exposing J this is
    checked on the root translation rows, where J^T f is the plain sum of the pyramid-edge forces f_e d_e over all contacts
    (edges d = (0, mu, 1), (0, -mu, 1), (-mu, 0, 1), (mu, 0, 1) in the floor frame), and on the sign of every row force."""
    m, d = O.Model(), O.Data()
    mu = float(m.z["friction"])
    rng = np.random.default_rng(42)
    seen = 0
    for case in range(12):
        q = m.qpos0.copy()
        q[2] = rng.uniform(0.3, 0.93)
        q[3:7] = np.array([0.7071068, 0.7071068, 0, 0]) + rng.normal(size=4) * 0.1
        q[3:7] /= np.linalg.norm(q[3:7])
        q[7:] = rng.uniform(-0.5, 0.5, 69)
        d.qpos[:], d.qvel[:], d.ctrl[:] = q, rng.normal(size=75) * 0.5, rng.normal(size=69) * 20
        d.qfrc_applied[:] = 0
        d.qacc_warm[:] = 0
        O.forward(m, d)
        if d.ncon == 0:
            continue
        seen += 1
        f = d.efc_force[:4 * d.ncon].reshape(d.ncon, 4)
        assert f.min() >= 0.0
        total = np.array([mu * (f[:, 3] - f[:, 2]).sum(), mu * (f[:, 0] - f[:, 1]).sum(), f.sum()])
        lhs = (d.M.reshape(75, 75) @ (d.qacc - d.qacc_smooth))[:3]
        assert np.abs(lhs - total).max() < 1e-6 * max(1.0, np.abs(total).max()), (case, lhs, total)
        assert total[2] > 0                                       # the floor only pushes
    assert seen >= 5


def test_joint_limit_row_matches_closed_form():
    """One hinge past a tightened limit, no contacts: the constrained acceleration has the closed form of a single soft unilateral row
    (min 1/2 (a - a_s)^T M (a - a_s) + 1/2 D (sg a_i - aref)_-^2  =>  r = (sg a_s,i - aref) / (1 + (M^-1)_ii D), a = a_s - M^-1 e_i sg D r), with
    MuJoCo's default solref / solimp and diagApprox = dof_invweight0; inside the range nothing changes."""
    z = np.load(O.MODEL_NPZ)
    rng = np.random.default_rng(5)
    jr = z["jnt_range"].copy()
    knee_x = 3 * 1 + 2                      # hinge index of L_Knee_x (body 2 = L_Knee: hinges 3..5, order z, y, x)
    jr[knee_x] = [-0.1, 2.0]
    om, d = O.Model(tables={"jnt_range": jr}), O.Data()
    for q_knee, sg in ((-0.3, 1.0), (2.25, -1.0), (0.7, 0.0)):
        q = om.qpos0.copy(); q[2] = 3.0
        q[7:] = rng.uniform(-0.2, 0.2, 69); q[7 + knee_x] = q_knee
        v = rng.normal(size=75) * 0.3
        d.qpos[:], d.qvel[:], d.ctrl[:] = q, v, rng.normal(size=69) * 5
        d.qfrc_applied[:] = 0; d.qacc_warm[:] = 0
        O.forward(om, d)
        assert d.ncon == 0
        a_s, M = np.array(d.qacc_smooth), np.array(d.M).reshape(75, 75)
        if sg == 0:
            np.testing.assert_allclose(d.qacc, a_s, atol=1e-12)
            continue
        i = 6 + knee_x
        dist = (q_knee - jr[knee_x, 0]) if sg > 0 else (jr[knee_x, 1] - q_knee)
        assert dist < 0
        dmin, dmax, width, mid, power = z["solimp"]; tc, dr = z["solref"]
        x = min(abs(dist) / width, 1.0)
        y = (x / mid) ** power * mid if x < mid else 1 - ((1 - x) / (1 - mid)) ** power * (1 - mid)
        imp = dmin + y * (dmax - dmin)
        D = 1.0 / ((1 - imp) * z["dof_invweight0"][i] / imp)
        kk, bb = 1.0 / (dmax * dmax * tc * tc * dr * dr), 2.0 / (dmax * tc)
        aref = -bb * sg * v[i] - kk * imp * dist
        Minv_i = np.linalg.solve(M, np.eye(75)[i])
        r = (sg * a_s[i] - aref) / (1 + Minv_i[i] * D)
        assert r < 0                                                        # the row is active: the limit pushes the joint back
        np.testing.assert_allclose(d.qacc, a_s - Minv_i * sg * D * r, rtol=0, atol=1e-8 * max(1.0, np.abs(d.qacc).max()))
        assert sg * (d.qacc[i] - a_s[i]) > 0
