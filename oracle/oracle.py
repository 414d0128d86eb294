"""ctypes front-end of the CPU oracle (oracle/uhc_oracle.c).  TEST INFRASTRUCTURE -- see the header of that file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libuhc_oracle.so")
MODEL_NPZ = os.path.join(_HERE, "..", "uhc_b200", "assets", "smpl_neutral_model.npz")

NB, NQ, NV, NU, OBS_DIM, ACT_DIM = 24, 76, 75, 69, 657, 105
_F = {"qpos": (0, NQ), "qvel": (1, NV), "qacc_warm": (2, NV), "xpos": (3, NB * 3), "xquat": (4, NB * 4),
      "xipos": (5, NB * 3), "M": (6, NV * NV), "C": (7, NV), "qacc": (8, NV), "ctrl": (9, NU),
      "qfrc_applied": (10, NV), "con_pos": (11, 96 * 3), "con_dist": (12, 96), "efc_force": (13, 384),
      "qacc_smooth": (14, NV), "xmat": (15, NB * 9)}


def build(force=False):
    src = os.path.join(_HERE, "uhc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.or_model_create.restype = C.c_void_p
        L.or_data_create.restype = C.c_void_p
        L.or_env_create.restype = C.c_void_p
        L.or_env_create.argtypes = [C.c_void_p]
        L.or_env_data.restype = C.c_void_p
        L.or_env_data.argtypes = [C.c_void_p]
        L.or_field.restype = C.POINTER(C.c_double)
        L.or_field.argtypes = [C.c_void_p, C.c_int]
        for f in ("or_env_bquat", "or_env_prev_bquat", "or_env_torque"):
            getattr(L, f).restype = C.POINTER(C.c_double)
            getattr(L, f).argtypes = [C.c_void_p]
        L.or_reward.restype = C.c_double
        L.or_body_diff.restype = C.c_double
        L.or_body_diff.argtypes = [C.c_void_p]
        L.or_energy.restype = C.c_double
        for f in ("or_forward", "or_step"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Model:
    def __init__(self, npz=MODEL_NPZ, tables=None):
        """tables: optional dict overriding body_offset / body_mass / body_ipos / body_inertia / body_invweight0 / hull_vert
        (a body-shape variant with the same topology)."""
        z0 = np.load(npz)
        z = {k: z0[k] for k in z0.files}
        if tables:
            z.update(tables)
        self.z = dict(z)
        f = lambda k: np.ascontiguousarray(z[k], dtype=np.float64)
        i = lambda k: np.ascontiguousarray(z[k], dtype=np.int32)
        self._keep = [i("parent"), f("body_offset"), f("body_mass"), f("body_ipos"), f("body_inertia"),
                      np.ascontiguousarray(z["body_invweight0"][:, 0]), f("armature"), f("jkp"), f("jkd"),
                      f("torque_lim"), f("diffw"), i("ee_body"), f("hull_vert"), i("hull_vadr"), i("hull_vnum"),
                      i("hull_nbr"), i("hull_nbradr"), f("solref"), f("solimp"), f("gravity")]
        k = self._keep
        self.h = C.c_void_p(lib().or_model_create(
            _p(k[0], C.c_int), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]),
            _p(k[10]), _p(k[11], C.c_int), C.c_int(len(k[12])), _p(k[12]), _p(k[13], C.c_int), _p(k[14], C.c_int),
            _p(k[15], C.c_int), _p(k[16], C.c_int), C.c_double(float(z["timestep"])), C.c_double(float(z["margin"])),
            C.c_double(float(z["friction"])), _p(k[17]), _p(k[18]), _p(k[19])))
        # joint limits: ranges from the model (xml: every hinge +-180 deg), overridable (tables["jnt_range"]) as smpl_robot.py:1087-1110 tightens them per shape
        if "jnt_range" in z and "dof_invweight0" in z:
            self._lim = (f("jnt_range"), f("dof_invweight0"))
            lib().or_model_set_limits(self.h, _p(self._lim[0]), _p(self._lim[1]))
        self.dt = float(z["timestep"])
        self.qpos0 = np.zeros(NQ)
        self.qpos0[:3] = z["body_gpos"][0]
        self.qpos0[3] = 1.0


class Data:
    """View over an OrData (owned unless `handle` is given)."""

    def __init__(self, handle=None):
        self.h = C.c_void_p(handle if handle is not None else lib().or_data_create())

    def __getattr__(self, name):
        if name in _F:
            fid, n = _F[name]
            return np.ctypeslib.as_array(lib().or_field(self.h, fid), shape=(n,))
        raise AttributeError(name)

    @property
    def ncon(self):
        return lib().or_ncon(self.h)

    def con_body(self, i):
        return lib().or_con_body(self.h, i)

    @property
    def newton_iters(self):
        return lib().or_newton_iters(self.h)


def forward(m, d):
    lib().or_forward(m.h, d.h)


def step(m, d):
    lib().or_step(m.h, d.h)


def energy(m, d):
    mom = np.zeros(3)
    e = lib().or_energy(m.h, d.h, _p(mom))
    return e, mom


class Env:
    """Single-environment imitation env on the oracle (mirrors HumanoidEnv.reset/step + reward)."""

    def __init__(self, model, expert, shape_obs=None, **cfg):
        self.m = model
        self.h = C.c_void_p(lib().or_env_create(model.h))
        self.d = Data(lib().or_env_data(self.h))
        self.configure(**cfg)
        self.load_expert(expert, shape_obs)

    def configure(self, base_rot=(0.7071, 0.7071, 0.0, 0.0), rfc_scale=100.0, rfc_lim=100.0, rfc_rate=1.0,
                  body_diff_thresh=0.5, meta_pd=1, env_episode_len=100000, trail_steps=0,
                  w=(0.3, 0.1, 0.45, 0.1, 0.05), k=(2.0, 0.005, 5.0, 100.0, 1.0)):
        br, w_, k_ = (np.array(x, dtype=np.float64) for x in (base_rot, w, k))
        lib().or_env_config(self.h, _p(br), C.c_double(rfc_scale), C.c_double(rfc_lim), C.c_double(rfc_rate),
                            C.c_double(body_diff_thresh), C.c_int(meta_pd), C.c_int(env_episode_len),
                            C.c_int(trail_steps), _p(w_), _p(k_))

    # residual-force slot order of the explicit mode: vf_bodies = SMPL_BONE_ORDER_NAMES (humanoid_im.py:236-237, smpl_parser.py:11-36)
    SMPL_BONE_ORDER = ["Pelvis", "L_Hip", "R_Hip", "Torso", "L_Knee", "R_Knee", "Spine", "L_Ankle", "R_Ankle", "Chest", "L_Toe", "R_Toe", "Neck", "L_Thorax",
                       "R_Thorax", "Head", "L_Shoulder", "R_Shoulder", "L_Elbow", "R_Elbow", "L_Wrist", "R_Wrist", "L_Hand", "R_Hand"]

    def set_rfc_mode(self, explicit):
        """residual_force_mode "explicit" (per-body contact point / force / torque, action dim 69 + 216 + 30) or "implicit" (root wrench, 105)"""
        names = [str(n) for n in self.m.z["body_names"]]
        vf_body = np.array([names.index(n) for n in self.SMPL_BONE_ORDER], dtype=np.int32)
        mode = 2 if explicit in (2, "none", None) else int(bool(explicit))         # "none": cfg.residual_force false (no residual-force dims)
        lib().or_env_set_rfc_mode(self.h, C.c_int(mode), vf_body.ctypes.data_as(C.POINTER(C.c_int)))
        self.vf_body = vf_body

    def set_has_shape(self, has_shape):
        lib().or_env_set_has_shape(self.h, C.c_int(int(bool(has_shape))))

    def set_reward_mul(self, on=True):
        """reward_id world_rfc_implicit_v1_mul: the product of the five terms (reward_function.py:174-250)"""
        lib().or_env_set_reward_mul(self.h, C.c_int(int(bool(on))))

    def set_term_body(self, term_body, head_body=13):
        """cfg.env_term_body (humanoid_im.py:1223-1229): "body" (default), "root" or "Head"; the height bounds are minima over the loaded expert"""
        lib().or_env_set_term_body(self.h, C.c_int({"body": 0, "root": 1, "Head": 2}[term_body]), C.c_int(int(head_body)))

    @property
    def action_dim(self):
        return lib().or_env_action_dim(self.h)

    def set_obs_v(self, obs_v, fut_frames=10, skip=10):
        """cfg.obs_v: 2 (default, 657 dims), 1 (get_full_obs_v1, 784 dims; needs expert["body_com"]) or 3 (get_full_obs_v3: fut_frames v2 blocks)"""
        self._obs_v = int(obs_v)
        lib().or_env_set_future(self.h, C.c_int(int(fut_frames)), C.c_int(int(skip)))
        bc = getattr(self, "_body_com", None)
        assert self._obs_v != 1 or bc is not None, "obs_v 1 needs expert['body_com']"
        lib().or_env_set_obs_v(self.h, C.c_int(self._obs_v), _p(bc) if bc is not None else None)

    @property
    def obs_dim(self):
        return lib().or_env_obs_dim(self.h)

    def load_expert(self, ex, shape_obs=None):
        keys = ["qpos", "qvel", "wbpos", "wbquat", "bquat", "bangvel", "ee_wpos", "com"]
        self._body_com = np.ascontiguousarray(np.asarray(ex["body_com"], dtype=np.float64).reshape(len(ex["qpos"]), 72)) if "body_com" in ex else None
        if getattr(self, "_obs_v", 2) == 1:
            self.set_obs_v(1)
        self._ex = [np.ascontiguousarray(ex[k], dtype=np.float64) for k in keys]
        so = np.zeros(17) if shape_obs is None else np.asarray(shape_obs, dtype=np.float64)
        self._so = np.ascontiguousarray(so)
        lib().or_env_set_expert(self.h, C.c_int(len(self._ex[0])), *[_p(a) for a in self._ex], _p(self._so))

    def reset(self, qpos=None, qvel=None):
        obs = np.zeros(self.obs_dim)
        qp = None if qpos is None else _p(np.ascontiguousarray(qpos, dtype=np.float64))
        qv = None if qvel is None else _p(np.ascontiguousarray(qvel, dtype=np.float64))
        lib().or_env_reset(self.h, qp, qv, _p(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(self.obs_dim)
        fail, end, pct = C.c_int(0), C.c_int(0), C.c_double(0)
        done = lib().or_env_step(self.h, _p(a), _p(obs), C.byref(fail), C.byref(end), C.byref(pct))
        cinfo = np.zeros(5)
        r = lib().or_reward(self.h, _p(a), _p(cinfo))
        return obs, r, bool(done), {"fail": bool(fail.value), "end": bool(end.value), "percent": pct.value,
                                    "c_info": cinfo}

    @property
    def cur_t(self):
        return lib().or_env_cur_t(self.h)

    @property
    def bquat(self):
        return np.ctypeslib.as_array(lib().or_env_bquat(self.h), shape=(96,))

    @property
    def prev_bquat(self):
        return np.ctypeslib.as_array(lib().or_env_prev_bquat(self.h), shape=(96,))

    @property
    def torque(self):
        return np.ctypeslib.as_array(lib().or_env_torque(self.h), shape=(15, NU))

    def body_diff(self):
        return lib().or_body_diff(self.h)
