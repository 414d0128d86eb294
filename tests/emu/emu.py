"""ctypes wrapper of the host emulation build of the product kernels (TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

from uhc_b200.model import HumanoidModel

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libuhc_emu.so")
ST = dict(Q=0, V=76, AW=152, C=228, IB=304, S=544, XPOS=996, XQUAT=1068, XIPOS=1164, BQUAT=1236, PBQUAT=1332, SIZE=1428)
EX_SIZE = 576


class UhcEnvCfg(C.Structure):
    _fields_ = [("base_rot", C.c_double * 4), ("rfc_scale", C.c_double), ("rfc_lim", C.c_double), ("rfc_rate", C.c_double),
                ("body_diff_thresh", C.c_double), ("meta_pd", C.c_int), ("env_episode_len", C.c_int), ("trail_steps", C.c_int),
                ("newton_max_iter", C.c_int), ("w", C.c_double * 5), ("k", C.c_double * 5), ("newton_tol", C.c_double),
                ("auto_reset", C.c_int), ("t_min", C.c_int), ("t_max", C.c_int), ("reactive_v", C.c_int), ("reset_seed", C.c_ulonglong),
                ("reactive_rate", C.c_double), ("rfc_mode", C.c_int), ("vf_slot", C.c_int * 24), ("obs_v", C.c_int), ("fut_frames", C.c_int), ("fut_skip", C.c_int), ("no_shape", C.c_int), ("term_body", C.c_int), ("head_body", C.c_int), ("reward_mul", C.c_int)]


def default_cfg(precision=32, **kw):
    c = UhcEnvCfg()
    c.base_rot = (C.c_double * 4)(0.7071, 0.7071, 0.0, 0.0)
    c.rfc_scale, c.rfc_lim, c.rfc_rate, c.body_diff_thresh = 100.0, 100.0, 1.0, 0.5
    c.meta_pd, c.env_episode_len, c.trail_steps = 1, 100000, 0
    c.newton_max_iter = 20 if precision == 64 else 12
    c.w = (C.c_double * 5)(0.3, 0.1, 0.45, 0.1, 0.05)
    c.k = (C.c_double * 5)(2.0, 0.005, 5.0, 100.0, 1.0)
    c.newton_tol = 1e-11 if precision == 64 else 1e-5
    c.vf_slot = (C.c_int * 24)(*range(24))
    c.obs_v = 2
    c.term_body, c.head_body = 0, 13
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def pack_expert(ex):
    """expert dict (qpos, qvel, wbpos, wbquat, bquat, bangvel, ee_wpos, com) -> [T][508] record array."""
    T = len(ex["qpos"])
    out = np.zeros((T, EX_SIZE))
    o = 0
    for k, n in (("qpos", 76), ("qvel", 75), ("wbpos", 72), ("wbquat", 96), ("bquat", 96), ("bangvel", 72), ("ee_wpos", 15), ("com", 3)):
        out[:, o:o + n] = np.asarray(ex[k]).reshape(T, n)
        o += n
    if "body_com" in ex:
        out[:, 502:574] = np.asarray(ex["body_com"]).reshape(T, 72)
    return out


def build():
    srcs = [os.path.join(_HERE, "emu.cpp"), os.path.join(_HERE, "..", "..", "uhc_b200", "csrc", "sim_core.h"),
            os.path.join(_HERE, "..", "..", "uhc_b200", "csrc", "env_step.h")]
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, srcs[0]])
    return _SO


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Emu:
    def __init__(self, precision=64, num_envs=1, model=None, **cfg):
        self.lib = C.CDLL(build())
        self.lib.emu_create.restype = C.c_void_p
        self.prec = precision
        self.model = model or HumanoidModel()
        self._ms = self.model.host_struct()
        self._cfg = default_cfg(precision, **cfg)
        blk = 640 if self._cfg.no_shape else 657
        self.obs_dim = 784 if self._cfg.obs_v == 1 else (blk * (self._cfg.fut_frames or 10) if self._cfg.obs_v == 3 else blk)
        if self._cfg.obs_v in (5, 6):
            self.obs_dim = (636 if self._cfg.obs_v == 5 else 384) + (0 if self._cfg.no_shape else 17)
        self.h = C.c_void_p(self.lib.emu_create(C.byref(self._ms), C.byref(self._cfg), C.c_int(num_envs), C.c_int(precision)))

    def load_clips(self, experts, shapes):
        lens = np.array([len(e["qpos"]) for e in experts], np.int32)
        frames = np.ascontiguousarray(np.concatenate([pack_expert(e) for e in experts]))
        shp = np.ascontiguousarray(np.asarray(shapes, dtype=np.float64).reshape(len(experts), 17))
        self.lens = lens
        self.lib.emu_load_clips(self.h, self.prec, len(experts), _p(lens, C.c_int), _p(frames), _p(shp))

    def reset(self, env=0, clip=0, start=0, length=None, qpos=None, qvel=None):
        obs = np.zeros(self.obs_dim)
        L = int(self.lens[clip]) - start if length is None else length      # Engine.reset: the rest of the clip from `start`
        q = None if qpos is None else np.ascontiguousarray(qpos, dtype=np.float64)
        v = None if qvel is None else np.ascontiguousarray(qvel, dtype=np.float64)
        self.lib.emu_reset(self.h, self.prec, env, clip, start, L, _p(q), _p(v), _p(obs))
        return obs

    def step(self, action, env=0):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs, rew, ci, pct, tq = np.zeros(self.obs_dim), np.zeros(1), np.zeros(5), np.zeros(1), np.zeros((15, 69))
        fail, end = C.c_int(0), C.c_int(0)
        done = self.lib.emu_step(self.h, self.prec, env, _p(a), _p(obs), _p(rew), _p(ci), C.byref(fail), C.byref(end), _p(pct), _p(tq))
        return obs, float(rew[0]), bool(done), {"fail": bool(fail.value), "end": bool(end.value), "percent": float(pct[0]), "c_info": ci, "torque": tq}

    def state(self, env=0):
        out, iout = np.zeros(ST["SIZE"]), np.zeros(8, np.int32)
        self.lib.emu_get_state(self.h, self.prec, env, _p(out), _p(iout, C.c_int))
        return out, iout

    def forward(self, qpos, qvel, tau=None, fapp=None, aw=None):
        tau = np.zeros(69) if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        fapp = np.zeros(6) if fapp is None else np.ascontiguousarray(fapp, dtype=np.float64)
        aw_ = None if aw is None else np.ascontiguousarray(aw, dtype=np.float64)
        Ms, Cc, qa, xp = np.zeros(1221), np.zeros(75), np.zeros(75), np.zeros(72)
        ncon, iters = C.c_int(0), C.c_int(0)
        self.lib.emu_forward(self.h, self.prec, _p(np.ascontiguousarray(qpos, dtype=np.float64)), _p(np.ascontiguousarray(qvel, dtype=np.float64)),
                             _p(tau), _p(fapp), _p(aw_), _p(Ms), _p(Cc), _p(qa), _p(xp), C.byref(ncon), C.byref(iters))
        return dict(C=Cc, qacc=qa, xpos=xp.reshape(24, 3), ncon=ncon.value, iters=iters.value)
