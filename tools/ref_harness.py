#!/usr/bin/env python
"""Run the reference's OWN Python (unmodified, imported from /root/reference) above the L0 boundary.

The reference needs mujoco-py/MuJoCo 2.1.0, gym, smplx, lxml, vtk, ... none of which exist in this container.
This harness (used ONLY by tools/make_golden.py, in this container) installs:
  * a functional fake `mujoco_py` whose MjSim is backed by the CPU oracle physics (oracle/uhc_oracle.c:
    or_forward / or_step) -- so every line of uhc/envs/humanoid_im.py (stable PD, RFC, obs, termination),
    uhc/losses/reward_function.py, uhc/smpllib/torch_smpl_humanoid.py, smpl_to_qpose, the dataset sampler and
    khrylib's PPO runs exactly as shipped;
  * inert stubs for the other missing third-party modules and for the SMPL robot builder
    (uhc/smpllib/smpl_robot.py needs the licence-gated SMPL files).
Nothing here is product code and nothing here runs on the GPU box.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types
from collections import namedtuple

import ctypes as C

import numpy as np

REF = os.environ.get("UHC_REFERENCE", "/root/reference")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

STUB_TOP = ["glfw", "smplx", "skimage", "imageio", "lxml", "stl", "vtk", "vtkmodules", "fasteners", "ipdb",
            "OpenGL", "mujoco", "pyvista", "open3d", "chumpy", "numpy_stl"]


class _AnyMeta(type):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _AnyMeta(n, (_Any,), {})


class _Any(metaclass=_AnyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Any,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_TOP or fullname.startswith("mujoco_py."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ----------------------------------------------------------------------------- fake mujoco_py on the oracle
class FakeOpt:
    def __init__(self, dt):
        self.timestep = dt


class FakeModel:
    def __init__(self, om):
        z = om.z
        names = [str(n) for n in z["body_names"]]
        self.om = om
        self.body_names = tuple(["world"] + names)
        self._body_name2id = {n: i for i, n in enumerate(self.body_names)}
        self.body_pos = np.vstack([np.zeros(3), z["body_offset"]])
        self.body_ipos = np.vstack([np.zeros(3), z["body_ipos"]])
        self.body_parentid = np.concatenate([[0], z["parent"] + 1]).astype(np.int32)
        self.body_jntadr = np.array([-1, 0] + [1 + 3 * b for b in range(23)], dtype=np.int32)
        self.body_jntnum = np.array([0, 1] + [3] * 23, dtype=np.int32)
        self.jnt_qposadr = np.array([0] + [7 + i for i in range(69)], dtype=np.int32)
        self.jnt_dofadr = np.array([0] + [6 + i for i in range(69)], dtype=np.int32)
        self.nq, self.nv, self.nu = 76, 75, 69
        self.actuator_ctrlrange = np.zeros((69, 2))
        self.actuator_names = tuple(f"{n}_{a}" for n in names[1:] for a in "zyx")
        self.joint_names = tuple([names[0]] + list(self.actuator_names))
        self.geom_bodyid = np.arange(25, dtype=np.int32)
        self.body_mass = np.concatenate([[0], z["body_mass"]])
        self.opt = FakeOpt(om.dt)
        self.jnt_stiffness = np.zeros(70)
        self.dof_damping = np.zeros(75)

        class _Stat:
            extent = 3.0
        self.stat = _Stat()


MjSimState = namedtuple("MjSimState", ["time", "qpos", "qvel", "act", "udd_state"])


class FakeData:
    def __init__(self, sim):
        self._sim = sim
        d = sim.d
        self.qpos, self.qvel, self.ctrl, self.qfrc_applied = d.qpos, d.qvel, d.ctrl, d.qfrc_applied
        self.qfrc_bias = d.C
        self.qM = sim  # token consumed by functions.mj_fullM
        self.time = 0.0

    def _w(self, a, n, first):
        return np.vstack([np.asarray(first, dtype=np.float64)[None], np.array(a).reshape(24, n)])

    @property
    def body_xpos(self):
        return self._w(self._sim.d.xpos, 3, [0, 0, 0])

    @property
    def body_xquat(self):
        return self._w(self._sim.d.xquat, 4, [1, 0, 0, 0])

    @property
    def xipos(self):
        return self._w(self._sim.d.xipos, 3, [0, 0, 0])

    def get_body_xipos(self, name):
        return self.xipos[self._sim.model._body_name2id[name]]

    def get_body_xpos(self, name):
        return self.body_xpos[self._sim.model._body_name2id[name]]

    def get_body_xquat(self, name):
        return self.body_xquat[self._sim.model._body_name2id[name]]

    def get_body_xmat(self, name):      # pose of the last forward pass, like data.body_xmat
        return np.array(self._sim.d.xmat).reshape(24, 3, 3)[self._sim.model._body_name2id[name] - 1].copy()

    @property
    def contact(self):                  # geom 0 = the floor, geom i = body i (FakeModel.geom_bodyid)
        d = self._sim.d
        Cn = namedtuple("Contact", ["geom1", "geom2"])
        return [Cn(0, int(d.con_body(i)) + 1) for i in range(d.ncon)]

    @property
    def ncon(self):
        return self._sim.d.ncon


class FakeSim:
    def __init__(self, model):
        self.model = model
        self.d = O.Data()
        self.data = FakeData(self)
        self.reset()

    def reset(self):  # mj_resetData: qpos <- qpos0, everything else zero; no forward pass
        for k in ("qpos", "qvel", "qacc_warm", "xpos", "xquat", "xipos", "M", "C", "qacc", "ctrl", "qfrc_applied"):
            getattr(self.d, k)[:] = 0
        self.d.qpos[:] = self.model.om.qpos0

    def forward(self):
        O.forward(self.model.om, self.d)

    def step(self):
        O.step(self.model.om, self.d)

    def get_state(self):
        return MjSimState(0.0, self.d.qpos.copy(), self.d.qvel.copy(), None, {})

    def set_state(self, s):
        self.d.qpos[:] = s.qpos
        self.d.qvel[:] = s.qvel


def _install_fake_mujoco(om):
    mp = types.ModuleType("mujoco_py")
    mp.__path__ = []  # a package: unknown submodules (builder, utils, ...) resolve to inert stubs
    fn = types.ModuleType("mujoco_py.functions")

    def mj_fullM(model, M, qM):
        M[:] = qM.d.M

    fn.mj_fullM = mj_fullM
    fn.mj_getTotalmass = lambda model: float(model.body_mass.sum())
    def mj_applyFT(model, data, force, torque, point, body_id, qfrc):
        """qfrc += J(point, body)^T [force; torque] from the kinematics of the last forward pass (oracle/uhc_oracle.c or_apply_ft)"""
        f, t, p = (np.ascontiguousarray(x, dtype=np.float64) for x in (force, torque, point))
        q = np.ascontiguousarray(qfrc, dtype=np.float64)
        O.lib().or_apply_ft(model.om.h, data._sim.d.h, O._p(f), O._p(t), O._p(p), C.c_int(int(body_id) - 1), O._p(q))
        qfrc[:] = q

    fn.mj_applyFT = mj_applyFT
    mp.functions = fn
    mp.load_model_from_xml = lambda xml: FakeModel(om)
    mp.load_model_from_path = lambda path: FakeModel(om)
    mp.MjSim = FakeSim
    mp.MjSimState = MjSimState
    for n in ("MjViewer", "MjRenderContextOffscreen", "MjViewerBasic", "const", "cymj", "generated"):
        setattr(mp, n, _Any)
    mp.generated = types.ModuleType("mujoco_py.generated")
    mp.generated.const = _StubModule("mujoco_py.generated.const")
    sys.modules["mujoco_py"] = mp
    sys.modules["mujoco_py.functions"] = fn
    sys.modules["mujoco_py.generated"] = mp.generated
    sys.modules["mujoco_py.generated.const"] = mp.generated.const


def _install_fake_gym():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high = low, high
            self.shape = np.shape(low) if shape is None else shape

    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.spaces, gym.utils = spaces, utils
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding})


def _install_fake_robot():
    m = types.ModuleType("uhc.smpllib.smpl_robot")

    class Robot:
        def __init__(self, cfg, data_dir=None, masterfoot=False):
            self.smpl_model = cfg.get("model", "smpl")
            self.weight = 80.29

        def export_xml_string(self):
            return b"<mujoco model='humanoid'/>"

        def export_vis_string(self):
            return b"<mujoco model='humanoid'/>"

        def load_from_skeleton(self, *a, **k):
            pass

    m.Robot = Robot
    m.in_hull = lambda *a, **k: False
    sys.modules["uhc.smpllib.smpl_robot"] = m


_WORK = None


def install(cwd=True):
    """Install all stubs, put the reference on sys.path, chdir into a scratch mirror of its data dirs."""
    global _WORK
    om = O.Model()
    sys.meta_path.insert(0, _StubFinder())
    _install_fake_mujoco(om)
    _install_fake_gym()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the repo root holds a drop-in `uhc` shim package: make sure the REFERENCE one wins inside the harness
    for k in [k for k in sys.modules if k == "uhc" or k.startswith("uhc.")]:
        del sys.modules[k]
    if ROOT in sys.path:
        sys.path.remove(ROOT)
        sys.path.append(ROOT)
    if "" in sys.path:
        sys.path.remove("")
    _install_fake_robot()
    if cwd:
        _WORK = tempfile.mkdtemp(prefix="uhc_ref_")
        for d in ("config", "assets", "sample_data"):
            os.symlink(os.path.join(REF, d), os.path.join(_WORK, d))
        os.chdir(_WORK)
    return om


def make_cfg(cfg_id="uhc_implicit_shape", data_file="sample_data/amass_copycat_take5_test_small.pkl"):
    from uhc.utils.config_utils.copycat_config import Config
    cfg = Config(cfg_id=cfg_id, create_dirs=False)
    cfg.data_specs["file_path"] = data_file
    cfg.no_log, cfg.render, cfg.num_threads, cfg.mode = True, False, 1, "train"
    return cfg


def make_env(cfg, expert_seq, mode="train"):
    from uhc.envs.humanoid_im import HumanoidEnv
    env = HumanoidEnv(cfg, init_expert=expert_seq, data_specs=cfg.data_specs, mode=mode, no_root=cfg.no_root)
    return env
