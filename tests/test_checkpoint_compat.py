"""CPU: the checkpoint wire format (agent_copycat.py:190-201, :249-260).  `running_state` is a pickled
uhc.khrylib.utils.zfilter.ZFilter; a file written by this repo must load into the REFERENCE's class and a file written by the reference
must load here.  The cross-check against the real reference runs where a checkout exists ($UHC_REFERENCE, default /root/reference)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("UHC_REFERENCE", "/root/reference")


def test_host_zfilter_matches_running_statistics():
    from uhc.khrylib.utils.zfilter import ZFilter
    rng = np.random.RandomState(0)
    x = rng.normal(1.0, 2.0, (50, 6))
    z = ZFilter((6,), clip=5)
    ys = [z(r) for r in x[:10]]                          # per-sample pushes (the reference's use) ...
    z(x[10:])                                            # ... and one batch merge give the statistics of all 50 rows
    assert z.rs.n == 50 and np.allclose(z.rs.mean, x.mean(0)) and np.allclose(z.rs.var, x.var(0, ddof=1))
    assert np.allclose(ys[0], 0.0)                       # first sample: (x - x) / (|x| + eps)
    z2 = ZFilter.from_stats(50, x.mean(0), ((x - x.mean(0)) ** 2).sum(0), clip=5.0)
    assert np.allclose(z2(x[3], update=False), z(x[3], update=False))
    z3 = pickle.loads(pickle.dumps(z2))
    assert z3.rs._n == 50 and np.allclose(z3.rs._M, z2.rs._M) and np.allclose(z3.rs._S, z2.rs._S) and z3.clip == 5.0


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "uhc", "khrylib", "utils", "zfilter.py")), reason="no reference checkout on this machine")
def test_running_state_round_trips_through_the_reference_class(tmp_path):
    from uhc.khrylib.utils.zfilter import ZFilter
    rng = np.random.RandomState(1)
    x = rng.normal(0.5, 3.0, (40, 657))
    ours = ZFilter.from_stats(40, x.mean(0), ((x - x.mean(0)) ** 2).sum(0), clip=5.0)
    p1, p2 = tmp_path / "ours.p", tmp_path / "ref.p"
    pickle.dump({"running_state": ours}, open(p1, "wb"))
    np.save(tmp_path / "x.npy", x)
    # in a clean interpreter with ONLY the reference on the path: load our pickle into the reference's class, use it, write one of its own
    code = f"""
import sys, pickle, numpy as np
sys.path.insert(0, {REF!r})
from uhc.khrylib.utils import zfilter
assert zfilter.__file__.startswith({REF!r})
cp = pickle.load(open({str(p1)!r}, 'rb'))
rs = cp['running_state']
assert type(rs).__module__ == 'uhc.khrylib.utils.zfilter' and rs.rs.n == 40
x = np.load({str(tmp_path / 'x.npy')!r})
y = rs(x[0], update=False)
np.save({str(tmp_path / 'y.npy')!r}, y)
z = zfilter.ZFilter((657,), clip=5)
for r in x: z(r)
pickle.dump({{'running_state': z}}, open({str(p2)!r}, 'wb'))
"""
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.allclose(np.load(tmp_path / "y.npy"), ours(x[0], update=False))
    theirs = pickle.load(open(p2, "rb"))["running_state"]          # unpickles into this repo's stand-in (same module path)
    assert type(theirs).__module__ == "uhc.khrylib.utils.zfilter" and theirs.rs._n == 40
    assert np.allclose(theirs.rs._M, x.mean(0)) and np.allclose(theirs.rs._S, ((x - x.mean(0)) ** 2).sum(0))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "train_uhc.py")), reason="no reference checkout on this machine")
def test_reference_train_script_file_reaches_the_agent_on_this_package(tmp_path):
    """The reference's OWN scripts/train_uhc.py, unmodified, executed with cwd = a tree holding this repo's `uhc` package and config layout
    (train_uhc.py:24 puts the cwd on sys.path).  Without a GPU the run must get through argument parsing, Config, the wandb / flags /
    agent_dict imports and torch.set_default_dtype(float64) and stop exactly where AgentCopycat asks for a CUDA device -- i.e. the whole
    import surface of the script resolves against this package.  (With a GPU the same file trains: tests/test_gpu_dropin.py.)"""
    import yaml
    from tests.helpers import write_synthetic_pkl
    for d in ("uhc", "uhc_b200", "mujoco_py", "assets"):
        if os.path.exists(os.path.join(ROOT, d)):
            os.symlink(os.path.join(ROOT, d), tmp_path / d)
    os.makedirs(tmp_path / "config")
    base = yaml.safe_load(open(os.path.join(ROOT, "config", "uhc_b200_default.yml")))
    base.update(policy_hsize=[64], value_hsize=[64], min_batch_size=256, num_optim_epoch=1, num_envs=8, num_epoch=1)
    base["data_specs"]["file_path"] = write_synthetic_pkl(str(tmp_path / "sample_data" / "clips.pkl"))
    yaml.safe_dump(base, open(tmp_path / "config" / "refscript.yml", "w"))
    env = dict({k: v for k, v in os.environ.items() if k != "PYTHONPATH"}, WANDB_MODE="disabled", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(REF, "scripts", "train_uhc.py"), "--cfg", "refscript", "--no_log"], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert "Using: cpu" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
    assert r.returncode != 0 and "needs a CUDA device" in r.stderr, r.stderr[-3000:]
