"""CPU: the C-ABI shared library loads without a GPU and exports every entry point include/*.h declares (no compute calls here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = {}
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"^[A-Za-z_][\w \t\*]*?\b(uhc_\w+)\s*\(", txt, flags=re.M):
            names[m.group(1)] = os.path.basename(h)
    return names


def test_library_exports_every_declared_entry_point():
    from uhc_b200 import build
    so = build.build()
    lib = ctypes.CDLL(so)
    names = declared_functions()
    assert len(names) >= 60 and {"uhc_env_step", "uhc_rollout", "uhc_policy_forward", "uhc_ppo_update", "uhc_linear_forward_tc"} <= set(names)
    missing = [f"{n} ({h})" for n, h in names.items() if not hasattr(lib, n)]
    assert not missing, "declared in include/ but not exported by libuhc_b200.so: " + ", ".join(missing)


def test_error_strings_are_callable_without_a_gpu():
    from uhc_b200 import build
    lib = ctypes.CDLL(build.build())
    for f in ("uhc_last_error", "uhc_nn_last_error", "uhc_tc_last_error", "uhc_rollout_last_error", "uhc_ppo_last_error"):
        fn = getattr(lib, f)
        fn.restype = ctypes.c_char_p
        assert isinstance(fn(), bytes)
