"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/crl_b200.h declares, the Python binding knows them all, and without a CUDA device the
product refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "crl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from competitive_rl_b200 import _native
    names = _header_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_native.EXPORTS) == names
    assert _native.load().crl_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from competitive_rl_b200 import make_envs
    with pytest.raises(RuntimeError, match="no CPU path"):
        make_envs("cPongDouble-v0", num_envs=2, frame_stack=None, log_dir=None)
    # straight through the C ABI as well
    from competitive_rl_b200 import _native
    lib = _native.load()
    cfg = _native.PongConfig(4, 2, 84, 4, 21, 0, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.crl_pong_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and b"no CPU path" in lib.crl_last_error()


def test_make_envs_surface():
    import inspect
    import competitive_rl_b200 as pkg
    sig = inspect.signature(pkg.make_envs)
    names = list(sig.parameters)[:8]
    assert names == ["env_id", "seed", "log_dir", "num_envs", "asynchronous", "resized_dim", "frame_stack",
                     "action_repeat"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(env_id="cPong-v0", seed=0, log_dir="data", num_envs=3, asynchronous=False, resized_dim=42,
                     frame_stack=4, action_repeat=None)   # make_envs.py:67-68
    from competitive_rl_b200 import registry
    pkg.register_competitive_envs()
    for i in ("cPong-v0", "cPongDouble-v0", "cCarRacing-v0", "cCarRacingDouble-v0"):
        assert i in registry.registered_ids()
    assert registry.spec("cPongDouble-v0")["kwargs"]["max_num_rounds"] == 21
