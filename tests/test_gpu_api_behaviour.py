"""Error behaviour and API conventions of the CUDA path through the C ABI / Python front end."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_c_abi_error_codes():
    from competitive_rl_b200 import _native
    lib = _native.load()
    h = ctypes.c_void_p()
    for bad in [_native.PongConfig(0, 2, 84, 4, 21, 0, 0, 0), _native.PongConfig(4, 3, 84, 4, 21, 0, 0, 0),
                _native.PongConfig(4, 2, 85, 4, 21, 0, 0, 0), _native.PongConfig(4, 2, 84, 9, 21, 0, 0, 0),
                _native.PongConfig(4, 2, 84, 4, 30, 0, 0, 0), _native.PongConfig(4, 2, 84, 4, 21, 99, 0, 0)]:
        assert lib.crl_pong_create(ctypes.byref(bad), ctypes.byref(h)) == -1 and lib.crl_last_error()
    cfg = _native.PongConfig(4, 2, 84, 4, 21, 0, 0, 0)
    assert lib.crl_pong_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    obs = torch.zeros((2, 4, 4, 84, 84), dtype=torch.uint8, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.crl_pong_reset(h, p(obs[0]), p(obs[1]), None) == -3            # atlas not loaded
    junk = np.zeros(100, np.uint8)
    assert lib.crl_pong_load_atlas(h, junk.ctypes.data, junk.nbytes, None) == -1
    atlas = np.load(_native.DEFAULT_ATLAS)["strips"]
    assert lib.crl_pong_load_atlas(h, atlas.ctypes.data, atlas.nbytes, None) == 0
    a = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
    r = torch.zeros((4, 2), device="cuda")
    d = torch.zeros((4,), dtype=torch.uint8, device="cuda")
    s = torch.zeros((4,), dtype=torch.int32, device="cuda")
    assert lib.crl_pong_step(h, p(a), p(obs[0]), p(obs[1]), p(r), p(d), p(s), p(r), None) == -3   # step before reset
    assert lib.crl_pong_reset(h, p(obs[0]), p(obs[1]), None) == 0
    assert lib.crl_pong_step(h, None, p(obs[0]), p(obs[1]), p(r), p(d), p(s), p(r), None) == -1   # null actions
    assert lib.crl_pong_step(h, p(a), p(obs[0]), p(obs[1]), p(r), p(d), p(s), p(r), None) == 0
    assert lib.crl_pong_destroy(h) == 0


def test_serve_table_overrun_is_reported():
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200._native import CrlError
    serves = np.tile(np.array([[[4.0, 1.5]]]), (2, 4, 1))      # only 4 serves per env
    envs = make_envs("cPongDouble-v0", num_envs=2, resized_dim=84, frame_stack=None, log_dir=None, serves=serves)
    envs.reset()
    for _ in range(60):
        envs.step(np.ones((2, 2), np.int32))
    with pytest.raises(CrlError, match="serve table exhausted"):
        envs.check()
    envs.close()


def test_return_conventions_dummy_vs_subproc():
    """F10: DummyVecEnv returns rew (N, A) float32 / done (N, A) bool; SubprocVecEnv rew (N, 2) float64 /
    done (N,) bool (single: rew (N,))."""
    from competitive_rl_b200 import make_envs
    N = 3
    for env_id, A in (("cPongDouble-v0", 2), ("cPong-v0", 1)):
        fs = None if A == 2 else 4
        dummy = make_envs(env_id, num_envs=N, resized_dim=42, frame_stack=fs, log_dir=None, return_numpy=True)
        sub = make_envs(env_id, num_envs=N, resized_dim=42, frame_stack=fs, log_dir=None, asynchronous=True,
                        return_numpy=True)
        a = np.ones((N, 2), np.int64) if A == 2 else np.ones((N,), np.int64)
        for envs, is_sub in ((dummy, False), (sub, True)):
            o = envs.reset()
            o, r, d, info = envs.step(a)
            first = o[0] if A == 2 else o
            assert isinstance(o, tuple) == (A == 2) and first.shape == (N, fs or 1, 42, 42) and first.dtype == np.uint8
            if is_sub:
                assert r.shape == ((N, 2) if A == 2 else (N,)) and r.dtype == np.float64
                assert d.shape == (N,) and d.dtype == bool and isinstance(info, tuple)
            else:
                assert r.shape == (N, A) and r.dtype == np.float32 and d.shape == (N, A) and isinstance(info, list)
            assert set(info[0].keys()) == {"real_reward", "num_steps"} and info[0]["num_steps"] == 1
            assert envs.seed(3) == [None] * N
            envs.close()
    # spaces (utils/atari_wrappers.py:12-23: CHW Box, Tuple for Double)
    e = make_envs("cPongDouble-v0", num_envs=1, resized_dim=84, frame_stack=None, log_dir=None)
    assert len(e.observation_space) == 2 and e.observation_space[0].shape == (1, 84, 84)
    assert e.action_space[0].n == 3
    img = e.envs[0].render("rgb_array")
    e.reset()
    assert e.envs[0].render("rgb_array").shape == (210, 160, 3) and img.shape == (210, 160, 3)
    e.close()


def test_tournament_wrapper_and_framestack_tensor():
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200.utils import FrameStackTensor
    N = 4
    t = make_envs("cPongTournament-v0", num_envs=N, resized_dim=42, log_dir=None)
    o = t.reset()
    assert tuple(o.shape) == (N, 1, 42, 42)
    fst = FrameStackTensor(N, (1, 42, 42), 4, "cuda")
    for k in range(5):
        o, r, d, info = t.step(np.zeros(N, np.int64))
        assert tuple(r.shape) == (N, 1) and tuple(d.shape) == (N, 1)
        stacked = fst.update(o, mask=1.0 - d.float())
    assert tuple(stacked.shape) == (N, 4, 42, 42) and torch.equal(stacked[:, -1], o[:, 0].float())
    t.close()
