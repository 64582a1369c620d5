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


def test_tournament_wrapper_shapes():
    """shapes only; the behaviour is pinned against the reference's own wrapper in test_gpu_tournament.py, and
    FrameStackTensor's zero-on-done stacking is a mode of the rasteriser (zero_on_done=True, test_gpu_ring_mode.py)"""
    from competitive_rl_b200 import make_envs
    N = 4
    t = make_envs("cPongTournament-v0", num_envs=N, resized_dim=42, log_dir=None)
    o = t.reset()
    assert tuple(o.shape) == (N, 1, 42, 42)
    for k in range(5):
        o, r, d, info = t.step(np.zeros(N, np.int64))
        assert tuple(r.shape) == (N, 1) and tuple(d.shape) == (N, 1)
    t.close()


def test_step_host_matches_device_step():
    """crl_pong_step_host (pinned host actions in, rewards / dones / counters out on a side stream behind the
    rasteriser, optional host copies of the observations) against the device-pointer call on a twin env."""
    from competitive_rl_b200 import _native, make_envs
    lib = _native.load()
    N, T = 512, 150
    kw = dict(seed=11, log_dir=None, num_envs=N, resized_dim=84, frame_stack=4, n_buffers=1)
    ea, eb = make_envs("cPongDouble-v0", **kw), make_envs("cPongDouble-v0", **kw)
    ea.reset(); eb.reset()
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    h_act = torch.zeros((N, 2), dtype=torch.int32).pin_memory()
    h_rew = torch.zeros((N, 2), dtype=torch.float32).pin_memory()
    h_done = torch.zeros((N,), dtype=torch.uint8).pin_memory()
    h_steps = torch.zeros((N,), dtype=torch.int32).pin_memory()
    h_real = torch.zeros((N, 2), dtype=torch.float32).pin_memory()
    h_o0 = torch.zeros((N, 4, 84, 84), dtype=torch.uint8).pin_memory()
    h_o1 = torch.zeros((N, 4, 84, 84), dtype=torch.uint8).pin_memory()
    o0, o1 = eb._obs
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(3)
    dones = 0
    for t in range(T):
        a = rng.integers(0, 3, (N, 2)).astype(np.int32)
        a[rng.random((N, 2)) < 0.2] = 999
        obs, rew, done, info = ea.step(torch.from_numpy(a).cuda())
        h_act.copy_(torch.from_numpy(a))
        with_obs = t % 25 == 0
        _native.check(lib.crl_pong_step_host(eb._h, p(h_act), p(o0), p(o1), p(h_o0) if with_obs else None,
                                             p(h_o1) if with_obs else None, p(h_rew), p(h_done), p(h_steps), p(h_real), sp))
        # the call returns with every host buffer filled: no further synchronisation here
        assert np.array_equal(h_rew.numpy(), rew.cpu().numpy()) and np.array_equal(h_done.numpy() != 0, done.cpu().numpy().astype(bool).reshape(N, -1)[:, 0])
        assert np.array_equal(h_steps.numpy(), info.num_steps.cpu().numpy()) and np.array_equal(h_real.numpy(), info.real_reward.cpu().numpy())
        assert torch.equal(o0, obs[0]) and torch.equal(o1, obs[1])
        if with_obs:
            assert torch.equal(h_o0, obs[0].cpu()) and torch.equal(h_o1, obs[1].cpu())
        dones += int(h_done.sum())
    ea.close(); eb.close()


def test_evaluate_two_policies_in_batch_matches_host_bookkeeping():
    """Device-side evaluate_two_policies_in_batch against the reference's per-env Python bookkeeping
    (pong/evaluate.py:53-88) replayed on a twin env."""
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200.evaluate import evaluate_two_policies_in_batch
    from competitive_rl_b200.builtin_policies import get_compute_action_function
    N, EPISODES = 128, 200
    kw = dict(num_envs=N, resized_dim=42, frame_stack=None, log_dir=None, seed=9, asynchronous=True)

    def lazy(obs):                      # deterministic opponent: always "stay"
        return torch.ones((N,), dtype=torch.int32, device="cuda")
    rule = get_compute_action_function("RULE_BASED", N, "cuda")
    g0, g1 = evaluate_two_policies_in_batch(rule, lazy, make_envs("cPongDouble-v0", **kw), EPISODES)
    # the reference's loop, on the host, over an identically seeded env
    envs = make_envs("cPongDouble-v0", **kw)
    r0, r1 = [0] * 4, [0] * 4
    ep = np.zeros((N, 2))
    total = 0
    obs = envs.reset()
    while total < EPISODES:
        a = torch.stack([rule(obs[0]), lazy(obs[1])], dim=1)
        obs, rew, done, _ = envs.step(a)
        ep += rew.cpu().numpy()
        for i, dn in enumerate(done.cpu().numpy().reshape(N, -1).all(axis=1)):
            if dn:
                k = 0 if ep[i, 0] > 0 else (1 if ep[i, 0] == 0 else 2)
                r0[k] += 1; r1[2 - k] += 1
                r0[3] += ep[i, 0]; r1[3] += ep[i, 1]
                total += 1
                ep[i] = 0
    assert g0 == r0 and g1 == r1
    assert g0[0] + g0[1] + g0[2] >= EPISODES and g0[0] > g0[2]      # the rule-based bat beats a bat that never moves
    envs.close()


def test_invalid_actions_are_rejected():
    """The reference asserts action_space.contains (cPong-v0) / indexes BAT_DIRECTIONS (cPongDouble-v0): host actions are
    rejected before the step; device-resident ones are played as "stay" and reported by check()."""
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200._native import CrlError
    single = make_envs("cPong-v0", num_envs=4, resized_dim=84, frame_stack=4, log_dir=None)
    single.reset()
    with pytest.raises(AssertionError):
        single.step(np.array([0, 1, 2, 3]))
    with pytest.raises(AssertionError):
        single.step(np.array([0, 999, 2, 1]))                  # the cheat code exists only in the two-player env
    single.step(np.array([0, 1, 2, 1]))
    single.check()
    single.close()
    double = make_envs("cPongDouble-v0", num_envs=4, resized_dim=84, frame_stack=None, log_dir=None)
    double.reset()
    with pytest.raises(IndexError):
        double.step(np.array([[0, 1], [2, 3], [1, 1], [1, 1]]))
    double.step(np.array([[0, 999], [999, 2], [1, 1], [1, 1]]))
    double.check()
    s0 = double.get_state().clone()
    double.step(torch.tensor([[1, 1], [1, 1], [7, 1], [1, -1]], dtype=torch.int32, device="cuda"))   # not validated on the host
    twin = make_envs("cPongDouble-v0", num_envs=4, resized_dim=84, frame_stack=None, log_dir=None)
    twin.reset()
    twin.set_state(s0)
    twin.step(np.ones((4, 2), np.int32))
    assert torch.equal(double.get_state()[:, 4:6], twin.get_state()[:, 4:6])      # the bad actions moved no bat
    with pytest.raises(CrlError, match="action outside"):
        double.check()
    double.close()
    twin.close()


def test_car_obs_rotation_is_checked_at_the_c_abi():
    """crl_car_set_obs_rotation: too few / duplicate buffers are refused, and once a rotation is registered a call given
    any buffer but the next one fails with CRL_E_INVALID instead of silently writing the wrong stack."""
    from competitive_rl_b200 import _native, make_envs
    lib = _native.load()
    N = 4
    envs = make_envs("cCarRacing-v0", num_envs=N, frame_stack=4, log_dir=None, seed=2, stack_mode="stack-shift")
    h = envs._h
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    bufs = [torch.zeros((N, 4, 96, 96), dtype=torch.uint8, device="cuda") for _ in range(5)]

    def register(tensors):
        arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return lib.crl_car_set_obs_rotation(h, arr, len(tensors), sp)

    assert register(bufs[:4]) == _native.CRL_E_INVALID                       # needs frame_stack + 1 buffers
    assert register(bufs[:4] + [bufs[0]]) == _native.CRL_E_INVALID           # the same buffer twice
    assert register(bufs) == 0
    b = envs._sets[0]
    act = torch.zeros((N, 2), dtype=torch.float32, device="cuda")
    step = lambda o: lib.crl_car_step(h, p(act), p(o), p(b["rew"]), p(b["done"]), p(b["steps"]), p(b["trunc"]), None, sp)  # noqa: E731
    assert step(bufs[1]) == _native.CRL_E_STATE                              # the stacks are rebuilt by a reset first
    assert lib.crl_car_reset(h, p(b["obs"]), sp) == _native.CRL_E_INVALID    # not a registered buffer
    assert lib.crl_car_reset(h, p(bufs[3]), sp) == 0                         # any registered buffer may start the rotation
    assert step(bufs[3]) == _native.CRL_E_INVALID and step(bufs[0]) == _native.CRL_E_INVALID
    assert step(bufs[4]) == 0 and step(bufs[0]) == 0 and step(bufs[1]) == 0
    torch.cuda.synchronize()
    assert torch.equal(bufs[1][:, 2], bufs[0][:, 3]) and torch.equal(bufs[1][:, 1], bufs[4][:, 3])   # the frames are where FrameStack puts them
    assert torch.equal(bufs[1][:, 0], bufs[3][:, 3])                         # ... down to the reset frame
    assert lib.crl_car_set_obs_rotation(h, None, 0, sp) == 0                 # back to the plain mode (after a reset)
    assert lib.crl_car_reset(h, p(b["obs"]), sp) == 0 and step(b["obs"]) == 0
    envs.close()
