"""stack_mode="ring" (double-write ring views) and zero_on_done stacking against the plain stack mode of the same CUDA
path -- which tests/test_gpu_pong_parity.py / test_gpu_car_parity.py pin to the reference fixtures and the oracle -- and
against an independent numpy restatement of FrameStackTensor (utils/utils.py:145-173; pinned to the reference class in
tests/test_frame_stack_tensor.py).  Bar: bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _np(o):
    if isinstance(o, tuple):
        return np.stack([x.cpu().numpy() for x in o])
    return o.cpu().numpy()[None]


@pytest.mark.parametrize("env_id,dim,fs", [("cPongDouble-v0", 84, 4), ("cPongDouble-v0", 42, 4), ("cPong-v0", 84, 4),
                                           ("cPong-v0", 42, 3), ("cPongDouble-v0", 84, 2)])
def test_pong_ring_views_equal_stack(env_id, dim, fs):
    from competitive_rl_b200 import make_envs
    N, T = 96, 260
    kw = dict(num_envs=N, resized_dim=dim, frame_stack=fs, log_dir=None, seed=11, max_num_rounds=2)   # short games: many resets
    a = make_envs(env_id, **kw)
    b = make_envs(env_id, stack_mode="ring", **kw)
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(_np(oa), _np(ob))
    assert np.array_equal(_np(b.render_obs_generic()), _np(ob))
    rng = np.random.default_rng(3)
    shape = (N, 2) if env_id == "cPongDouble-v0" else (N,)
    n_done = 0
    for t in range(T):
        act = rng.integers(0, 3, shape)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(_np(oa), _np(ob)), t
        assert torch.equal(ra, rb) and torch.equal(da, db), t
        if t % 37 == 0:
            assert np.array_equal(_np(b.render_obs_generic()), _np(ob)), t       # the whole ring, rewritten from the frame specs
        d = da.reshape(N, -1)[:, 0].cpu().numpy()
        for i in np.nonzero(d)[0][:2]:
            ta, tb = ia[int(i)]["terminal_observation"], ib[int(i)]["terminal_observation"]
            assert np.array_equal(_np(ta), _np(tb)), (t, i)
        n_done += int(d.sum())
    assert n_done > 20
    assert b.bytes_per_env_step * fs == a.bytes_per_env_step * 2
    a.close()
    b.close()


@pytest.mark.parametrize("stack_mode", ["stack", "ring"])
def test_zero_on_done_is_frame_stack_tensor(stack_mode):
    from competitive_rl_b200 import make_envs
    N, T, C = 48, 200, 4
    kw = dict(num_envs=N, resized_dim=42, log_dir=None, seed=5, max_num_rounds=2)
    single = make_envs("cPongDouble-v0", frame_stack=None, **kw)                      # what the reference feeds FrameStackTensor
    fused = make_envs("cPongDouble-v0", frame_stack=C, zero_on_done=True, stack_mode=stack_mode, **kw)
    f0, o0 = single.reset(), fused.reset()
    frames = [[_np(f0)[k][:, 0]] for k in range(2)]
    dones, got = [], [_np(o0)]
    rng = np.random.default_rng(9)
    for t in range(T):
        act = rng.integers(0, 3, (N, 2))
        f, _, d, _ = single.step(act)
        o, _, d2, _ = fused.step(act)
        assert torch.equal(d, d2)
        for k in range(2):
            frames[k].append(_np(f)[k][:, 0])
        dones.append(d.reshape(N, -1)[:, 0].cpu().numpy())
        got.append(_np(o))
    got = np.stack(got)                                   # (T+1, 2, N, C, D, D)
    dn = np.stack([np.zeros(N, bool)] + dones)            # done flag that came WITH frame t (t = 0: the reset frame)
    assert dn.sum() > 10
    for k in range(2):
        # FrameStackTensor starts zeroed and its first update appends the reset frame: [0, 0, 0, reset]
        want = frame_stack_tensor_numpy(np.stack(frames[k]), dn, C)
        assert np.array_equal(got[:, k], want), k
    single.close()
    fused.close()


def frame_stack_tensor_numpy(frames, done_with_frame, n_stack):
    """`frames[t]` arrives with `done_with_frame[t]` (the auto-reset observation of a finished env arrives with done=True):
    mask = 1 - done zeroes the stack BEFORE that frame is appended."""
    T, N = frames.shape[:2]
    stack = np.zeros((N, n_stack) + frames.shape[2:], np.uint8)
    out = []
    for t in range(T):
        stack[done_with_frame[t]] = 0
        stack = np.roll(stack, -1, axis=1)
        stack[:, -1] = frames[t]
        out.append(stack.copy())
    return np.stack(out)


@pytest.mark.parametrize("env_id", ["cCarRacing-v0", "cCarRacingDouble-v0"])
def test_car_ring_views_equal_stack(env_id):
    from competitive_rl_b200 import make_envs
    N, T, C = 12, 70, 4
    P = 2 if "Double" in env_id else 1
    kw = dict(num_envs=N, frame_stack=C, log_dir=None, seed=21, max_episode_steps=23, asynchronous=True)   # resets every 23 steps
    a = make_envs(env_id, **kw)
    b = make_envs(env_id, stack_mode="ring", **kw)
    oa, ob = a.reset(), b.reset()
    assert tuple(ob.shape) == ((N, C, 96, 96) if P == 1 else (N, 2, C, 96, 96))
    assert np.array_equal(oa.cpu().numpy(), ob.reshape(N, P * C, 96, 96).cpu().numpy())
    rng = np.random.default_rng(4)
    n_done = 0
    for t in range(T):
        act = rng.uniform(-1, 1, (N, 2) if P == 1 else (N, 2, 2)).astype(np.float32)
        act[..., 1] = np.abs(act[..., 1])      # drive, so that the frames change
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa.cpu().numpy(), ob.reshape(N, P * C, 96, 96).cpu().numpy()), t
        assert torch.equal(ra, rb) and torch.equal(da, db), t
        for i in np.nonzero(da.cpu().numpy())[0]:
            assert torch.equal(ia[int(i)]["terminal_observation"], ib[int(i)]["terminal_observation"]), (t, i)
            n_done += 1
    assert n_done >= N * 2
    a.check()
    b.check()
    a.close()
    b.close()
