"""Parity of the CUDA path (through the C ABI, via competitive_rl_b200.make_envs) against
(1) fixtures recorded from the reference's own files and (2) the CPU oracle on larger seeded
inputs.  Bar: bit-exact for states (incl. fp64 ball velocity), rewards, dones, info and
uint8 observations."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _make(env_id, n, dim, fs, serves, **kw):
    from competitive_rl_b200 import make_envs
    return make_envs(env_id, num_envs=n, resized_dim=dim, frame_stack=fs, log_dir=None, serves=serves, **kw)


def _stack(o, double):
    if double:
        return np.stack([o[0].cpu().numpy(), o[1].cpu().numpy()])
    return o.cpu().numpy()[None]


GOLDEN_CASES = ["pong_double_84", "pong_double_42", "pong_single_84_fs4", "pong_single_42_fs4",
                "pong_double_84_cheat"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_matches_reference_fixture(name):
    g = load_golden(name)
    env_id = str(g["env_id"])
    double = env_id == "cPongDouble-v0"
    fs = int(g["frame_stack"]) or None
    actions = g["actions"]
    T, N = actions.shape[:2]
    A = 2 if double else 1
    envs = _make(env_id, N, int(g["dim"]), fs, g["serves"])
    obs = envs.reset()
    assert np.array_equal(_stack(obs, double), g["reset_obs"])
    assert np.array_equal(envs.get_state().cpu().numpy(), g["state0"])
    term = {tuple(k): i for i, k in enumerate(g["term_idx"].tolist())}
    n_term = 0
    for t in range(T):
        obs, rew, done, info = envs.step(actions[t])
        assert np.array_equal(envs.get_state().cpu().numpy(), g["state"][t]), (name, t)
        assert np.array_equal(rew.cpu().numpy().reshape(N, -1)[:, :A], g["rew"][t]), (name, t)
        d = done.cpu().numpy().reshape(N, -1)[:, 0]
        assert np.array_equal(d, g["done"][t]), (name, t)
        assert np.array_equal(info.num_steps.cpu().numpy(), g["num_steps"][t]), (name, t)
        assert np.array_equal(info.real_reward.cpu().numpy()[:, :A], g["real_reward"][t]), (name, t)
        assert np.array_equal(_stack(obs, double), g["obs"][t]), (name, t)
        for i in np.nonzero(d)[0]:
            to = info[int(i)]["terminal_observation"]
            to = np.stack([x.cpu().numpy() for x in to]) if double else to.cpu().numpy()[None]
            assert np.array_equal(to, g["term_obs"][term[(t, int(i))]]), (name, t, i)
            n_term += 1
    assert n_term == len(term)
    envs.check()
    envs.close()


def _actions(T, N, double, seed, p_cheat=0.03):
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (T, N, 2) if double else (T, N)
    a = rng.integers(0, 3, shape).astype(np.int32)
    hold = rng.random(shape) < 0.5
    for t in range(1, T):
        a[t][hold[t]] = a[t - 1][hold[t]]
    if double:
        a[rng.random(shape) < p_cheat] = 999
    return a


def test_state_parity_4096_envs_2000_steps():
    """BASELINE config 2: cPongDouble, N=4096, T=2000, injected serves: state trajectories,
    rewards, dones, num_steps bit-exact against the oracle after every env-step."""
    from pong_oracle import PongOracleVec, make_serve_table, set_threads
    N, T = 4096, 2000
    serves = make_serve_table(N, 700, seed=3)
    actions = _actions(T, N, True, 12345)
    set_threads(8)
    orc = PongOracleVec("cPongDouble-v0", N, 84, None, 21, None, serves, render=False)
    envs = _make("cPongDouble-v0", N, 84, None, serves)
    orc.reset()
    envs.reset()
    assert np.array_equal(envs.get_state().cpu().numpy(), orc.get_state())
    n_done = 0
    for t in range(T):
        _, r_o, d_o, i_o = orc.step(actions[t])
        _, r_g, d_g, i_g = envs.step(actions[t])
        assert np.array_equal(envs.get_state().cpu().numpy(), orc.get_state()), t
        assert np.array_equal(r_g.cpu().numpy(), r_o), t
        assert np.array_equal(d_g.cpu().numpy()[:, 0], d_o), t
        assert np.array_equal(i_g.num_steps.cpu().numpy(), i_o["num_steps"]), t
        assert np.array_equal(i_g.real_reward.cpu().numpy(), i_o["real_reward"]), t
        n_done += int(d_o.sum())
    assert n_done > 4096 * 5   # ~13 episodes per env
    envs.check()
    envs.close()


@pytest.mark.parametrize("env_id,dim,fs,N,T", [
    ("cPongDouble-v0", 84, 4, 96, 500),     # the BASELINE observation format: per-agent 4-stack
    ("cPongDouble-v0", 42, 4, 64, 300),
    ("cPongDouble-v0", 84, None, 64, 300),
    ("cPong-v0", 84, 4, 64, 400),
    ("cPong-v0", 42, None, 64, 300),
])
def test_obs_parity_vs_oracle(env_id, dim, fs, N, T, atlas):
    from pong_oracle import PongOracleVec, make_serve_table, set_threads
    double = env_id == "cPongDouble-v0"
    serves = make_serve_table(N, 300, seed=5)
    actions = _actions(T, N, double, 777)
    set_threads(8)
    orc = PongOracleVec(env_id, N, dim, fs, 21, atlas, serves)
    envs = _make(env_id, N, dim, fs, serves)
    o_o, o_g = orc.reset(), envs.reset()
    assert np.array_equal(_stack(o_g, double), np.stack(o_o) if double else o_o[None])
    n_done = 0
    for t in range(T):
        o_o, r_o, d_o, i_o = orc.step(actions[t])
        o_g, r_g, d_g, i_g = envs.step(actions[t])
        assert np.array_equal(_stack(o_g, double), np.stack(o_o) if double else o_o[None]), t
        assert np.array_equal(d_g.cpu().numpy().reshape(N, -1)[:, 0], d_o), t
        if d_o.any():
            tg = i_g.terminal_observation()
            tg = np.stack([x.cpu().numpy() for x in tg]) if double else tg.cpu().numpy()[None]
            to = i_o["terminal_observation"]
            to = np.stack(to) if double else to[None]
            idx = np.nonzero(d_o)[0]
            assert np.array_equal(tg[:, idx], to[:, idx]), t
            n_done += len(idx)
        # the fast rasteriser against the in-library one-thread-per-pixel rasteriser
        if t % 50 == 0:
            gen = envs.render_obs_generic()
            assert np.array_equal(_stack(gen, double), _stack(o_g, double)), t
    assert n_done > 0
    envs.close()


def test_fast_vs_generic_rasteriser_random_states(atlas):
    """Arbitrary (also unreachable) states: every ball/bat position class incl. ball above the arena,
    at the walls, overlapping bats; scores anywhere in the atlas."""
    rng = np.random.default_rng(11)
    # 42x42 runs the four-frames-per-warp kernel: full stacks, frame_stack None (four envs per quad), and sizes that
    # leave a partial last quad (8190 * 3 and 8189 * 1 frames per agent are not multiples of 4)
    for env_id, dim, fs, N in [("cPongDouble-v0", 84, 4, 8192), ("cPongDouble-v0", 42, None, 8192), ("cPong-v0", 84, None, 8192),
                               ("cPongDouble-v0", 42, 4, 8192), ("cPong-v0", 42, 3, 8190), ("cPongDouble-v0", 42, None, 8189)]:
        envs = _make(env_id, N, dim, fs, None, seed=1)
        envs.reset()
        for it in range(3):
            st = envs.get_state().cpu().numpy()
            st[:, 0] = rng.integers(0, 157, N)
            st[:, 1] = rng.integers(30, 191, N)
            st[:, 4] = rng.integers(34, 180, N)
            st[:, 5] = rng.integers(34, 180, N)
            st[:, 6] = rng.integers(0, 11, N)
            st[:, 7] = rng.integers(0, 11, N)
            st[:, 8] = 0
            near = rng.random(N) < 0.3     # ball hugging a bat
            st[near, 0] = np.where(rng.random(near.sum()) < 0.5, 21, 135) + rng.integers(-3, 4, near.sum())
            envs.set_state(st)
            a = rng.integers(0, 3, (N, 2) if env_id == "cPongDouble-v0" else (N,))
            obs, _, _, _ = envs.step(a)
            gen = envs.render_obs_generic()
            double = env_id == "cPongDouble-v0"
            assert np.array_equal(_stack(gen, double), _stack(obs, double)), (env_id, it)
        envs.close()


def test_rng_mode_sharding_invariance():
    """Without injected serves the serve RNG is keyed by the GLOBAL env index: two shards of 256
    envs reproduce one batch of 512."""
    N, T = 512, 200
    actions = _actions(T, N, True, 99, p_cheat=0.0)
    whole = _make("cPongDouble-v0", N, 84, None, None, seed=123)
    lo = _make("cPongDouble-v0", N // 2, 84, None, None, seed=123, first_env=0)
    hi = _make("cPongDouble-v0", N // 2, 84, None, None, seed=123, first_env=N // 2)
    for e in (whole, lo, hi):
        e.reset()
    for t in range(T):
        ow, rw, dw, _ = whole.step(actions[t])
        ol, rl, dl, _ = lo.step(actions[t][:N // 2])
        oh, rh, dh, _ = hi.step(actions[t][N // 2:])
        assert torch.equal(ow[0], torch.cat([ol[0], oh[0]])) and torch.equal(ow[1], torch.cat([ol[1], oh[1]]))
        assert torch.equal(rw, torch.cat([rl, rh])) and torch.equal(dw, torch.cat([dl, dh]))
    s = whole.get_state().cpu().numpy()
    assert len(np.unique(s[:, 3])) > N // 4   # serves really differ between envs
    for e in (whole, lo, hi):
        e.close()


def test_full_size_properties():
    """BASELINE config 3 size (65 536 envs, 84x84x4 per agent): size-independent properties."""
    N = 65536
    envs = _make("cPongDouble-v0", N, 84, 4, None, seed=7)
    o = envs.reset()
    # reset observation: 4 identical copies of one constant frame, same for every env
    assert torch.equal(o[0][:, 0], o[0][:, 3]) and torch.equal(o[0][0], o[0][N - 1])
    first = o[0][0, 0].clone()
    gen = torch.Generator(device="cuda").manual_seed(0)
    prev = None
    for t in range(40):
        a = torch.randint(0, 3, (N, 2), generator=gen, device="cuda", dtype=torch.int32)
        o, r, d, info = envs.step(a)
        # FrameStack: channels 0..2 of this step are channels 1..3 of the previous one (no done yet)
        if prev is not None:
            assert torch.equal(o[0][:, :3], prev[0][:, 1:]) and torch.equal(o[1][:, :3], prev[1][:, 1:])
        prev = (o[0].clone(), o[1].clone())
        # zero-sum rewards; border rows are white, agent 1's arena rows are agent 0's mirrored
        assert torch.equal(r[:, 0], -r[:, 1])
        assert int(o[0][:, :, :3].min()) == 255 and int(o[0][:, :, 78:].min()) == 255
        assert torch.equal(o[1][:, :, 14:77], o[0][:, :, 14:77].flip(-1))
        assert torch.equal(o[1][:, :, :9], o[0][:, :, :9])
    assert not torch.equal(o[0][0, 3], first)
    envs.close()


def test_survey_known_answer_rollout():
    """SURVEY.md section 8(c) known-answer rollout (hand-derived from PongGame): the CUDA path through the C ABI."""
    from test_oracle_golden import kat_rollout
    kat_rollout(lambda serves: _make("cPongDouble-v0", 1, 84, None, serves))
