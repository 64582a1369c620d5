"""oracle/car_oracle.c against fixtures recorded by running the reference's own car-racing Python
(car_dynamics.py, car_racing_multi_players.py) on the stand-in Box2D (tests/golden/car_*.npz,
generator oracle/gen_golden_car.py).  Pins the Python-level logic bit for bit: _create_track,
Car.step, gas/brake/steer, process_action, CarRacing.step rewards/dones, FrictionDetector."""
import numpy as np
import pytest

from conftest import load_golden
import car_oracle as C

CASES = ["car_single_seed123", "car_single_seed5_rep2", "car_double"]


@pytest.mark.parametrize("name", CASES)
def test_track_generator_matches_reference(name):
    g = load_golden(name)
    all_draws = g["all_draws"].reshape(-1, 24)
    # every failed attempt of the reference fails here too, the last one succeeds
    for d in all_draws[:-1]:
        assert C.create_track(d) is None
    track, border = C.create_track(all_draws[-1])
    assert np.array_equal(track, g["track"])          # float64, bit-exact (same libm as CPython)
    assert int(border.sum()) == len(g["kerbs"])


@pytest.mark.parametrize("name", CASES)
def test_rollout_matches_reference(name):
    g = load_golden(name)
    n = int(g["n_players"])
    track, border = C.create_track(g["draws"])
    env = C.CarOracleEnv(n, int(g["action_repeat"]), None, render=False)
    env.reset(track, border, g["birth"])
    assert np.array_equal(env.get_state(), g["state0"])
    for t, a in enumerate(g["actions"]):
        _, rew, done, _ = env.step(a)
        assert np.array_equal(env.get_state(), g["states"][t]), (name, t)
        assert np.array_equal(rew, g["rewards"][t]), (name, t)
        assert np.array_equal(done, g["dones"][t]), (name, t)


def test_track_generator_statistics():
    """Shape sanity of generated tracks (closed loop, tile spacing, retry rate)."""
    rng = np.random.RandomState(1)
    fails = 0
    for _ in range(40):
        while True:
            d = C.draw_track_uniforms(rng)
            t = C.create_track(d)
            if t is not None:
                break
            fails += 1
        track, border = t
        assert 100 < len(track) <= C.MAX_TRACK
        step = np.hypot(np.diff(track[:, 2]), np.diff(track[:, 3]))
        assert np.allclose(step, 21 / 6.0, atol=1e-9)
        assert np.hypot(track[0, 2] - track[-1, 2], track[0, 3] - track[-1, 3]) < 3 * 21 / 6.0
    assert fails < 80


def test_renderer_matches_reference_frames():
    """The C renderer against frames produced by the reference's own get_observation / camera_view /
    draw_for_pygame / render_indicators_for_pygame running on the pygame stand-in: bit-exact."""
    import os
    import conftest
    g = load_golden("car_frames")
    glyphs = C.load_glyphs(os.path.join(conftest.ROOT, "competitive-rl_b200", "data", "car_hud_glyphs.npz"))
    for ci in range(int(g["n_cases"])):
        P = int(g["c%d_players" % ci])
        track, border = C.create_track(g["c%d_draws" % ci])
        env = C.CarOracleEnv(P, 1, glyphs, render=False)
        env.reset(track, border, g["c%d_birth" % ci])
        frames, steps = g["c%d_frames" % ci], g["c%d_steps" % ci].tolist()
        assert np.array_equal(np.stack(env.observe()), frames[0]), ci
        k = 1
        for t, a in enumerate(g["c%d_actions" % ci]):
            env.step(a)
            if k < len(steps) and steps[k] == t:
                assert np.array_equal(np.stack(env.observe()), frames[k]), (ci, t)
                k += 1
        assert k == len(steps)


def test_render_is_deterministic_and_plausible():
    glyphs = C.load_glyphs(__import__("os").path.join(__import__("conftest").ROOT, "competitive-rl_b200", "data",
                                                     "car_hud_glyphs.npz"))
    rng = np.random.RandomState(0)
    track, border, _ = C.make_track(rng)
    env = C.CarOracleEnv(1, 1, glyphs)
    o0 = env.reset(track, border)[0]
    assert o0.shape == (96, 96) and o0.dtype == np.uint8
    assert (o0[86:] == 0).mean() > 0.7                 # HUD bar
    assert set(np.unique(o0[:86])) <= {0, 60, 101, 103, 107, 161, 176, 255, 76, 29}
    assert (o0[70:82, 42:54] == 60).any()              # own car (204,0,0) -> 60
    for _ in range(5):
        o, r, d, _ = env.step(np.array([[0.0, 0.5]]))
    assert not np.array_equal(o[0], o0)


def test_replayed_track_matches_reference():
    """reset(use_local_track=<json>) (car_racing_multi_players.py:376-381): the reference's own replay run on the stand-in."""
    g = load_golden("car_replay")
    track = g["track_json"]
    env = C.CarOracleEnv(1, 1, None, render=False)
    env.reset(track, C.track_border(track), [0])
    assert np.array_equal(env.get_state(), g["state0"])
    for t, a in enumerate(g["actions"]):
        _, rew, done, _ = env.step(a[None])
        assert np.array_equal(env.get_state(), g["states"][t]), t
        assert np.array_equal(rew, g["rewards"][t]), t
