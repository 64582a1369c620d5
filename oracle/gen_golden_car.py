"""Generate tests/golden/car_*.npz by running the REFERENCE's own car-racing Python on the
stand-in Box2D (oracle/ref_car_loader.py).  Build container only.

Pins the reference's Python logic (Car.step, CarRacing.step/reset/_create_track, process_action,
FrictionDetector._contact); Box2D numerics are oracle/car_oracle.c's own (see its header)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_car_loader as RC  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def actions_for(T, n, seed, amp=0.6):
    rng = np.random.default_rng(seed)
    a = np.zeros((T, n, 2))
    for k in range(n):
        steer = 0.0
        for t in range(T):
            if t % 25 == 0:
                steer = rng.uniform(-amp, amp)
            gas = 0.9 if (t // 40) % 3 != 2 else -0.7
            a[t, k] = (steer + 0.1 * rng.standard_normal(), gas)
    a[T // 2: T // 2 + 5] = 3.0    # exercises the clipping in process_action
    return a


def run(name, n_players, seed, T, action_repeat=None, amp=0.6, actions=None, min_contact_steps=0):
    M = RC.load_car_racing()
    env = M.CarRacing(num_player=n_players, verbose=0, action_repeat=action_repeat)
    env.seed(seed)
    rec = RC.RecordingRandom(env.np_random)
    env.np_random = rec
    np.random.seed(seed)
    env.reset()
    np.random.seed(seed)
    birth = np.arange(n_players)
    np.random.shuffle(birth)
    draws = np.array(rec.draws[-24:])
    n_attempts = len(rec.draws) // 24
    track = np.array(env.track)
    kerbs = np.array([np.array(p).reshape(-1) for p, c in env.road_poly if len(p) == 4])
    tiles = np.array([np.array(p).reshape(-1) for p, c in env.road_poly if len(p) == 5])
    if actions is None:
        actions = actions_for(T, n_players, seed + 1, amp)
    contacts = np.zeros((T,), np.int32)
    states = np.zeros((T, n_players, 24))
    rewards = np.zeros((T, n_players))
    dones = np.zeros((T, n_players), bool)
    state0 = np.array([RC.car_state(env, k) for k in range(n_players)])
    for t in range(T):
        a = actions[t, 0] if n_players == 1 else {k: actions[t, k] for k in range(n_players)}
        o, r, d, info = env.step(a)
        contacts[t] = env.world.contact_count()
        for k in range(n_players):
            states[t, k] = RC.car_state(env, k)
            rewards[t, k] = r if n_players == 1 else r[k]
            dones[t, k] = d if n_players == 1 else d[k]
    if (contacts > 0).sum() < min_contact_steps:
        return False
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, n_players=n_players, seed=seed, draws=draws, all_draws=np.array(rec.draws), contacts=contacts,
                        n_attempts=n_attempts, track=track, kerbs=kerbs, tiles=tiles, birth=birth, actions=actions,
                        state0=state0, states=states, rewards=rewards, dones=dones,
                        action_repeat=action_repeat or 1)
    print("%-22s track=%d attempts=%d kerbs=%d tiles_visited=%s return=%s  %d KiB" % (
        name, len(track), n_attempts, len(kerbs), states[-1, :, 23], rewards.sum(0).round(2), os.path.getsize(path) // 1024))
    return True


def run_replay(name, seed_track, seed_env, T):
    """reset(record_track_to=...) of one env, then reset(use_local_track=<that json>) of another (:376-381, 447-451)."""
    import glob
    import json
    import tempfile
    M = RC.load_car_racing()
    rec_env = M.CarRacing(num_player=1, verbose=0)
    rec_env.seed(seed_track)
    np.random.seed(seed_track)
    with tempfile.TemporaryDirectory() as d:
        rec_env.reset(record_track_to=d)
        path = glob.glob(os.path.join(d, "*_track.json"))[0]
        rows = np.array(json.load(open(path)), np.float64)
        assert np.array_equal(rows, np.array(rec_env.track))
        env = M.CarRacing(num_player=1, verbose=0)
        env.seed(seed_env)
        np.random.seed(seed_env)
        env.reset(use_local_track=path)
    assert np.array_equal(np.array(env.track), rows)
    actions = actions_for(T, 1, seed_env + 1, 0.4)
    states, rewards = np.zeros((T, 1, 24)), np.zeros((T, 1))
    state0 = np.array([RC.car_state(env, 0)])
    for t in range(T):
        o, r, dn, info = env.step(actions[t, 0])
        states[t, 0], rewards[t, 0] = RC.car_state(env, 0), r
    out = os.path.join(OUT, name + ".npz")
    np.savez_compressed(out, track_json=rows, actions=actions[:, 0], state0=state0, states=states, rewards=rewards)
    print("%-22s track=%d tiles_visited=%s return=%.2f  %d KiB" % (name, len(rows), states[-1, :, 23], rewards.sum(), os.path.getsize(out) // 1024))


def run_collision(name, seed, T):
    """Two cars steered into each other (the cars spawn 5 units apart side by side): the car-car contact
    path of world.Step under the reference's own CarRacing.step.  The steering sign that closes the gap depends
    on which car got which birth place, so both are tried."""
    for sign in (+1.0, -1.0):
        a = np.zeros((T, 2, 2))
        a[:, 0] = (sign * 0.3, 0.5)
        a[:, 1] = (-sign * 0.3, 0.5)
        a[T // 2:, :, 0] *= -0.5      # later pull apart again
        try:
            if run(name, 2, seed, T, actions=a, min_contact_steps=20):
                return True
        except AttributeError as exc:
            print("seed", seed, "-> reference crashed:", exc)
    return False


def run_frames(name, cases, T=60, every=12):
    """Observation fixtures: the reference's OWN renderer (get_observation and everything under it) run on the
    pygame stand-in.  Pins camera, scales, colours, paint order, car polygons and HUD of the observation."""
    M = RC.load_car_racing(render=True)
    out = {}
    for ci, (n_players, seed) in enumerate(cases):
        env = M.CarRacing(num_player=n_players, verbose=0)
        env.seed(seed)
        rec = RC.RecordingRandom(env.np_random)
        env.np_random = rec
        np.random.seed(seed)
        o = env.reset()
        np.random.seed(seed)
        birth = np.arange(n_players)
        np.random.shuffle(birth)

        def frames_of(o):
            return np.stack([o[:, :, 0]] if n_players == 1 else [o[k][:, :, 0] for k in range(n_players)])
        frames, steps = [frames_of(o)], [-1]
        actions = np.zeros((T, n_players, 2))
        for t in range(T):
            for k in range(n_players):
                actions[t, k] = (0.3 * np.sin(t / 9.0 + seed + k), 0.6 if (t // 25) % 2 == 0 else -0.4)
            o, r, d, info = env.step(actions[t, 0] if n_players == 1 else {k: actions[t, k] for k in range(n_players)})
            if t % every == every - 1:
                frames.append(frames_of(o))
                steps.append(t)
        out["c%d_players" % ci] = n_players
        out["c%d_draws" % ci] = np.array(rec.draws[-24:])
        out["c%d_birth" % ci] = birth
        out["c%d_actions" % ci] = actions
        out["c%d_frames" % ci] = np.array(frames, np.uint8)
        out["c%d_steps" % ci] = np.array(steps)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, n_cases=len(cases), **out)
    print("%-22s cases=%d  %d KiB" % (name, len(cases), os.path.getsize(path) // 1024))
    RC.load_car_racing(render=False)


if __name__ == "__main__":
    run_frames("car_frames", [(1, 123), (1, 31), (2, 12)])
    run_replay("car_replay", 31, 4, 80)
    run("car_single_seed123", 1, 123, 400)
    run("car_single_seed5_rep2", 1, 5, 150, action_repeat=2)
    # The reference raises AttributeError inside FrictionDetector._contact (it reads self.verbose,
    # :146) when a car touches a tile >= 50 blocks ahead of its last one, e.g. after spinning back
    # over the start line; fixtures must stay clear of that crash, so the two-car run steers gently.
    for seed in range(3, 40):
        if run_collision("car_double_collision", seed, 120):
            break
    for seed in range(7, 40):
        try:
            run("car_double", 2, seed, 300, amp=0.15)
            break
        except AttributeError as exc:    # the reference's own crash (see above): try the next track
            print("seed", seed, "-> reference crashed:", exc)
