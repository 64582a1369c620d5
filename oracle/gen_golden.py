"""Generate tests/golden/*.npz by running the REFERENCE's own Pong path
(/root/reference, via oracle/ref_loader.py's stand-in gym/pygame + real cv2).

TEST INFRASTRUCTURE ONLY; runs only in the build container.  The fixtures it
writes are committed, so the GPU box (which has no /root/reference) checks the
oracle and the CUDA path against the reference's recorded behaviour.

    python oracle/gen_golden.py            # regenerates every fixture

Each fixture holds the injected serve table, the action sequence and, per
env-step: observations, clipped rewards, dones, info["num_steps"],
info["real_reward"], every info["terminal_observation"], and the PongGame
state after the step (ball x, y, vx, vy, bat ys, scores, rounds, steps).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from pong_oracle import make_serve_table  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [
    # name, env_id, N, T, dim, frame_stack, serve_seed, action_seed, p_cheat
    ("pong_double_84", "cPongDouble-v0", 4, 420, 84, None, 7, 12345, 0.05),
    ("pong_double_42", "cPongDouble-v0", 2, 260, 42, None, 8, 12346, 0.05),
    ("pong_single_84_fs4", "cPong-v0", 2, 320, 84, 4, 9, 12347, 0.0),
    ("pong_single_42_fs4", "cPong-v0", 2, 260, 42, 4, 10, 12348, 0.0),
    ("pong_double_84_cheat", "cPongDouble-v0", 2, 300, 84, None, 11, 12349, 0.7),
]


def make_actions(T, N, double, seed, p_cheat):
    rng = np.random.Generator(np.random.PCG64(seed))
    if double:
        a = rng.integers(0, 3, (T, N, 2)).astype(np.int32)
        cheat = rng.random((T, N, 2)) < p_cheat
        a[cheat] = 999
        # persistent runs of the same action make rallies (and bat hits with spin) likelier
        hold = rng.random((T, N, 2)) < 0.6
        for t in range(1, T):
            a[t][hold[t]] = a[t - 1][hold[t]]
    else:
        a = rng.integers(0, 3, (T, N)).astype(np.int32)
        hold = rng.random((T, N)) < 0.6
        for t in range(1, T):
            a[t][hold[t]] = a[t - 1][hold[t]]
    return a


def run_case(name, env_id, N, T, dim, fs, serve_seed, action_seed, p_cheat):
    double = env_id == "cPongDouble-v0"
    serves = make_serve_table(N, 400, seed=serve_seed)
    envs, inj = ref_loader.make_reference_vec_env(env_id, N, serves, resized_dim=dim, frame_stack=fs)
    actions = make_actions(T, N, double, action_seed, p_cheat)
    A = 2 if double else 1

    def split(o):
        return list(o) if double else [o]

    reset_obs = np.stack(split(envs.reset()))                     # (A, N, C, D, D)
    state0 = np.array([ref_loader.game_state(e) for e in envs.envs], np.float64)
    obs = np.zeros((T,) + reset_obs.shape, np.uint8)
    rew = np.zeros((T, N, A), np.float32)
    done = np.zeros((T, N), bool)
    num_steps = np.zeros((T, N), np.int32)
    real = np.zeros((T, N, A), np.float32)
    state = np.zeros((T, N, 10), np.float64)
    term_idx, term_obs = [], []
    for t in range(T):
        o, r, d, info = envs.step(actions[t])
        obs[t] = np.stack(split(o))
        rew[t] = np.asarray(r).reshape(N, A)
        done[t] = np.asarray(d).reshape(N, -1).all(axis=1)
        for i in range(N):
            num_steps[t, i] = info[i]["num_steps"]
            real[t, i] = np.asarray(info[i]["real_reward"], np.float32).reshape(A)
            if done[t, i]:
                term_idx.append((t, i))
                term_obs.append(np.stack(split(info[i]["terminal_observation"])))  # (A, C, D, D)
        state[t] = np.array([ref_loader.game_state(e) for e in envs.envs], np.float64)
    assert obs.dtype == np.uint8
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(
        path, env_id=env_id, dim=dim, frame_stack=0 if fs is None else fs, serves=serves, actions=actions,
        reset_obs=reset_obs, state0=state0, obs=obs, rew=rew, done=done, num_steps=num_steps, real_reward=real,
        state=state, term_idx=np.array(term_idx, np.int32).reshape(-1, 2),
        term_obs=np.array(term_obs, np.uint8) if term_obs else np.zeros((0,) + reset_obs[:, 0].shape, np.uint8),
        serve_count=np.array(inj.count, np.int32))
    print("%-24s T=%d N=%d dones=%d max_rounds=%d hits(vy changes)=%d  %d KiB" % (
        name, T, N, int(done.sum()), int(state[:, :, 8].max()),
        int((np.abs(np.diff(np.abs(state[:, :, 3]), axis=0)) > 1e-12).sum()), os.path.getsize(path) // 1024))


def raw_frame_case():
    """Raw 210x160x3 frames (agent 0 and the agent-1 mirrored copy) for a few game states:
    pins the renderer restatement independently of the cv2 stage."""
    ref_loader.install()
    import competitive_rl.pong.base_pong_env as B
    serves = make_serve_table(1, 64, seed=3)
    B.random = ref_loader.ServeInjector(serves)
    env = B.PongDoublePlayerEnv(max_num_rounds=21)
    env.reset()
    rng = np.random.default_rng(5)
    frames0, frames1, states = [], [], []
    for t in range(600):
        (o0, o1), _, d, _ = env.step(tuple(int(x) for x in rng.integers(0, 3, 2)))
        if t % 25 == 0 or t in (1, 2, 3):
            g = env._game
            frames0.append(o0.copy())
            frames1.append(o1.copy())
            states.append((g._ball._rect.x, g._ball._rect.y, g._left_bat._rect.y, g._right_bat._rect.y,
                           g._score_left, g._score_right))
        if d:
            env.reset()
    path = os.path.join(OUT, "pong_raw_frames.npz")
    np.savez_compressed(path, frames0=np.array(frames0), frames1=np.array(frames1), states=np.array(states, np.int32))
    print("pong_raw_frames %d frames %d KiB" % (len(frames0), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)
    if not only or "pong_raw_frames" in only:
        raw_frame_case()
