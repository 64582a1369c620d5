"""Generate tests/golden/pong_*_f32.npz: the reference's Pong path as a STOCK gym install runs it.  gym's
Box without a dtype defaults to float32 (pong/base_pong_env.py:22-24), so MaxAndSkipEnv's buffers are float32
(utils/atari_wrappers.py:106-115), cv2.cvtColor / cv2.resize take their float paths and the observation is the UNROUNDED
fp32 area sum -- except the frames that come from reset(), which MaxAndSkipEnv passes through un-pooled as the raw uint8
array (:162-163): those go through cv2's uint8 path and enter the float32 stack as rounded integers.
Build container only (real cv2; gym / pygame stand-ins with GYM_SHIM_BOX_FLOAT32=1)."""
import os
import sys

import numpy as np

os.environ["GYM_SHIM_BOX_FLOAT32"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from gen_golden import make_actions  # noqa: E402
from pong_oracle import make_serve_table  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run_case(name, env_id, N, T, dim, fs, serve_seed, action_seed):
    double = env_id == "cPongDouble-v0"
    serves = make_serve_table(N, 200, seed=serve_seed)
    envs, inj = ref_loader.make_reference_vec_env(env_id, N, serves, resized_dim=dim, frame_stack=fs)
    actions = make_actions(T, N, double, action_seed, 0.03 if double else 0.0)

    def split(o):
        return list(o) if double else [o]
    reset_obs = np.stack(split(envs.reset()))
    assert reset_obs.dtype == np.float32, reset_obs.dtype
    obs = np.zeros((T,) + reset_obs.shape, np.float32)
    done = np.zeros((T, N), bool)
    term_idx, term_obs = [], []
    for t in range(T):
        o, r, d, info = envs.step(actions[t])
        obs[t] = np.stack(split(o))
        done[t] = np.asarray(d).reshape(N, -1).all(axis=1)
        for i in range(N):
            if done[t, i]:
                term_idx.append((t, i))
                term_obs.append(np.stack(split(info[i]["terminal_observation"])))
    frac = float((obs != np.rint(obs)).mean())
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, env_id=env_id, dim=dim, frame_stack=0 if fs is None else fs, serves=serves, actions=actions,
                        reset_obs=reset_obs, obs=obs, done=done, term_idx=np.array(term_idx, np.int32).reshape(-1, 2),
                        term_obs=np.array(term_obs, np.float32) if term_obs else np.zeros((0,) + reset_obs[:, 0].shape, np.float32))
    print("%-22s T=%d N=%d dones=%d non-integer pixels %.4f %%  %d KiB" % (name, T, N, int(done.sum()), 100 * frac, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    run_case("pong_double_84_f32", "cPongDouble-v0", 2, 220, 84, None, 31, 4242)
    run_case("pong_single_42_fs4_f32", "cPong-v0", 2, 220, 42, 4, 32, 4243)
